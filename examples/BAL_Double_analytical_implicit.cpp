#define BAL_T double
#define BAL_IMPLICIT 1
#include "BAL_analytical.inc"
