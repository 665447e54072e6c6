#define BAL_T double
#define BAL_IMPLICIT 0
#include "BAL_analytical.inc"
