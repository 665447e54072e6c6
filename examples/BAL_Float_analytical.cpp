#define BAL_T float
#define BAL_IMPLICIT 0
#include "BAL_analytical.inc"
