// BAL driver, single precision, with the residual written on JetVectors (automatic differentiation), as the
// reference's examples/BAL_Float.cpp: forward() is recorded once and runs on the GPU as one fused bytecode program.
#define BAL_T float
#define BAL_IMPLICIT 0
#define BAL_AUTODIFF 1
#include "BAL_analytical.inc"
