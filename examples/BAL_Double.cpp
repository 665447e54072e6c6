// BAL driver with the residual written on JetVectors (automatic differentiation), as the reference's
// examples/BAL_Double.cpp:15-33: forward() is recorded once and runs on the GPU as one fused bytecode program.
#define BAL_T double
#define BAL_IMPLICIT 0
#define BAL_AUTODIFF 1
#include "BAL_analytical.inc"
