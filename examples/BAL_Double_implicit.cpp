// BAL driver with the residual written on JetVectors and the matrix-free (implicit) Schur complement, as the
// reference's examples/BAL_Double_implicit.cpp.
#define BAL_T double
#define BAL_IMPLICIT 1
#define BAL_AUTODIFF 1
#include "BAL_analytical.inc"
