// megba/solver/implicit_schur_pcg_solver.h -- include-path compatibility with the reference
// (include/solver/implicit_schur_pcg_solver.h); the class lives in schur_pcg_solver.h.
#pragma once
#include "schur_pcg_solver.h"
