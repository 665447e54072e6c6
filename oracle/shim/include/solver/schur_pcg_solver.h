// shim of include/solver/{base_solver,schur_pcg_solver}.h: options + the one method the .cu file defines
#pragma once
#include "common.h"
#include "edge/base_edge.h"
namespace MegBA {
template <typename T>
struct SchurPCGSolver {
  SchurPCGSolver(const ProblemOption &p, const SolverOption &s) : problemOption(p), solverOption(s) {}
  void solve(const BaseLinearSystem<T> &baseLinearSystem);
  const ProblemOption &problemOption;
  const SolverOption &solverOption;
};
}  // namespace MegBA
