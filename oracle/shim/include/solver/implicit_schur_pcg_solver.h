// shim of include/solver/implicit_schur_pcg_solver.h: options + the one method implicit_schur_pcg_solver.cu defines
#pragma once
#include "common.h"
#include "edge/base_edge.h"
namespace MegBA {
template <typename T>
struct ImplicitSchurPCGSolver {
  ImplicitSchurPCGSolver(const ProblemOption &p, const SolverOption &s) : problemOption(p), solverOption(s) {}
  void solve(const BaseLinearSystem<T> &baseLinearSystem, const EdgeVector<T> &edges, const JVD<T> &jetEstimation);
  const ProblemOption &problemOption;
  const SolverOption &solverOption;
};
}  // namespace MegBA
