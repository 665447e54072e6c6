// megba/solver/schur_pcg_solver.h -- BaseSolver / SchurPCGSolver / ImplicitSchurPCGSolver
// (reference: include/solver/{base_solver,schur_pcg_solver,implicit_schur_pcg_solver}.h).
// The solve is ba_solve (Schur complement + block-Jacobi PCG on the GPU); these classes select
// explicit vs implicit and carry SolverOption.
#pragma once
#include "../common.h"

namespace MegBA {
template <typename T>
struct BaseSolver {
  virtual ~BaseSolver() = default;
  virtual SolverKind solverKind() const { return BASE_SOLVER; }
  virtual ComputeKind computeKind() const { return EXPLICIT; }
  const ProblemOption &problemOption;
  const SolverOption &solverOption;

 protected:
  BaseSolver(const ProblemOption &p, const SolverOption &s) : problemOption(p), solverOption(s) {}
};

template <typename T>
struct SchurPCGSolver : public BaseSolver<T> {
  SchurPCGSolver(const ProblemOption &p, const SolverOption &s) : BaseSolver<T>(p, s) {}
  SolverKind solverKind() const final { return PCG; }
};

template <typename T>
struct ImplicitSchurPCGSolver : public BaseSolver<T> {
  ImplicitSchurPCGSolver(const ProblemOption &p, const SolverOption &s) : BaseSolver<T>(p, s) {}
  SolverKind solverKind() const final { return PCG; }
  ComputeKind computeKind() const final { return IMPLICIT; }
};
}  // namespace MegBA
