// megba/linear_system/schur_LM_linear_system.h -- BaseLinearSystem / SchurLMLinearSystem /
// ImplicitSchurLMLinearSystem (reference: include/linear_system/*.h).  The device arrays the reference's
// classes own (Hpp/Hll/Hpl/g/deltaX, include/linear_system/base_linear_system.h:17-59) live inside the
// ba_solver arena; this class composes the solver exactly like the reference's constructor injection.
#pragma once
#include <memory>
#include <utility>

#include "../common.h"
#include "../solver/schur_pcg_solver.h"

namespace MegBA {
template <typename T>
struct BaseLinearSystem {
  virtual ~BaseLinearSystem() = default;
  virtual ComputeKind computeKind() const { return solver->computeKind(); }
  const ProblemOption &problemOption;
  const std::unique_ptr<BaseSolver<T>> solver;

 protected:
  BaseLinearSystem(const ProblemOption &p, std::unique_ptr<BaseSolver<T>> s) : problemOption(p), solver(std::move(s)) {}
};

template <typename T>
struct SchurLMLinearSystem : public BaseLinearSystem<T> {
  SchurLMLinearSystem(const ProblemOption &p, std::unique_ptr<BaseSolver<T>> s) : BaseLinearSystem<T>(p, std::move(s)) {}
  ComputeKind computeKind() const override { return EXPLICIT; }
};

template <typename T>
struct ImplicitSchurLMLinearSystem : public BaseLinearSystem<T> {
  ImplicitSchurLMLinearSystem(const ProblemOption &p, std::unique_ptr<BaseSolver<T>> s) : BaseLinearSystem<T>(p, std::move(s)) {}
  ComputeKind computeKind() const override { return IMPLICIT; }
};
}  // namespace MegBA
