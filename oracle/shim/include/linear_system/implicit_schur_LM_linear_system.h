// shim of include/linear_system/implicit_schur_{linear_system,LM_linear_system}.h: the container the implicit hot-path
// files read (implicit_schur_linear_system.h:23-26); the LM hooks of the implicit system are not compiled here (its
// processDiag twin, implicit_schur_LM_linear_system.cu:93-116, repeats schur_LM_linear_system.cu:162-185 which is).
#pragma once
#include <array>
#include <cstddef>
#include <vector>
#include "linear_system/schur_LM_linear_system.h"
namespace MegBA {
template <typename T>
struct ImplicitSchurLinearSystem : public BaseLinearSystem<T> {
  using BaseLinearSystem<T>::BaseLinearSystem;
  ComputeKind computeKind() const override { return IMPLICIT; }
  struct EquationContainer {
    std::array<T *, 2> csrVal{nullptr, nullptr};
    std::array<std::size_t, 2> nnz{0, 0};
  };
  std::vector<EquationContainer> implicitEquationContainers;
};
template <typename T>
struct ImplicitSchurLMLinearSystem : public ImplicitSchurLinearSystem<T> {
  using ImplicitSchurLinearSystem<T>::ImplicitSchurLinearSystem;
  void solve() const override {}
};
}  // namespace MegBA
