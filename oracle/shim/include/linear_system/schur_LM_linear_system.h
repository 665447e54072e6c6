// shim of include/linear_system/{base,LM,schur,schur_LM}_linear_system.h: the data contract the hot-path files read
// (base_linear_system.h:17-59, schur_linear_system.h:22-27).
#pragma once
#include <array>
#include <cstddef>
#include <vector>
#include "common.h"
namespace MegBA {
template <typename T>
struct BaseLinearSystem {
  explicit BaseLinearSystem(const ProblemOption &o) : problemOption(o) {}
  virtual ~BaseLinearSystem() = default;
  const ProblemOption &problemOption;
  std::vector<T *> deltaXPtr{};
  std::vector<T *> g{};
  std::array<int, 2> num{};
  std::array<int, 2> dim{};
};
template <typename T>
struct SchurLinearSystem : public BaseLinearSystem<T> {
  using BaseLinearSystem<T>::BaseLinearSystem;
  struct EquationContainer {
    std::array<int *, 2> csrRowPtr{};
    std::array<T *, 4> csrVal{};
    std::array<int *, 2> csrColInd{};
    std::array<std::size_t, 4> nnz{};
  };
  std::vector<EquationContainer> equationContainers{};
};
template <typename T>
struct SchurLMLinearSystem : public SchurLinearSystem<T> {
  using SchurLinearSystem<T>::SchurLinearSystem;
};
}  // namespace MegBA
