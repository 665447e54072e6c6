// shim of include/linear_system/{base,LM,schur,schur_LM}_linear_system.h: the data contract the hot-path files read
// (base_linear_system.h:17-59, schur_linear_system.h:22-27).
#pragma once
#include <array>
#include <cstddef>
#include <vector>
#include "common.h"
#include "operator/jet_vector.h"
namespace MegBA {
template <typename T>
struct BaseLinearSystem {
  explicit BaseLinearSystem(const ProblemOption &o) : problemOption(o) {}
  virtual ~BaseLinearSystem() = default;
  virtual ComputeKind computeKind() const { return EXPLICIT; }
  // BaseLinearSystem::solve() forwards to the composed solver (src/linear_system/base_linear_system.cpp:34-45);
  // the driver implements it with the reference's SchurPCGSolver::solve
  virtual void solve() const = 0;
  virtual void solve(const EdgeVector<T> &, const JVD<T> &) const {}
  const ProblemOption &problemOption;
  std::vector<T *> deltaXPtr{};
  std::vector<T *> g{};
  std::array<int, 2> num{};
  std::array<int, 2> dim{};
  std::size_t getHessianShape() const { return (std::size_t)num[0] * dim[0] + (std::size_t)num[1] * dim[1]; }
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
// LM hooks (include/linear_system/LM_linear_system.h:13-41)
template <typename T>
struct LMLinearSystem : public SchurLinearSystem<T> {
  using SchurLinearSystem<T>::SchurLinearSystem;
  mutable std::vector<T *> deltaXPtrBackup{};
  mutable std::vector<T *> gBackup{};
  mutable std::vector<std::array<T *, 2>> extractedDiag{};
  virtual void processDiag(const AlgoStatus::AlgoStatusLM &lmAlgoStatus) const = 0;
  virtual void backup() const = 0;
  virtual void rollback() const = 0;
  virtual void applyUpdate(T *xPtr) const = 0;
};
// the members src/linear_system/schur_LM_linear_system.cu defines
template <typename T>
struct SchurLMLinearSystem : public LMLinearSystem<T> {
  using LMLinearSystem<T>::LMLinearSystem;
  void allocateResourceCUDA();
  void processDiag(const AlgoStatus::AlgoStatusLM &lmAlgoStatus) const override;
  void backup() const override;
  void rollback() const override;
  void applyUpdate(T *xPtr) const override;
  const void *solverHook{nullptr};  // driver-owned: the reference SchurPCGSolver used by solve()
  void solve() const override;      // defined in ref_driver.cu
};
}  // namespace MegBA
