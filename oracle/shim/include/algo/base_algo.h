// shim of include/algo/base_algo.h: options + status the LM loop reads (base_algo.h:13-36)
#pragma once
#include "common.h"
namespace MegBA {
template <typename T>
struct BaseAlgo {
  virtual ~BaseAlgo() = default;
  virtual AlgoKind algoKind() const { return BASE_ALGO; }
  virtual void solveCUDA(const BaseLinearSystem<T> &baseLinearSystem, const EdgeVector<T> &edges, T *xPtr) = 0;

 protected:
  BaseAlgo(const ProblemOption &p, const AlgoOption &a) : problemOption(p), algoOption(a) {}
  const ProblemOption &problemOption;
  const AlgoOption &algoOption;
  AlgoStatus algoStatus{};
};
}  // namespace MegBA
