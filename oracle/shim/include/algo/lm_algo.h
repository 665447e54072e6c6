// shim of include/algo/lm_algo.h (lm_algo.h:11-21): the constructor is defined by the reference's src/algo/lm_algo.cpp,
// solveCUDA by the reference's src/algo/lm_algo.cu -- both compiled unmodified.
#pragma once
#include "base_algo.h"
namespace MegBA {
template <typename T>
struct LMAlgo : public BaseAlgo<T> {
  AlgoKind algoKind() const override { return LM; }
  explicit LMAlgo(const ProblemOption &problemOption, const AlgoOption &algoOption);
  void solveCUDA(const BaseLinearSystem<T> &baseLinearSystem, const EdgeVector<T> &edges, T *xPtr) override;
};
}  // namespace MegBA
