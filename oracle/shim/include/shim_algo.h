// shim_algo.h -- stand-in declarations for the reference's include/algo/{base_algo,lm_algo}.h.
//
// Only what src/algo/lm_algo.cu (LMAlgo<T>::solveCUDA and its scalar kernels) and src/algo/lm_algo.cpp (the LMAlgo
// constructor) need in order to compile unmodified: the option / status members they read and the two member functions
// they define.  The dispatching BaseAlgo::solve of the reference (src/algo/base_algo.cpp) is not part of the hot path
// and is not declared here; the shim driver calls solveCUDA directly.
#pragma once
#include "common.h"

namespace MegBA {

template <typename T>
struct BaseAlgo {
  // state shared with the derived algorithm: the user options (held by reference, as upstream does) and the LM status
  // block (trust region + recoverDiag flag) that LMAlgo::solveCUDA mutates while it iterates
  const ProblemOption &problemOption;
  const AlgoOption &algoOption;
  AlgoStatus algoStatus{};

  BaseAlgo(const ProblemOption &problem, const AlgoOption &algo) : problemOption(problem), algoOption(algo) {}
  virtual ~BaseAlgo() {}

  virtual AlgoKind algoKind() const { return BASE_ALGO; }
  virtual void solveCUDA(const BaseLinearSystem<T> &, const EdgeVector<T> &, T *) = 0;
};

template <typename T>
struct LMAlgo : BaseAlgo<T> {
  // defined by the reference: the constructor in src/algo/lm_algo.cpp:11-16 (sets region = initialRegion),
  // the loop in src/algo/lm_algo.cu:138-223
  LMAlgo(const ProblemOption &problemOption, const AlgoOption &algoOption);
  void solveCUDA(const BaseLinearSystem<T> &baseLinearSystem, const EdgeVector<T> &edges, T *xPtr) override;
  AlgoKind algoKind() const override { return LM; }
};

}  // namespace MegBA
