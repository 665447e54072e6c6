// megba/algo/lm_algo.h -- BaseAlgo / LMAlgo (reference: include/algo/base_algo.h, include/algo/lm_algo.h).
// The LM control law itself runs in ba_lm_solve (megba_b200/csrc/ba_api.cu), a restatement of
// LMAlgo::solveCUDA (src/algo/lm_algo.cu:139-223); this class only carries the options.
#pragma once
#include "../common.h"

namespace MegBA {
template <typename T>
struct BaseAlgo {
  virtual ~BaseAlgo() = default;
  virtual AlgoKind algoKind() const { return BASE_ALGO; }
  const ProblemOption &problemOption;
  const AlgoOption &algoOption;

 protected:
  BaseAlgo(const ProblemOption &p, const AlgoOption &a) : problemOption(p), algoOption(a) {}
};

template <typename T>
struct LMAlgo : public BaseAlgo<T> {
  AlgoKind algoKind() const override { return LM; }
  LMAlgo(const ProblemOption &p, const AlgoOption &a) : BaseAlgo<T>(p, a) {}
};
}  // namespace MegBA
