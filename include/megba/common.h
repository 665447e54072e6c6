// megba/common.h -- option structs and enums of the g2o-style facade.
// API-compatible with the reference's include/common.h:17-63 (same names and fields, so user code that
// fills ProblemOption / SolverOption / AlgoOption compiles unchanged); everything behind it is this
// repository's C-ABI (include/megba_b200.h).
#pragma once
#include <cstdint>
#include <vector>

namespace MegBA {
enum Device { CPU, CUDA };
enum AlgoKind { BASE_ALGO, LM };
enum LinearSystemKind { BASE_LINEAR_SYSTEM, SCHUR };
enum ComputeKind { EXPLICIT, IMPLICIT };
enum SolverKind { BASE_SOLVER, PCG };

struct SolverOption {
  struct SolverOptionPCG {
    int maxIter{100};
    double tol{1e-1};
    double refuseRatio{1e0};
  } solverOptionPCG;
};

struct AlgoOption {
  struct AlgoOptionLM {
    int maxIter{20};
    double initialRegion{1e3};
    double epsilon1{1};
    double epsilon2{1e-10};
  } algoOptionLM;
};

struct ProblemOption {
  bool useSchur{true};
  Device device{Device::CUDA};
  std::vector<int> deviceUsed{};  // one solver (one host thread, one NCCL rank) per entry
  int N{-1};                      // kept for source compatibility; the analytical BAL edge has N = 12
  int64_t nItem{-1};              // kept for source compatibility; the edge count is taken from appendEdge
  AlgoKind algoKind{LM};
  LinearSystemKind linearSystemKind{SCHUR};
  SolverKind solverKind{PCG};
};

template <typename T> class BaseProblem;
template <typename T> struct BaseVertex;
template <typename T> class BaseEdge;
}  // namespace MegBA
