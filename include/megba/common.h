// megba/common.h -- the option structs and enums user code fills in before building a problem.
//
// Source-compatible with the reference's include/common.h:17-63: the same type and field names and the same
// defaults, so `MegBA::ProblemOption`, `SolverOption::solverOptionPCG`, `AlgoOption::algoOptionLM` in an existing
// MegBA program compile unchanged.  Everything behind them is this repository's C-ABI (include/megba_b200.h);
// the mapping is: deviceUsed[i] -> one ba_solver (device = deviceUsed[i], rank = i), solverOptionPCG ->
// ba_pcg_options, algoOptionLM -> ba_lm_options.
#pragma once
#include <cstdint>
#include <vector>

namespace MegBA {

// ---- kinds (the reference implements exactly one value of each besides the BASE_* placeholders)
enum Device { CPU, CUDA };                               // only CUDA has an implementation, upstream and here
enum ComputeKind { EXPLICIT, IMPLICIT };                 // stored Hpl blocks vs matrix-free operator
enum AlgoKind { BASE_ALGO, LM };
enum SolverKind { BASE_SOLVER, PCG };
enum LinearSystemKind { BASE_LINEAR_SYSTEM, SCHUR };

// ---- linear solver: block-Jacobi preconditioned CG on the Schur complement
struct SolverOption {
  struct SolverOptionPCG {
    int maxIter = 100;          // upper bound of the do-while loop
    double tol = 1e-1;          // absolute threshold on rho = r.z
    double refuseRatio = 1e0;   // stop and restore the previous iterate when rho > refuseRatio * min(rho)
  };
  SolverOptionPCG solverOptionPCG;
};

// ---- outer algorithm: Levenberg-Marquardt
struct AlgoOption {
  struct AlgoOptionLM {
    int maxIter = 20;             // counts rejected iterations too
    double initialRegion = 1e3;   // trust region; the damping factor is 1 + 1/region
    double epsilon1 = 1;          // stop when ||g||_inf <= epsilon1
    double epsilon2 = 1e-10;      // stop when ||dx|| <= epsilon2 (||x|| + epsilon1)
  };
  AlgoOptionLM algoOptionLM;
};

// ---- problem
struct ProblemOption {
  bool useSchur = true;
  Device device = Device::CUDA;
  std::vector<int> deviceUsed;                      // one host thread + one NCCL rank per entry
  int N = -1;                                       // gradient width of the reference's JetVectors; 12 for the BAL edge
  int64_t nItem = -1;                               // upstream needs the edge count up front; ignored here
  AlgoKind algoKind = LM;
  LinearSystemKind linearSystemKind = SCHUR;
  SolverKind solverKind = PCG;
};

template <typename T> struct BaseVertex;
template <typename T> class BaseEdge;
template <typename T> class BaseProblem;

}  // namespace MegBA
