/*
 * megba_b200.h -- C-ABI of the Blackwell-native LM / Schur-PCG hot path.
 *
 * This is the drop-in boundary.  MegBA itself has no FFI
 * seam: its seams are C++ virtual interfaces between static libraries that own raw
 * device pointers.  Each entry point below replaces one of those seams and cites
 * it (paths relative to the MegBA checkout).  INTEGRATION.md shows the binding a
 * MegBA maintainer adds on the reference side; include/megba/ is the C++ facade
 * that reproduces the reference's class names on top of these functions.
 *
 * Conventions: plain C types only; every function returns BA_OK (0) or a negative
 * ba_status and never throws; `void*` numeric buffers are HOST arrays of the
 * solver's dtype (double for BA_F64, float for BA_F32); one ba_solver is bound to
 * one CUDA device and one rank (one process per GPU; several solvers in one
 * process, one host thread each, also work); all device memory is allocated in
 * ba_create; there is no CPU fallback -- without a CUDA device ba_create fails with
 * BA_ERR_NO_DEVICE.
 */
#ifndef MEGBA_B200_H_
#define MEGBA_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MEGBA_B200_ABI_VERSION 2

typedef enum {
  BA_OK = 0,
  BA_ERR_ARG = -1,       /* bad argument / inconsistent graph                    */
  BA_ERR_NO_DEVICE = -2, /* no usable CUDA device (the product has no CPU path)  */
  BA_ERR_CUDA = -3,      /* a CUDA runtime call or kernel failed                 */
  BA_ERR_NCCL = -4,      /* an NCCL call failed                                  */
  BA_ERR_STATE = -5,     /* call sequence violated (e.g. solve before linearize) */
  BA_ERR_ALLOC = -6
} ba_status;

typedef enum { BA_F64 = 0, BA_F32 = 1 } ba_dtype;
/* ComputeKind{EXPLICIT, IMPLICIT}, include/common.h:21 */
typedef enum { BA_EXPLICIT = 0, BA_IMPLICIT = 1 } ba_compute_kind;

/* How the observations are laid out in HBM and what the S*p pass streams per observation.
 * 0 selects the default (ELL layout, factored stream) for both compute kinds; the other values exist for
 * measurements and parity tests.  BA_LAYOUT_TILE stores what the compute kind names (Hpl blocks / Jacobians).   */
typedef enum { BA_LAYOUT_AUTO = 0, BA_LAYOUT_TILE = 1, BA_LAYOUT_ELL = 2 } ba_layout_kind;
typedef enum {
  BA_STORE_AUTO = 0,
  BA_STORE_W = 1,    /* 9x3 Hpl block per observation, 27 scalars (tile layout, explicit kind only)            */
  BA_STORE_J = 2,    /* 2x12 Jacobian, 24 scalars                                                               */
  BA_STORE_FACT = 3  /* dr/dX (2x3) + normalised image point (2), 8 scalars: W = Jc^T Omega Jp is applied through
                        a per-camera 9x9 block-diagonal map, see ba_internal.h (ELL layout only)                */
} ba_operator_store;
#define BA_LAYOUT_DEFAULT BA_LAYOUT_ELL

/* ProblemOption (include/common.h:44-53) restricted to what the hot path needs.
 * deviceUsed[i] becomes one solver with device = deviceUsed[i], rank = i.         */
typedef struct {
  int32_t dtype;          /* ba_dtype                                              */
  int32_t compute_kind;   /* ba_compute_kind                                       */
  int32_t device;         /* CUDA ordinal                                          */
  int32_t rank;           /* 0 .. world_size-1                                     */
  int32_t world_size;     /* number of GPUs = ProblemOption::deviceUsed.size()     */
  int32_t layout;         /* ba_layout_kind, 0 = default                           */
  int32_t operator_store; /* ba_operator_store, 0 = default                        */
  int32_t reserved[1];
} ba_options;

/* SolverOption::SolverOptionPCG, include/common.h:27-33 */
typedef struct {
  int32_t max_iter;
  int32_t reserved;
  double tol;
  double refuse_ratio;
} ba_pcg_options;

/* AlgoOption::AlgoOptionLM, include/common.h:35-42 */
typedef struct {
  int32_t max_iter;
  int32_t reserved;
  double initial_region;
  double epsilon1;
  double epsilon2;
} ba_lm_options;

/* One row per LM iteration; row 0 is the "Start with error" line
 * (src/algo/lm_algo.cu:149-150,190-191,207).                                      */
typedef struct {
  double cost;      /* F = sum r^2 after this iteration (console prints F/2)     */
  int32_t accepted; /* 1 accepted, 0 "Iter k failed", -1 stopped on step size    */
  int32_t pcg_iters;
  double region;
  double dx_norm;
  double rho;
  double elapsed_ms; /* cumulative, like the reference's trace                   */
} ba_lm_trace;

/* Device-side timing / launch accounting since the last ba_reset_stats.
 * ms values are CUDA-event times on the solver's stream.                          */
typedef struct {
  int64_t launches;          /* kernels of this library launched                 */
  int64_t lm_iters;          /* LM iterations executed (accepted + rejected)     */
  int64_t pcg_iters;         /* PCG iterations executed (n of schur_pcg :405)    */
  int64_t schur_apply_calls; /* launches of the S*p edge kernel                  */
  double schur_apply_ms;     /* summed device time of the TIMED launches (timing on) */
  int64_t linearize_calls;
  double linearize_ms;       /* summed device time of ba_linearize (timing on)   */
  double pcg_loop_ms;        /* device time inside the PCG do-while bodies       */
  double solve_ms;           /* device time of ba_solve (inverse+MakeV+PCG+SolveW)*/
  double pcg_cam_ms;         /* summed device time of the TIMED camera-space PCG kernel launches   */
  double schur_apply_timed;  /* how many S*p launches carry an event pair (see ba_set_timing)      */
  double pcg_cam_timed;      /* same for the camera-space kernel                                   */
  double reserved[1];
} ba_stats;

typedef struct ba_solver ba_solver;

int ba_abi_version(void);

/* Host-only helper, no GPU needed.  Partition rule replacing MemoryPool::getItemNum
 * (include/resource/memory_pool.h:48-63) + the device walk in BaseProblem::appendEdge
 * (src/problem/base_problem.cpp:116-121): points are cut into world_size contiguous
 * ranges balanced by edge count, so every edge of a point lives on one rank.
 * pt_degree[nP] = observations per point; pt_begin[world_size+1] receives the cuts. */
int ba_partition_points(int64_t nP, const int64_t *pt_degree, int32_t world_size,
                        int64_t *pt_begin);

/* Host-only, no GPU needed: the observation layout ba_create builds for (opt->rank, opt->world_size) -- the
 * counterpart of what EdgeVector::buildPositionContainer (src/edge/base_edge.cpp:224-262) and
 * SchurLMLinearSystem::buildIndex (src/linear_system/schur_LM_linear_system.cpp:53-84) leave in the reference's
 * CSR index arrays, exposed so that tools and the no-GPU tests can inspect it.  Every output pointer may be NULL;
 * call once with only `sizes` to learn the array lengths.
 *   tile[n_tiles][4]     = {first slot (== 256 * tile index), observations, first local point, points}
 *   edge_of_slot[n_slots] = observation number stored in the slot, -1 for padding
 *   slot_cam / slot_lp[n_slots] = camera and tile-local point index of the slot
 *   pt_slot[n_local_points]    = first slot of each local point (its observations are contiguous, camera-ascending)
 *   chunk[n_chunks][4]   = {camera, begin, end, 0}: ranges of the camera-major view, at most 1024 entries each
 *   cm_slot[n_local_edges]     = slot of each camera-major entry                                              */
typedef struct ba_layout_sizes {
  int64_t n_tiles, n_slots, n_local_edges, n_local_points, n_chunks, pt_begin;
} ba_layout_sizes;
int ba_layout(const ba_options *opt, int64_t nC, int64_t nP, int64_t nE, const int32_t *cam_idx,
              const int32_t *pt_idx, ba_layout_sizes *sizes, int32_t *tile, int64_t *edge_of_slot,
              int32_t *slot_cam, int32_t *slot_lp, int32_t *pt_slot, int32_t *chunk, int32_t *cm_slot);

/* The same for the ELL layout (the default).  Local points are sorted by observation count, descending ("device
 * order"); points with more than 32 observations come first and get a warp each, the rest are cut into slices of
 * 32 points (lane = point).  Observation r of the slice's l-th point sits in slot (row0 + r) * 32 + l; observation
 * i of a heavy point in slot row0 * 32 + i.
 *   warp[n_warps][4]      = {first row, rows, first point (device order), kind | points << 8 | stage row << 16},
 *                           kind 1 = heavy point; stage row = the slice's place in its CTA's shared-memory stage
 *   perm[n_local_points]  = local point (global id - pt_begin) at each device position
 *   edge_of_slot / slot_cam[n_slots] = observation number (-1 padding) and camera (-1 padding) of the slot
 *   chunk / cm_slot       = camera-major view as in ba_layout                                                  */
typedef struct ba_ell_sizes {
  int64_t n_warps, n_rows, n_slots, n_local_edges, n_local_points, n_chunks, pt_begin, n_heavy;
} ba_ell_sizes;
int ba_layout_ell(const ba_options *opt, int64_t nC, int64_t nP, int64_t nE, const int32_t *cam_idx,
                  const int32_t *pt_idx, ba_ell_sizes *sizes, int32_t *warp, int32_t *perm, int64_t *edge_of_slot,
                  int32_t *slot_cam, int32_t *chunk, int32_t *cm_slot);

/* BaseProblem::appendVertex/appendEdge + buildIndex (src/problem/base_problem.cpp:
 * 82-125,183-214), EdgeVector::buildPositionContainer/allocateResource
 * (src/edge/base_edge.cpp:224-262, src/edge/base_edge.cu:47-103) and
 * SchurLMLinearSystem::buildIndex/allocateResourceCUDA
 * (src/linear_system/schur_LM_linear_system.cpp:53-84, .cu:29-110).
 * cam_idx/pt_idx: vertex numbers of each observation (any order, every rank passes
 * the full list); obs[nE][2] in the solver dtype.
 * Limits (BA_ERR_ARG otherwise): 0 < nC < 2^24, 0 < nP <= 2^30, and fewer than 2^31
 * observation slots per rank after padding (256-slot tiles / 32-slot ELL rows).
 * PCG: max_iter < 1 behaves like 1 (the reference's do-while, schur_pcg_solver.cu:265-407).   */
int ba_create(const ba_options *opt, int64_t nC, int64_t nP, int64_t nE,
              const int32_t *cam_idx, const int32_t *pt_idx, const void *obs,
              ba_solver **out);
void ba_destroy(ba_solver *s);
const char *ba_last_error(const ba_solver *s);

/* HandleManager NCCL set-up (src/resource/handle_manager.cpp:17-21), one process per
 * GPU: rank 0 calls ba_comm_unique_id, ships the 128 bytes to every rank by any
 * means, every rank calls ba_comm_init.  Not needed for world_size == 1.           */
int ba_comm_unique_id(void *id128);
int ba_comm_init(ba_solver *s, const void *id128);

/* All ranks are solvers of THIS process (one host thread per rank afterwards, as with ba_comm_init): no NCCL; the
 * ranks may share a device, so the partition count can differ from the device count (e.g. 4 shards on one GPU).
 * solvers[r] must have been created with rank = r, world_size = n.  Call once, from one thread.                  */
int ba_comm_init_local(ba_solver **solvers, int32_t n);

/* BaseProblem::setAbsolutePosition (H2D of x, src/problem/base_problem.cpp:217-247)
 * and writeBack (:250-272).  cam[nC][9], pt[nP][3].                                */
int ba_set_params(ba_solver *s, const void *cam, const void *pt);
int ba_get_params(ba_solver *s, void *cam, void *pt);

/* Device-pointer variants of the two calls above -- the reference's seams hand raw device pointers around
 * (include/solver/base_solver.h:28-32, src/solver/schur_pcg_solver.cu:599-608), and a caller that already holds x
 * on the GPU must not bounce through the host.  cam / pt are DEVICE buffers on the solver's device, same shapes and
 * vertex order as above; nothing is synchronised with the host.  `stream` (a cudaStream_t, may be NULL) is the
 * caller's stream: the copies are ordered after the work already enqueued on it, and work enqueued on it afterwards
 * is ordered after them.  ba_get_params_device is for single-rank solvers.                                        */
int ba_set_params_device(ba_solver *s, const void *cam_dev, const void *pt_dev, void *stream);
int ba_get_params_device(ba_solver *s, void *cam_dev, void *pt_dev, void *stream);

/* Optional per-edge 2x2 information matrices, info[nE][4] row-major
 * (BaseEdge::setInformation; JMulInfo/makeHSchurWithInfo,
 * src/edge/build_linear_system.cu:149-239).  NULL restores identity.  In the default
 * layout both compute kinds honour them (H = J^T Omega J as in the reference's explicit
 * build; its implicit build has an empty branch there, build_implicit_linear_system.cu:
 * 272-274); the tile layout refuses them in implicit mode.                        */
int ba_set_information(ba_solver *s, const void *info);

/* Fixed vertices (BaseVertex::fixed, include/vertex/base_vertex.h:34 -- declared upstream, never honoured there,
 * src/edge/base_edge.cpp:27-36).  cam_fixed[nC] / pt_fixed[nP]: non-zero = the vertex keeps its parameters; either
 * pointer may be NULL (none fixed).  Its Jacobian columns leave the normal equations (identity diagonal block, zero
 * gradient, no coupling): the PCG iterates are those of the system reduced to the free vertices.  Default layout
 * only.  Call before ba_linearize.                                                                                 */
int ba_set_fixed(ba_solver *s, const uint8_t *cam_fixed, const uint8_t *pt_fixed);

/* User-defined residuals: the reference's plug-in seam BaseEdge::forward() on JetVectors (include/edge/base_edge.h:43,
 * include/operator/jet_vector.h:23-171, one kernel launch per operator).  Here the residual is a register bytecode --
 * the facade records forward() once on a tape and compiles it (include/megba/operator/jet_vector.h) -- from which the
 * library generates ONE fused kernel (forward-mode automatic differentiation written out as straight-line code, built
 * with NVRTC for this GPU; a device interpreter of the same program is the fallback).  The Jacobians go into the
 * Jacobian operator stream, everything downstream (assembly, Schur PCG, LM) is the path of the built-in residual.
 *   code[n_instr][6] = {op, dst, a, b, c, d} (register numbers < n_slots <= BA_PROGRAM_MAX_SLOTS);
 *   ops: 0 INPUT (a = input: 0..8 camera, 9..11 point, 12..13 observation), 1 CONST (a = index into consts),
 *        2 ADD, 3 SUB, 4 MUL, 5 DIV, 6 NEG, 7 SQRT, 8 SIN, 9 COS, 10 SELGT (a > b ? c : d), 11 ABS;
 *   out_slot[2] = registers holding the two residual components at the end.
 * Instructions whose value depends on the camera alone are evaluated once per camera, not once per observation
 * (the library splits the program).  n_instr = 0 restores the built-in analytical BAL residual.  Default layout only. */
#define BA_PROGRAM_MAX_SLOTS 48
int ba_set_residual_program(ba_solver *s, int32_t n_instr, const uint8_t *code, int32_t n_const, const double *consts,
                            int32_t n_slots, const int32_t *out_slot);

/* How the residual of this handle is evaluated: 0 the built-in analytical BAL kernels, 1 the device interpreter,
 * 2 a kernel generated from the program and compiled with NVRTC for this GPU when ba_set_residual_program was called
 * (the default when NVRTC is usable; BA_PROG_JIT=0 in the environment selects the interpreter).                       */
int ba_residual_program_mode(ba_solver *s);

/* Host-only (no GPU needed): the CUDA source ba_set_residual_program generates for this program, written to src (at most
 * src_cap bytes incl. the terminator; *src_len = full length), compiled with NVRTC for sm_100a when cubin_bytes is not
 * NULL (*cubin_bytes = size of the image; the compiler log goes to ba_last_error(NULL) on failure).  dtype: BA_F64/BA_F32. */
int ba_residual_program_source(int32_t n_instr, const uint8_t *code, int32_t n_const, const double *consts, int32_t n_slots,
                               const int32_t *out_slot, int32_t dtype, char *src, int64_t src_cap, int64_t *src_len,
                               int64_t *cubin_bytes);

/* HP1.  EdgeVector::forward + buildLinearSystem (src/edge/base_edge.cpp:160-163,
 * src/geo/analytical_derivatives.cu:162-322, src/edge/build_linear_system.cu:243-423
 * or build_implicit_linear_system.cu:208-322) + computeResidualNorm
 * (src/algo/lm_algo.cu:26-51): fills Hpp, Hll, Hpl(=W) or J, g at the current
 * parameters; *cost = sum r^2 over all ranks.                                      */
int ba_linearize(ba_solver *s, double *cost);

/* LMLinearSystem::processDiag (src/linear_system/schur_LM_linear_system.cu:162-185).
 * The undamped blocks are kept; damping (1 + 1/region) is applied when the blocks
 * are consumed, so recover != 0 ("recoverDiag") needs no saved diagonal.           */
int ba_process_diag(ba_solver *s, double region, int recover);

/* HP2.  BaseLinearSystem::solve -> SchurPCGSolver::solve /
 * ImplicitSchurPCGSolver::solve (src/solver/schur_pcg_solver.cu:599-679,
 * src/solver/implicit_schur_pcg_solver.cu:661-720): block inverses, Schur RHS,
 * warm-started PCG, back-substitution.  Result in the solver's deltaX.            */
int ba_solve(ba_solver *s, const ba_pcg_options *opt, int32_t *iters, int32_t *done);

/* The same solve without host synchronisation: ba_solve_async enqueues all launches (max_iter PCG iterations; those
 * past the end of the reference's loop are no-ops on the device) on the solver's stream and returns; ba_solve_wait
 * blocks until they are done and reports the iteration count.  Exactly one ba_solve_wait per ba_solve_async, before
 * any other call that reads the step.                                                                             */
int ba_solve_async(ba_solver *s, const ba_pcg_options *opt);
int ba_solve_wait(ba_solver *s, int32_t *iters, int32_t *done);

/* ||deltaX||_2 and ||x||_2 (l2NormPow2, src/algo/lm_algo.cu:54-58,172-175).       */
int ba_step_norms(ba_solver *s, double *dx_norm, double *x_norm);

/* EdgeVector::update + forward at the trial point + computeResidualNorm +
 * computeRhoDenominator (src/edge/update.cu:45-62, src/algo/lm_algo.cu:61-126,
 * 180-185): *cost_new = sum r(x+dx)^2, *model_sq = sum (J dx + r)^2 (caller
 * subtracts F).  x itself is not modified.                                        */
int ba_eval_trial(ba_solver *s, double *cost_new, double *model_sq);

/* Accept: applyUpdate (x += dx, schur_LM_linear_system.cu:212-218), backup of
 * deltaX (:188-197); the caller then re-linearises.  *g_inf = ||g||_inf of the
 * CURRENT g (call after ba_linearize; linfNorm, lm_algo.cu:129-136,203-205).
 * Reject: rollback of deltaX (:200-209).                                          */
int ba_accept_step(ba_solver *s);
int ba_reject_step(ba_solver *s);
int ba_grad_inf_norm(ba_solver *s, double *g_inf);

/* LMAlgo::solveCUDA (src/algo/lm_algo.cu:139-223) on top of the calls above.
 * trace may be NULL; at most trace_cap rows are written, *n_trace = rows produced.
 * verbose != 0 prints the reference's console trace.
 * Two guards the reference lacks: a non-finite gain ratio (model decrease rounding to 0 at a
 * converged point) counts as a fully successful step, and the trust region is kept inside
 * [1e-20, 1e20] (the reference lets it underflow to 0 after ~45 consecutive rejections and
 * then damps with inf); neither can change an iteration the reference completes sanely.    */
int ba_lm_solve(ba_solver *s, const ba_lm_options *lm, const ba_pcg_options *pcg,
                ba_lm_trace *trace, int32_t trace_cap, int32_t *n_trace, int32_t verbose);

/* Parity / inspection access.  Buffers are host arrays in the solver dtype, indexed
 * in the CALLER's edge and vertex numbering, reference layouts:
 *   BA_BUF_HPP [nC][9][9]   equationContainers.csrVal[2]  (row-major blocks)
 *   BA_BUF_HLL [nP][3][3]   csrVal[3]
 *   BA_BUF_W   [nE][9][3]   csrVal[0] = Hpl block of each observation (explicit)
 *   BA_BUF_J   [nE][2][12]  JVD grad slabs (implicit mode, or via ba_eval_edges)
 *   BA_BUF_G   [9nC+3nP]    g
 *   BA_BUF_DX  [9nC+3nP]    deltaXPtr
 * On world_size > 1 vertex/edge rows owned by other ranks read as zero for HLL/W/J
 * and the point parts of G/DX.                                                     */
typedef enum { BA_BUF_HPP = 0, BA_BUF_HLL = 1, BA_BUF_W = 2, BA_BUF_J = 3, BA_BUF_G = 4, BA_BUF_DX = 5 } ba_buffer;
int ba_read(ba_solver *s, int32_t which, void *host_out);
int ba_write_dx(ba_solver *s, const void *dx); /* warm start for ba_solve */

/* Residuals r[nE][2] and (if J != NULL) Jacobians J[nE][2][12] at the current
 * parameters (analytical_derivatives.cu:162-285), caller's edge order.            */
int ba_eval_edges(ba_solver *s, void *r, void *J);

/* q = S p for a host vector p[9nC] with the current damped blocks:
 * one operator application of the PCG loop (schur_pcg_solver.cu:315-366).          */
int ba_schur_apply(ba_solver *s, const void *p, void *q);

/* on = 0: no events.  on = 1: CUDA-event pair around every hot-kernel launch.  on = N > 1: around the launches of
 * every N-th PCG iteration only (sampling keeps the event overhead out of the measured throughput).              */
int ba_set_timing(ba_solver *s, int32_t on);
int ba_reset_stats(ba_solver *s);
int ba_get_stats(ba_solver *s, ba_stats *out);
/* Algorithmic bytes of one S*p edge-kernel launch and one linearisation on THIS
 * rank (SURVEY.md section 8d formulas).                             */
int ba_algorithmic_bytes(ba_solver *s, double *schur_apply_bytes, double *linearize_bytes);
/* on != 0: ba_solve runs the whole PCG solve (MakeV + S*x0, the do-while, back-substitution) as ONE persistent
 * cooperative kernel -- edge phases and camera phases separated by grid barriers -- where it is available (default
 * layout and stream, ranks on distinct devices).  Off by default: on one B200 it measured 210 us per PCG iteration
 * against 194 us for two launches per iteration chained by programmatic dependent launch (profiles/README.md).
 * Results are identical up to the summation order of the scatter.  Environment override: BA_PERSIST=0/1.           */
int ba_set_persistent(ba_solver *s, int32_t on);
/* The CUDA stream (cudaStream_t) all of this solver's work is enqueued on.         */
void *ba_stream(ba_solver *s);

#ifdef __cplusplus
}
#endif
#endif /* MEGBA_B200_H_ */
