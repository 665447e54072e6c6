// ba_internal.h -- device-side data model shared by ba_kernels.cu and ba_api.cu.
//
// HBM layout (one rank).  Observations ("edges") owned by this rank are sorted
// point-major and packed into TILES: a tile is a run of whole points whose edges fit
// one slab of TE slots; a point with more than TE edges gets a run of slabs of its
// own.  Slot s of slab b is edge-array index b*TE + i.  Per-slot arrays:
//   e_cam[slot] int32, e_lp[slot] uint16 (point index inside the tile),
//   e_obs[slot][2], W/J blocked SoA: value k of slot (b,i) at  b*K*TE + k*TE + i
//   (K = 27 for the explicit Hpl block, 24 for the implicit Jacobian), so a warp reads
//   32 consecutive scalars per k and a whole slab is one contiguous K*TE*sizeof(T) run.
// Points of this rank are numbered locally 0..nPl-1 (global id - pt_begin).
// A second, camera-major view of the same edges (cm_*) feeds the atomic-free
// accumulation of Hpp / g_c.
//
// ELL layout (the default; the tile layout above is kept as the measured baseline).  Local points are
// sorted by observation count (descending) -- "device order" -- and cut into SLICES of 32 points; a slice
// is the work of one warp, lane = point.  The observations of a slice are stored as D rows of 32 slots
// (row r, lane l = r-th observation of the slice's l-th point, D = largest count in the slice, padding
// slots carry camera -1), so that every per-observation stream is read with fully coalesced loads while
// all per-point quantities (X, Hll^-1, t, u) stay in the registers of ONE thread: no shared-memory
// segment sums, no block barriers.  A point with more than ELL_HEAVY observations gets a warp of its own
// (lane = observation, row r holds its observations 32r .. 32r+31).  Per-slot arrays:
//   r_cam[row][32] int32, r_obs[row][32][2], r_info[row][32][4] (optional),
//   operator stream, row-blocked:  value pair j of slot (row, l) at  ((row*K/2 + j)*32 + l)*2, i.e. a warp
//   reads K/2 consecutive 512-byte (fp64) runs per row and a slice is one contiguous run of D*K*32 scalars.
// Operator stream flavours (Dev::store):
//   ST_W (tile layout only)  K = 27  the Hpl block W = Jc^T Omega Jp of the observation
//   ST_J                     K = 24  the 2x12 Jacobian
//   ST_FACT (ELL only)       K = 8   Jp (2x3) and the normalised image point B (2): with the per-camera
//        block-diagonal map  T_c = diag(Jr(omega), R^T, E(f,k1,k2))  one has  Jc = [Jp (.)x X-form | ...]:
//          Jc x = Jp (psi x X + xt) + B (e0 + l2 e1 + l2^2 e2),   (psi, xt, e) = T_c x,   l2 = |B|^2
//          Jc^T z = T_c^T ( X x v, v, (B.z)(1, l2, l2^2) ),        v = Jp^T z
//        (Jr = right Jacobian of SO(3): R^T dR(omega)[x] = [Jr x]x), so the S*p pass streams 64 instead
//        of 216 bytes per observation and the camera kernel applies T_c / T_c^T to its 9-vectors.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ba {

constexpr int TE = 256;       // slots per slab == threads of the edge kernels
constexpr int CM_CHUNK = 1024;  // edges per camera-major work item
constexpr int CM_THREADS = 128;
constexpr int NTRI = 45;       // upper triangle of a 9x9 block
constexpr int NREP = 16;        // replicas of the camera-space scatter target (spreads REDs over L2 slices)
constexpr int AUX_GRID = 296;   // 2 CTAs per SM for the small grid-stride kernels
constexpr int ELL_HEAVY = 32;   // points with more observations than this get a warp of their own
#ifndef BA_ELL_WARPS
#define BA_ELL_WARPS 2
#endif
constexpr int ELL_WARPS = BA_ELL_WARPS;  // warps (slices) per CTA of the ELL kernels
constexpr int CAMC = 16;        // per-camera constants for the edge kernels: cam[9], c, a, b, a', b', small, pad
constexpr int CAMT = 27;        // per-camera transform blocks of the factored operator: Jr[9], R[9], E[9]
enum Store { ST_W = 1, ST_J = 2, ST_FACT = 3 };

template <typename T> struct CamPad;
template <> struct CamPad<double> { static constexpr int N = 10; using vec = double2; };  // 80 B rows
template <> struct CamPad<float> { static constexpr int N = 12; using vec = float4; };    // 48 B rows

struct PcgState {
  double rho, rho_prev, rho_min, alpha, pq;
  double tol, refuse;
  int n, max_iter, active, done, refused, pad;
};

// Peer-memory exchange descriptor of the fused allreduce inside the camera-space PCG kernel (world_size > 1).
// data[r] -> rank r's inbox [2 parity][ws senders][2 vectors][9nC] (IPC-mapped for r != rank), flag[r] -> its
// [2 parity][8 senders] 64-bit epoch flags.  Ranks PUSH their partial vectors and flags into the peers' memory and
// poll / read only their own.  ws == 1 disables the exchange.
struct PeerExchange {
  void *data[8];
  unsigned long long *flag[8];
  unsigned long long epoch;
  int ws, rank, parity;
  int pull;   // 0: push (write into the peers' inboxes, poll / read local memory); 1: pull (publish in the own slot and
              // raise the own flag locally, poll / read the peers' memory) -- same buffer layout, BA_XCHG=push|pull
};

// device scalar slots (double) read back by the host
enum Scal { S_COST = 0, S_TRIAL_COST = 1, S_MODEL_SQ = 2, S_DX2_PT = 3, S_X2_PT = 4, S_DX2_CAM = 5,
            S_X2_CAM = 6, S_GINF = 7, S_ANY_SMALL = 8, S_MODEL_PT = 9, S_MODEL_CAM = 10, S_XCHG_ERR = 15, S_COUNT = 16 };

template <typename T>
struct Dev {
  // sizes
  int nC, nPl, nT, nSlabs, nChunks, nEl;
  int implicit;
  int no_pdl;                   // several ranks share this device (in-process communicator): no programmatic dependent launch,
                                // an early-launched grid parked at griddepcontrol.wait would starve the peer it waits for
  // point-major static graph
  const int4 *tile;             // {slot0, ne, pt0, np}
  const int *e_cam;
  const unsigned short *e_lp;
  const T *e_obs;               // [slot][2]
  const T *e_info;              // [slot][4] or nullptr
  const int *pt_slot;           // [nPl] first slot of the point
  // ELL static graph (ell != 0; the tile fields above are then unused)
  int ell, store, nW, nRows, nPlPad, nCta, tma_rows;
  const int4 *wdesc;            // per warp: {first row, rows D, first point (device order), kind | np << 8 | stage row << 16}
  const int2 *cdesc;            // CTAs of the TMA-staged edge pass: {first warp, warps}
  const int *r_cam;             // [nRows][32] camera of the slot, -1 = padding
  const T *r_obs;               // [nRows][32][2]
  const T *r_info;              // [nRows][32][4] or nullptr
  T *camc;                      // [nC][CAMC]
  T *camc2;                     // [nC][CAMC] the same for the trial cameras cam + dx_c
  T *camT;                      // [nC][CAMT]
  T *HllInvP, *wP, *XP;         // planes [9][nPlPad], [3][nPlPad], [3][nPlPad] (device order)
  // fixed vertices (BaseVertex::fixed): their columns leave the system, their parameters never change
  const unsigned char *cam_fixed;  // [nC] or nullptr
  const unsigned char *pt_fixed;   // [nPl] device order, or nullptr
  // user-defined residual (ba_set_residual_program, ba_program.cuh); nullptr = the analytical BAL model
  const uint2 *prog_code;
  const T *prog_const;
  int prog_n, prog_out0, prog_out1;
  const uint2 *prog_cam_code;  // camera-only part, run once per camera (k_prog_cam); fills prog_tab
  int prog_cam_n, prog_nf;     // its length; frontier values per camera
  T *prog_tab, *prog_tab2;     // [nC][prog_nf][10] at the current / at the trial parameters
  int prog_jit;                // 1: the point side of the linearisation and the trial cost run in the NVRTC-generated
                               // kernels (ba_program_jit.h); k_lin_cm<PROG> / k_trial_ell<PROG> read what they left
  int split;           // warps sharing one slice in the S*p pass (1, 2 or 4; ImplT::init chooses by the shard's size)
  const int *cm_slot;  // [nEl] slot of each camera-major entry: the camera side reads r and J the point side left
  T *r_slots;          // [nSlots][2] residuals of the current linearisation
  const int *slot_cm;  // [nSlots] inverse of cm_slot (-1: empty slot), generated kernels only
  T *cm_jr;            // [nEl][20] camera-major copy {J row 0 camera part, J row 1 camera part, r} written by prog_fill
  // camera-major static graph
  const int4 *chunk;            // {cam, beg, end, 0}
  const int *cam_chunk_ptr;     // [nC+1]
  const int *cm_pt;             // [nEl] local point index
  const T *cm_obs;              // [nEl][2]
  const T *cm_info;             // [nEl][4] or nullptr
  // parameters
  T *cam;                       // [nC][9]
  T *pt;                        // [nPl][3]
  // linear system (undamped)
  T *Hpp, *gc;                  // [nC][81], [nC][9]   (contiguous: Hpp then gc)
  T *Hll, *gl;                  // [nPl][9], [nPl][3]
  T *WJ;                        // blocked SoA, K = 27 or 24
  T *cm_part;                   // [nChunks][54]
  // solve
  T *HppD, *HppInv;             // [nC][81] damped block and its inverse
  T *HllInv, *w;                // [nPl][9], [nPl][3]
  T *dxc, *dxl;                 // step: [9nC], [3nPl]
  T *dxc_bk, *dxl_bk;           // LM backup of the step (warm start after a reject)
  T *sA, *sB;                   // sA[NREP][9nC], sB[2][NREP][9nC]: edge-kernel scatter targets (replicated;
                                // sB has two sets used alternately so the idle one can be cleared by the edge pass)
  T *r, *p, *q, *z, *xbk;       // PCG vectors [9nC]
  T *xpad;                      // [nC][CamPad::N] padded copy of the vector the edge kernel gathers
  cudaTextureObject_t wj_tex;   // the operator stream as a linear texture of Vec2<T> texels (same experiment)
  cudaTextureObject_t camc_tex, camc2_tex;  // camc / camc2 as linear textures of 16-byte texels
  cudaTextureObject_t xpad_tex; // the same buffer as a linear texture of 16-byte texels (experiment BA_ELL_TEX: gather
                                // through the TEX pipe instead of the LSU pipe)
  PcgState *st;
  int *pk_counters;             // [4] slice counters of the persistent PCG kernel
  double *pk_partials;          // [2][#SMs] per-CTA partial dot products of the persistent PCG kernel
  double *pk_timers;            // [4] device-timer sums of the persistent PCG kernel (edge ns, edge phases, camera ns, camera phases)
  double *scal;                 // [S_COUNT]
  double *partials;             // per-CTA partial sums
  double *model_partials;       // [nW] or nullptr: the back-substitution pass also leaves the point part of the model
                                // decrease  dx^T H dx - 2 g^T dx  there (computeRhoDenominator without an edge pass)
  double damp;                  // 1/region
};

// ---- launchers (ba_kernels.cu); all enqueue on `st` and return cudaGetLastError ----
template <typename T> cudaError_t launch_linearize(const Dev<T> &d, cudaStream_t st);
template <typename T> cudaError_t launch_lin_cm(const Dev<T> &d, cudaStream_t st, int dispatch);
enum SchurMode { SM_FULL = 0, SM_MAKEV = 1, SM_BACKSUB = 2, SM_INIT = 3 };
template <typename T> cudaError_t launch_schur_edge(const Dev<T> &d, int mode, const T *xin, T *sout, T *zero_set,
                                                    int check_active, cudaStream_t st);
template <typename T> cudaError_t launch_prep(const Dev<T> &d, cudaStream_t st);
template <typename T> cudaError_t launch_pcg_init(const Dev<T> &d, T *sB, int max_iter, double tol, double refuse, int nrep,
                                                  const PeerExchange &px, cudaStream_t st);
template <typename T> cudaError_t launch_pcg_step(const Dev<T> &d, T *sB, int nrep, const PeerExchange &px, cudaStream_t st);
template <typename T> bool pcg_reg_kernel_available(const Dev<T> &d);
template <typename T> cudaError_t launch_fold_replicas(const Dev<T> &d, T *s, int nvec, cudaStream_t st);
template <typename T> cudaError_t launch_pad_cam(const Dev<T> &d, const T *src, cudaStream_t st);
int pcg_cluster_size();
template <typename U> cudaError_t launch_sum_ranks(U *dst, const void *const *srcs, int nr, size_t n, int op, cudaStream_t st);
template <typename T> cudaError_t launch_trial(const Dev<T> &d, cudaStream_t st);
template <typename T> cudaError_t launch_norms(const Dev<T> &d, cudaStream_t st);
template <typename T> cudaError_t launch_apply_step(const Dev<T> &d, cudaStream_t st);
template <typename T> cudaError_t launch_ginf(const Dev<T> &d, cudaStream_t st);
template <typename T> cudaError_t launch_eval_edges(const Dev<T> &d, T *r_slots, T *J_slots, cudaStream_t st);
template <typename T> cudaError_t launch_cam_matvec_sub(const Dev<T> &d, const T *p, const T *s, T *q, cudaStream_t st);
int launches_of_linearize();
int launches_of_prep();
// ---- ELL layout (ba_ell.cu)
template <typename T> cudaError_t launch_cam_consts(const Dev<T> &d, cudaStream_t st);
template <typename T> cudaError_t launch_linearize_ell(const Dev<T> &d, cudaStream_t st);
template <typename T> cudaError_t launch_schur_ell(const Dev<T> &d, int mode, const T *xin, T *sout, T *zero_set,
                                                   int check_active, cudaStream_t st);
template <typename T> cudaError_t launch_prep_ell(const Dev<T> &d, cudaStream_t st);
template <typename T> cudaError_t launch_trial_ell(const Dev<T> &d, cudaStream_t st);
template <typename T> void launch_prog_cam(const Dev<T> &d, int trial, cudaStream_t st);
template <typename T> cudaError_t launch_reduce_partials(const Dev<T> &d, int ncomp, int scal_index, cudaStream_t st);
template <typename T> cudaError_t launch_trial_cost_ell(const Dev<T> &d, cudaStream_t st);
template <typename T> cudaError_t launch_model_terms(const Dev<T> &d, cudaStream_t st);
template <typename T> cudaError_t launch_eval_edges_ell(const Dev<T> &d, T *r_slots, T *J_slots, cudaStream_t st);
template <typename T> cudaError_t launch_pad_cam_T(const Dev<T> &d, const T *src, cudaStream_t st);
template <typename T> cudaError_t launch_permute3(const T *src, T *dst, const int *perm, int n, int gather, cudaStream_t st);
int launches_of_linearize_ell();
template <typename T> bool pcg_persistent_available(const Dev<T> &d);
template <typename T> cudaError_t launch_pcg_persistent(const Dev<T> &d, int max_iter, double tol, double refuse, const PeerExchange &px, cudaStream_t st);
bool pdl_enabled();

}  // namespace ba
