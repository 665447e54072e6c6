// ba_api.cu -- host side of the C-ABI in include/megba_b200.h: graph preprocessing
// (point-major tiling, camera-major view, point partition), the device arena, the
// Schur-PCG driver, the LM loop and NCCL plumbing.  No CPU compute path exists here:
// every numeric result is produced by the kernels in ba_kernels.cu.
#include <cuda_runtime.h>
#include <nccl.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "../../include/megba_b200.h"
#include "ba_internal.h"
#include "ba_program.cuh"
#include "ba_program_host.h"
#include "ba_program_jit.h"

namespace ba {

#define BA_CUDA(expr)                                                                         \
  do {                                                                                        \
    cudaError_t e__ = (expr);                                                                 \
    if (e__ != cudaSuccess) return fail(BA_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e__)); \
  } while (0)
#define BA_NCCL(expr)                                                                          \
  do {                                                                                         \
    ncclResult_t e__ = (expr);                                                                 \
    if (e__ != ncclSuccess) return fail(BA_ERR_NCCL, "%s: %s", #expr, ncclGetErrorString(e__)); \
  } while (0)
#define BA_TRY(expr)              \
  do {                            \
    int rc__ = (expr);            \
    if (rc__ != BA_OK) return rc__; \
  } while (0)

template <typename T> struct NcclType;
template <> struct NcclType<double> { static constexpr ncclDataType_t v = ncclDouble; };
template <> struct NcclType<float> { static constexpr ncclDataType_t v = ncclFloat; };

// In-process communicator (ba_comm_init_local): the ranks are solvers of ONE process, one host thread each, on the same
// or on different devices -- NCCL cannot put two ranks on one device, and a partition count different from the device
// count (SURVEY.md section 4 vi) needs exactly that.  Collectives outside the PCG loop rendezvous on a host barrier and
// sum the ranks' buffers with a small kernel over (peer) device pointers, in rank order; the per-iteration exchange is
// the same fused peer-memory exchange as with NCCL ranks.
struct LocalComm {
  int n = 0;
  std::mutex mu;
  std::condition_variable cv;
  int arrived = 0;
  unsigned long long gen = 0;
  const void *ptr[8] = {};
  int dev[8] = {};
  int okflag[8] = {};
  void barrier() {
    std::unique_lock<std::mutex> lk(mu);
    const unsigned long long g = gen;
    if (++arrived == n) { arrived = 0; ++gen; cv.notify_all(); }
    else cv.wait(lk, [&] { return gen != g; });
  }
};

// Host-side, dtype-independent result of graph preprocessing.
struct HostGraph {
  int64_t nC = 0, nP = 0, nE = 0;
  int64_t pt_begin = 0, pt_end = 0;  // this rank's points
  int nPl = 0, nT = 0, nSlabs = 0, nChunks = 0, nEl = 0;
  std::vector<int4> tile;
  std::vector<int> e_cam;
  std::vector<unsigned short> e_lp;
  std::vector<int> pt_slot;
  std::vector<int64_t> edge_of_slot;   // -1 for padding
  std::vector<int32_t> slot_of_edge;   // [nE], -1 if the edge lives on another rank
  std::vector<int4> chunk;
  std::vector<int> cam_chunk_ptr;
  std::vector<int> cm_pt;
  std::vector<int> cm_slot;            // slot of each camera-major entry
  std::vector<int64_t> cuts;           // point cuts of all ranks [world_size + 1]
  // ELL layout (ba_internal.h): device order = local points sorted by observation count, descending
  bool ell = false;
  int nW = 0, nRows = 0, nHeavy = 0;
  std::vector<int4> wdesc;             // {first row, rows, first point (device order), kind | np << 8}
  std::vector<int> perm;               // device position -> local point
  std::vector<int2> cdesc;             // CTAs of the TMA-staged edge pass: {first warp, warps}; staged rows per CTA <= tma_rows
  int tma_rows = 0;
};

// ba_options::layout / operator_store (0 = auto) and the BA_LAYOUT / BA_STORE environment overrides (experiments).
static int resolve_layout(const ba_options &opt, int *layout, int *store) {
  int l = opt.layout, st = opt.operator_store;
  if (const char *e = getenv("BA_LAYOUT")) l = !strcmp(e, "tile") ? BA_LAYOUT_TILE : !strcmp(e, "ell") ? BA_LAYOUT_ELL : l;
  if (const char *e = getenv("BA_STORE")) st = !strcmp(e, "w") ? BA_STORE_W : !strcmp(e, "j") ? BA_STORE_J : !strcmp(e, "fact") ? BA_STORE_FACT : st;
  if (l == 0) l = (st == BA_STORE_W) ? BA_LAYOUT_TILE : BA_LAYOUT_DEFAULT;
  if (l == BA_LAYOUT_TILE) {
    // the tile layout stores what the compute kind says: Hpl blocks (explicit) or Jacobians (implicit)
    const int want = opt.compute_kind == BA_IMPLICIT ? BA_STORE_J : BA_STORE_W;
    if (st != 0 && st != want) return BA_ERR_ARG;
    st = want;
  } else if (l == BA_LAYOUT_ELL) {
    if (st == 0) st = BA_STORE_FACT;
    if (st == BA_STORE_W) return BA_ERR_ARG;
  } else {
    return BA_ERR_ARG;
  }
  *layout = l;
  *store = st == BA_STORE_W ? ST_W : st == BA_STORE_J ? ST_J : ST_FACT;
  return BA_OK;
}

static int partition_points(int64_t nP, const int64_t *deg, int ws, int64_t *cuts) {
  if (nP < 0 || ws < 1 || !cuts || (nP > 0 && !deg)) return BA_ERR_ARG;
  int64_t total = 0;
  for (int64_t p = 0; p < nP; ++p) total += deg[p];
  cuts[0] = 0;
  int64_t p = 0, acc = 0;
  for (int r = 1; r < ws; ++r) {
    const double target = (double)total * r / ws;
    while (p < nP && (double)acc + 0.5 * (double)deg[p] < target) acc += deg[p++];
    cuts[r] = p;
  }
  cuts[ws] = nP;
  return BA_OK;
}

// Camera-major view shared by both layouts: entries of a camera in ascending slot order, cut into chunks.
static void build_camera_major(int64_t nC, size_t nSlots, HostGraph &g) {
  std::vector<int> cptr(nC + 1, 0);
  for (size_t s = 0; s < nSlots; ++s)
    if (g.edge_of_slot[s] >= 0) cptr[g.e_cam[s] + 1]++;
  for (int64_t c = 0; c < nC; ++c) cptr[c + 1] += cptr[c];
  g.chunk.clear();
  g.cam_chunk_ptr.assign(nC + 1, 0);
  for (int64_t c = 0; c < nC; ++c) {
    for (int b = cptr[c]; b < cptr[c + 1]; b += CM_CHUNK)
      g.chunk.push_back(make_int4((int)c, b, std::min(b + CM_CHUNK, cptr[c + 1]), 0));
    g.cam_chunk_ptr[c + 1] = (int)g.chunk.size();
  }
  g.nChunks = (int)g.chunk.size();
}

// ELL layout: see ba_internal.h.  ptr/order = this rank's observations grouped by local point, camera-ascending.
static int build_ell(int64_t nC, int64_t nE, const int32_t *cam_idx, const std::vector<int64_t> &ptr,
                     const std::vector<int64_t> &order, HostGraph &g, std::string &err) {
  const int nPl = g.nPl;
  g.perm.resize(nPl);
  for (int p = 0; p < nPl; ++p) g.perm[p] = p;
  std::stable_sort(g.perm.begin(), g.perm.end(), [&](int a, int b) { return ptr[a + 1] - ptr[a] > ptr[b + 1] - ptr[b]; });
  auto deg = [&](int k) { return ptr[g.perm[k] + 1] - ptr[g.perm[k]]; };
  g.wdesc.clear();
  int64_t rows = 0;
  int k = 0;
  for (; k < nPl && deg(k) > ELL_HEAVY; ++k) {  // heavy points: one warp each, lane = observation
    const int64_t D = (deg(k) + 31) / 32;
    g.wdesc.push_back(make_int4((int)rows, (int)D, k, 1 | (1 << 8)));
    rows += D;
    if (rows * 32 >= (1ll << 31)) { err = "too many observation slots on one rank"; return BA_ERR_ARG; }
  }
  g.nHeavy = k;
  for (; k < nPl; k += 32) {  // slices of 32 points, lane = point; sorted, so deg(k) is the largest of the slice
    const int np = std::min(32, nPl - k);
    const int64_t D = deg(k);
    g.wdesc.push_back(make_int4((int)rows, (int)D, k, 0 | (np << 8)));
    rows += D;
    if (rows * 32 >= (1ll << 31)) { err = "too many observation slots on one rank"; return BA_ERR_ARG; }
  }
  g.nW = (int)g.wdesc.size();
  g.nRows = (int)rows;
  // CTAs of the TMA-staged edge pass: consecutive warps whose slices' rows fit the CTA's shared-memory stage
  // (tma_rows rows, at most ELL_WARPS warps); heavy points and slices longer than the stage stream from global memory
  // and take no stage rows.  The warp's first stage row goes into bits 16.. of the descriptor (0xffff = not staged).
  {
    static const int rows_env = getenv("BA_TMA_ROWS") ? atoi(getenv("BA_TMA_ROWS")) : 0;
    g.tma_rows = rows_env > 0 ? std::min(rows_env, 64) : 24;
    g.cdesc.clear();
    int first = 0, used = 0;
    for (int w = 0; w < g.nW; ++w) {
      int4 &wd = g.wdesc[w];
      const bool staged = (wd.w & 0xff) == 0 && wd.y <= g.tma_rows;
      const int need = staged ? wd.y : 0;
      if (w - first == ELL_WARPS || used + need > g.tma_rows) {
        g.cdesc.push_back(make_int2(first, w - first));
        first = w;
        used = 0;
      }
      wd.w |= (staged ? used : 0xffff) << 16;
      used += need;
    }
    if (g.nW > first) g.cdesc.push_back(make_int2(first, g.nW - first));
  }
  const size_t nSlots = (size_t)rows * 32;
  g.e_cam.assign(nSlots, -1);
  g.edge_of_slot.assign(nSlots, -1);
  g.slot_of_edge.assign(nE, -1);
  g.cm_pt.assign(g.nEl, 0);
  g.cm_slot.assign(g.nEl, 0);
  std::vector<int> slot_pt(nSlots, 0);
  for (const int4 &wd : g.wdesc) {
    const bool heavy = (wd.w & 0xff) != 0;
    const int np = (wd.w >> 8) & 0xff;
    for (int l = 0; l < (heavy ? 1 : np); ++l) {
      const int kd = wd.z + l, pn = g.perm[kd];
      for (int64_t i = 0; i < ptr[pn + 1] - ptr[pn]; ++i) {
        const size_t slot = heavy ? ((size_t)wd.x + i / 32) * 32 + i % 32 : ((size_t)wd.x + i) * 32 + l;
        const int64_t e = order[ptr[pn] + i];
        g.e_cam[slot] = cam_idx[e];
        g.edge_of_slot[slot] = e;
        g.slot_of_edge[e] = (int32_t)slot;
        slot_pt[slot] = kd;
      }
    }
  }
  build_camera_major(nC, nSlots, g);
  {
    std::vector<int> cf(nC, 0);
    for (int64_t c = 0; c < nC; ++c) cf[c] = g.chunk.empty() || g.cam_chunk_ptr[c] == g.cam_chunk_ptr[c + 1] ? 0 : g.chunk[g.cam_chunk_ptr[c]].y;
    for (size_t s = 0; s < nSlots; ++s) {
      if (g.edge_of_slot[s] < 0) continue;
      const int kk = cf[g.e_cam[s]]++;
      g.cm_pt[kk] = slot_pt[s];
      g.cm_slot[kk] = (int)s;
    }
  }
  return BA_OK;
}

static int build_host_graph(const ba_options &opt, int64_t nC, int64_t nP, int64_t nE, const int32_t *cam_idx,
                            const int32_t *pt_idx, HostGraph &g, std::string &err) {
  g.nC = nC; g.nP = nP; g.nE = nE;
  if (nC <= 0 || nP <= 0 || nE < 0 || nC >= (1 << 24) || nP > (1ll << 30)) { err = "bad problem size"; return BA_ERR_ARG; }
  std::vector<int64_t> deg(nP, 0);
  for (int64_t e = 0; e < nE; ++e) {
    if (cam_idx[e] < 0 || cam_idx[e] >= nC || pt_idx[e] < 0 || pt_idx[e] >= nP) { err = "vertex index out of range"; return BA_ERR_ARG; }
    deg[pt_idx[e]]++;
  }
  std::vector<int64_t> cuts(opt.world_size + 1);
  partition_points(nP, deg.data(), opt.world_size, cuts.data());
  g.pt_begin = cuts[opt.rank];
  g.pt_end = cuts[opt.rank + 1];
  g.nPl = (int)(g.pt_end - g.pt_begin);
  // local edges grouped by point (counting sort), then camera-ascending inside a point
  std::vector<int64_t> ptr(g.nPl + 1, 0);
  for (int p = 0; p < g.nPl; ++p) ptr[p + 1] = ptr[p] + deg[g.pt_begin + p];
  const int64_t nEl = ptr[g.nPl];
  if (nEl > (1ll << 31) - 4 * TE) { err = "too many observations on one rank"; return BA_ERR_ARG; }
  g.nEl = (int)nEl;
  std::vector<int64_t> order(nEl), fill(ptr.begin(), ptr.end() - 1);
  for (int64_t e = 0; e < nE; ++e) {
    const int64_t p = pt_idx[e];
    if (p >= g.pt_begin && p < g.pt_end) order[fill[p - g.pt_begin]++] = e;
  }
  for (int p = 0; p < g.nPl; ++p)
    std::sort(order.begin() + ptr[p], order.begin() + ptr[p + 1], [&](int64_t a, int64_t b) {
      return cam_idx[a] != cam_idx[b] ? cam_idx[a] < cam_idx[b] : a < b;
    });
  g.cuts = cuts;
  if (g.ell) return build_ell(nC, nE, cam_idx, ptr, order, g, err);
  // tiles
  g.tile.clear();
  g.pt_slot.assign(g.nPl, 0);
  int64_t slab = 0;  // padding can double the slot count of a pathological graph: checked against int range below
  int cur_pt0 = 0, cur_np = 0, cur_ne = 0;
  auto close_tile = [&]() {
    if (cur_np > 0) {
      g.tile.push_back(make_int4((int)(slab * TE), cur_ne, cur_pt0, cur_np));
      slab += 1;
    }
    cur_np = 0; cur_ne = 0;
  };
  for (int p = 0; p < g.nPl; ++p) {
    const int64_t dg = ptr[p + 1] - ptr[p];
    if (dg > TE) {
      close_tile();
      g.pt_slot[p] = (int)(slab * TE);
      g.tile.push_back(make_int4((int)(slab * TE), (int)dg, p, 1));
      // the point's further slabs get empty tiles: tile index == slab index is what lets the edge kernels address a
      // CTA's slots from blockIdx alone, before the descriptor has been read
      const int nslab = (int)((dg + TE - 1) / TE);
      for (int k = 1; k < nslab; ++k) g.tile.push_back(make_int4((int)((slab + k) * TE), 0, p, 0));
      slab += nslab;
      continue;
    }
    if (cur_np > 0 && (cur_ne + dg > TE || cur_np == TE)) close_tile();
    if (cur_np == 0) cur_pt0 = p;
    g.pt_slot[p] = (int)(slab * TE) + cur_ne;
    cur_np += 1;
    cur_ne += (int)dg;
  }
  close_tile();
  if ((slab + 1) * TE >= (1ll << 31)) { err = "too many observation slots on one rank"; return BA_ERR_ARG; }
  g.nT = (int)g.tile.size();
  g.nSlabs = (int)slab;
  const size_t nSlots = (size_t)slab * TE;
  g.e_cam.assign(nSlots, 0);
  g.e_lp.assign(nSlots, 0);
  g.edge_of_slot.assign(nSlots, -1);
  g.slot_of_edge.assign(nE, -1);
  for (int t = 0; t < g.nT; ++t) {
    const int4 td = g.tile[t];
    for (int lp = 0; lp < td.w; ++lp) {
      const int p = td.z + lp;
      const int s0 = g.pt_slot[p];
      for (int64_t k = ptr[p]; k < ptr[p + 1]; ++k) {
        const int slot = s0 + (int)(k - ptr[p]);
        const int64_t e = order[k];
        g.e_cam[slot] = cam_idx[e];
        g.e_lp[slot] = (unsigned short)lp;
        g.edge_of_slot[slot] = e;
        g.slot_of_edge[e] = slot;
      }
    }
  }
  // camera-major view of the local edges (slot-ascending inside a camera)
  std::vector<int> cptr(nC + 1, 0);
  for (size_t s = 0; s < nSlots; ++s)
    if (g.edge_of_slot[s] >= 0) cptr[g.e_cam[s] + 1]++;
  for (int64_t c = 0; c < nC; ++c) cptr[c + 1] += cptr[c];
  g.cm_pt.assign(nEl, 0);
  g.cm_slot.assign(nEl, 0);
  {
    std::vector<int> cf(cptr.begin(), cptr.end() - 1);
    for (int t = 0; t < g.nT; ++t) {
      const int4 td = g.tile[t];
      for (int i = 0; i < td.y; ++i) {
        const int slot = td.x + i;
        const int c = g.e_cam[slot];
        const int k = cf[c]++;
        g.cm_pt[k] = td.z + g.e_lp[slot];
        g.cm_slot[k] = slot;
      }
    }
  }
  g.chunk.clear();
  g.cam_chunk_ptr.assign(nC + 1, 0);
  for (int64_t c = 0; c < nC; ++c) {
    for (int b = cptr[c]; b < cptr[c + 1]; b += CM_CHUNK)
      g.chunk.push_back(make_int4((int)c, b, std::min(b + CM_CHUNK, cptr[c + 1]), 0));
    g.cam_chunk_ptr[c + 1] = (int)g.chunk.size();
  }
  g.nChunks = (int)g.chunk.size();
  return BA_OK;
}

struct EventPair { cudaEvent_t a, b; };

struct Impl {
  ba_options opt{};
  HostGraph g;
  std::string err;
  cudaStream_t stream = nullptr;
  ncclComm_t comm = nullptr;
  bool comm_ready = false;
  std::shared_ptr<LocalComm> local;  // in-process communicator instead of NCCL
  bool timing = false;
  int timing_stride = 1;  // per-launch events are recorded for every timing_stride-th PCG iteration
  ba_stats stats{};
  bool linearized = false;
  double region = 1e3;
  int store_cfg = ST_W;  // configured operator stream; Dev::store is the one in effect for the current linearisation
  // ba_lm_solve fuses the scalar kernels behind the big ones so that an LM iteration costs three host round trips
  // (solve + step norms, trial evaluation, linearisation + gradient norm) instead of five
  bool fuse_norms = false, fuse_ginf = false, norms_valid = false, ginf_valid = false;
  bool persistent = false;      // whole PCG solve in one cooperative launch (ba_set_persistent / BA_PERSIST=1); measured slower
                                // than two launches per iteration on one GPU (profiles/README.md), hence off by default
  bool persistent_used = false; // the current / last solve took that path
  virtual ~Impl() {}
  int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    err = buf;
    return code;
  }
  virtual int init(const void *obs) = 0;
  virtual int set_params(const void *cam, const void *pt) = 0;
  virtual int get_params(void *cam, void *pt) = 0;
  virtual int set_information(const void *info) = 0;
  virtual int set_fixed(const unsigned char *cam_fixed, const unsigned char *pt_fixed) = 0;
  virtual int set_residual_program(int n_instr, const uint8_t *code, int n_const, const double *consts, int n_slots,
                                   const int32_t *out_slot) = 0;
  virtual int residual_program_mode() = 0;
  virtual int linearize(double *cost) = 0;
  virtual int process_diag(double region, int recover) = 0;
  virtual int solve(const ba_pcg_options *o, int *iters, int *done) = 0;
  virtual int solve_async(const ba_pcg_options *o) = 0;
  virtual int solve_wait(int *iters, int *done) = 0;
  virtual int set_params_device(const void *cam, const void *pt, void *caller_stream) = 0;
  virtual int get_params_device(void *cam, void *pt, void *caller_stream) = 0;
  virtual int step_norms(double *dx, double *x) = 0;
  virtual int eval_trial(double *cost_new, double *model_sq) = 0;
  virtual int accept_step() = 0;
  virtual int reject_step() = 0;
  virtual int backup_step() = 0;
  virtual int grad_inf(double *g) = 0;
  virtual int read(int which, void *out) = 0;
  virtual int write_dx(const void *dx) = 0;
  virtual int eval_edges(void *r, void *J) = 0;
  virtual int schur_apply(const void *p, void *q) = 0;
  virtual int get_stats(ba_stats *out) = 0;
  virtual int reset_stats() = 0;
  virtual int algorithmic_bytes(double *a, double *b) = 0;
  virtual int setup_p2p() = 0;
};

template <typename T>
struct ImplT : Impl {
  Dev<T> d{};
  void *arena = nullptr;
  size_t arena_bytes = 0;
  T *e_info_dev = nullptr, *cm_info_dev = nullptr;
  T *pt_full = nullptr;  // [3 nP] gather buffer, world_size > 1 only
  void *coll_scratch = nullptr;  // result buffer of the in-process allreduce (>= 90 nC scalars)
  unsigned char *cam_fixed_dev = nullptr, *pt_fixed_dev = nullptr;
  uint2 *prog_code_dev = nullptr;  // user-defined residual (ba_set_residual_program)
  T *prog_const_dev = nullptr;
  cudaLibrary_t jit_lib = nullptr;  // NVRTC-generated kernels of the residual program (ba_program_jit.h)
  cudaKernel_t jit_fill = nullptr, jit_cost = nullptr;
  T *prog_tab_dev = nullptr;  // frontier tables of the per-camera program (current | trial parameters)
  int *cm_slot_dev = nullptr;
  int *slot_cm_dev = nullptr;  // generated kernels: inverse of cm_slot, and the camera-major {J camera part, r} copy
  T *cm_jr_dev = nullptr;
  T *r_slots_dev = nullptr;
  std::string jit_note;            // why the generated kernels are not in use (empty: they are)
  int store_before_prog = -1;      // the configured stream to go back to when the program is cleared
  double *model_part = nullptr;  // [nW] point partials of the model decrease (left by the back-substitution pass)
  bool model_valid = false;      // scal[S_MODEL_PT / S_MODEL_CAM] describe the step of the last solve
  double last_cost = 0.;         // F of the current linearisation
  // ELL layout
  int *perm_dev = nullptr; // [nPl] device position -> local point
  T *pt_nat = nullptr;     // [3 nPl] staging in the caller's point order
  T *r_info_dev = nullptr;
  double *h_scal = nullptr;
  PcgState *h_st = nullptr;  // two pinned slots (chunk k is polled while chunk k+1 runs)
  cudaEvent_t st_event[2] = {nullptr, nullptr};
  std::vector<EventPair> ev_pool;
  size_t ev_used = 0;
  std::vector<size_t> ev_schur, ev_lin, ev_pcg, ev_solve, ev_cam;
  int pcg_chunk = 8;
  int launched = 0;             // PCG iterations enqueued by the current / last solve
  int last_iters = 1 << 20;     // iterations the previous solve executed
  bool solve_pending = false;   // ba_solve_async enqueued, ba_solve_wait not yet called
  std::vector<std::pair<size_t, int>> loc_schur, loc_cam;  // (event pair, iteration) of the sampled launches
  // fused peer-memory exchange of the camera kernel (world_size > 1)
  void *xchg = nullptr;
  void *ipc_opened[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  PeerExchange px{};
  bool p2p = false;
  unsigned long long px_epoch = 0;

  ~ImplT() override {
    if (stream) cudaStreamSynchronize(stream);
    for (auto &e : ev_pool) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    if (d.xpad_tex) cudaDestroyTextureObject(d.xpad_tex);
    if (d.wj_tex) cudaDestroyTextureObject(d.wj_tex);
    if (d.camc_tex) cudaDestroyTextureObject(d.camc_tex);
    if (d.camc2_tex) cudaDestroyTextureObject(d.camc2_tex);
    if (arena) cudaFree(arena);
    if (e_info_dev) cudaFree(e_info_dev);
    if (cm_info_dev) cudaFree(cm_info_dev);
    if (r_info_dev) cudaFree(r_info_dev);
    if (cam_fixed_dev) cudaFree(cam_fixed_dev);
    if (prog_code_dev) cudaFree(prog_code_dev);
    if (prog_const_dev) cudaFree(prog_const_dev);
    if (cm_slot_dev) cudaFree(cm_slot_dev);
    if (slot_cm_dev) cudaFree(slot_cm_dev);
    if (cm_jr_dev) cudaFree(cm_jr_dev);
    if (prog_tab_dev) cudaFree(prog_tab_dev);
    if (jit_lib) cudaLibraryUnload(jit_lib);
    if (r_slots_dev) cudaFree(r_slots_dev);
    if (pt_fixed_dev) cudaFree(pt_fixed_dev);
    if (pt_full) cudaFree(pt_full);
    if (coll_scratch) cudaFree(coll_scratch);
    for (auto *q : ipc_opened) if (q) cudaIpcCloseMemHandle(q);
    if (xchg) cudaFree(xchg);
    if (h_scal) cudaFreeHost(h_scal);
    if (h_st) cudaFreeHost(h_st);
    for (auto &e : st_event) if (e) cudaEventDestroy(e);
    if (comm_ready) ncclCommDestroy(comm);
    if (stream) cudaStreamDestroy(stream);
  }

  // ---- timing helpers -------------------------------------------------------
  size_t ev_begin() {
    if (ev_used == ev_pool.size()) {
      EventPair p;
      cudaEventCreate(&p.a);
      cudaEventCreate(&p.b);
      ev_pool.push_back(p);
    }
    cudaEventRecord(ev_pool[ev_used].a, stream);
    return ev_used++;
  }
  void ev_end(size_t i) { cudaEventRecord(ev_pool[i].b, stream); }
  double ev_sum(std::vector<size_t> &v) {
    double tot = 0;
    for (size_t i : v) {
      float ms = 0;
      cudaEventElapsedTime(&ms, ev_pool[i].a, ev_pool[i].b);
      tot += ms;
    }
    v.clear();
    return tot;
  }
  void fold_events() {
    cudaStreamSynchronize(stream);
    stats.schur_apply_timed += (double)ev_schur.size();
    stats.pcg_cam_timed += (double)ev_cam.size();
    stats.schur_apply_ms += ev_sum(ev_schur);
    stats.linearize_ms += ev_sum(ev_lin);
    stats.pcg_loop_ms += ev_sum(ev_pcg);
    stats.solve_ms += ev_sum(ev_solve);
    stats.pcg_cam_ms += ev_sum(ev_cam);  // summed device time of the camera-space PCG kernel
    ev_used = 0;
  }

  // ---- collectives ----------------------------------------------------------
  // in-process flavour: every rank sums all ranks' buffers (rank order) into its own scratch, then copies back
  template <typename U>
  int local_allreduce(U *p, size_t n, int op) {
    BA_CUDA(cudaStreamSynchronize(stream));
    local->ptr[opt.rank] = p;
    local->barrier();
    const void *srcs[8];
    for (int r = 0; r < 8; ++r) srcs[r] = r < local->n ? local->ptr[r] : nullptr;
    U *tmp = reinterpret_cast<U *>(coll_scratch);
    BA_CUDA(launch_sum_ranks<U>(tmp, srcs, local->n, n, op, stream));
    stats.launches += 1;
    BA_CUDA(cudaStreamSynchronize(stream));
    local->barrier();  // everybody has read every source
    BA_CUDA(cudaMemcpyAsync(p, tmp, sizeof(U) * n, cudaMemcpyDeviceToDevice, stream));
    return BA_OK;
  }
  int allreduce_T(T *p, size_t n) {
    if (opt.world_size == 1) return BA_OK;
    if (local) return local_allreduce<T>(p, n, 0);
    if (!comm_ready) return fail(BA_ERR_STATE, "world_size > 1 but ba_comm_init was not called");
    BA_NCCL(ncclAllReduce(p, p, n, NcclType<T>::v, ncclSum, comm, stream));
    return BA_OK;
  }
  int allreduce_scal(int first, int n, ncclRedOp_t op) {
    if (opt.world_size == 1) return BA_OK;
    if (local) return local_allreduce<double>(d.scal + first, (size_t)n, op == ncclMax ? 1 : op == ncclMin ? 2 : 0);
    if (!comm_ready) return fail(BA_ERR_STATE, "world_size > 1 but ba_comm_init was not called");
    BA_NCCL(ncclAllReduce(d.scal + first, d.scal + first, n, ncclDouble, op, comm, stream));
    return BA_OK;
  }
  int fetch_scalars() {
    BA_CUDA(cudaMemcpyAsync(h_scal, d.scal, sizeof(double) * S_COUNT, cudaMemcpyDeviceToHost, stream));
    BA_CUDA(cudaStreamSynchronize(stream));
    return BA_OK;
  }

  // Maps every rank's inbox (see PeerExchange) into this rank's address space: CUDA IPC across processes, plain
  // (peer) pointers between solvers of one process.  Any failure leaves p2p = false and the fold + allreduce path in use.
  int setup_p2p() override {
    if (opt.world_size == 1 || opt.world_size > 8 || (!comm_ready && !local)) return BA_OK;
    BA_CUDA(cudaSetDevice(opt.device));
    const size_t n9 = 9 * (size_t)g.nC, bytes = sizeof(T) * 2 * (size_t)opt.world_size * 2 * n9, off = (bytes + 255) / 256 * 256;
    struct Info { cudaIpcMemHandle_t h; long long pid; unsigned long long ptr; int dev, ok; char pad[128 - sizeof(cudaIpcMemHandle_t) - 24]; };
    static_assert(sizeof(Info) == 128, "Info is one 128-byte record");
    Info mine{};
    int ok = 1;
    if (cudaMalloc(&xchg, off + 256) != cudaSuccess) { cudaGetLastError(); xchg = nullptr; ok = 0; }
    if (ok && cudaMemset(xchg, 0, off + 256) != cudaSuccess) ok = 0;
    if (ok && !local && cudaIpcGetMemHandle(&mine.h, xchg) != cudaSuccess) { cudaGetLastError(); ok = 0; }
    mine.pid = (long long)getpid(); mine.ptr = (unsigned long long)xchg; mine.dev = opt.device; mine.ok = ok;
    std::vector<Info> all(opt.world_size);
    if (local) {
      static std::mutex info_mu;
      static std::vector<Info> board(8);
      { std::lock_guard<std::mutex> lk(info_mu); board[opt.rank] = mine; }
      local->barrier();
      { std::lock_guard<std::mutex> lk(info_mu); for (int r = 0; r < opt.world_size; ++r) all[r] = board[r]; }
      local->barrier();
    } else {
      char *scratch = reinterpret_cast<char *>(d.partials);  // >= 8 * 128 bytes
      BA_CUDA(cudaMemcpyAsync(scratch + 128 * opt.rank, &mine, 128, cudaMemcpyHostToDevice, stream));
      BA_NCCL(ncclAllGather(scratch + 128 * opt.rank, scratch, 128, ncclChar, comm, stream));
      BA_CUDA(cudaMemcpyAsync(all.data(), scratch, 128 * opt.world_size, cudaMemcpyDeviceToHost, stream));
      BA_CUDA(cudaStreamSynchronize(stream));
    }
    px = PeerExchange{};
    bool shared_dev = false;
    for (int r = 0; r < opt.world_size; ++r) {
      void *ptr = nullptr;
      if (!all[r].ok) { ok = 0; continue; }
      if (r == opt.rank) {
        ptr = xchg;
      } else if (all[r].pid == mine.pid) {  // another solver of this process: direct (peer) pointers
        if (all[r].dev != opt.device) {
          cudaError_t e = cudaDeviceEnablePeerAccess(all[r].dev, 0);
          if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) ok = 0;
          cudaGetLastError();
        } else {
          shared_dev = true;
        }
        ptr = (void *)all[r].ptr;
      } else if (cudaIpcOpenMemHandle(&ptr, all[r].h, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess) {
        ipc_opened[r] = ptr;
      } else {
        cudaGetLastError();
        ok = 0;
      }
      px.data[r] = ptr;
      px.flag[r] = ptr ? reinterpret_cast<unsigned long long *>((char *)ptr + off) : nullptr;
    }
    d.no_pdl = shared_dev ? 1 : 0;
    // every rank must take the same path: agree on min(ok)
    double okd = ok;
    BA_CUDA(cudaMemcpyAsync(d.scal + S_XCHG_ERR, &okd, sizeof(double), cudaMemcpyHostToDevice, stream));
    BA_TRY(allreduce_scal(S_XCHG_ERR, 1, ncclMin));
    BA_CUDA(cudaMemcpyAsync(&okd, d.scal + S_XCHG_ERR, sizeof(double), cudaMemcpyDeviceToHost, stream));
    BA_CUDA(cudaMemsetAsync(d.scal + S_XCHG_ERR, 0, sizeof(double), stream));
    BA_CUDA(cudaStreamSynchronize(stream));
    p2p = okd > 0.5;
    px.ws = opt.world_size;
    px.rank = opt.rank;
    {
      const char *e = getenv("BA_XCHG");
      px.pull = e ? (!strcmp(e, "pull") ? 1 : 0) : 0;
    }
    return BA_OK;
  }

  // ---- set-up -----------------------------------------------------------------
  int init(const void *obs_v) override {
    const T *obs = (const T *)obs_v;
    BA_CUDA(cudaSetDevice(opt.device));
    BA_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    const bool ell = g.ell;
    const int K = ell ? 24 : (opt.compute_kind == BA_IMPLICIT ? 24 : 27);  // ELL: room for the Jacobian stream (small-angle cameras)
    const size_t nSlots = ell ? (size_t)g.nRows * 32 : (size_t)g.nSlabs * TE, nC = g.nC, nPl = g.nPl, nEl = g.nEl;
    const size_t nPlPad = (nPl + 255) / 256 * 256;
    const size_t nPart = std::max<size_t>({(size_t)g.nT * 2, (size_t)g.nW * 2, (size_t)AUX_GRID * 4, 64});
    // arena carving
    struct Item { void **ptr; size_t bytes; };
    std::vector<Item> items;
    int4 *tile, *chunk; int *e_cam, *pt_slot, *cam_chunk_ptr, *cm_pt; unsigned short *e_lp; T *e_obs, *cm_obs;
    auto add = [&](auto **p, size_t count) { items.push_back({(void **)p, count * sizeof(**p)}); };
    int4 *wdesc = nullptr;
    int2 *cdesc = nullptr;
    add(&tile, g.nT); add(&e_cam, nSlots); add(&e_lp, ell ? 0 : nSlots); add(&e_obs, 2 * nSlots); add(&pt_slot, ell ? 0 : nPl);
    if (ell) {
      add(&wdesc, g.nW); add(&cdesc, g.cdesc.size()); add(&d.camc, (size_t)CAMC * nC); add(&d.camc2, (size_t)CAMC * nC); add(&d.camT, (size_t)CAMT * nC);
      add(&d.HllInvP, 9 * nPlPad); add(&d.wP, 3 * nPlPad); add(&d.XP, 3 * nPlPad);
      add(&perm_dev, nPl); add(&pt_nat, 3 * nPl);
    }
    add(&chunk, g.nChunks); add(&cam_chunk_ptr, nC + 1); add(&cm_pt, nEl); add(&cm_obs, 2 * nEl);
    add(&d.cam, 9 * nC); add(&d.pt, 3 * nPl);
    add(&d.Hpp, 90 * nC);  // Hpp followed by gc
    add(&d.Hll, 9 * nPl); add(&d.gl, 3 * nPl);
    add(&d.WJ, (size_t)K * nSlots);
    add(&d.cm_part, (size_t)g.nChunks * 54);
    add(&d.HppD, 81 * nC); add(&d.HppInv, 81 * nC); add(&d.HllInv, ell ? 0 : 9 * nPl); add(&d.w, ell ? 0 : 3 * nPl);
    add(&d.dxc, 9 * nC); add(&d.dxl, 3 * nPl); add(&d.dxc_bk, 9 * nC); add(&d.dxl_bk, 3 * nPl);
    add(&d.sA, 3 * (size_t)NREP * 9 * nC);  // sA[NREP][9nC] followed by sB[2][NREP][9nC]
    add(&d.r, 9 * nC); add(&d.p, 9 * nC); add(&d.q, 9 * nC); add(&d.z, 9 * nC); add(&d.xbk, 9 * nC);
    add(&d.xpad, (size_t)CamPad<T>::N * nC);
    add(&d.st, 1); add(&d.scal, S_COUNT); add(&d.partials, nPart);
    add(&d.pk_counters, 4); add(&d.pk_partials, 2 * 512); add(&d.pk_timers, 4);
    add(&model_part, std::max<size_t>(g.nW, 1));
    size_t total = 0;
    for (auto &it : items) total += (it.bytes + 511) / 512 * 512 + 512;
    BA_CUDA(cudaMalloc(&arena, total));
    arena_bytes = total;
    BA_CUDA(cudaMemsetAsync(arena, 0, total, stream));
    size_t off = 0;
    for (auto &it : items) {
      *it.ptr = (char *)arena + off;
      off += (it.bytes + 511) / 512 * 512 + 512;
    }
    d.gc = d.Hpp + 81 * nC;
    {
      // linear texture over xpad (16-byte texels); only read by the BA_ELL_TEX experiment build
      cudaResourceDesc rd{};
      rd.resType = cudaResourceTypeLinear;
      rd.res.linear.devPtr = d.xpad;
      rd.res.linear.sizeInBytes = sizeof(T) * (size_t)CamPad<T>::N * nC;
      rd.res.linear.desc = sizeof(T) == 8 ? cudaCreateChannelDesc<int4>() : cudaCreateChannelDesc<float4>();
      cudaTextureDesc td{};
      td.readMode = cudaReadModeElementType;
      d.xpad_tex = 0;
      if (cudaCreateTextureObject(&d.xpad_tex, &rd, &td, nullptr) != cudaSuccess) { cudaGetLastError(); d.xpad_tex = 0; }
      d.camc_tex = d.camc2_tex = 0;
      if (ell) {
        rd.res.linear.sizeInBytes = sizeof(T) * (size_t)CAMC * nC;
        rd.res.linear.devPtr = d.camc;
        if (cudaCreateTextureObject(&d.camc_tex, &rd, &td, nullptr) != cudaSuccess) { cudaGetLastError(); d.camc_tex = 0; }
        rd.res.linear.devPtr = d.camc2;
        if (cudaCreateTextureObject(&d.camc2_tex, &rd, &td, nullptr) != cudaSuccess) { cudaGetLastError(); d.camc2_tex = 0; }
      }
      d.wj_tex = 0;
      const size_t wj_bytes = sizeof(T) * (size_t)K * nSlots;
      if (ell && wj_bytes / (2 * sizeof(T)) < (1ull << 27)) {
        rd.res.linear.devPtr = d.WJ;
        rd.res.linear.sizeInBytes = wj_bytes;
        rd.res.linear.desc = sizeof(T) == 8 ? cudaCreateChannelDesc<int4>() : cudaCreateChannelDesc<float2>();
        if (cudaCreateTextureObject(&d.wj_tex, &rd, &td, nullptr) != cudaSuccess) { cudaGetLastError(); d.wj_tex = 0; }
      }
    }
    d.sB = d.sA + (size_t)NREP * 9 * nC;
    d.nC = (int)nC; d.nPl = (int)nPl; d.nT = g.nT; d.nSlabs = g.nSlabs; d.nChunks = g.nChunks; d.nEl = g.nEl;
    d.implicit = opt.compute_kind == BA_IMPLICIT;
    d.ell = ell; d.nW = g.nW; d.nRows = g.nRows; d.nPlPad = (int)nPlPad;
    d.split = 1;
    if (ell) {  // few slices (a small problem, or one shard of many): several warps per slice, see k_schur_ell
      // measured on B200 (profiles/r02_edge_variants.txt): an eighth of Venice (3 884 slices, about one wave of the
      // 148 x 24 resident warps) 33.6 -> 29.5 us with two warps per slice (31.5 with four); a quarter (7 768 slices)
      // 49.3 -> 52.1 us; Trafalgar-257 (2 036 slices) 21.3 -> 17.9 -> 17.1 us; Ladybug-49 (243) 16.4 -> 12.9 -> 10.8 us
      const char *e = getenv("BA_ELL_SPLIT");
      const int want = e ? atoi(e) : 0;
      if (want == 1 || want == 2 || want == 4) d.split = want;
      else if (g.nW < 2500) d.split = 4;
      else if (g.nW < 4500) d.split = 2;
    }
    d.wdesc = wdesc; d.cdesc = cdesc; d.nCta = (int)g.cdesc.size(); d.tma_rows = g.tma_rows; d.r_cam = e_cam; d.r_obs = e_obs; d.r_info = nullptr;
    if (ell) {
      // the factored stream needs the register-resident camera kernel (it applies T_c / T_c^T)
      if (store_cfg == ST_FACT && !pcg_reg_kernel_available<T>(d)) store_cfg = ST_J;
      d.implicit = 1;  // nothing in the ELL path materialises Hpl
    }
    d.store = store_cfg;
    d.tile = tile; d.e_cam = e_cam; d.e_lp = e_lp; d.e_obs = e_obs; d.e_info = nullptr; d.pt_slot = pt_slot;
    d.chunk = chunk; d.cam_chunk_ptr = cam_chunk_ptr; d.cm_pt = cm_pt; d.cm_obs = cm_obs; d.cm_info = nullptr;
    d.damp = 1. / region;
    // upload the static graph
    std::vector<T> h_obs(2 * nSlots, T(0)), h_cm_obs(2 * nEl);
    for (size_t s = 0; s < nSlots; ++s) {
      const int64_t e = g.edge_of_slot[s];
      if (e >= 0) { h_obs[2 * s] = obs[2 * e]; h_obs[2 * s + 1] = obs[2 * e + 1]; }
    }
    for (size_t k = 0; k < nEl; ++k) {
      const int64_t e = g.edge_of_slot[g.cm_slot[k]];
      h_cm_obs[2 * k] = obs[2 * e];
      h_cm_obs[2 * k + 1] = obs[2 * e + 1];
    }
    auto up = [&](void *dst, const void *src, size_t bytes) -> cudaError_t {
      return bytes ? cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream) : cudaSuccess;
    };
    BA_CUDA(up(tile, g.tile.data(), sizeof(int4) * g.nT));
    BA_CUDA(up(e_cam, g.e_cam.data(), sizeof(int) * nSlots));
    BA_CUDA(up(e_obs, h_obs.data(), sizeof(T) * 2 * nSlots));
    if (ell) {
      BA_CUDA(up(wdesc, g.wdesc.data(), sizeof(int4) * g.nW));
      BA_CUDA(up(cdesc, g.cdesc.data(), sizeof(int2) * g.cdesc.size()));
      BA_CUDA(up(perm_dev, g.perm.data(), sizeof(int) * nPl));
    } else {
      BA_CUDA(up(e_lp, g.e_lp.data(), sizeof(unsigned short) * nSlots));
      BA_CUDA(up(pt_slot, g.pt_slot.data(), sizeof(int) * nPl));
    }
    BA_CUDA(up(chunk, g.chunk.data(), sizeof(int4) * g.nChunks));
    BA_CUDA(up(cam_chunk_ptr, g.cam_chunk_ptr.data(), sizeof(int) * (nC + 1)));
    BA_CUDA(up(cm_pt, g.cm_pt.data(), sizeof(int) * nEl));
    BA_CUDA(up(cm_obs, h_cm_obs.data(), sizeof(T) * 2 * nEl));
    BA_CUDA(cudaHostAlloc(&h_scal, sizeof(double) * S_COUNT, cudaHostAllocDefault));
    BA_CUDA(cudaHostAlloc(&h_st, 2 * sizeof(PcgState), cudaHostAllocDefault));
    BA_CUDA(cudaEventCreateWithFlags(&st_event[0], cudaEventDisableTiming));
    BA_CUDA(cudaEventCreateWithFlags(&st_event[1], cudaEventDisableTiming));
    if (opt.world_size > 1) {
      BA_CUDA(cudaMalloc(&pt_full, sizeof(T) * 3 * g.nP));
      BA_CUDA(cudaMalloc(&coll_scratch, std::max<size_t>(sizeof(T) * 90 * nC, sizeof(double) * S_COUNT)));
    }
    BA_CUDA(cudaStreamSynchronize(stream));
    // free the big host staging vectors that are no longer needed
    std::vector<int>().swap(g.e_cam);
    std::vector<unsigned short>().swap(g.e_lp);
    std::vector<int>().swap(g.cm_pt);
    return BA_OK;
  }

  int set_params(const void *cam_v, const void *pt_v) override {
    if (!cam_v || !pt_v) return fail(BA_ERR_ARG, "null parameter buffer");
    BA_CUDA(cudaSetDevice(opt.device));
    const T *cam = (const T *)cam_v, *pt = (const T *)pt_v;
    BA_CUDA(cudaMemcpyAsync(d.cam, cam, sizeof(T) * 9 * g.nC, cudaMemcpyHostToDevice, stream));
    if (g.nPl) {
      BA_CUDA(cudaMemcpyAsync(g.ell ? pt_nat : d.pt, pt + 3 * g.pt_begin, sizeof(T) * 3 * g.nPl, cudaMemcpyHostToDevice, stream));
      if (g.ell) BA_CUDA(launch_permute3<T>(pt_nat, d.pt, perm_dev, g.nPl, 1, stream));
    }
    if (g.ell) { BA_CUDA(launch_cam_consts<T>(d, stream)); stats.launches += g.nPl ? 2 : 1; }
    // a new parameter vector is a new problem: deltaX starts at zero (allocateResourceCUDA memset,
    // schur_LM_linear_system.cu:37)
    BA_CUDA(cudaMemsetAsync(d.dxc, 0, sizeof(T) * 9 * g.nC, stream));
    BA_CUDA(cudaMemsetAsync(d.dxc_bk, 0, sizeof(T) * 9 * g.nC, stream));
    if (g.nPl) {
      BA_CUDA(cudaMemsetAsync(d.dxl, 0, sizeof(T) * 3 * g.nPl, stream));
      BA_CUDA(cudaMemsetAsync(d.dxl_bk, 0, sizeof(T) * 3 * g.nPl, stream));
    }
    BA_CUDA(cudaStreamSynchronize(stream));
    linearized = false;
    return BA_OK;
  }

  int get_params(void *cam_v, void *pt_v) override {
    if (!cam_v || !pt_v) return fail(BA_ERR_ARG, "null parameter buffer");
    BA_CUDA(cudaSetDevice(opt.device));
    T *cam = (T *)cam_v, *pt = (T *)pt_v;
    BA_CUDA(cudaMemcpyAsync(cam, d.cam, sizeof(T) * 9 * g.nC, cudaMemcpyDeviceToHost, stream));
    const T *own = d.pt;  // this rank's points in the caller's order
    if (g.ell && g.nPl) {
      BA_CUDA(launch_permute3<T>(d.pt, pt_nat, perm_dev, g.nPl, 0, stream));
      stats.launches += 1;
      own = pt_nat;
    }
    if (opt.world_size == 1) {
      BA_CUDA(cudaMemcpyAsync(pt, own, sizeof(T) * 3 * g.nP, cudaMemcpyDeviceToHost, stream));
    } else {
      // every rank owns a contiguous point range: an all-gather with per-rank counts = one broadcast per rank, grouped
      if (local) {
        BA_CUDA(cudaStreamSynchronize(stream));
        local->ptr[opt.rank] = own;
        local->barrier();
        for (int r = 0; r < opt.world_size; ++r) {
          const size_t cnt = 3 * (size_t)(g.cuts[r + 1] - g.cuts[r]);
          if (cnt) BA_CUDA(cudaMemcpyAsync(pt + 3 * g.cuts[r], local->ptr[r], sizeof(T) * cnt, cudaMemcpyDeviceToHost, stream));
        }
        BA_CUDA(cudaStreamSynchronize(stream));
        local->barrier();
        return BA_OK;
      }
      if (!comm_ready) return fail(BA_ERR_STATE, "world_size > 1 but ba_comm_init was not called");
      if (g.nPl)
        BA_CUDA(cudaMemcpyAsync(pt_full + 3 * g.pt_begin, own, sizeof(T) * 3 * g.nPl, cudaMemcpyDeviceToDevice, stream));
      BA_NCCL(ncclGroupStart());
      for (int r = 0; r < opt.world_size; ++r) {
        const size_t cnt = 3 * (size_t)(g.cuts[r + 1] - g.cuts[r]);
        if (cnt) BA_NCCL(ncclBroadcast(pt_full + 3 * g.cuts[r], pt_full + 3 * g.cuts[r], cnt, NcclType<T>::v, r, comm, stream));
      }
      BA_NCCL(ncclGroupEnd());
      BA_CUDA(cudaMemcpyAsync(pt, pt_full, sizeof(T) * 3 * g.nP, cudaMemcpyDeviceToHost, stream));
    }
    BA_CUDA(cudaStreamSynchronize(stream));
    return BA_OK;
  }

  // Device-pointer variants (the reference's raw-pointer seam, include/solver/base_solver.h:28-32): buffers live on THIS
  // solver's device, in the caller's vertex order; no host copy, no host synchronisation.  `caller_stream` (may be
  // null) is ordered with the solver's stream through events: the copies start after the work already enqueued on it,
  // and work enqueued on it afterwards sees the results.
  int link_in(void *caller_stream) {
    if (!caller_stream) return BA_OK;
    BA_CUDA(cudaEventRecord(st_event[0], (cudaStream_t)caller_stream));
    BA_CUDA(cudaStreamWaitEvent(stream, st_event[0], 0));
    return BA_OK;
  }
  int link_out(void *caller_stream) {
    if (!caller_stream) return BA_OK;
    BA_CUDA(cudaEventRecord(st_event[1], stream));
    BA_CUDA(cudaStreamWaitEvent((cudaStream_t)caller_stream, st_event[1], 0));
    return BA_OK;
  }
  int set_params_device(const void *cam_v, const void *pt_v, void *caller_stream) override {
    if (!cam_v || !pt_v) return fail(BA_ERR_ARG, "null parameter buffer");
    if (solve_pending) return fail(BA_ERR_STATE, "ba_solve_async is still pending");
    BA_CUDA(cudaSetDevice(opt.device));
    BA_TRY(link_in(caller_stream));
    const T *cam = (const T *)cam_v, *pt = (const T *)pt_v;
    BA_CUDA(cudaMemcpyAsync(d.cam, cam, sizeof(T) * 9 * g.nC, cudaMemcpyDeviceToDevice, stream));
    if (g.nPl) {
      if (g.ell) BA_CUDA(launch_permute3<T>(pt + 3 * g.pt_begin, d.pt, perm_dev, g.nPl, 1, stream));
      else BA_CUDA(cudaMemcpyAsync(d.pt, pt + 3 * g.pt_begin, sizeof(T) * 3 * g.nPl, cudaMemcpyDeviceToDevice, stream));
    }
    if (g.ell) { BA_CUDA(launch_cam_consts<T>(d, stream)); stats.launches += g.nPl ? 2 : 1; }
    BA_CUDA(cudaMemsetAsync(d.dxc, 0, sizeof(T) * 9 * g.nC, stream));
    BA_CUDA(cudaMemsetAsync(d.dxc_bk, 0, sizeof(T) * 9 * g.nC, stream));
    if (g.nPl) {
      BA_CUDA(cudaMemsetAsync(d.dxl, 0, sizeof(T) * 3 * g.nPl, stream));
      BA_CUDA(cudaMemsetAsync(d.dxl_bk, 0, sizeof(T) * 3 * g.nPl, stream));
    }
    BA_TRY(link_out(caller_stream));  // the caller may reuse its buffers once its stream passes this point
    linearized = false;
    return BA_OK;
  }
  int get_params_device(void *cam_v, void *pt_v, void *caller_stream) override {
    if (!cam_v || !pt_v) return fail(BA_ERR_ARG, "null parameter buffer");
    if (opt.world_size > 1) return fail(BA_ERR_STATE, "ba_get_params_device: single-rank solvers only (use ba_get_params)");
    BA_CUDA(cudaSetDevice(opt.device));
    BA_TRY(link_in(caller_stream));
    T *cam = (T *)cam_v, *pt = (T *)pt_v;
    BA_CUDA(cudaMemcpyAsync(cam, d.cam, sizeof(T) * 9 * g.nC, cudaMemcpyDeviceToDevice, stream));
    if (g.nPl) {
      if (g.ell) { BA_CUDA(launch_permute3<T>(d.pt, pt, perm_dev, g.nPl, 0, stream)); stats.launches += 1; }
      else BA_CUDA(cudaMemcpyAsync(pt, d.pt, sizeof(T) * 3 * g.nPl, cudaMemcpyDeviceToDevice, stream));
    }
    return link_out(caller_stream);
  }

  // BaseVertex::fixed (include/vertex/base_vertex.h): declared upstream, never honoured there.  Semantics here: a fixed
  // vertex keeps its parameters; its Jacobian columns leave the normal equations (identity diagonal block, zero
  // gradient, no coupling), so the PCG iterates are those of the system reduced to the free vertices.
  int set_fixed(const unsigned char *cf, const unsigned char *pf) override {
    BA_CUDA(cudaSetDevice(opt.device));
    if (!g.ell) return fail(BA_ERR_STATE, "fixed vertices need the default (ELL) layout");
    d.cam_fixed = nullptr;
    d.pt_fixed = nullptr;
    if (cf) {
      if (!cam_fixed_dev) BA_CUDA(cudaMalloc(&cam_fixed_dev, std::max<size_t>(g.nC, 1)));
      BA_CUDA(cudaMemcpyAsync(cam_fixed_dev, cf, g.nC, cudaMemcpyHostToDevice, stream));
      d.cam_fixed = cam_fixed_dev;
    }
    if (pf) {
      if (!pt_fixed_dev) BA_CUDA(cudaMalloc(&pt_fixed_dev, std::max<size_t>(g.nPl, 1)));
      std::vector<unsigned char> h(std::max<size_t>(g.nPl, 1), 0);
      for (int k = 0; k < g.nPl; ++k) h[k] = pf[g.pt_begin + g.perm[k]];
      BA_CUDA(cudaMemcpyAsync(pt_fixed_dev, h.data(), g.nPl, cudaMemcpyHostToDevice, stream));
      BA_CUDA(cudaStreamSynchronize(stream));
      d.pt_fixed = pt_fixed_dev;
    }
    BA_CUDA(launch_cam_consts<T>(d, stream));  // T_c of the fixed cameras
    BA_CUDA(cudaStreamSynchronize(stream));
    linearized = false;
    return BA_OK;
  }

  // The residual of every edge as a register program (see ba_program.cuh for the instruction set).  The program
  // replaces the analytical BAL model in the linearisation, the trial-cost evaluation and ba_eval_edges; the
  // operator stream becomes the Jacobian (the factored stream is a property of the BAL projection).
  int set_residual_program(int n_instr, const uint8_t *code, int n_const, const double *consts, int n_slots,
                           const int32_t *out_slot) override {
    BA_CUDA(cudaSetDevice(opt.device));
    if (!g.ell) return fail(BA_ERR_STATE, "a residual program needs the default (ELL) layout");
    if (n_instr == 0 || !code) {  // back to the analytical model
      d.prog_code = nullptr;
      d.prog_const = nullptr;
      d.prog_n = 0;
      d.prog_jit = 0;
      if (store_before_prog >= 0) store_cfg = store_before_prog;
      store_before_prog = -1;
      d.store = store_cfg;
      linearized = false;
      return BA_OK;
    }
    if (n_instr < 0 || n_const < 0 || (n_const > 0 && !consts) || !out_slot)
      return fail(BA_ERR_ARG, "residual program: bad argument");
    if (n_slots < 1 || n_slots > PROG_SLOTS)
      return fail(BA_ERR_ARG, "residual program: %d registers, the interpreter has %d", n_slots, PROG_SLOTS);
    std::vector<ProgInstr> parsed;
    int out_def[2];
    const std::string why = prog_parse(n_instr, code, n_const, n_slots, out_slot, &parsed, out_def);
    if (!why.empty()) return fail(BA_ERR_ARG, "residual program: %s", why.c_str());
    static const bool hoist_env = !(getenv("BA_PROG_HOIST") && atoi(getenv("BA_PROG_HOIST")) == 0);
    const ProgDevice P = prog_lower(parsed, out_def, out_slot, hoist_env);
    std::vector<uint2> h = P.obs_code;
    const size_t n_obs_code = h.size();
    h.insert(h.end(), P.cam_code.begin(), P.cam_code.end());  // one allocation: [per observation | per camera]
    std::vector<T> hc((size_t)std::max(n_const, 1), T(0));
    for (int k = 0; k < n_const; ++k) hc[k] = (T)consts[k];
    if (prog_code_dev) { cudaFree(prog_code_dev); prog_code_dev = nullptr; }
    if (prog_const_dev) { cudaFree(prog_const_dev); prog_const_dev = nullptr; }
    BA_CUDA(cudaMalloc(&prog_code_dev, sizeof(uint2) * h.size()));
    BA_CUDA(cudaMalloc(&prog_const_dev, sizeof(T) * hc.size()));
    if (!cm_slot_dev) {  // the camera side of the linearisation reads the point side's r and J through this map
      BA_CUDA(cudaMalloc(&cm_slot_dev, sizeof(int) * std::max<size_t>(g.cm_slot.size(), 1)));
      BA_CUDA(cudaMalloc(&r_slots_dev, sizeof(T) * 2 * std::max<size_t>(g.edge_of_slot.size(), 1)));
      BA_CUDA(cudaMemcpyAsync(cm_slot_dev, g.cm_slot.data(), sizeof(int) * g.cm_slot.size(), cudaMemcpyHostToDevice, stream));
      d.cm_slot = cm_slot_dev;
      d.r_slots = r_slots_dev;
    }
    BA_CUDA(cudaMemcpyAsync(prog_code_dev, h.data(), sizeof(uint2) * h.size(), cudaMemcpyHostToDevice, stream));
    BA_CUDA(cudaMemcpyAsync(prog_const_dev, hc.data(), sizeof(T) * hc.size(), cudaMemcpyHostToDevice, stream));
    BA_CUDA(cudaStreamSynchronize(stream));
    d.prog_code = prog_code_dev;
    d.prog_const = prog_const_dev;
    d.prog_n = (int)n_obs_code;
    d.prog_cam_code = prog_code_dev + n_obs_code;
    d.prog_cam_n = (int)P.cam_code.size();
    d.prog_nf = P.n_frontier;
    if (prog_tab_dev) { cudaFree(prog_tab_dev); prog_tab_dev = nullptr; }
    BA_CUDA(cudaMalloc(&prog_tab_dev, sizeof(T) * 2 * std::max<size_t>((size_t)g.nC * PROG_FSTRIDE * P.n_frontier, 1)));
    d.prog_tab = prog_tab_dev;
    d.prog_tab2 = prog_tab_dev + (size_t)g.nC * PROG_FSTRIDE * P.n_frontier;
    d.prog_out0 = out_slot[0];
    d.prog_out1 = out_slot[1];
    if (store_before_prog < 0) store_before_prog = store_cfg;
    store_cfg = ST_J;
    d.store = ST_J;
    linearized = false;
    // the per-observation program as a kernel of its own (NVRTC); the interpreter stays for ba_eval_edges, for the
    // per-camera program and as the fallback when NVRTC is not usable on this machine
    d.prog_jit = 0;
    jit_fill = jit_cost = nullptr;
    if (jit_lib) { cudaLibraryUnload(jit_lib); jit_lib = nullptr; }
    jit_note.clear();
    const char *je = getenv("BA_PROG_JIT");
    if (!(je && atoi(je) == 0)) {
      std::vector<double> kc(consts, consts + n_const);
      const std::string src = jit_source(P, kc, n_slots, out_slot, sizeof(T) == 8);
      std::vector<char> cubin;
      std::string why = jit_compile(src, &cubin);
      if (why.empty()) {
        cudaError_t e = cudaLibraryLoadData(&jit_lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
        if (e == cudaSuccess) e = cudaLibraryGetKernel(&jit_fill, jit_lib, "prog_fill");
        if (e == cudaSuccess) e = cudaLibraryGetKernel(&jit_cost, jit_lib, "prog_cost");
        if (e != cudaSuccess) { why = std::string("loading the generated kernels: ") + cudaGetErrorString(e); cudaGetLastError(); }
      }
      if (why.empty() && !slot_cm_dev) {
        std::vector<int> inv(std::max<size_t>(g.edge_of_slot.size(), 1), -1);
        for (size_t k = 0; k < g.cm_slot.size(); ++k) inv[g.cm_slot[k]] = (int)k;
        cudaError_t e = cudaMalloc(&slot_cm_dev, sizeof(int) * inv.size());
        if (e == cudaSuccess) e = cudaMalloc(&cm_jr_dev, sizeof(T) * 20 * std::max<size_t>(g.cm_slot.size(), 1));
        if (e == cudaSuccess) e = cudaMemcpyAsync(slot_cm_dev, inv.data(), sizeof(int) * inv.size(), cudaMemcpyHostToDevice, stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess) { why = std::string("buffers of the generated kernels: ") + cudaGetErrorString(e); cudaGetLastError(); }
        d.slot_cm = slot_cm_dev;
        d.cm_jr = cm_jr_dev;
      }
      if (why.empty()) d.prog_jit = 1;
      else jit_note = why;  // ba_residual_program_mode reports 1; the reason is kept for ba_last_error
    }
    if (!jit_note.empty()) err = "residual program runs in the interpreter: " + jit_note;
    return BA_OK;
  }

  int residual_program_mode() override { return !d.prog_code ? 0 : d.prog_jit ? 2 : 1; }

  // the point side of the linearisation by the generated kernel (in place of k_lin_ell<PROG>)
  int jit_fill_stream() {
    launch_prog_cam<T>(d, 0, stream);
    if (d.nW > 0) {
      int nW = d.nW, stride = PROG_FSTRIDE * d.prog_nf;
      const void *wdesc = d.wdesc, *r_cam = d.r_cam, *r_obs = d.r_obs, *cam = d.camc, *pt = d.pt, *tab = d.prog_tab;
      void *WJ = d.WJ, *rs = d.r_slots, *cmjr = d.cm_jr, *Hll = d.Hll, *gl = d.gl, *partials = d.partials;
      const void *scm = d.slot_cm, *info = d.r_info, *cfix = d.cam_fixed, *pfix = d.pt_fixed;
      void *args[] = {&wdesc, &nW, &r_cam, &r_obs, &cam, &pt, &tab, &stride, &WJ, &rs, &scm, &cmjr, &info, &cfix, &pfix, &Hll, &gl,
                      &partials};
      BA_CUDA(cudaLaunchKernel((const void *)jit_fill, dim3((nW + 3) / 4), dim3(128), args, 0, stream));
    }
    stats.launches += 2;
    return BA_OK;
  }
  int jit_trial_cost(int pstride) {  // partials[w * pstride] = sum of squared trial residuals of warp w
    launch_prog_cam<T>(d, 1, stream);
    if (d.nW > 0) {
      int nW = d.nW, stride = PROG_FSTRIDE * d.prog_nf;
      // camc2: the trial cameras cam + dx_c, written by k_prog_cam just above
      const void *wdesc = d.wdesc, *r_cam = d.r_cam, *r_obs = d.r_obs, *cam = d.camc2, *pt = d.pt, *dxl = d.dxl, *tab = d.prog_tab2;
      void *partials = d.partials;
      void *args[] = {&wdesc, &nW, &r_cam, &r_obs, &cam, &pt, &dxl, &tab, &stride, &partials, &pstride};
      BA_CUDA(cudaLaunchKernel((const void *)jit_cost, dim3((nW + 3) / 4), dim3(128), args, 0, stream));
    }
    stats.launches += 2;
    return BA_OK;
  }

  int set_information(const void *info_v) override {
    BA_CUDA(cudaSetDevice(opt.device));
    if (!info_v) {
      d.e_info = nullptr;
      d.r_info = nullptr;
      d.cm_info = nullptr;
      linearized = false;
      return BA_OK;
    }
    if (g.ell) {
      // Both compute kinds apply W = Jc^T Omega Jp through the stored factors, so information matrices work in
      // implicit mode too (upstream's implicit build has an empty branch there, build_implicit_linear_system.cu:272-274;
      // the semantics chosen here are those of its explicit build, build_linear_system.cu:149-239).
      const T *info = (const T *)info_v;
      const size_t nSlots = (size_t)g.nRows * 32, nEl = g.nEl;
      std::vector<T> h_e(4 * nSlots, T(0)), h_c(4 * nEl);
      for (size_t sl = 0; sl < nSlots; ++sl) {
        const int64_t e = g.edge_of_slot[sl];
        if (e >= 0) for (int k = 0; k < 4; ++k) h_e[4 * sl + k] = info[4 * e + k];
      }
      for (size_t k = 0; k < nEl; ++k) {
        const int64_t e = g.edge_of_slot[g.cm_slot[k]];
        for (int q = 0; q < 4; ++q) h_c[4 * k + q] = info[4 * e + q];
      }
      if (!r_info_dev) {
        BA_CUDA(cudaMalloc(&r_info_dev, sizeof(T) * std::max<size_t>(4 * nSlots, 4)));
        BA_CUDA(cudaMalloc(&cm_info_dev, sizeof(T) * std::max<size_t>(4 * nEl, 4)));
      }
      BA_CUDA(cudaMemcpyAsync(r_info_dev, h_e.data(), sizeof(T) * 4 * nSlots, cudaMemcpyHostToDevice, stream));
      BA_CUDA(cudaMemcpyAsync(cm_info_dev, h_c.data(), sizeof(T) * 4 * nEl, cudaMemcpyHostToDevice, stream));
      BA_CUDA(cudaStreamSynchronize(stream));
      d.r_info = r_info_dev;
      d.cm_info = cm_info_dev;
      linearized = false;
      return BA_OK;
    }
    if (opt.compute_kind == BA_IMPLICIT)
      return fail(BA_ERR_ARG, "information matrices are not supported in implicit mode "
                              "(the reference's implicit build ignores them, build_implicit_linear_system.cu:272-274)");
    const T *info = (const T *)info_v;
    const size_t nSlots = (size_t)g.nSlabs * TE, nEl = g.nEl;
    std::vector<T> h_e(4 * nSlots, T(0)), h_c(4 * nEl);
    for (size_t s = 0; s < nSlots; ++s) {
      const int64_t e = g.edge_of_slot[s];
      if (e >= 0) for (int k = 0; k < 4; ++k) h_e[4 * s + k] = info[4 * e + k];
    }
    for (size_t k = 0; k < nEl; ++k) {
      const int64_t e = g.edge_of_slot[g.cm_slot[k]];
      for (int q = 0; q < 4; ++q) h_c[4 * k + q] = info[4 * e + q];
    }
    if (!e_info_dev) {
      BA_CUDA(cudaMalloc(&e_info_dev, sizeof(T) * std::max<size_t>(4 * nSlots, 4)));
      BA_CUDA(cudaMalloc(&cm_info_dev, sizeof(T) * std::max<size_t>(4 * nEl, 4)));
    }
    BA_CUDA(cudaMemcpyAsync(e_info_dev, h_e.data(), sizeof(T) * 4 * nSlots, cudaMemcpyHostToDevice, stream));
    BA_CUDA(cudaMemcpyAsync(cm_info_dev, h_c.data(), sizeof(T) * 4 * nEl, cudaMemcpyHostToDevice, stream));
    BA_CUDA(cudaStreamSynchronize(stream));
    d.e_info = e_info_dev;
    d.cm_info = cm_info_dev;
    linearized = false;
    return BA_OK;
  }

  // ---- HP1 --------------------------------------------------------------------
  int linearize(double *cost) override {
    BA_CUDA(cudaSetDevice(opt.device));
    size_t ev = 0;
    if (timing) ev = ev_begin();
    if (g.ell) {
      d.store = store_cfg;  // the launcher wants the configured store; the one in effect is known after the fetch below
      if (d.prog_code && d.prog_jit) BA_TRY(jit_fill_stream());
      BA_CUDA(launch_linearize_ell<T>(d, stream));
      // factored store: both point-side flavours + both camera-side flavours (device-side dispatch) + the cost reduction;
      // Jacobian store: k_lin_ell, k_lin_cm, k_lin_cam_reduce, k_reduce; a residual program adds k_prog_cam (interpreter) or
      // replaces k_lin_ell by the generated kernel (counted in jit_fill_stream)
      stats.launches += store_cfg == ST_FACT ? launches_of_linearize_ell() : 4 + (d.prog_code ? (d.prog_jit ? -1 : 1) : 0);
    } else {
      BA_CUDA(launch_linearize<T>(d, stream));
      stats.launches += launches_of_linearize();
    }
    stats.linearize_calls += 1;
    BA_TRY(allreduce_T(d.Hpp, 90 * (size_t)g.nC));
    BA_TRY(allreduce_scal(S_COST, 1, ncclSum));
    if (timing) { ev_end(ev); ev_lin.push_back(ev); }
    if (fuse_ginf) {  // LM loop: ||g||_inf of the new gradient travels with the cost (one round trip less)
      BA_CUDA(launch_ginf<T>(d, stream));
      stats.launches += 2;
      BA_TRY(allreduce_scal(S_GINF, 1, ncclMax));
    }
    BA_TRY(fetch_scalars());
    ginf_valid = fuse_ginf;
    if (cost) *cost = h_scal[S_COST];
    last_cost = h_scal[S_COST];
    // a camera in the small-angle branch (theta^2 <= eps) makes this linearisation stream the Jacobian instead
    if (g.ell) d.store = (store_cfg == ST_FACT && h_scal[S_ANY_SMALL] == 0.) ? ST_FACT : ST_J;
    linearized = true;
    return BA_OK;
  }

  int process_diag(double region_, int /*recover*/) override {
    if (!(region_ > 0)) return fail(BA_ERR_ARG, "region must be positive (got %g)", region_);
    region = region_;
    d.damp = 1. / region_;
    return BA_OK;
  }

  // ---- HP2 --------------------------------------------------------------------
  int schur_edge(int mode, const T *xin, T *sout, int check_active, bool timed, T *zero_set = nullptr) {
    size_t ev = 0;
    if (timing && timed) ev = ev_begin();
    if (g.ell) BA_CUDA(launch_schur_ell<T>(d, mode, xin, sout, zero_set, check_active, stream));
    else BA_CUDA(launch_schur_edge<T>(d, mode, xin, sout, zero_set, check_active, stream));
    if (timing && timed) { ev_end(ev); ev_schur.push_back(ev); }
    stats.launches += (d.nT > 0 || d.nW > 0) ? 1 : 0;
    if (timed) stats.schur_apply_calls += 1;
    return BA_OK;
  }

  // block inverses (+ w, plane copies) and the padded / transformed copy of a camera vector the edge pass gathers
  int prep() {
    if (g.ell) { BA_CUDA(launch_prep_ell<T>(d, stream)); BA_CUDA(launch_prep<T>(d, stream)); }
    else BA_CUDA(launch_prep<T>(d, stream));
    stats.launches += launches_of_prep();
    return BA_OK;
  }
  int pad_cam(const T *src) {
    if (d.store == ST_FACT) BA_CUDA(launch_pad_cam_T<T>(d, src, stream));
    else BA_CUDA(launch_pad_cam<T>(d, src, stream));
    stats.launches += 1;
    return BA_OK;
  }

  // ba_solve = enqueue + finish.  ba_solve_async enqueues every launch of the solve (no polling: iterations past the
  // end of the reference's do-while are no-ops on the device) and returns without a host synchronisation;
  // ba_solve_wait fetches the final PCG state.
  int solve(const ba_pcg_options *o, int *iters, int *done) override {
    BA_TRY(solve_enqueue(o, true));
    return solve_finish(iters, done);
  }
  int solve_async(const ba_pcg_options *o) override { return solve_enqueue(o, false); }
  int solve_wait(int *iters, int *done) override {
    if (!solve_pending) return fail(BA_ERR_STATE, "ba_solve_wait without ba_solve_async");
    return solve_finish(iters, done);
  }

  int solve_enqueue(const ba_pcg_options *o, bool polling) {
    if (!o) return fail(BA_ERR_ARG, "null pcg options");
    if (!linearized) return fail(BA_ERR_STATE, "ba_solve before ba_linearize");
    if (solve_pending) return fail(BA_ERR_STATE, "ba_solve_async is still pending: call ba_solve_wait first");
    BA_CUDA(cudaSetDevice(opt.device));
    size_t ev_s = 0, ev_p = 0;
    if (timing) ev_s = ev_begin();
    const int max_iter = std::max(1, (int)o->max_iter);  // the reference's do-while always runs its first iteration
    const bool multi = opt.world_size > 1;
    // multi-GPU: either the fused peer-memory exchange inside the camera kernel, or fold + NCCL allreduce
    static const bool no_p2p_env = getenv("BA_NO_P2P") != nullptr;
    const bool fused = multi && p2p && !no_p2p_env && pcg_reg_kernel_available<T>(d);
    const int nrep = (multi && !fused) ? 1 : NREP;
    PeerExchange pxl = px;
    if (!fused) pxl.ws = 1;
    auto next_epoch = [&]() { if (fused) { pxl.epoch = ++px_epoch; pxl.parity = (int)(px_epoch & 1); } };
    BA_CUDA(cudaMemsetAsync(d.sA, 0, sizeof(T) * 3 * NREP * 9 * g.nC, stream));
    if (fused) {
      // a time-out of an earlier solve must not poison this one, and no rank may start publishing into an exchange
      // buffer that a slower peer is still reading in the last iteration of ITS previous solve: one tiny allreduce
      // re-aligns the ranks (the fused path has no other collective between solves)
      BA_CUDA(cudaMemsetAsync(d.scal + S_XCHG_ERR, 0, sizeof(double), stream));
      BA_TRY(allreduce_scal(S_XCHG_ERR, 1, ncclMax));
    }
    const size_t set_sz = (size_t)NREP * 9 * g.nC;  // sB set k lives at d.sB + k * set_sz
    BA_TRY(prep());
    // MakeV and the operator application on the warm start x0 share ONE pass over W:
    //   sB = Hpl (Hll^-1 Hlp x0 - w)  so that  r0 = v - S x0 = g_c - Hpp x0 + sB   (schur_pcg_solver.cu:430-510,202-258);
    // sA (the reference's separate Hpl w) stays zero.
    BA_TRY(pad_cam(d.dxc));
    static const int persist_env = getenv("BA_PERSIST") ? atoi(getenv("BA_PERSIST")) : -1;  // -1: follow ba_set_persistent
    persistent_used = (persist_env < 0 ? persistent : persist_env != 0) && (!multi || fused) && pcg_persistent_available<T>(d);
    if (persistent_used) {
      // ---- the whole solve (MakeV + S*x0, the do-while, back-substitution) in one cooperative launch
      BA_CUDA(cudaMemsetAsync(d.pk_counters, 0, 4 * sizeof(int), stream));
      BA_CUDA(cudaMemsetAsync(d.pk_timers, 0, 4 * sizeof(double), stream));
      if (fused) { pxl.epoch = px_epoch + 1; px_epoch += (unsigned long long)max_iter + 2; }
      if (timing) ev_p = ev_begin();
      BA_CUDA(launch_pcg_persistent<T>(d, max_iter, o->tol, o->refuse_ratio, pxl, stream));
      if (timing) { ev_end(ev_p); ev_pcg.push_back(ev_p); ev_end(ev_s); ev_solve.push_back(ev_s); }
      stats.launches += 1;
      model_valid = false;
      launched = 0;
      loc_schur.clear();
      loc_cam.clear();
      if (fuse_norms) {
        BA_CUDA(launch_norms<T>(d, stream));
        stats.launches += 2;
        BA_TRY(allreduce_scal(S_DX2_PT, 2, ncclSum));
      }
      solve_pending = true;
      return BA_OK;
    }
    BA_TRY(schur_edge(SM_INIT, d.xpad, d.sB, 0, true));
    if (multi && !fused) {
      BA_CUDA(launch_fold_replicas<T>(d, d.sA, 2, stream));
      stats.launches += 1;
      BA_TRY(allreduce_T(d.sA, 9 * (size_t)g.nC));
      BA_TRY(allreduce_T(d.sB, 9 * (size_t)g.nC));
    }
    next_epoch();
    BA_CUDA(launch_pcg_init<T>(d, d.sB, max_iter, o->tol, o->refuse_ratio, nrep, pxl, stream));
    stats.launches += 1;
    if (timing) ev_p = ev_begin();
    // The loop is enqueued in chunks; the device-side `active` flag turns the remaining launches of a
    // chunk into no-ops once the reference's do-while would have exited.  The state of chunk k is polled
    // only after chunk k+1 has been enqueued, so the GPU never waits for the host.
    int chunk_id = 0;
    launched = 0;
    struct Pending { int slot; bool valid; } pend{0, false};
    loc_schur.clear();
    loc_cam.clear();
    bool finished = false;
    auto poll = [&](const Pending &pd) -> int {
      BA_CUDA(cudaEventSynchronize(st_event[pd.slot]));
      const PcgState &hs = h_st[pd.slot];
      if (!hs.active) finished = true;
      return BA_OK;
    };
    // first chunk: what the previous solve needed (+1), so that a loop that exits after two or three iterations
    // (README tolerances) is not followed by a tail of no-op launches; later chunks: pcg_chunk
    const int first_chunk = std::max(2, std::min(pcg_chunk, last_iters + 1));
    while (launched < max_iter && !finished) {
      const int n = std::min(launched == 0 ? first_chunk : pcg_chunk, max_iter - launched);
      for (int i = 0; i < n; ++i) {
        // iteration it accumulates into set (it+1)&1 and clears set it&1 (read by the previous camera kernel)
        const int it = launched + i;
        T *cur = d.sB + (size_t)((it + 1) & 1) * set_sz, *other = d.sB + (size_t)(it & 1) * set_sz;
        const bool sample = timing && (it % timing_stride) == 0;
        size_t eve = 0;
        if (sample) eve = ev_begin();
        BA_TRY(schur_edge(SM_FULL, d.xpad, cur, 1, false, other));  // xpad holds p (written by the camera kernel)
        if (sample) { ev_end(eve); loc_schur.push_back({eve, it}); }
        stats.schur_apply_calls += 1;
        if (multi && !fused) {
          BA_CUDA(launch_fold_replicas<T>(d, cur, 1, stream));
          stats.launches += 1;
          BA_TRY(allreduce_T(cur, 9 * (size_t)g.nC));
        }
        next_epoch();
        size_t evc = 0;
        if (sample) evc = ev_begin();
        BA_CUDA(launch_pcg_step<T>(d, cur, nrep, pxl, stream));
        if (sample) { ev_end(evc); loc_cam.push_back({evc, it}); }
        stats.launches += 1;
      }
      launched += n;
      const int slot = chunk_id++ & 1;
      BA_CUDA(cudaMemcpyAsync(&h_st[slot], d.st, sizeof(PcgState), cudaMemcpyDeviceToHost, stream));
      BA_CUDA(cudaEventRecord(st_event[slot], stream));
      if (pend.valid && polling) BA_TRY(poll(pend));
      pend = Pending{slot, true};
    }
    if (timing) { ev_end(ev_p); ev_pcg.push_back(ev_p); }
    // Back-substitution (and, inside the LM loop, the step norms) are enqueued BEHIND the PCG launches without waiting
    // for them: launches past the end of the loop are no-ops, so one host round trip at the very end fetches the final
    // PCG state and the scalars together.
    BA_TRY(pad_cam(d.dxc));
    // Without information matrices H = J^T J exactly, so the model decrease ||J dx + r||^2 - ||r||^2 of the trial step
    // is  dx^T H dx - 2 g^T dx : the back-substitution pass has Hlp dx_c per point in hand and leaves the point part,
    // a one-block kernel adds the camera part -- ba_eval_trial then only needs the trial cost.
    static const bool model_env = !(getenv("BA_MODEL_QUAD") && atoi(getenv("BA_MODEL_QUAD")) == 0);
    const bool quad = g.ell && (d.store == ST_FACT || (d.prog_code && d.prog_jit)) && !d.r_info && model_env && fuse_norms;
    d.model_partials = quad ? model_part : nullptr;
    BA_TRY(schur_edge(SM_BACKSUB, d.xpad, nullptr, 0, false));  // dx_l = w - Hll^-1 Hlp dx_c
    d.model_partials = nullptr;
    if (quad) {
      Dev<T> dm = d;
      dm.model_partials = model_part;
      BA_CUDA(launch_model_terms<T>(dm, stream));
      stats.launches += 2;
      BA_TRY(allreduce_scal(S_MODEL_PT, 1, ncclSum));
    }
    model_valid = quad;
    if (timing) { ev_end(ev_s); ev_solve.push_back(ev_s); }
    if (fuse_norms) {
      BA_CUDA(launch_norms<T>(d, stream));
      stats.launches += 2;
      BA_TRY(allreduce_scal(S_DX2_PT, 2, ncclSum));
    }
    solve_pending = true;
    return BA_OK;
  }

  int solve_finish(int *iters, int *done) {
    BA_CUDA(cudaSetDevice(opt.device));
    solve_pending = false;
    {
      // final state (also covers the over-enqueued chunk)
      BA_CUDA(cudaMemcpyAsync(&h_st[0], d.st, sizeof(PcgState), cudaMemcpyDeviceToHost, stream));
      BA_CUDA(cudaMemcpyAsync(h_scal, d.scal, sizeof(double) * S_COUNT, cudaMemcpyDeviceToHost, stream));
      BA_CUDA(cudaStreamSynchronize(stream));
      if (h_scal[S_XCHG_ERR] != 0.) return fail(BA_ERR_NCCL, "peer-memory exchange timed out (a rank did not arrive)");
      norms_valid = fuse_norms;
    }
    if (persistent_used) {
      // per-phase device time from %globaltimer (CTA 0): the edge phases are the launches of the S*p pass
      double tm[4];
      BA_CUDA(cudaMemcpy(tm, d.pk_timers, sizeof tm, cudaMemcpyDeviceToHost));
      stats.schur_apply_calls += h_st[0].n + 1;
      if (timing) {
        stats.schur_apply_ms += tm[0] * 1e-6;
        stats.schur_apply_timed += tm[1];
        stats.pcg_cam_ms += tm[2] * 1e-6;
        stats.pcg_cam_timed += tm[3];
      }
    }
    // inactive edge launches return immediately; only executed iterations are counted
    if (!persistent_used) stats.schur_apply_calls -= (launched - h_st[0].n);
    // launches issued after the loop ended returned immediately: keep only the samples of iterations that ran
    for (auto &e : loc_schur) if (e.second < h_st[0].n) ev_schur.push_back(e.first);
    for (auto &e : loc_cam) if (e.second < h_st[0].n) ev_cam.push_back(e.first);
    stats.pcg_iters += h_st[0].n;
    last_iters = h_st[0].n;
    if (iters) *iters = h_st[0].n;
    if (done) *done = h_st[0].done;
    return BA_OK;
  }

  int step_norms(double *dxn, double *xn) override {
    BA_CUDA(cudaSetDevice(opt.device));
    if (!norms_valid) {  // otherwise ba_solve computed and fetched them already (LM loop)
      BA_CUDA(launch_norms<T>(d, stream));
      stats.launches += 2;
      BA_TRY(allreduce_scal(S_DX2_PT, 2, ncclSum));
      BA_TRY(fetch_scalars());
    }
    norms_valid = false;
    if (dxn) *dxn = std::sqrt(h_scal[S_DX2_PT] + h_scal[S_DX2_CAM]);
    if (xn) *xn = std::sqrt(h_scal[S_X2_PT] + h_scal[S_X2_CAM]);
    return BA_OK;
  }

  int eval_trial(double *cost_new, double *model_sq) override {
    BA_CUDA(cudaSetDevice(opt.device));
    if (g.ell && model_valid && (d.store == ST_FACT || (d.prog_code && d.prog_jit))) {
      // the model term came with the solve (see solve_enqueue); only the trial cost needs the observations
      model_valid = false;
      const double msq = last_cost + h_scal[S_MODEL_PT] + h_scal[S_MODEL_CAM];
      if (d.prog_code) {  // user-defined residual: the generated prog_cost kernel, one partial per warp
        BA_TRY(jit_trial_cost(1));
        BA_CUDA(launch_reduce_partials<T>(d, 1, S_TRIAL_COST, stream));
        stats.launches += 1;
      } else {
        BA_CUDA(launch_trial_cost_ell<T>(d, stream));
        stats.launches += 3;
      }
      BA_TRY(allreduce_scal(S_TRIAL_COST, 1, ncclSum));
      BA_TRY(fetch_scalars());
      if (cost_new) *cost_new = h_scal[S_TRIAL_COST];
      if (model_sq) *model_sq = msq;
      return BA_OK;
    }
    if (g.ell) {
      if (d.store == ST_FACT) BA_TRY(pad_cam(d.dxc));  // the factored trial kernel gathers T_c dx_c
      if (d.prog_code && d.prog_jit) BA_TRY(jit_trial_cost(2));
      BA_CUDA(launch_trial_ell<T>(d, stream));
    } else {
      BA_CUDA(launch_trial<T>(d, stream));
    }
    stats.launches += 2;
    BA_TRY(allreduce_scal(S_TRIAL_COST, 2, ncclSum));
    BA_TRY(fetch_scalars());
    if (cost_new) *cost_new = h_scal[S_TRIAL_COST];
    if (model_sq) *model_sq = h_scal[S_MODEL_SQ];
    return BA_OK;
  }

  int accept_step() override {
    BA_CUDA(cudaSetDevice(opt.device));
    BA_CUDA(launch_apply_step<T>(d, stream));
    stats.launches += 1;
    if (g.ell) { BA_CUDA(launch_cam_consts<T>(d, stream)); stats.launches += 1; }
    linearized = false;
    return BA_OK;
  }

  int reject_step() override {
    BA_CUDA(cudaSetDevice(opt.device));
    BA_CUDA(cudaMemcpyAsync(d.dxc, d.dxc_bk, sizeof(T) * 9 * g.nC, cudaMemcpyDeviceToDevice, stream));
    if (g.nPl) BA_CUDA(cudaMemcpyAsync(d.dxl, d.dxl_bk, sizeof(T) * 3 * g.nPl, cudaMemcpyDeviceToDevice, stream));
    return BA_OK;
  }

  int backup_step() override {
    BA_CUDA(cudaSetDevice(opt.device));
    BA_CUDA(cudaMemcpyAsync(d.dxc_bk, d.dxc, sizeof(T) * 9 * g.nC, cudaMemcpyDeviceToDevice, stream));
    if (g.nPl) BA_CUDA(cudaMemcpyAsync(d.dxl_bk, d.dxl, sizeof(T) * 3 * g.nPl, cudaMemcpyDeviceToDevice, stream));
    return BA_OK;
  }

  int grad_inf(double *gi) override {
    BA_CUDA(cudaSetDevice(opt.device));
    if (!ginf_valid) {
      BA_CUDA(launch_ginf<T>(d, stream));
      stats.launches += 2;
      BA_TRY(allreduce_scal(S_GINF, 1, ncclMax));
      BA_TRY(fetch_scalars());
    }
    ginf_valid = false;
    if (gi) *gi = h_scal[S_GINF];
    return BA_OK;
  }

  // ---- inspection -------------------------------------------------------------
  int read(int which, void *out_v) override {
    if (!out_v) return fail(BA_ERR_ARG, "null output buffer");
    BA_CUDA(cudaSetDevice(opt.device));
    T *out = (T *)out_v;
    const size_t nC = g.nC, nP = g.nP, nPl = g.nPl, pb = g.pt_begin;
    if (g.ell && which != BA_BUF_HPP) return read_ell(which, out);
    switch (which) {
      case BA_BUF_HPP:
        BA_CUDA(cudaMemcpyAsync(out, d.Hpp, sizeof(T) * 81 * nC, cudaMemcpyDeviceToHost, stream));
        break;
      case BA_BUF_HLL:
        memset(out, 0, sizeof(T) * 9 * nP);
        if (nPl) BA_CUDA(cudaMemcpyAsync(out + 9 * pb, d.Hll, sizeof(T) * 9 * nPl, cudaMemcpyDeviceToHost, stream));
        break;
      case BA_BUF_G:
        memset(out, 0, sizeof(T) * (9 * nC + 3 * nP));
        BA_CUDA(cudaMemcpyAsync(out, d.gc, sizeof(T) * 9 * nC, cudaMemcpyDeviceToHost, stream));
        if (nPl) BA_CUDA(cudaMemcpyAsync(out + 9 * nC + 3 * pb, d.gl, sizeof(T) * 3 * nPl, cudaMemcpyDeviceToHost, stream));
        break;
      case BA_BUF_DX:
        memset(out, 0, sizeof(T) * (9 * nC + 3 * nP));
        BA_CUDA(cudaMemcpyAsync(out, d.dxc, sizeof(T) * 9 * nC, cudaMemcpyDeviceToHost, stream));
        if (nPl) BA_CUDA(cudaMemcpyAsync(out + 9 * nC + 3 * pb, d.dxl, sizeof(T) * 3 * nPl, cudaMemcpyDeviceToHost, stream));
        break;
      case BA_BUF_W:
      case BA_BUF_J: {
        const int K = d.implicit ? 24 : 27;
        if ((which == BA_BUF_W) == (bool)d.implicit)
          return fail(BA_ERR_STATE, "BA_BUF_W exists in explicit mode, BA_BUF_J in implicit mode");
        const size_t nSlots = (size_t)g.nSlabs * TE;
        std::vector<T> h((size_t)K * nSlots);
        if (nSlots) BA_CUDA(cudaMemcpyAsync(h.data(), d.WJ, sizeof(T) * K * nSlots, cudaMemcpyDeviceToHost, stream));
        BA_CUDA(cudaStreamSynchronize(stream));
        for (int64_t e = 0; e < g.nE; ++e) {
          const int slot = g.slot_of_edge[e];
          for (int k = 0; k < K; ++k)
            out[(size_t)K * e + k] = slot < 0 ? T(0) : h[(size_t)(slot / TE) * K * TE + (size_t)k * TE + slot % TE];
        }
        break;
      }
      default:
        return fail(BA_ERR_ARG, "unknown buffer id %d", which);
    }
    BA_CUDA(cudaStreamSynchronize(stream));
    return BA_OK;
  }

  // ELL: per-point device arrays are in device order (g.perm); W / J are produced from the Jacobian kernel
  int read_ell(int which, T *out) {
    const size_t nC = g.nC, nP = g.nP, nPl = g.nPl, pb = g.pt_begin;
    auto fetch_points = [&](const T *dev, int width, T *dst /* [nP][width] region start */) -> int {
      std::vector<T> h(width * nPl);
      if (nPl) BA_CUDA(cudaMemcpyAsync(h.data(), dev, sizeof(T) * width * nPl, cudaMemcpyDeviceToHost, stream));
      BA_CUDA(cudaStreamSynchronize(stream));
      for (size_t k = 0; k < nPl; ++k)
        for (int j = 0; j < width; ++j) dst[width * (pb + g.perm[k]) + j] = h[width * k + j];
      return BA_OK;
    };
    switch (which) {
      case BA_BUF_HLL:
        memset(out, 0, sizeof(T) * 9 * nP);
        BA_TRY(fetch_points(d.Hll, 9, out));
        break;
      case BA_BUF_G:
        memset(out, 0, sizeof(T) * (9 * nC + 3 * nP));
        BA_CUDA(cudaMemcpyAsync(out, d.gc, sizeof(T) * 9 * nC, cudaMemcpyDeviceToHost, stream));
        BA_TRY(fetch_points(d.gl, 3, out + 9 * nC));
        break;
      case BA_BUF_DX:
        memset(out, 0, sizeof(T) * (9 * nC + 3 * nP));
        BA_CUDA(cudaMemcpyAsync(out, d.dxc, sizeof(T) * 9 * nC, cudaMemcpyDeviceToHost, stream));
        BA_TRY(fetch_points(d.dxl, 3, out + 9 * nC));
        break;
      case BA_BUF_W:
      case BA_BUF_J: {
        if ((which == BA_BUF_W) == (opt.compute_kind == BA_IMPLICIT))
          return fail(BA_ERR_STATE, "BA_BUF_W exists in explicit mode, BA_BUF_J in implicit mode");
        // the ELL path never materialises Hpl; both buffers are views of the per-observation Jacobian
        std::vector<T> r(2 * (size_t)g.nE), J(24 * (size_t)g.nE);
        BA_TRY(eval_edges(r.data(), J.data()));
        if (which == BA_BUF_J) { memcpy(out, J.data(), sizeof(T) * 24 * g.nE); break; }
        std::vector<T> h_info;
        if (d.r_info) {
          h_info.resize(4 * (size_t)g.nRows * 32);
          BA_CUDA(cudaMemcpyAsync(h_info.data(), d.r_info, sizeof(T) * h_info.size(), cudaMemcpyDeviceToHost, stream));
          BA_CUDA(cudaStreamSynchronize(stream));
        }
        for (int64_t e = 0; e < g.nE; ++e) {
          const int slot = g.slot_of_edge[e];
          const T *Je = J.data() + 24 * e;
          T om[4] = {T(1), T(0), T(0), T(1)};
          if (d.r_info && slot >= 0) for (int k = 0; k < 4; ++k) om[k] = h_info[4 * (size_t)slot + k];
          for (int y = 0; y < 9; ++y) {
            const T L0 = Je[y] * om[0] + Je[12 + y] * om[2], L1 = Je[y] * om[1] + Je[12 + y] * om[3];
            for (int q = 0; q < 3; ++q) out[27 * e + 3 * y + q] = slot < 0 ? T(0) : L0 * Je[9 + q] + L1 * Je[21 + q];
          }
        }
        break;
      }
      default:
        return fail(BA_ERR_ARG, "unknown buffer id %d", which);
    }
    BA_CUDA(cudaStreamSynchronize(stream));
    return BA_OK;
  }

  int write_dx(const void *dx_v) override {
    if (!dx_v) return fail(BA_ERR_ARG, "null dx");
    BA_CUDA(cudaSetDevice(opt.device));
    const T *dx = (const T *)dx_v;
    BA_CUDA(cudaMemcpyAsync(d.dxc, dx, sizeof(T) * 9 * g.nC, cudaMemcpyHostToDevice, stream));
    if (g.ell && g.nPl) {
      BA_CUDA(cudaMemcpyAsync(pt_nat, dx + 9 * g.nC + 3 * g.pt_begin, sizeof(T) * 3 * g.nPl, cudaMemcpyHostToDevice, stream));
      BA_CUDA(launch_permute3<T>(pt_nat, d.dxl, perm_dev, g.nPl, 1, stream));
    } else if (g.nPl)
      BA_CUDA(cudaMemcpyAsync(d.dxl, dx + 9 * g.nC + 3 * g.pt_begin, sizeof(T) * 3 * g.nPl, cudaMemcpyHostToDevice, stream));
    BA_CUDA(cudaStreamSynchronize(stream));
    return BA_OK;
  }

  int eval_edges(void *r_v, void *J_v) override {
    if (!r_v) return fail(BA_ERR_ARG, "null residual buffer");
    BA_CUDA(cudaSetDevice(opt.device));
    const size_t nSlots = g.ell ? (size_t)g.nRows * 32 : (size_t)g.nSlabs * TE;
    T *dr = nullptr, *dJ = nullptr;
    BA_CUDA(cudaMalloc(&dr, sizeof(T) * std::max<size_t>(2 * nSlots, 2)));
    if (J_v) BA_CUDA(cudaMalloc(&dJ, sizeof(T) * std::max<size_t>(24 * nSlots, 24)));
    cudaError_t e = g.ell ? launch_eval_edges_ell<T>(d, dr, dJ, stream) : launch_eval_edges<T>(d, dr, dJ, stream);
    stats.launches += 1;
    std::vector<T> hr(2 * nSlots), hJ(J_v ? 24 * nSlots : 0);
    if (e == cudaSuccess && nSlots) e = cudaMemcpyAsync(hr.data(), dr, sizeof(T) * 2 * nSlots, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess && J_v && nSlots) e = cudaMemcpyAsync(hJ.data(), dJ, sizeof(T) * 24 * nSlots, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    cudaFree(dr);
    if (dJ) cudaFree(dJ);
    BA_CUDA(e);
    T *r = (T *)r_v, *J = (T *)J_v;
    for (int64_t ed = 0; ed < g.nE; ++ed) {
      const int slot = g.slot_of_edge[ed];
      for (int k = 0; k < 2; ++k) r[2 * ed + k] = slot < 0 ? T(0) : hr[2 * (size_t)slot + k];
      if (J) for (int k = 0; k < 24; ++k) J[24 * ed + k] = slot < 0 ? T(0) : hJ[24 * (size_t)slot + k];
    }
    return BA_OK;
  }

  int schur_apply(const void *p_v, void *q_v) override {
    if (!p_v || !q_v) return fail(BA_ERR_ARG, "null vector");
    if (!linearized) return fail(BA_ERR_STATE, "ba_schur_apply before ba_linearize");
    BA_CUDA(cudaSetDevice(opt.device));
    BA_TRY(prep());
    BA_CUDA(cudaMemsetAsync(d.sB, 0, sizeof(T) * NREP * 9 * g.nC, stream));
    BA_CUDA(cudaMemcpyAsync(d.z, p_v, sizeof(T) * 9 * g.nC, cudaMemcpyHostToDevice, stream));
    BA_TRY(pad_cam(d.z));
    BA_TRY(schur_edge(SM_FULL, d.xpad, d.sB, 0, true));
    BA_CUDA(launch_fold_replicas<T>(d, d.sB, 1, stream));
    BA_TRY(allreduce_T(d.sB, 9 * (size_t)g.nC));
    BA_CUDA(launch_cam_matvec_sub<T>(d, d.z, d.sB, d.q, stream));
    stats.launches += 2;
    BA_CUDA(cudaMemcpyAsync(q_v, d.q, sizeof(T) * 9 * g.nC, cudaMemcpyDeviceToHost, stream));
    BA_CUDA(cudaMemsetAsync(d.sB, 0, sizeof(T) * NREP * 9 * g.nC, stream));
    BA_CUDA(cudaStreamSynchronize(stream));
    return BA_OK;
  }

  int get_stats(ba_stats *out) override {
    if (!out) return fail(BA_ERR_ARG, "null stats");
    BA_CUDA(cudaSetDevice(opt.device));
    fold_events();
    *out = stats;
    return BA_OK;
  }
  int reset_stats() override {
    BA_CUDA(cudaSetDevice(opt.device));
    fold_events();
    stats = ba_stats{};
    return BA_OK;
  }

  // A_pcg = nE(27s+4) + nP(9s+4) + nC 216 s  (24s for implicit);
  // A_lin = nE(8+29s) + nP 15 s + nC 99 s  (26s implicit).  Per rank: local nE, nP.
  int algorithmic_bytes(double *a, double *b) override {
    const double s = sizeof(T), nE = g.nEl, nP = g.nPl, nC = g.nC;
    if (g.ell) {
      // ELL: per observation the operator stream (K scalars) + its camera index; per point Hll^-1 (9) and,
      // for the factored stream, X (3); the camera term as in the tile layout.  Linearisation: camera index + two
      // copies of the observation (point- and camera-major) + point index in, stream out.
      const double kw = d.store == ST_FACT ? 8. : 24.;
      if (a) *a = nE * (kw * s + 4.) + nP * ((d.store == ST_FACT ? 12. : 9.) * s) + nC * 216. * s;
      if (b) *b = nE * (8. + (kw + 4.) * s) + nP * 15. * s + nC * 99. * s;
      return BA_OK;
    }
    const double kw = d.implicit ? 24. : 27.;
    if (a) *a = nE * (kw * s + 4.) + nP * (9. * s + 4.) + nC * 216. * s;
    if (b) *b = nE * (8. + (kw + 2.) * s) + nP * 15. * s + nC * 99. * s;
    return BA_OK;
  }
};

}  // namespace ba

// ==================================================================================
// C-ABI
struct ba_solver {
  std::unique_ptr<ba::Impl> impl;
};

static thread_local std::string g_create_error;

extern "C" {

int ba_abi_version(void) { return MEGBA_B200_ABI_VERSION; }

int ba_partition_points(int64_t nP, const int64_t *pt_degree, int32_t world_size, int64_t *pt_begin) {
  return ba::partition_points(nP, pt_degree, world_size, pt_begin);
}

int ba_layout(const ba_options *opt, int64_t nC, int64_t nP, int64_t nE, const int32_t *cam_idx, const int32_t *pt_idx,
              ba_layout_sizes *sizes, int32_t *tile, int64_t *edge_of_slot, int32_t *slot_cam, int32_t *slot_lp,
              int32_t *pt_slot, int32_t *chunk, int32_t *cm_slot) {
  if (!opt || (nE > 0 && (!cam_idx || !pt_idx))) return BA_ERR_ARG;
  if (opt->world_size < 1 || opt->rank < 0 || opt->rank >= opt->world_size) return BA_ERR_ARG;
  ba::HostGraph g;
  std::string err;
  const int rc = ba::build_host_graph(*opt, nC, nP, nE, cam_idx, pt_idx, g, err);
  if (rc != BA_OK) return rc;
  const size_t nSlots = (size_t)g.nSlabs * ba::TE;
  if (sizes) {
    sizes->n_tiles = g.nT; sizes->n_slots = (int64_t)nSlots; sizes->n_local_edges = g.nEl;
    sizes->n_local_points = g.nPl; sizes->n_chunks = g.nChunks; sizes->pt_begin = g.pt_begin;
  }
  for (int t = 0; tile && t < g.nT; ++t) {
    tile[4 * t] = g.tile[t].x; tile[4 * t + 1] = g.tile[t].y; tile[4 * t + 2] = g.tile[t].z; tile[4 * t + 3] = g.tile[t].w;
  }
  for (size_t i = 0; i < nSlots; ++i) {
    if (edge_of_slot) edge_of_slot[i] = g.edge_of_slot[i];
    if (slot_cam) slot_cam[i] = g.e_cam[i];
    if (slot_lp) slot_lp[i] = g.e_lp[i];
  }
  for (int p = 0; pt_slot && p < g.nPl; ++p) pt_slot[p] = g.pt_slot[p];
  for (int k = 0; chunk && k < g.nChunks; ++k) {
    chunk[4 * k] = g.chunk[k].x; chunk[4 * k + 1] = g.chunk[k].y; chunk[4 * k + 2] = g.chunk[k].z; chunk[4 * k + 3] = g.chunk[k].w;
  }
  for (int k = 0; cm_slot && k < g.nEl; ++k) cm_slot[k] = g.cm_slot[k];
  return BA_OK;
}

int ba_layout_ell(const ba_options *opt, int64_t nC, int64_t nP, int64_t nE, const int32_t *cam_idx, const int32_t *pt_idx,
                  ba_ell_sizes *sizes, int32_t *warp, int32_t *perm, int64_t *edge_of_slot, int32_t *slot_cam,
                  int32_t *chunk, int32_t *cm_slot) {
  if (!opt || (nE > 0 && (!cam_idx || !pt_idx))) return BA_ERR_ARG;
  if (opt->world_size < 1 || opt->rank < 0 || opt->rank >= opt->world_size) return BA_ERR_ARG;
  ba::HostGraph g;
  g.ell = true;
  std::string err;
  const int rc = ba::build_host_graph(*opt, nC, nP, nE, cam_idx, pt_idx, g, err);
  if (rc != BA_OK) return rc;
  const size_t nSlots = (size_t)g.nRows * 32;
  if (sizes) {
    sizes->n_warps = g.nW; sizes->n_rows = g.nRows; sizes->n_slots = (int64_t)nSlots; sizes->n_local_edges = g.nEl;
    sizes->n_local_points = g.nPl; sizes->n_chunks = g.nChunks; sizes->pt_begin = g.pt_begin; sizes->n_heavy = g.nHeavy;
  }
  for (int w = 0; warp && w < g.nW; ++w) {
    warp[4 * w] = g.wdesc[w].x; warp[4 * w + 1] = g.wdesc[w].y; warp[4 * w + 2] = g.wdesc[w].z; warp[4 * w + 3] = g.wdesc[w].w;
  }
  for (int p = 0; perm && p < g.nPl; ++p) perm[p] = g.perm[p];
  for (size_t i = 0; i < nSlots; ++i) {
    if (edge_of_slot) edge_of_slot[i] = g.edge_of_slot[i];
    if (slot_cam) slot_cam[i] = g.e_cam[i];
  }
  for (int k = 0; chunk && k < g.nChunks; ++k) {
    chunk[4 * k] = g.chunk[k].x; chunk[4 * k + 1] = g.chunk[k].y; chunk[4 * k + 2] = g.chunk[k].z; chunk[4 * k + 3] = g.chunk[k].w;
  }
  for (int k = 0; cm_slot && k < g.nEl; ++k) cm_slot[k] = g.cm_slot[k];
  return BA_OK;
}

int ba_create(const ba_options *opt, int64_t nC, int64_t nP, int64_t nE, const int32_t *cam_idx,
              const int32_t *pt_idx, const void *obs, ba_solver **out) {
  if (!opt || !out || (nE > 0 && (!cam_idx || !pt_idx || !obs))) return BA_ERR_ARG;
  *out = nullptr;
  if (opt->world_size < 1 || opt->rank < 0 || opt->rank >= opt->world_size) return BA_ERR_ARG;
  if (opt->dtype != BA_F64 && opt->dtype != BA_F32) return BA_ERR_ARG;
  if (opt->compute_kind != BA_EXPLICIT && opt->compute_kind != BA_IMPLICIT) return BA_ERR_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
    cudaGetLastError();
    g_create_error = "no CUDA device: megba_b200 has no CPU fallback";
    return BA_ERR_NO_DEVICE;
  }
  if (opt->device < 0 || opt->device >= ndev) return BA_ERR_ARG;
  std::unique_ptr<ba::Impl> impl;
  if (opt->dtype == BA_F64) impl.reset(new ba::ImplT<double>());
  else impl.reset(new ba::ImplT<float>());
  impl->opt = *opt;
  {
    int layout = 0, store = 0;
    int rcfg = ba::resolve_layout(*opt, &layout, &store);
    if (rcfg != BA_OK) { g_create_error = "unsupported layout / operator_store combination"; return rcfg; }
    impl->g.ell = layout == BA_LAYOUT_ELL;
    impl->store_cfg = store;
  }
  int rc = ba::build_host_graph(*opt, nC, nP, nE, cam_idx, pt_idx, impl->g, impl->err);
  if (rc != BA_OK) { g_create_error = impl->err; return rc; }
  rc = impl->init(obs);
  if (rc != BA_OK) { g_create_error = impl->err; return rc; }
  *out = new ba_solver{std::move(impl)};
  return BA_OK;
}

void ba_destroy(ba_solver *s) {
  if (!s) return;
  if (s->impl) cudaSetDevice(s->impl->opt.device);
  delete s;
}

const char *ba_last_error(const ba_solver *s) { return s ? s->impl->err.c_str() : g_create_error.c_str(); }

int ba_comm_unique_id(void *id128) {
  if (!id128) return BA_ERR_ARG;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  if (ncclGetUniqueId(&id) != ncclSuccess) return BA_ERR_NCCL;
  memcpy(id128, &id, sizeof id);
  return BA_OK;
}

int ba_comm_init(ba_solver *s, const void *id128) {
  if (!s || !id128) return BA_ERR_ARG;
  ba::Impl *m = s->impl.get();
  if (m->opt.world_size == 1) return BA_OK;
  if (cudaSetDevice(m->opt.device) != cudaSuccess) return m->fail(BA_ERR_CUDA, "cudaSetDevice");
  ncclUniqueId id;
  memcpy(&id, id128, sizeof id);
  ncclResult_t r = ncclCommInitRank(&m->comm, m->opt.world_size, id, m->opt.rank);
  if (r != ncclSuccess) return m->fail(BA_ERR_NCCL, "ncclCommInitRank: %s", ncclGetErrorString(r));
  m->comm_ready = true;
  return m->setup_p2p();
}

int ba_comm_init_local(ba_solver **solvers, int32_t n) {
  if (!solvers || n < 1 || n > 8) return BA_ERR_ARG;
  auto lc = std::make_shared<ba::LocalComm>();
  lc->n = n;
  for (int r = 0; r < n; ++r) {
    if (!solvers[r]) return BA_ERR_ARG;
    ba::Impl *m = solvers[r]->impl.get();
    if (m->opt.world_size != n || m->opt.rank != r || m->comm_ready) return BA_ERR_ARG;
  }
  if (n == 1) return BA_OK;
  // the peer mapping rendezvous on the communicator's barrier: one thread per rank for the set-up, like every later call
  std::vector<int> rcs(n, BA_OK);
  std::vector<std::thread> th;
  for (int r = 0; r < n; ++r) {
    solvers[r]->impl->local = lc;
    th.emplace_back([&, r] { rcs[r] = solvers[r]->impl->setup_p2p(); });
  }
  for (auto &t : th) t.join();
  for (int r = 0; r < n; ++r) if (rcs[r] != BA_OK) return rcs[r];
  return BA_OK;
}

#define BA_FWD(call) \
  if (!s) return BA_ERR_ARG; \
  return s->impl->call

int ba_set_params(ba_solver *s, const void *cam, const void *pt) { BA_FWD(set_params(cam, pt)); }
int ba_get_params(ba_solver *s, void *cam, void *pt) { BA_FWD(get_params(cam, pt)); }
int ba_set_information(ba_solver *s, const void *info) { BA_FWD(set_information(info)); }
int ba_set_fixed(ba_solver *s, const uint8_t *cam_fixed, const uint8_t *pt_fixed) { BA_FWD(set_fixed(cam_fixed, pt_fixed)); }
int ba_set_residual_program(ba_solver *s, int32_t n_instr, const uint8_t *code, int32_t n_const, const double *consts,
                            int32_t n_slots, const int32_t *out_slot) {
  BA_FWD(set_residual_program(n_instr, code, n_const, consts, n_slots, out_slot));
}
int ba_residual_program_mode(ba_solver *s) { return s && s->impl ? s->impl->residual_program_mode() : BA_ERR_ARG; }
int ba_residual_program_source(int32_t n_instr, const uint8_t *code, int32_t n_const, const double *consts, int32_t n_slots,
                               const int32_t *out_slot, int32_t dtype, char *src, int64_t src_cap, int64_t *src_len,
                               int64_t *cubin_bytes) {
  using namespace ba;
  if (n_instr <= 0 || !code || n_const < 0 || (n_const > 0 && !consts) || !out_slot || n_slots < 1 || n_slots > PROG_SLOTS ||
      (dtype != BA_F64 && dtype != BA_F32)) {
    g_create_error = "ba_residual_program_source: bad argument";
    return BA_ERR_ARG;
  }
  std::vector<ProgInstr> parsed;
  int out_def[2];
  const std::string why = prog_parse(n_instr, code, n_const, n_slots, out_slot, &parsed, out_def);
  if (!why.empty()) { g_create_error = "residual program: " + why; return BA_ERR_ARG; }
  const ProgDevice P = prog_lower(parsed, out_def, out_slot, true);
  const std::string text = jit_source(P, std::vector<double>(consts, consts + n_const), n_slots, out_slot, dtype == BA_F64);
  if (src_len) *src_len = (int64_t)text.size();
  if (src && src_cap > 0) {
    const size_t n = std::min<size_t>(text.size(), (size_t)src_cap - 1);
    memcpy(src, text.data(), n);
    src[n] = 0;
  }
  if (cubin_bytes) {
    std::vector<char> cubin;
    const std::string cw = jit_compile(text, &cubin);
    if (!cw.empty()) { g_create_error = cw; return BA_ERR_STATE; }
    *cubin_bytes = (int64_t)cubin.size();
  }
  return BA_OK;
}
int ba_linearize(ba_solver *s, double *cost) { BA_FWD(linearize(cost)); }
int ba_process_diag(ba_solver *s, double region, int recover) { BA_FWD(process_diag(region, recover)); }
int ba_solve(ba_solver *s, const ba_pcg_options *o, int32_t *iters, int32_t *done) { BA_FWD(solve(o, iters, done)); }
int ba_solve_async(ba_solver *s, const ba_pcg_options *o) { BA_FWD(solve_async(o)); }
int ba_solve_wait(ba_solver *s, int32_t *iters, int32_t *done) { BA_FWD(solve_wait(iters, done)); }
int ba_set_params_device(ba_solver *s, const void *cam, const void *pt, void *stream) { BA_FWD(set_params_device(cam, pt, stream)); }
int ba_get_params_device(ba_solver *s, void *cam, void *pt, void *stream) { BA_FWD(get_params_device(cam, pt, stream)); }
int ba_step_norms(ba_solver *s, double *a, double *b) { BA_FWD(step_norms(a, b)); }
int ba_eval_trial(ba_solver *s, double *a, double *b) { BA_FWD(eval_trial(a, b)); }
int ba_accept_step(ba_solver *s) { BA_FWD(accept_step()); }
int ba_reject_step(ba_solver *s) { BA_FWD(reject_step()); }
int ba_grad_inf_norm(ba_solver *s, double *g) { BA_FWD(grad_inf(g)); }
int ba_read(ba_solver *s, int32_t which, void *out) { BA_FWD(read(which, out)); }
int ba_write_dx(ba_solver *s, const void *dx) { BA_FWD(write_dx(dx)); }
int ba_eval_edges(ba_solver *s, void *r, void *J) { BA_FWD(eval_edges(r, J)); }
int ba_schur_apply(ba_solver *s, const void *p, void *q) { BA_FWD(schur_apply(p, q)); }
int ba_get_stats(ba_solver *s, ba_stats *out) { BA_FWD(get_stats(out)); }
int ba_reset_stats(ba_solver *s) { BA_FWD(reset_stats()); }
int ba_algorithmic_bytes(ba_solver *s, double *a, double *b) { BA_FWD(algorithmic_bytes(a, b)); }

int ba_set_timing(ba_solver *s, int32_t on) {
  if (!s) return BA_ERR_ARG;
  s->impl->timing = on != 0;
  s->impl->timing_stride = on > 1 ? on : 1;
  return BA_OK;
}

int ba_set_persistent(ba_solver *s, int32_t on) {
  if (!s) return BA_ERR_ARG;
  s->impl->persistent = on != 0;
  return BA_OK;
}

void *ba_stream(ba_solver *s) { return s ? (void *)s->impl->stream : nullptr; }

// LMAlgo::solveCUDA, src/algo/lm_algo.cu:139-223 (line numbers in comments below)
// Trust-region bounds of ba_lm_solve (the reference has none; see the reject branch).
static constexpr double kMinRegion = 1e-20, kMaxRegion = 1e20;

int ba_lm_solve(ba_solver *s, const ba_lm_options *lm, const ba_pcg_options *pcg, ba_lm_trace *trace,
                int32_t trace_cap, int32_t *n_trace, int32_t verbose) {
  if (!s || !lm || !pcg) return BA_ERR_ARG;
  ba::Impl *m = s->impl.get();
  const auto t0 = std::chrono::steady_clock::now();
  auto elapsed = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); };
  int nt = 0;
  auto push = [&](double cost, int acc, int it, double region, double dxn, double rho) {
    if (trace && nt < trace_cap) trace[nt] = ba_lm_trace{cost, acc, it, region, dxn, rho, elapsed()};
    ++nt;
  };
  const bool talk = verbose && m->opt.rank == 0;
  double F_new = 0., F = 0.;
  struct Fuse {  // scoped: the step-by-step entry points keep their one-call-one-result behaviour
    ba::Impl *m;
    explicit Fuse(ba::Impl *mm) : m(mm) { m->fuse_norms = m->fuse_ginf = true; }
    ~Fuse() { m->fuse_norms = m->fuse_ginf = m->norms_valid = m->ginf_valid = false; }
  } fuse(m);
  BA_TRY(m->linearize(&F_new));                                               // :145-148
  double region = lm->initial_region;                                         // lm_algo.cpp:15
  if (talk) printf("Start with error: %g, log error: %g, elapsed %lld ms\n", F_new / 2, std::log10(F_new / 2), (long long)elapsed());
  push(F_new, 1, 0, region, 0., 0.);
  bool stop = false, recover = false;
  int k = 0;
  double v = 2.;
  BA_TRY(m->backup_step());                                                  // :155
  while (!stop && k < lm->max_iter) {                                         // :163
    ++k;
    m->stats.lm_iters += 1;
    BA_TRY(m->process_diag(region, recover));                                 // :165
    int it = 0, done = 0;
    BA_TRY(m->solve(pcg, &it, &done));                                        // :166-169
    double dxn = 0., xn = 0.;
    BA_TRY(m->step_norms(&dxn, &xn));                                         // :172-175
    if (dxn <= lm->epsilon2 * (xn + lm->epsilon1)) {                          // :176-179
      push(F_new, -1, it, region, dxn, 0.);
      break;
    }
    double F_trial = 0., model_sq = 0.;
    BA_TRY(m->eval_trial(&F_trial, &model_sq));                               // :180-185
    const double rho_den = model_sq - F_new;                                  // :181-182
    F = F_new;
    F_new = F_trial;
    const double rho = -(F - F_new) / rho_den;                                // :186
    if (F > F_new) {                                                          // :187
      BA_TRY(m->accept_step());                                               // :192-194 (x += dx, backup)
      double F_chk = 0.;
      BA_TRY(m->linearize(&F_chk));                                           // :189
      if (talk) printf("Iter %d error: %g, log error: %g", k, F_new / 2, std::log10(F_new / 2));
      {
        // :197-198.  At a converged point rho_den can round to 0 (rho = +-inf / NaN); the reference would then set
        // region to 0 or NaN and damp with inf.  Treat a non-finite ratio as a fully successful step instead.
        const double f = std::max(1. / 3., 1. - std::pow(2. * rho - 1., 3));
        region /= std::isfinite(f) && f > 0. ? f : 1. / 3.;
        region = std::min(region, kMaxRegion);  // 1 + 1/region == 1 long before this in either precision
      }
      v = 2.;
      recover = false;
      double ginf = 0.;
      BA_TRY(m->grad_inf(&ginf));                                             // :203-204
      stop = ginf <= lm->epsilon1;                                            // :205
      push(F_new, 1, it, region, dxn, rho);
    } else {
      if (talk) printf("Iter %d failed", k);                                  // :207
      BA_TRY(m->reject_step());                                               // :208-209
      F_new = F;                                                              // :210
      region /= v;                                                            // :211
      v *= 2.;
      // After ~45 consecutive rejections (a converged problem driven on with epsilon1 = epsilon2 = 0) the reference's
      // region underflows to 0 and it damps with inf.  Keep both finite; the bounds cannot change an accepted step.
      region = std::max(region, kMinRegion);
      v = std::min(v, 1e300);
      recover = true;
      // H and g are untouched by our solve, so nothing else needs restoring
      m->linearized = true;
      push(F_new, 0, it, region, dxn, rho);
    }
    if (talk) printf(", elapsed %lld ms\n", (long long)elapsed());
  }
  if (talk) printf("Finished\n");
  if (n_trace) *n_trace = nt;
  return BA_OK;
}

}  // extern "C"
