// ba_kernels.cu -- sm_100a kernels of the LM / Schur-PCG hot path (fp64 and fp32).
//
// Reference functions replaced (see DESIGN.md for the full map):
//   k_lin_pm / k_lin_cm / k_lin_cam_reduce   <- AnalyticalDerivativesKernelGradKernel
//       (src/geo/analytical_derivatives.cu:162-285) fused with makeHSchur /
//       makeHSchurWithInfo / makeHppHllSchur (src/edge/build_linear_system.cu:88-239,
//       src/edge/build_implicit_linear_system.cu:65-111) and computeResidualNorm
//       (src/algo/lm_algo.cu:26-51).  No global atomics: the point side is reduced
//       inside a tile, the camera side by a second, camera-major pass.
//   k_prep_points / k_prep_cams   <- processDiag kernels
//       (src/linear_system/schur_LM_linear_system.cu:114-185) + cublasXmatinvBatched
//       (src/solver/schur_pcg_solver.cu:61-97) + w = Hll^-1 g_l (:467-476).
//   k_schur_edge   <- the two cusparseSpMV calls + oursGgemvBatched(Hll^-1) of one
//       operator application (schur_pcg_solver.cu:315-357), MakeV (:430-510) and
//       SolveW (:513-596); implicit operators implicitETMulx / implicitEMulx
//       (src/solver/implicit_schur_pcg_solver.cu:21-90).  W (or J) is read ONCE per pass;
//       modes FULL (S*p), INIT (MakeV fused with S*x0), BACKSUB, MAKEV.
//   k_pcg_cam_reg   <- all camera-space vector work of a PCG iteration
//       (schur_pcg_solver.cu:265-313,359-406): one 16-CTA cluster (DSMEM reductions) or, for
//       large camera counts, a cooperative grid; rows stay in registers; scalars stay on
//       the device (the reference syncs the host twice per iteration); with several GPUs the
//       allreduce of the camera vector is fused in over NVLink peer memory (no NCCL call).
//       k_pcg_cam is the global-memory fallback for sizes neither flavour covers.
//   k_trial   <- updateDeltaXTwoVertices + forward + JdxpF + reductions
//       (src/edge/update.cu:14-41, src/algo/lm_algo.cu:61-126).
// The two loop kernels chain through programmatic dependent launch (griddepcontrol).
#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <atomic>

#include "ba_geometry.cuh"
#include "ba_internal.h"
#include "ba_device_utils.cuh"

namespace cg = cooperative_groups;

namespace ba {

#ifndef BA_EARLY
#define BA_EARLY 1  // edge pass: slot streams requested before the tile descriptor has arrived
#endif
#ifndef BA_TPP
#define BA_TPP 1    // edge pass: up to 4 threads per point in the point phase
#endif
#ifndef BA_DEFAULT_VAR
#define BA_DEFAULT_VAR 5
#endif

// per-edge evaluation shared by the linearisation kernels
template <typename T>
__device__ __forceinline__ void load_cam(const T *__restrict__ cam, int c, T *out) {
#pragma unroll
  for (int i = 0; i < 9; ++i) out[i] = __ldg(cam + 9 * (size_t)c + i);
}

// L = Omega * J rows (JMulInfo, build_linear_system.cu:149-176); identity when info == nullptr
template <typename T>
__device__ __forceinline__ void left_factor(const T *J, const T *info4, T *L) {
#pragma unroll
  for (int y = 0; y < 12; ++y) {
    L[y] = J[y] * info4[0] + J[12 + y] * info4[2];
    L[12 + y] = J[y] * info4[1] + J[12 + y] * info4[3];
  }
}

// --------------------------------------------------------------------------------
// HP1-A: point-major linearisation.  One CTA per tile, one thread per slot.
template <typename T, bool IMPLICIT>
__global__ void __launch_bounds__(TE, 3) k_lin_pm(Dev<T> d) {
  constexpr int K = IMPLICIT ? 24 : 27;
  __shared__ T sh[12][TE];
  __shared__ double red[32];
  const int4 td = d.tile[blockIdx.x];
  const int slot0 = td.x, ne = td.y, pt0 = td.z, np = td.w;
  const int tid = threadIdx.x;
  double cost = 0.;
  T acc[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) acc[k] = T(0);
  const bool big = ne > TE;
  for (int base = 0; base < ne; base += TE) {
    const int i = base + tid;
    T hl[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) hl[k] = T(0);
    if (i < ne) {
      const int slot = slot0 + i;
      const int c = d.e_cam[slot];
      const int p = pt0 + (int)d.e_lp[slot];
      T cm[9];
      load_cam(d.cam, c, cm);
      const T X0 = d.pt[3 * (size_t)p], X1 = d.pt[3 * (size_t)p + 1], X2 = d.pt[3 * (size_t)p + 2];
      const typename Vec2<T>::type ob = reinterpret_cast<const typename Vec2<T>::type *>(d.e_obs)[slot];
      T J[24], r0, r1;
      const RotConst<T> rc = rot_const<T>(cm[0], cm[1], cm[2]);
      edge_residual_jacobian<T>(cm, rc, X0, X1, X2, ob.x, ob.y, r0, r1, J);
      cost += (double)(r0 * r0) + (double)(r1 * r1);
      T Lb[24];
      const T *L = J;
      if (d.e_info) {
        T om[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) om[k] = d.e_info[4 * (size_t)slot + k];
        left_factor<T>(J, om, Lb);
        L = Lb;
      }
      T *wj = d.WJ + wj_base<T>(slot, K);
      if (IMPLICIT) {
#pragma unroll
        for (int k = 0; k < 24; ++k) wj[(size_t)k * TE] = J[k];
      } else {
#pragma unroll
        for (int y = 0; y < 9; ++y)
#pragma unroll
          for (int q = 0; q < 3; ++q) wj[(size_t)(3 * y + q) * TE] = L[y] * J[9 + q] + L[12 + y] * J[21 + q];
      }
#pragma unroll
      for (int y = 0; y < 3; ++y) {
#pragma unroll
        for (int q = 0; q < 3; ++q) hl[3 * y + q] = L[9 + y] * J[9 + q] + L[21 + y] * J[21 + q];
        hl[9 + y] = -(L[9 + y] * r0 + L[21 + y] * r1);
      }
    }
    if (big) {
#pragma unroll
      for (int k = 0; k < 12; ++k) acc[k] += hl[k];
    } else {
#pragma unroll
      for (int k = 0; k < 12; ++k) sh[k][tid] = hl[k];
    }
  }
  if (!big) {
    __syncthreads();
    if (tid < np) {
      const int p = pt0 + tid;
      const int s0 = d.pt_slot[p] - slot0;
      const int s1 = (tid + 1 < np) ? d.pt_slot[p + 1] - slot0 : ne;
      T sum[12];
#pragma unroll
      for (int k = 0; k < 12; ++k) sum[k] = T(0);
      for (int s = s0; s < s1; ++s) {
#pragma unroll
        for (int k = 0; k < 12; ++k) sum[k] += sh[k][s];
      }
#pragma unroll
      for (int k = 0; k < 9; ++k) d.Hll[9 * (size_t)p + k] = sum[k];
#pragma unroll
      for (int k = 0; k < 3; ++k) d.gl[3 * (size_t)p + k] = sum[9 + k];
    }
  } else {
    T *smT = &sh[0][0];
#pragma unroll
    for (int k = 0; k < 12; ++k) {
      const T tot = block_sum<T>(acc[k], smT);
      if (tid == 0) {
        if (k < 9) d.Hll[9 * (size_t)pt0 + k] = tot;
        else d.gl[3 * (size_t)pt0 + (k - 9)] = tot;
      }
    }
  }
  const double ctot = block_sum<double>(cost, red);
  if (tid == 0) d.partials[blockIdx.x] = ctot;
}

// HP1-B: camera-major accumulation of Hpp (upper triangle) and g_c.  One CTA per
// (camera, chunk of <= CM_CHUNK edges); residual/Jacobian are recomputed from the
// L2-resident point array instead of being re-read from HBM.
template <typename T, bool PROG>
__global__ void __launch_bounds__(CM_THREADS, 3) k_lin_cm(Dev<T> d, int dispatch) {
  if (dispatch == 2 && d.scal[S_ANY_SMALL] == 0.) return;  // the factored flavour (ba_ell.cu) did the work
  constexpr int NW = CM_THREADS / 32;
  __shared__ T sh[NW][NTRI + 9];
  const int4 ch = d.chunk[blockIdx.x];
  const int c = ch.x, beg = ch.y, end = ch.z;
  T cm[9];
  load_cam(d.cam, c, cm);
  const RotConst<T> rc = rot_const<T>(cm[0], cm[1], cm[2]);
  T acc[NTRI + 9];
#pragma unroll
  for (int k = 0; k < NTRI + 9; ++k) acc[k] = T(0);
  for (int k = beg + (int)threadIdx.x; k < end; k += CM_THREADS) {
    T J[24], r0, r1;
    if (PROG && d.prog_jit) {  // the generated kernel left a camera-major copy of what this pass needs: 160 contiguous bytes
      using V = typename Vec2<T>::type;
      const V *src = reinterpret_cast<const V *>(d.cm_jr) + 10 * (size_t)k;
      T v[20];
#pragma unroll
      for (int j = 0; j < 10; ++j) {
        const V x = __ldg(src + j);
        v[2 * j] = x.x;
        v[2 * j + 1] = x.y;
      }
      const bool cfix = d.cam_fixed && d.cam_fixed[c];
#pragma unroll
      for (int q = 0; q < 9; ++q) { J[q] = cfix ? T(0) : v[q]; J[12 + q] = cfix ? T(0) : v[9 + q]; }
#pragma unroll
      for (int q = 9; q < 12; ++q) J[q] = J[12 + q] = T(0);  // the point columns play no part on the camera side
      r0 = v[18]; r1 = v[19];
    } else if (PROG) {  // interpreter: read what the point-side pass (k_lin_ell) evaluated, from its row-blocked stream
      using V = typename Vec2<T>::type;
      const int slot = d.cm_slot[k];
      const V *src = reinterpret_cast<const V *>(d.WJ) + (size_t)(slot >> 5) * (12 * 32) + (slot & 31);
#pragma unroll
      for (int j = 0; j < 12; ++j) {
        const V x = __ldg(src + j * 32);
        J[2 * j] = x.x;
        J[2 * j + 1] = x.y;
      }
      const V rv = __ldg(reinterpret_cast<const V *>(d.r_slots) + slot);
      r0 = rv.x; r1 = rv.y;
    } else {
      const int p = d.cm_pt[k];
      const T X0 = d.pt[3 * (size_t)p], X1 = d.pt[3 * (size_t)p + 1], X2 = d.pt[3 * (size_t)p + 2];
      const typename Vec2<T>::type ob = reinterpret_cast<const typename Vec2<T>::type *>(d.cm_obs)[k];
      edge_residual_jacobian<T>(cm, rc, X0, X1, X2, ob.x, ob.y, r0, r1, J);
    }
    T Lb[24];
    const T *L = J;
    if (d.cm_info) {
      T om[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) om[q] = d.cm_info[4 * (size_t)k + q];
      left_factor<T>(J, om, Lb);
      L = Lb;
    }
    int idx = 0;
#pragma unroll
    for (int y = 0; y < 9; ++y) {
#pragma unroll
      for (int i = y; i < 9; ++i) acc[idx++] += L[y] * J[i] + L[12 + y] * J[12 + i];
      acc[NTRI + y] -= L[y] * r0 + L[12 + y] * r1;
    }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NTRI + 9; ++k) {
    const T v = warp_sum<T>(acc[k]);
    if (lane == 0) sh[wid][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < NTRI + 9) {
    T v = T(0);
#pragma unroll
    for (int w = 0; w < NW; ++w) v += sh[w][threadIdx.x];
    d.cm_part[(size_t)blockIdx.x * (NTRI + 9) + threadIdx.x] = v;
  }
}

// Sum the chunk partials of one camera (fixed order) and expand the triangle.
template <typename T>
__global__ void __launch_bounds__(64) k_lin_cam_reduce(Dev<T> d, int dispatch) {
  if (dispatch == 2 && d.scal[S_ANY_SMALL] == 0.) return;
  const int c = blockIdx.x, k = threadIdx.x;
  if (k >= NTRI + 9) return;
  T v = T(0);
  for (int ch = d.cam_chunk_ptr[c]; ch < d.cam_chunk_ptr[c + 1]; ++ch) v += d.cm_part[(size_t)ch * (NTRI + 9) + k];
  const bool fixed = d.cam_fixed && d.cam_fixed[c];
  if (k >= NTRI) {
    d.gc[9 * (size_t)c + (k - NTRI)] = fixed ? T(0) : v;
  } else {
    int y = 0, rem = k;
    while (rem >= 9 - y) { rem -= 9 - y; ++y; }
    const int i = y + rem;
    if (fixed) v = y == i ? T(1) : T(0);  // a fixed camera keeps an identity block: its step stays zero
    d.Hpp[81 * (size_t)c + 9 * y + i] = v;
    d.Hpp[81 * (size_t)c + 9 * i + y] = v;
  }
}

// --------------------------------------------------------------------------------
// Solve preparation
template <typename T>
__global__ void __launch_bounds__(256) k_prep_points(Dev<T> d) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= d.nPl) return;
  const T a1 = T(d.damp) + T(1);  // (a + 1) * diag, schur_LM_linear_system.cu:123-124,149
  T h[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) h[k] = d.Hll[9 * (size_t)p + k];
  h[0] *= a1; h[4] *= a1; h[8] *= a1;
  const T c00 = h[4] * h[8] - h[5] * h[7], c01 = h[5] * h[6] - h[3] * h[8], c02 = h[3] * h[7] - h[4] * h[6];
  const T det = h[0] * c00 + h[1] * c01 + h[2] * c02;
  const T id = det != T(0) ? T(1) / det : T(0);
  T inv[9];
  inv[0] = c00 * id;
  inv[1] = (h[2] * h[7] - h[1] * h[8]) * id;
  inv[2] = (h[1] * h[5] - h[2] * h[4]) * id;
  inv[3] = c01 * id;
  inv[4] = (h[0] * h[8] - h[2] * h[6]) * id;
  inv[5] = (h[2] * h[3] - h[0] * h[5]) * id;
  inv[6] = c02 * id;
  inv[7] = (h[1] * h[6] - h[0] * h[7]) * id;
  inv[8] = (h[0] * h[4] - h[1] * h[3]) * id;
#pragma unroll
  for (int k = 0; k < 9; ++k) d.HllInv[9 * (size_t)p + k] = inv[k];
  const T g0 = d.gl[3 * (size_t)p], g1 = d.gl[3 * (size_t)p + 1], g2 = d.gl[3 * (size_t)p + 2];
#pragma unroll
  for (int y = 0; y < 3; ++y) d.w[3 * (size_t)p + y] = inv[3 * y] * g0 + inv[3 * y + 1] * g1 + inv[3 * y + 2] * g2;
}

template <typename T>
__global__ void __launch_bounds__(64) k_prep_cams(Dev<T> d) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d.nC) return;
  const T a1 = T(d.damp) + T(1);
  T A[81];
#pragma unroll
  for (int k = 0; k < 81; ++k) A[k] = d.Hpp[81 * (size_t)c + k];
#pragma unroll
  for (int k = 0; k < 9; ++k) A[10 * k] *= a1;
#pragma unroll
  for (int k = 0; k < 81; ++k) d.HppD[81 * (size_t)c + k] = A[k];
  // in-place Gauss-Jordan (SPD block: no pivoting needed)
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const T piv = A[10 * k] != T(0) ? T(1) / A[10 * k] : T(0);
#pragma unroll
    for (int j = 0; j < 9; ++j)
      if (j != k) A[9 * k + j] *= piv;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      if (i == k) continue;
      const T f = A[9 * i + k];
#pragma unroll
      for (int j = 0; j < 9; ++j)
        if (j != k) A[9 * i + j] -= f * A[9 * k + j];
      A[9 * i + k] = -f * piv;
    }
    A[10 * k] = piv;
  }
#pragma unroll
  for (int k = 0; k < 81; ++k) d.HppInv[81 * (size_t)c + k] = A[k];
}

// --------------------------------------------------------------------------------
// HP2 edge kernel: for every point j of the tile
//   t_j = sum_e W_e^T x[cam(e)]            (Hlp x, point-local segmented reduction)
//   u_j = Hll^-1 t_j | w_j | w_j - Hll^-1 t_j   (FULL | MAKEV | BACKSUB)
//   s[cam(e)] += W_e u_j                   (Hpl u, scatter by RED.ADD to global)
// W_e (or J_e) is loaded once and stays in registers between the two halves.
template <typename T, bool IMPLICIT>
__device__ __forceinline__ void edge_t(const T *wj, const T *x, T *t) {
  if (!IMPLICIT) {
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      T a = T(0);
#pragma unroll
      for (int y = 0; y < 9; ++y) a += wj[3 * y + q] * x[y];
      t[q] = a;
    }
  } else {
    T a0 = T(0), a1 = T(0);
#pragma unroll
    for (int y = 0; y < 9; ++y) { a0 += wj[y] * x[y]; a1 += wj[12 + y] * x[y]; }
#pragma unroll
    for (int q = 0; q < 3; ++q) t[q] = wj[9 + q] * a0 + wj[21 + q] * a1;
  }
}

template <typename T, bool IMPLICIT>
__device__ __forceinline__ void edge_scatter(const T *wj, const T *u, T *out9) {
  if (!IMPLICIT) {
#pragma unroll
    for (int y = 0; y < 9; ++y) out9[y] = wj[3 * y] * u[0] + wj[3 * y + 1] * u[1] + wj[3 * y + 2] * u[2];
  } else {
    const T b0 = wj[9] * u[0] + wj[10] * u[1] + wj[11] * u[2];
    const T b1 = wj[21] * u[0] + wj[22] * u[1] + wj[23] * u[2];
#pragma unroll
    for (int y = 0; y < 9; ++y) out9[y] = wj[y] * b0 + wj[12 + y] * b1;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) k_pad_cam(const T *__restrict__ src, T *dst, int nC) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 9 * nC) return;
  const int c = i / 9, k = i - 9 * c;
  dst[(size_t)CamPad<T>::N * c + k] = src[i];
}

// VAR bits.  The shipped kernel is VAR = 5; VAR = 0 (env BA_VAR=0) is kept as the measured baseline
// (profiles/r01_edge_variants_padded_gather.txt).
//   1 = camera values gathered with 16-byte vector loads from the padded copy (5 instead of 9 tag lookups per lane)
//   2 = W/J loads use the evict-first (streaming) policy (measured: no effect; not instantiated)
//   4 = the tile's Hll^-1 blocks / w vectors are prefetched into shared memory with cp.async at kernel start
template <typename T, int MODE, int VAR>
__host__ __device__ constexpr size_t schur_smem_bytes() {
  // t (3), u (3), staging / Hll^-1 prefetch (9) scalars per slot, the w prefetch (3) only in the modes that read w.
  // Keeping this small matters: what is not carved out as shared memory is L1, which the p gather lives in
  // (edge pass 272 -> 257 us when the prefetch buffer stopped having storage of its own).
  return sizeof(T) * (6 * TE + 9 * TE + (((VAR & 4) && MODE != SM_FULL) ? 3 * TE : 0)) + sizeof(int) * TE;
}

template <typename T, bool IMPLICIT, int MODE, int VAR>
__global__ void __launch_bounds__(TE, sizeof(T) == 4 ? 5 : 3) k_schur_edge(Dev<T> d, const T *__restrict__ xin, T *sout, T *zero_set, int check_active) {
  constexpr int K = IMPLICIT ? 24 : 27;
  constexpr bool VEC = (VAR & 1) != 0, LDCS = (VAR & 2) != 0, PREF = (VAR & 4) != 0;
  // Requesting the slot streams before the tile descriptor has arrived pays for the fp32 kernels (-3 %); the fp64
  // explicit kernel measured 257 us early / 256 us late with the 4-thread point phase off / on but 277 us with both,
  // the fp64 matrix-free kernel 245 us early / 242 us late (profiles/r01_edge_variants_early_tpp.txt): fp64 keeps the
  // late order.
  constexpr bool EARLY = BA_EARLY && sizeof(T) == 4;
  // Programmatic dependent launch: let the next kernel of the stream (the camera-space PCG kernel) be scheduled as
  // soon as every CTA of this grid has started; it waits for our completion itself before it reads our sums.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T *sh_t = reinterpret_cast<T *>(smem_raw);  // [3][TE]
  T *sh_u = sh_t + 3 * TE;                    // [3][TE]
  T *sv = sh_u + 3 * TE;                      // [TE*9] scatter staging (second half of the kernel)
  T *hv = sv;                                 // [TE*9] prefetched Hll^-1 (first half): same storage, a barrier apart
  T *hw = sv + 9 * TE;                        // [TE*3] prefetched w      (PREF, modes that read w)
  int *sc = reinterpret_cast<int *>(hw + ((PREF && MODE != SM_FULL) ? 3 * TE : 0));
  const int tid = threadIdx.x;
  // Tile index == slab index (the graph builder pads a point that spans several slabs with empty tiles), so this
  // thread's slot is known before the tile descriptor arrives: the descriptor, the slot's camera / point indices and
  // its W (or J) block are requested together instead of in three dependent round trips.  Slots past `ne` are padding
  // (zero blocks, camera 0): loading them is harmless and they take no part in the sums.
  const int4 td = __ldg(d.tile + blockIdx.x);
  const int slot_mine = (int)blockIdx.x * TE + tid;
  // camera (< 2^24, checked by ba_create) and local point index (< TE = 2^8) share one register until they are used
  int c_lp = 0;
  T wj[K];
  auto load_slot = [&]() {
    c_lp = __ldg(d.e_cam + slot_mine) | ((int)__ldg(d.e_lp + slot_mine) << 24);
    const T *src = d.WJ + (size_t)blockIdx.x * (K * TE) + tid;
#pragma unroll
    for (int k = 0; k < K; ++k) wj[k] = LDCS ? __ldcs(src + (size_t)k * TE) : __ldg(src + (size_t)k * TE);
  };
  if (EARLY) load_slot();
  const int slot0 = td.x, ne = td.y, pt0 = td.z, np = td.w;
  T *dst = MODE != SM_BACKSUB ? sout + (size_t)(blockIdx.x % NREP) * (9 * (size_t)d.nC) : nullptr;
  // Everything up to `dependent_ready` touches only data that is constant during a solve (tile descriptors, W/J,
  // Hll^-1, w): when this grid was launched programmatically it overlaps the tail of the preceding camera kernel.
  auto dependent_ready = [&]() -> bool {
    if (check_active) {
      asm volatile("griddepcontrol.wait;" ::: "memory");  // the preceding kernel (p, state, replica sets) is complete
      if (!d.st->active) return false;
    }
    if (MODE == SM_FULL && zero_set) {  // clear the replica set the NEXT edge pass will accumulate into
      const unsigned tot = (unsigned)NREP * 9u * (unsigned)d.nC;  // nC < 2^24 (ba_create): fits 32 bits
      for (unsigned i = blockIdx.x * TE + tid; i < tot; i += gridDim.x * TE) zero_set[i] = T(0);
    }
    return true;
  };
  if (ne <= TE) {
    const bool valid = tid < ne;
    if (PREF) {
      if (MODE != SM_MAKEV) {
        const T *src = d.HllInv + 9 * (size_t)pt0;
        for (int i = tid; i < 9 * np; i += TE) cp_async<sizeof(T)>(hv + i, src + i);
      }
      if (MODE != SM_FULL) {
        const T *src = d.w + 3 * (size_t)pt0;
        for (int i = tid; i < 3 * np; i += TE) cp_async<sizeof(T)>(hw + i, src + i);
      }
      cp_async_commit();
    }
    // point phase: 4 / 2 / 1 threads per point, so that the whole CTA is not parked at a barrier behind np <= TE/4
    // threads that walk their observation lists alone
    const int tsh = !BA_TPP ? 0 : np <= TE / 4 ? 2 : np <= TE / 2 ? 1 : 0;
    const int tpp = 1 << tsh, pi = tid >> tsh, sub = tid & (tpp - 1);
    const bool pact = pi < np;
    int s01 = 0;  // first slot | (end slot << 16) of this thread's point, relative to the tile (both <= TE)
    if (MODE != SM_MAKEV && pact) {
      const int s0 = __ldg(d.pt_slot + pt0 + pi) - slot0;
      const int s1 = (pi + 1 < np) ? __ldg(d.pt_slot + pt0 + pi + 1) - slot0 : ne;
      s01 = s0 | (s1 << 16);
    }
    if (!EARLY && valid) load_slot();
    if (!dependent_ready()) return;
    if (MODE != SM_MAKEV) {
      T t[3];
      if (valid) {
        __align__(16) T x[CamPad<T>::N];
        gather_cam<T, VEC>(xin, c_lp & 0xffffff, x);
        edge_t<T, IMPLICIT>(wj, x, t);
      }
      if (valid) { sh_t[tid] = t[0]; sh_t[TE + tid] = t[1]; sh_t[2 * TE + tid] = t[2]; }
    }
    if (PREF) cp_async_wait_all();
    if (MODE != SM_MAKEV || PREF) __syncthreads();
    {
      const int p = pt0 + pi;
      if (MODE != SM_MAKEV) {
        T t0 = T(0), t1 = T(0), t2 = T(0);
        if (pact)
          for (int s = (s01 & 0xffff) + sub, s1 = s01 >> 16; s < s1; s += tpp) { t0 += sh_t[s]; t1 += sh_t[TE + s]; t2 += sh_t[2 * TE + s]; }
        if (tsh >= 1) {  // CTA-uniform: every lane of every warp takes part in the exchange
          t0 += __shfl_xor_sync(0xffffffffu, t0, 1); t1 += __shfl_xor_sync(0xffffffffu, t1, 1); t2 += __shfl_xor_sync(0xffffffffu, t2, 1);
        }
        if (tsh >= 2) {
          t0 += __shfl_xor_sync(0xffffffffu, t0, 2); t1 += __shfl_xor_sync(0xffffffffu, t1, 2); t2 += __shfl_xor_sync(0xffffffffu, t2, 2);
        }
        if (pact)
          for (int y = sub; y < 3; y += tpp) {  // row y of u = Hll^-1 t
            const T h0 = PREF ? hv[9 * pi + 3 * y] : __ldg(d.HllInv + 9 * (size_t)p + 3 * y);
            const T h1 = PREF ? hv[9 * pi + 3 * y + 1] : __ldg(d.HllInv + 9 * (size_t)p + 3 * y + 1);
            const T h2 = PREF ? hv[9 * pi + 3 * y + 2] : __ldg(d.HllInv + 9 * (size_t)p + 3 * y + 2);
            T uy = h0 * t0 + h1 * t1 + h2 * t2;
            if (MODE == SM_INIT || MODE == SM_BACKSUB) {
              const T wy = PREF ? hw[3 * pi + y] : d.w[3 * (size_t)p + y];
              if (MODE == SM_INIT) uy -= wy;  // u = Hll^-1 Hlp x0 - w: one pass yields  Hpl Hll^-1 Hlp x0 - Hpl w
              else d.dxl[3 * (size_t)p + y] = wy - uy;
            }
            if (MODE != SM_BACKSUB) sh_u[y * TE + pi] = uy;
          }
      } else if (pact) {
        for (int y = sub; y < 3; y += tpp) sh_u[y * TE + pi] = PREF ? hw[3 * pi + y] : d.w[3 * (size_t)p + y];
      }
    }
    if (MODE != SM_BACKSUB) {
      __syncthreads();
      T o[9];
      if (valid) {
        const int lp = (unsigned)c_lp >> 24;
        const T u[3] = {sh_u[lp], sh_u[TE + lp], sh_u[2 * TE + lp]};
        edge_scatter<T, IMPLICIT>(wj, u, o);
      }
      warp_scatter9<T>(sv, sc, o, valid ? (c_lp & 0xffffff) : -1, dst);
    }
  } else {
    // a single point with more than TE observations: two passes over its slabs (rare; simple path)
    if (!dependent_ready()) return;
    T u[3];
    if (MODE != SM_MAKEV) {
      T t[3] = {T(0), T(0), T(0)};
      for (int i = tid; i < ne; i += TE) {
        const int slot = slot0 + i;
        const int c = d.e_cam[slot];
        const T *src = d.WJ + wj_base<T>(slot, K);
        T wj[K], te[3];
        __align__(16) T x[CamPad<T>::N];
#pragma unroll
        for (int k = 0; k < K; ++k) wj[k] = __ldg(src + (size_t)k * TE);
        gather_cam<T, false>(xin, c, x);
        edge_t<T, IMPLICIT>(wj, x, te);
        t[0] += te[0]; t[1] += te[1]; t[2] += te[2];
      }
      const T t0 = block_sum<T>(t[0], sh_t), t1 = block_sum<T>(t[1], sh_t), t2 = block_sum<T>(t[2], sh_t);
      T hi[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) hi[k] = __ldg(d.HllInv + 9 * (size_t)pt0 + k);
#pragma unroll
      for (int y = 0; y < 3; ++y) u[y] = hi[3 * y] * t0 + hi[3 * y + 1] * t1 + hi[3 * y + 2] * t2;
      if (MODE == SM_INIT) {
#pragma unroll
        for (int y = 0; y < 3; ++y) u[y] -= d.w[3 * (size_t)pt0 + y];
      }
      if (MODE == SM_BACKSUB) {
        if (tid == 0) {
#pragma unroll
          for (int y = 0; y < 3; ++y) d.dxl[3 * (size_t)pt0 + y] = d.w[3 * (size_t)pt0 + y] - u[y];
        }
      }
    } else {
#pragma unroll
      for (int y = 0; y < 3; ++y) u[y] = d.w[3 * (size_t)pt0 + y];
    }
    if (MODE != SM_BACKSUB) {
      for (int base = 0; base < ne; base += TE) {
        const int i = base + tid;
        int c = -1;
        T o[9];
        if (i < ne) {
          const int slot = slot0 + i;
          c = d.e_cam[slot];
          const T *src = d.WJ + wj_base<T>(slot, K);
          T wj[K];
#pragma unroll
          for (int k = 0; k < K; ++k) wj[k] = __ldg(src + (size_t)k * TE);
          edge_scatter<T, IMPLICIT>(wj, u, o);
        }
        warp_scatter9<T>(sv, sc, o, c, dst);
      }
    }
  }
}

// --------------------------------------------------------------------------------
// Camera-space PCG kernel.  One cluster of CL CTAs; cameras are split contiguously
// over the CTAs; the two dot products are reduced through distributed shared memory.
constexpr int PCG_CL = 8;
constexpr int PCG_THREADS = 512;

__device__ __forceinline__ double cluster_sum(double v, double *red_warp, double *slot, cg::cluster_group &cl) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = warp_sum<double>(v);
  __syncthreads();
  if (lane == 0) red_warp[wid] = v;
  __syncthreads();
  if (wid == 0) {
    double x = lane < nw ? red_warp[lane] : 0.;
    x = warp_sum<double>(x);
    if (lane == 0) *slot = x;
  }
  cl.sync();
  // all distributed-shared-memory loads are issued before the first add (a rolled loop pays one DSMEM round trip,
  // ~0.15 us, per CTA of the cluster); summation stays in CTA-rank order, identical on every CTA
  const int n = cl.num_blocks();
  double part[16];
#pragma unroll
  for (int r = 0; r < 16; ++r) part[r] = r < n ? *cl.map_shared_rank(slot, r) : 0.;
  double tot = 0.;
#pragma unroll
  for (int r = 0; r < 16; ++r) tot += part[r];
  return tot;
}

// INIT = 1: r = (g_c - sA) - (HppD x0 - sB); z = M^-1 r; rho; p = z      (:182-258,268-313 at n = 0)
// INIT = 0: bottom half of iteration n + top half of iteration n+1          (:359-406, :268-313)
template <typename T, int INIT>
__global__ void __launch_bounds__(PCG_THREADS) k_pcg_cam(Dev<T> d, T *sB, int max_iter, double tol, double refuse, int nrep) {
  cg::cluster_group cl = cg::this_cluster();
  __shared__ double red_warp[32];
  __shared__ double slots[2];
  // launched with the programmatic-serialisation attribute like the register-resident flavour: the preceding edge
  // pass (replica sums, p, state) must be complete before anything below is read
  asm volatile("griddepcontrol.wait;" ::: "memory");
  PcgState st;
  if (INIT) {
    st.rho = 0.; st.rho_prev = 0.; st.rho_min = INFINITY; st.alpha = 0.; st.pq = 0.;
    st.tol = tol; st.refuse = refuse; st.n = 0; st.max_iter = max_iter > 1 ? max_iter : 1; st.active = 1; st.done = 0; st.refused = 0;
  } else {
    st = *d.st;
    if (!st.active) return;  // uniform over the cluster: st is only written after the last cluster.sync
  }
  const int ncta = cl.num_blocks(), rank = cl.block_rank();
  const int per = (d.nC + ncta - 1) / ncta;
  const int c0 = min(d.nC, rank * per), c1 = min(d.nC, c0 + per);
  const int i0 = 9 * c0, i1 = 9 * c1;
  const int tid = threadIdx.x;
  const size_t rs = 9 * (size_t)d.nC;
  T *x = d.dxc;
  // fold the replicas of the scatter targets into replica 0 (own rows only)
  for (int i = i0 + tid; i < i1; i += PCG_THREADS) {
    T a = sB[i];
    for (int rep = 1; rep < nrep; ++rep) a += sB[rep * rs + i];
    sB[i] = a;
    if (INIT) {
      T b = d.sA[i];
      for (int rep = 1; rep < nrep; ++rep) b += d.sA[rep * rs + i];
      d.sA[i] = b;
    }
  }
  __syncthreads();
  if (INIT) {
    for (int i = i0 + tid; i < i1; i += PCG_THREADS) {
      const int c = i / 9, k = i - 9 * c;
      T a = T(0);
#pragma unroll
      for (int j = 0; j < 9; ++j) a += d.HppD[81 * (size_t)c + 9 * k + j] * x[9 * c + j];
      const T v = d.gc[i] - d.sA[i];            // v = g_c - Hpl w          (:478-493)
      d.r[i] = v - (a - sB[i]);               // r = v - (Hpp x0 - Hpl Hll^-1 Hlp x0)
    }
  } else {
    double acc = 0.;
    for (int i = i0 + tid; i < i1; i += PCG_THREADS) {
      const int c = i / 9, k = i - 9 * c;
      T a = T(0);
#pragma unroll
      for (int j = 0; j < 9; ++j) a += d.HppD[81 * (size_t)c + 9 * k + j] * d.p[9 * c + j];
      const T qi = a - sB[i];                 // q = Hpp p - Hpl Hll^-1 Hlp p   (:359-366)
      d.q[i] = qi;
      acc += (double)(d.p[i] * qi);
    }
    const double pq = cluster_sum(acc, red_warp, &slots[0], cl);
    const T alpha = pq != 0. ? (T)(st.rho / pq) : T(0);  // p.q = 0 only for a vanished search direction           // :385
    st.alpha = (double)alpha; st.pq = pq;
    for (int i = i0 + tid; i < i1; i += PCG_THREADS) {
      const T xi = x[i];
      d.xbk[i] = xi;                            // :389
      x[i] = xi + alpha * d.p[i];               // :391
      d.r[i] = d.r[i] + (-alpha) * d.q[i];      // :400
    }
  }
  __syncthreads();
  double acc = 0.;
  for (int i = i0 + tid; i < i1; i += PCG_THREADS) {
    const int c = i / 9, k = i - 9 * c;
    T a = T(0);
#pragma unroll
    for (int j = 0; j < 9; ++j) a += d.HppInv[81 * (size_t)c + 9 * k + j] * d.r[9 * c + j];
    d.z[i] = a;                                 // z = M^-1 r              (:273-275)
    acc += (double)(d.r[i] * a);
  }
  const double rho_new = (double)(T)cluster_sum(acc, red_warp, &slots[1], cl);  // rhoN is a T in the reference
  bool cont = true;
  if (!INIT) {
    st.n += 1;                                  // :405
    st.done = fabs(st.rho) < st.tol;            // :406 (rho of the iteration just finished)
    cont = !st.done && st.n < st.max_iter;      // :407
  }  // INIT: the reference's do-while always enters its first iteration
  if (cont) {
    if (rho_new > st.refuse * st.rho_min) {     // :288-295
      st.refused = 1;
      cont = false;
      if (!INIT)
        for (int i = i0 + tid; i < i1; i += PCG_THREADS) x[i] = d.xbk[i];
    } else {
      st.rho_min = fmin(st.rho_min, rho_new);   // :296
      if (!INIT) {
        const T beta = st.rho != 0. ? (T)(rho_new / st.rho) : T(0);   // :299
        for (int i = i0 + tid; i < i1; i += PCG_THREADS) {
          const T pv = d.z[i] + beta * d.p[i];
          d.p[i] = pv;
          d.xpad[(size_t)CamPad<T>::N * (i / 9) + i % 9] = pv;
        }
      } else {
        for (int i = i0 + tid; i < i1; i += PCG_THREADS) {  // :310
          d.p[i] = d.z[i];
          d.xpad[(size_t)CamPad<T>::N * (i / 9) + i % 9] = d.z[i];
        }
      }
      st.rho_prev = st.rho;
      st.rho = rho_new;
    }
  }
  st.active = cont ? 1 : 0;
  cl.sync();  // nobody leaves while peers may still read its slots; st is written after every read
  if (rank == 0 && tid == 0) *d.st = st;
}

// Register-resident variant (the one normally used): every thread owns RPT rows of the camera-space
// vectors and keeps their Hpp / Hpp^-1 rows, p, r, x in registers for the whole kernel, so the kernel
// is one L2 round trip + two cluster reductions instead of a chain of dependent global round trips.
// The set of CTAs that cooperate on one camera-space PCG step: a cluster or the whole (cooperative) grid.
template <bool GRID> struct Span;
template <> struct Span<false> {
  cg::cluster_group cl = cg::this_cluster();
  __device__ int num_blocks() const { return (int)cl.num_blocks(); }
  __device__ int block_rank() const { return (int)cl.block_rank(); }
  __device__ void sync() { cl.sync(); }
  __device__ double sum(double v, double *red_warp, double *slot, double *) { return cluster_sum(v, red_warp, slot, cl); }
  // peers may still read this CTA's slots: relaxed arrive/wait, so it does not wait for outstanding stores
  __device__ void lifetime_barrier() { asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory"); }
};
template <> struct Span<true> {
  cg::grid_group gr = cg::this_grid();
  __device__ int num_blocks() const { return (int)gridDim.x; }
  __device__ int block_rank() const { return (int)blockIdx.x; }
  __device__ void sync() { gr.sync(); }
  __device__ double sum(double v, double *red_warp, double *, double *partials) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    v = warp_sum<double>(v);
    __syncthreads();
    if (lane == 0) red_warp[wid] = v;
    __syncthreads();
    if (wid == 0) {
      double x = lane < nw ? red_warp[lane] : 0.;
      x = warp_sum<double>(x);
      if (lane == 0) partials[blockIdx.x] = x;
    }
    gr.sync();
    // fixed reduction tree (lane-strided partial sums, then a warp butterfly): the same on every CTA
    double mine = 0.;
    for (int r = lane; r < (int)gridDim.x; r += 32) mine += __ldcg(partials + r);
    return warp_sum<double>(mine);
  }
  __device__ void lifetime_barrier() { gr.sync(); }  // the state word is written only after every CTA has read it
};

template <typename T, int INIT, int RPT, bool PRE_INV, bool GRID>
__global__ void __launch_bounds__(PCG_THREADS, 1) k_pcg_cam_reg(Dev<T> d, const T *__restrict__ sB, int max_iter, double tol, double refuse, int nrep, PeerExchange px) {
  // GRID = false: one thread-block cluster (<= 16 CTAs), reductions through distributed shared memory.
  // GRID = true : cooperative launch over up to one CTA per SM (large camera counts), reductions through global
  //               partials + grid barrier.  Both sum the per-CTA partials in CTA order on every CTA (deterministic).
  Span<GRID> cl;
  extern __shared__ __align__(16) unsigned char smraw[];
  __shared__ double red_warp[32];
  __shared__ double slots[2];
  // the next edge pass may be scheduled now; it waits for our completion before it reads p / the state
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int ncta = cl.num_blocks(), rank = cl.block_rank();
  const int per = (d.nC + ncta - 1) / ncta;
  const int c0 = min(d.nC, rank * per), c1 = min(d.nC, c0 + per);
  const int i0 = 9 * c0, nrows = 9 * (c1 - c0);
  const int tid = threadIdx.x;
  const size_t rs = 9 * (size_t)d.nC;
  T *p_s = reinterpret_cast<T *>(smraw);
  T *r_s = p_s + RPT * PCG_THREADS;
  T *hi_s = r_s + RPT * PCG_THREADS;  // [RPT*PCG_THREADS][9], only when !PRE_INV
  T hd[RPT][9], hi[PRE_INV ? RPT : 1][9], pj[RPT], rj[RPT], xj[RPT], xb[RPT], sj[RPT], qj[RPT], zj[RPT];
  // Factored operator stream (ba_internal.h): the edge pass gathers T_c p and scatters Y with s = T_c^T Y, T_c block
  // diagonal (3x3 blocks Jr, R^T, E).  tin = row k of T_c, tout = column k of T_c (within the row's block).
  const bool fact = d.store == ST_FACT;
  T tin[RPT][3], tout[RPT][3];
  // ---- loads that do not depend on the preceding edge pass: the rows of the damped Hpp and of its inverse
#pragma unroll
  for (int j = 0; j < RPT; ++j) {
    const int l = tid + j * PCG_THREADS;
#pragma unroll
    for (int m = 0; m < 3; ++m) tin[j][m] = tout[j][m] = T(0);
    if (l < nrows) {
      const int i = i0 + l, c = i / 9, k = i - 9 * c;
      if (fact) {
        const int bk = k / 3, kk = k - 3 * bk;
        const T *tb = d.camT + (size_t)CAMT * c + 9 * bk;
#pragma unroll
        for (int m = 0; m < 3; ++m) {
          const T rowv = __ldg(tb + 3 * kk + m), colv = __ldg(tb + 3 * m + kk);
          tin[j][m] = bk == 1 ? colv : rowv;   // block 1 of T_c is R^T, stored as R
          tout[j][m] = bk == 1 ? rowv : colv;
        }
      }
#pragma unroll
      for (int m = 0; m < 9; ++m) hd[j][m] = __ldg(d.HppD + 81 * (size_t)c + 9 * k + m);
      if (PRE_INV) {
#pragma unroll
        for (int m = 0; m < 9; ++m) hi[j][m] = __ldg(d.HppInv + 81 * (size_t)c + 9 * k + m);
      } else {
#pragma unroll
        for (int m = 0; m < 9; ++m) cp_async<sizeof(T)>(hi_s + 9 * (size_t)l + m, d.HppInv + 81 * (size_t)c + 9 * k + m);
      }
    }
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");  // the edge pass (replica sums) and the previous state are complete
  PcgState st;
  if (INIT) {
    st.rho = 0.; st.rho_prev = 0.; st.rho_min = INFINITY; st.alpha = 0.; st.pq = 0.;
    st.tol = tol; st.refuse = refuse; st.n = 0; st.max_iter = max_iter > 1 ? max_iter : 1; st.active = 1; st.done = 0; st.refused = 0;
  } else {
    st = *d.st;
    if (!st.active) return;
  }
  // ---- the wave of loads that depend on it
#pragma unroll
  for (int j = 0; j < RPT; ++j) {
    const int l = tid + j * PCG_THREADS;
    pj[j] = rj[j] = xj[j] = xb[j] = sj[j] = qj[j] = zj[j] = T(0);
    if (l < nrows) {
      const int i = i0 + l;
      xj[j] = d.dxc[i];
      // all replica loads are issued before the first add (a rolled loop would serialise 16 L2 round trips)
      T sbv[NREP];
#pragma unroll
      for (int rep = 0; rep < NREP; ++rep) sbv[rep] = rep < nrep ? sB[rep * rs + i] : T(0);
      T sb = T(0);
#pragma unroll
      for (int rep = 0; rep < NREP; ++rep) sb += sbv[rep];
      sj[j] = sb;
      T sa = T(0);
      if (INIT) {
#pragma unroll
        for (int rep = 0; rep < NREP; ++rep) sbv[rep] = rep < nrep ? d.sA[rep * rs + i] : T(0);
#pragma unroll
        for (int rep = 0; rep < NREP; ++rep) sa += sbv[rep];
        pj[j] = xj[j];               // operand of Hpp x0
        rj[j] = d.gc[i] - sa;        // v = g_c - Hpl w                      (:478-493)
      } else {
        pj[j] = d.p[i];
        rj[j] = d.r[i];
      }
      p_s[l] = pj[j];
      if (px.ws > 1) {  // push this rank's partial sums into every rank's inbox (posted stores over NVLink);
                        // pull flavour: into the own slot only
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          if (r < px.ws && (!px.pull || r == px.rank)) {
            T *slot = reinterpret_cast<T *>(px.data[r]) + ((size_t)(px.parity * px.ws + px.rank) * 2) * rs;
            slot[i] = sb;
            if (INIT) slot[rs + i] = sa;
          }
        }
      }
    }
  }
  if (!PRE_INV) cp_async_commit();
  if (px.ws > 1) {
    // ---- fused allreduce over NVLink peer memory (replaces fold kernel + ncclAllReduce + a launch), push flavour:
    // every rank has written its partial camera vector into all inboxes; once those stores are ordered (system fence
    // by the storing threads, barrier, then a release store) it raises its flag IN EVERY PEER'S memory.  A rank then
    // polls only its own, local flags and sums its local inbox in rank order -- the same order on every rank, so the
    // result is bit-identical everywhere -- i.e. one NVLink traversal on the critical path instead of a remote flag
    // poll followed by remote loads.  Inboxes and flags are double-buffered by epoch parity: a rank can be at most one
    // exchange ahead of a peer (it needs the peer's flag of exchange k to finish k), so parity k is never rewritten
    // while someone still reads it.
    __threadfence_system();
    cl.sync();
    if (rank == 0 && tid < px.ws && (!px.pull || tid == px.rank))
      asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(px.flag[tid] + px.parity * 8 + px.rank), "l"(px.epoch) : "memory");
    if (tid < px.ws) {
      // push: sender tid's flag in MY memory; pull: sender tid's flag in ITS memory
      const unsigned long long *f = px.flag[px.pull ? tid : px.rank] + px.parity * 8 + tid;
      unsigned long long v = 0;
      const long long t0 = clock64();
      do {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
      } while (v < px.epoch && clock64() - t0 < 4000000000ll);  // ~2 s guard: never hang the GPU
      if (v < px.epoch) d.scal[S_XCHG_ERR] = 1.;  // surfaced to the host as BA_ERR_NCCL
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      const int l = tid + j * PCG_THREADS;
      if (l < nrows) {
        const int i = i0 + l;
        // all inbox loads are issued before the first add
        T pb[8], pa[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          pb[r] = pa[r] = T(0);
          if (r < px.ws) {
            const T *slot = reinterpret_cast<const T *>(px.data[px.pull ? r : px.rank]) + ((size_t)(px.parity * px.ws + r) * 2) * rs;
            pb[r] = __ldcv(slot + i);
            if (INIT) pa[r] = __ldcv(slot + rs + i);
          }
        }
        T sb = T(0), sa = T(0);
#pragma unroll
        for (int r = 0; r < 8; ++r) { sb += pb[r]; sa += pa[r]; }  // rank order: identical on every rank
        sj[j] = sb;
        if (INIT) rj[j] = d.gc[i] - sa;
      }
    }
  }
  if (fact) {  // s = T_c^T Y: the three Y rows of this row's block come through shared memory
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      const int l = tid + j * PCG_THREADS;
      if (l < nrows) r_s[l] = sj[j];
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      const int l = tid + j * PCG_THREADS;
      if (l < nrows) {
        const int lb = l - ((i0 + l) % 9) % 3;
        sj[j] = tout[j][0] * r_s[lb] + tout[j][1] * r_s[lb + 1] + tout[j][2] * r_s[lb + 2];
      }
    }
  }
  __syncthreads();
  double acc = 0.;
#pragma unroll
  for (int j = 0; j < RPT; ++j) {
    const int l = tid + j * PCG_THREADS;
    if (l < nrows) {
      const int lc = l - (i0 + l) % 9;
      T a = T(0);
#pragma unroll
      for (int m = 0; m < 9; ++m) a += hd[j][m] * p_s[lc + m];
      qj[j] = a - sj[j];                       // q = Hpp p - Hpl Hll^-1 Hlp p  (:359-366)
      if (INIT) rj[j] = rj[j] - qj[j];         // r = v - S x0                  (:182-258)
      else acc += (double)(pj[j] * qj[j]);
    }
  }
  if (!INIT) {
    const double pq = cl.sum(acc, red_warp, &slots[0], d.partials);
    const T alpha = pq != 0. ? (T)(st.rho / pq) : T(0);  // p.q = 0 only for a vanished search direction          // :385
    st.alpha = (double)alpha; st.pq = pq;
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      xb[j] = xj[j];                           // :389
      xj[j] = xj[j] + alpha * pj[j];           // :391
      rj[j] = rj[j] + (-alpha) * qj[j];        // :400
    }
  }
#pragma unroll
  for (int j = 0; j < RPT; ++j) {
    const int l = tid + j * PCG_THREADS;
    if (l < nrows) r_s[l] = rj[j];
  }
  if (!PRE_INV) cp_async_wait_all();  // each thread reads back only the rows it copied itself
  __syncthreads();
  acc = 0.;
#pragma unroll
  for (int j = 0; j < RPT; ++j) {
    const int l = tid + j * PCG_THREADS;
    if (l < nrows) {
      const int lc = l - (i0 + l) % 9;
      T a = T(0);
#pragma unroll
      for (int m = 0; m < 9; ++m) {
        const T h = PRE_INV ? hi[j][m] : hi_s[9 * (size_t)l + m];
        a += h * r_s[lc + m];
      }
      zj[j] = a;                               // z = M^-1 r                    (:273-275)
      acc += (double)(rj[j] * a);
    }
  }
  const double rho_new = (double)(T)cl.sum(acc, red_warp, &slots[1], d.partials + 512);
  bool cont = true, refused = false;
  if (!INIT) {
    st.n += 1;                                 // :405
    st.done = fabs(st.rho) < st.tol;           // :406
    cont = !st.done && st.n < st.max_iter;     // :407
  }  // INIT: the reference's do-while always enters its first iteration (max_iter < 1 behaves like 1)
  T beta = T(0);
  if (cont) {
    if (rho_new > st.refuse * st.rho_min) {    // :288-295
      st.refused = 1;
      refused = true;
      cont = false;
    } else {
      st.rho_min = fmin(st.rho_min, rho_new);  // :296
      if (!INIT) beta = st.rho != 0. ? (T)(rho_new / st.rho) : T(0); // :299
      st.rho_prev = st.rho;
      st.rho = rho_new;
    }
  }
  st.active = cont ? 1 : 0;
#pragma unroll
  for (int j = 0; j < RPT; ++j) {
    const int l = tid + j * PCG_THREADS;
    if (l < nrows) {
      const int i = i0 + l;
      d.r[i] = rj[j];
      if (!INIT) {
        d.dxc[i] = refused ? xb[j] : xj[j];
        d.xbk[i] = xb[j];
      }
      if (cont) {
        const T pv = INIT ? zj[j] : zj[j] + beta * pj[j];  // :310 / :299-306
        d.p[i] = pv;
        if (fact) p_s[l] = pv;
        else d.xpad[(size_t)CamPad<T>::N * (i / 9) + i % 9] = pv;
      }
    }
  }
  if (fact) {  // the edge pass gathers T_c p (cont and fact are uniform over the cluster / grid)
    __syncthreads();
    if (cont) {
#pragma unroll
      for (int j = 0; j < RPT; ++j) {
        const int l = tid + j * PCG_THREADS;
        if (l < nrows) {
          const int i = i0 + l, lb = l - (i % 9) % 3;
          d.xpad[(size_t)CamPad<T>::N * (i / 9) + i % 9] = tin[j][0] * p_s[lb] + tin[j][1] * p_s[lb + 1] + tin[j][2] * p_s[lb + 2];
        }
      }
    }
  }
  // lifetime barrier only (peers may still read this CTA's slots): relaxed, so it does not wait for the stores above
  cl.lifetime_barrier();
  if (rank == 0 && tid == 0) *d.st = st;
}

// s[0][i] += sum_{r>=1} s[r][i]; replicas r >= 1 are cleared.  Used before an NCCL allreduce.
template <typename T>
__global__ void __launch_bounds__(256) k_fold_replicas(T *s, int n, int nvec, size_t vec_stride) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int v = 0; v < nvec; ++v) {
    T *b = s + v * vec_stride;
    T a = b[i];
    for (int rep = 1; rep < NREP; ++rep) {
      a += b[(size_t)rep * n + i];
      b[(size_t)rep * n + i] = T(0);
    }
    b[i] = a;
  }
}

// --------------------------------------------------------------------------------
// LM scalars
template <typename T>
__global__ void __launch_bounds__(TE, 3) k_trial(Dev<T> d) {
  __shared__ double red[32];
  const int4 td = d.tile[blockIdx.x];
  const int slot0 = td.x, ne = td.y, pt0 = td.z;
  double a_model = 0., a_cost = 0.;
  for (int i = threadIdx.x; i < ne; i += TE) {
    const int slot = slot0 + i;
    const int c = d.e_cam[slot];
    const int p = pt0 + (int)d.e_lp[slot];
    T cm[9], dc[9];
    load_cam(d.cam, c, cm);
#pragma unroll
    for (int k = 0; k < 9; ++k) dc[k] = d.dxc[9 * (size_t)c + k];
    const T X0 = d.pt[3 * (size_t)p], X1 = d.pt[3 * (size_t)p + 1], X2 = d.pt[3 * (size_t)p + 2];
    const T D0 = d.dxl[3 * (size_t)p], D1 = d.dxl[3 * (size_t)p + 1], D2 = d.dxl[3 * (size_t)p + 2];
    const typename Vec2<T>::type ob = reinterpret_cast<const typename Vec2<T>::type *>(d.e_obs)[slot];
    T J[24], r0, r1;
    RotConst<T> rc = rot_const<T>(cm[0], cm[1], cm[2]);
    edge_residual_jacobian<T>(cm, rc, X0, X1, X2, ob.x, ob.y, r0, r1, J);
    T m0 = T(0), m1 = T(0);  // J dx  (JdxpF, lm_algo.cu:61-79)
#pragma unroll
    for (int k = 0; k < 9; ++k) { m0 += J[k] * dc[k]; m1 += J[12 + k] * dc[k]; }
    m0 += J[9] * D0 + J[10] * D1 + J[11] * D2;
    m1 += J[21] * D0 + J[22] * D1 + J[23] * D2;
    a_model += (double)((m0 + r0) * (m0 + r0)) + (double)((m1 + r1) * (m1 + r1));
#pragma unroll
    for (int k = 0; k < 9; ++k) cm[k] += dc[k];  // updateDeltaXTwoVertices, update.cu:14-41
    rc = rot_const<T>(cm[0], cm[1], cm[2]);
    T s0, s1;
    edge_residual<T>(cm, rc, X0 + D0, X1 + D1, X2 + D2, ob.x, ob.y, s0, s1);
    a_cost += (double)(s0 * s0) + (double)(s1 * s1);
  }
  const double tc = block_sum<double>(a_cost, red);
  const double tm = block_sum<double>(a_model, red);
  if (threadIdx.x == 0) {
    d.partials[2 * (size_t)blockIdx.x] = tc;
    d.partials[2 * (size_t)blockIdx.x + 1] = tm;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) k_norms(Dev<T> d) {
  __shared__ double red[32];
  double a[4] = {0., 0., 0., 0.};  // dx_pt^2, x_pt^2, dx_cam^2, x_cam^2
  const size_t stride = (size_t)gridDim.x * blockDim.x, g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (size_t i = g; i < 3 * (size_t)d.nPl; i += stride) {
    a[0] += (double)(d.dxl[i] * d.dxl[i]);
    a[1] += (double)(d.pt[i] * d.pt[i]);
  }
  for (size_t i = g; i < 9 * (size_t)d.nC; i += stride) {
    a[2] += (double)(d.dxc[i] * d.dxc[i]);
    a[3] += (double)(d.cam[i] * d.cam[i]);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const double t = block_sum<double>(a[j], red);
    if (threadIdx.x == 0) d.partials[4 * (size_t)blockIdx.x + j] = t;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) k_ginf(Dev<T> d) {
  __shared__ double red[32];
  double m = 0.;
  const size_t stride = (size_t)gridDim.x * blockDim.x, g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (size_t i = g; i < 3 * (size_t)d.nPl; i += stride) m = fmax(m, fabs((double)d.gl[i]));
  for (size_t i = g; i < 9 * (size_t)d.nC; i += stride) m = fmax(m, fabs((double)d.gc[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (threadIdx.x == 0) d.partials[blockIdx.x] = m;
  }
}

// x += dx (applyUpdate, schur_LM_linear_system.cu:212-218) and deltaX backup (:188-197)
template <typename T>
__global__ void __launch_bounds__(256) k_apply_step(Dev<T> d) {
  const size_t stride = (size_t)gridDim.x * blockDim.x, g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (size_t i = g; i < 3 * (size_t)d.nPl; i += stride) {
    const T v = d.dxl[i];
    d.pt[i] += v;
    d.dxl_bk[i] = v;
  }
  for (size_t i = g; i < 9 * (size_t)d.nC; i += stride) {
    const T v = d.dxc[i];
    d.cam[i] += v;
    d.dxc_bk[i] = v;
  }
}

// debug / parity: r[slot][2], J[slot][24]
template <typename T>
__global__ void __launch_bounds__(TE) k_eval_edges(Dev<T> d, T *r_out, T *J_out) {
  const int4 td = d.tile[blockIdx.x];
  const int slot0 = td.x, ne = td.y, pt0 = td.z;
  for (int i = threadIdx.x; i < ne; i += TE) {
    const int slot = slot0 + i;
    const int c = d.e_cam[slot];
    const int p = pt0 + (int)d.e_lp[slot];
    T cm[9];
    load_cam(d.cam, c, cm);
    const T X0 = d.pt[3 * (size_t)p], X1 = d.pt[3 * (size_t)p + 1], X2 = d.pt[3 * (size_t)p + 2];
    const T o0 = d.e_obs[2 * (size_t)slot], o1 = d.e_obs[2 * (size_t)slot + 1];
    T J[24], r0, r1;
    const RotConst<T> rc = rot_const<T>(cm[0], cm[1], cm[2]);
    edge_residual_jacobian<T>(cm, rc, X0, X1, X2, o0, o1, r0, r1, J);
    r_out[2 * (size_t)slot] = r0;
    r_out[2 * (size_t)slot + 1] = r1;
    if (J_out) {
#pragma unroll
      for (int k = 0; k < 24; ++k) J_out[24 * (size_t)slot + k] = J[k];
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) k_cam_matvec_sub(Dev<T> d, const T *p, const T *s, T *q) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 9 * d.nC) return;
  const int c = i / 9, k = i - 9 * c;
  T a = T(0);
#pragma unroll
  for (int j = 0; j < 9; ++j) a += d.HppD[81 * (size_t)c + 9 * k + j] * p[9 * c + j];
  T sv = s[i];
  if (d.store == ST_FACT) {  // s holds Y; apply T_c^T (see k_pcg_cam_reg)
    const int bk = k / 3, kk = k - 3 * bk;
    const T *tb = d.camT + (size_t)CAMT * c + 9 * bk;
    const T *y = s + 9 * (size_t)c + 3 * bk;
    sv = bk == 1 ? tb[3 * kk] * y[0] + tb[3 * kk + 1] * y[1] + tb[3 * kk + 2] * y[2]
                 : tb[kk] * y[0] + tb[3 + kk] * y[1] + tb[6 + kk] * y[2];
  }
  q[i] = a - sv;
}

// --------------------------------------------------------------------------------
// launchers
// camera side of the linearisation (layout-independent: it walks the camera-major view)
template <typename T>
cudaError_t launch_lin_cm(const Dev<T> &d, cudaStream_t st, int dispatch) {
  if (d.nChunks > 0) {
    if (d.prog_code) k_lin_cm<T, true><<<d.nChunks, CM_THREADS, 0, st>>>(d, dispatch);
    else k_lin_cm<T, false><<<d.nChunks, CM_THREADS, 0, st>>>(d, dispatch);
  }
  k_lin_cam_reduce<T><<<d.nC, 64, 0, st>>>(d, dispatch);
  return cudaGetLastError();
}

template <typename T>
cudaError_t launch_linearize(const Dev<T> &d, cudaStream_t st) {
  if (d.nT > 0) {
    if (d.implicit) k_lin_pm<T, true><<<d.nT, TE, 0, st>>>(d);
    else k_lin_pm<T, false><<<d.nT, TE, 0, st>>>(d);
  }
  launch_lin_cm<T>(d, st, 0);
  k_reduce<0><<<1, 1024, 0, st>>>(d.partials, d.nT, 1, d.scal + S_COST);
  return cudaGetLastError();
}
int launches_of_linearize() { return 4; }

template <typename T>
cudaError_t launch_prep(const Dev<T> &d, cudaStream_t st) {
  if (d.nPl > 0 && !d.ell) k_prep_points<T><<<(d.nPl + 255) / 256, 256, 0, st>>>(d);  // ELL: launch_prep_ell
  k_prep_cams<T><<<(d.nC + 63) / 64, 64, 0, st>>>(d);
  return cudaGetLastError();
}
int launches_of_prep() { return 2; }

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) { const char *e = getenv("BA_PDL"); v = e ? atoi(e) : 1; }
  return v != 0;
}

// Function attributes (dynamic shared memory size, non-portable cluster size) are per DEVICE; one process may drive
// several GPUs from several host threads, so the "already set" flags are kept per device ordinal.
static bool *device_flag(bool (&flags)[64]) {
  int dev = 0;
  cudaGetDevice(&dev);
  return &flags[dev & 63];
}

int schur_variant() {
  static const int v = [] {  // initialised once, thread-safe (the environment does not change under a running solver)
    const char *e = getenv("BA_VAR");
    return e ? (atoi(e) & 7) : BA_DEFAULT_VAR;
  }();
  return v;
}

template <typename T, bool IMPLICIT, int MODE, int VAR>
static void launch_schur_edge_v(const Dev<T> &d, const T *xin, T *sout, T *zero_set, int check_active, cudaStream_t st) {
  auto kern = k_schur_edge<T, IMPLICIT, MODE, VAR>;
  constexpr size_t smem = schur_smem_bytes<T, MODE, VAR>();
  static bool prepared_dev[64] = {};
  bool *prepared = device_flag(prepared_dev);
  if (!*prepared) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    *prepared = true;
  }
  if (check_active && pdl_enabled() && !d.no_pdl) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(d.nT);
    cfg.blockDim = dim3(TE);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kern, d, xin, sout, zero_set, check_active);
  } else {
    kern<<<d.nT, TE, smem, st>>>(d, xin, sout, zero_set, check_active);
  }
}

template <typename T, bool IMPLICIT, int MODE>
static void launch_schur_edge_m(const Dev<T> &d, const T *xin, T *sout, T *zero_set, int check_active, cudaStream_t st) {
  switch (schur_variant()) {
#define BA_V(n) case n: launch_schur_edge_v<T, IMPLICIT, MODE, n>(d, xin, sout, zero_set, check_active, st); break;
    BA_V(0) BA_V(5)
    default: launch_schur_edge_v<T, IMPLICIT, MODE, BA_DEFAULT_VAR>(d, xin, sout, zero_set, check_active, st); break;
#undef BA_V
  }
}

template <typename T, bool IMPLICIT>
static void launch_schur_edge_t(const Dev<T> &d, int mode, const T *xin, T *sout, T *zero_set, int check_active, cudaStream_t st) {
  switch (mode) {
    case SM_FULL: launch_schur_edge_m<T, IMPLICIT, SM_FULL>(d, xin, sout, zero_set, check_active, st); break;
    case SM_MAKEV: launch_schur_edge_m<T, IMPLICIT, SM_MAKEV>(d, xin, sout, zero_set, check_active, st); break;
    case SM_INIT: launch_schur_edge_m<T, IMPLICIT, SM_INIT>(d, xin, sout, zero_set, check_active, st); break;
    default: launch_schur_edge_m<T, IMPLICIT, SM_BACKSUB>(d, xin, sout, zero_set, check_active, st); break;
  }
}

template <typename T>
cudaError_t launch_schur_edge(const Dev<T> &d, int mode, const T *xin, T *sout, T *zero_set, int check_active, cudaStream_t st) {
  if (d.nT == 0) return cudaSuccess;
  if (d.implicit) launch_schur_edge_t<T, true>(d, mode, xin, sout, zero_set, check_active, st);
  else launch_schur_edge_t<T, false>(d, mode, xin, sout, zero_set, check_active, st);
  return cudaGetLastError();
}

template <typename K, typename T, typename P, typename... X>
static cudaError_t launch_cluster(K kernel, int cl, size_t smem, const Dev<T> &d, P sB, int max_iter, double tol,
                                  double refuse, int nrep, cudaStream_t st, X... extra) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cl);
  cfg.blockDim = dim3(PCG_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cl;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // may be scheduled during the edge pass's tail
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (pdl_enabled() && !d.no_pdl) ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kernel, d, sB, max_iter, tol, refuse, nrep, extra...);
}

template <typename T, int INIT, int RPT>
static cudaError_t launch_pcg_reg(const Dev<T> &d, T *sB, int cl, int max_iter, double tol, double refuse, int nrep,
                                  const PeerExchange &px, cudaStream_t st) {
  constexpr bool PRE = RPT <= 1 || sizeof(T) == 4;
  auto kern = k_pcg_cam_reg<T, INIT, RPT, PRE, false>;
  const size_t smem = (2 + (PRE ? 0 : 9)) * (size_t)RPT * PCG_THREADS * sizeof(T);
  static bool prepared_dev[64] = {};  // per instantiation and device
  bool *prepared = device_flag(prepared_dev);
  if (!*prepared) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    *prepared = true;
  }
  return launch_cluster(kern, cl, smem, d, (const T *)sB, max_iter, tol, refuse, nrep, st, px);
}

// 16-CTA clusters are "non-portable": ask the occupancy API once whether this device can host one.
int pcg_cluster_size() {
  // per device ordinal: the facade drives one host thread per GPU, and the answer is a property of the device
  static std::atomic<int> cached_dev[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  std::atomic<int> &cached = cached_dev[dev & 63];
  if (cached.load(std::memory_order_relaxed)) return cached.load(std::memory_order_relaxed);
  auto kern = k_pcg_cam_reg<double, 0, 2, false, false>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(16);
  cfg.blockDim = dim3(PCG_THREADS);
  cfg.dynamicSmemBytes = 11 * 2 * PCG_THREADS * sizeof(double);
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.dynamicSmemBytes);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 16;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  cudaError_t e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
  if (e != cudaSuccess) cudaGetLastError();
  const int v = (e == cudaSuccess && n >= 1) ? 16 : PCG_CL;
  cached.store(v, std::memory_order_relaxed);
  return v;
}

// Cooperative-grid flavour for camera counts that do not fit one cluster at <= 2 rows per thread.
static int pcg_grid_ctas(int nC) {
  static std::atomic<int> sms_dev[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  int sms = sms_dev[dev & 63].load(std::memory_order_relaxed);
  if (!sms) {
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    sms_dev[dev & 63].store(sms, std::memory_order_relaxed);
  }
  const int want = (9 * nC + 2 * PCG_THREADS - 1) / (2 * PCG_THREADS);
  return want < sms ? want : sms;
}

template <typename T, int INIT, int RPT>
static cudaError_t launch_pcg_grid(const Dev<T> &d, T *sB, int ncta, int max_iter, double tol, double refuse, int nrep,
                                   const PeerExchange &px, cudaStream_t st) {
  constexpr bool PRE = RPT <= 1;
  auto kern = k_pcg_cam_reg<T, INIT, RPT, PRE, true>;
  const size_t smem = (2 + (PRE ? 0 : 9)) * (size_t)RPT * PCG_THREADS * sizeof(T);
  static bool prepared_dev[64] = {};
  bool *prepared = device_flag(prepared_dev);
  if (!*prepared) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    *prepared = true;
  }
  Dev<T> dd = d;
  const T *sBc = sB;
  PeerExchange pxx = px;
  void *args[] = {&dd, &sBc, &max_iter, &tol, &refuse, &nrep, &pxx};
  return cudaLaunchCooperativeKernel((void *)kern, dim3(ncta), dim3(PCG_THREADS), args, smem, st);
}

template <typename T, int INIT>
static cudaError_t launch_pcg_cam(const Dev<T> &d, T *sB, int max_iter, double tol, double refuse, int nrep,
                                  const PeerExchange &px, cudaStream_t st) {
  static int clenv = -1;
  if (clenv < 0) { const char *e = getenv("BA_CL"); clenv = e ? atoi(e) : 0; }
  const int cl = clenv ? clenv : pcg_cluster_size();
  const int rows = 9 * ((d.nC + cl - 1) / cl);
  const int rpt = (rows + PCG_THREADS - 1) / PCG_THREADS;
  if (rpt <= 1) return launch_pcg_reg<T, INIT, 1>(d, sB, cl, max_iter, tol, refuse, nrep, px, st);
  if (rpt <= 2) return launch_pcg_reg<T, INIT, 2>(d, sB, cl, max_iter, tol, refuse, nrep, px, st);
  {
    const int ncta = pcg_grid_ctas(d.nC);
    const int grpt = (9 * ((d.nC + ncta - 1) / ncta) + PCG_THREADS - 1) / PCG_THREADS;
    if (grpt <= 1) return launch_pcg_grid<T, INIT, 1>(d, sB, ncta, max_iter, tol, refuse, nrep, px, st);
    if (grpt <= 2) return launch_pcg_grid<T, INIT, 2>(d, sB, ncta, max_iter, tol, refuse, nrep, px, st);
    if (grpt <= 4) return launch_pcg_grid<T, INIT, 4>(d, sB, ncta, max_iter, tol, refuse, nrep, px, st);
  }
  if (rpt <= 4) return launch_pcg_reg<T, INIT, 4>(d, sB, cl, max_iter, tol, refuse, nrep, px, st);
  if (rpt <= 8) return launch_pcg_reg<T, INIT, 8>(d, sB, cl, max_iter, tol, refuse, nrep, px, st);
  if (px.ws > 1) return cudaErrorNotSupported;  // callers fall back to fold + NCCL before reaching this
  return launch_cluster(k_pcg_cam<T, INIT>, PCG_CL, 0, d, sB, max_iter, tol, refuse, nrep, st);
}

template <typename T>
cudaError_t launch_pcg_init(const Dev<T> &d, T *sB, int max_iter, double tol, double refuse, int nrep,
                            const PeerExchange &px, cudaStream_t st) {
  return launch_pcg_cam<T, 1>(d, sB, max_iter, tol, refuse, nrep, px, st);
}
template <typename T>
cudaError_t launch_pcg_step(const Dev<T> &d, T *sB, int nrep, const PeerExchange &px, cudaStream_t st) {
  return launch_pcg_cam<T, 0>(d, sB, 0, 0., 0., nrep, px, st);
}
// true when the register-resident camera kernel (the only one with the fused peer exchange) covers this size
template <typename T>
bool pcg_reg_kernel_available(const Dev<T> &d) {
  const int cl = pcg_cluster_size();
  const int rows = 9 * ((d.nC + cl - 1) / cl);
  if ((rows + PCG_THREADS - 1) / PCG_THREADS <= 2) return true;
  const int ncta = pcg_grid_ctas(d.nC);
  return (9 * ((d.nC + ncta - 1) / ncta) + PCG_THREADS - 1) / PCG_THREADS <= 4;
}

template <typename T>
cudaError_t launch_pad_cam(const Dev<T> &d, const T *src, cudaStream_t st) {
  k_pad_cam<T><<<(9 * d.nC + 255) / 256, 256, 0, st>>>(src, d.xpad, d.nC);
  return cudaGetLastError();
}

template <typename T>
cudaError_t launch_fold_replicas(const Dev<T> &d, T *s, int nvec, cudaStream_t st) {
  const int n = 9 * d.nC;
  k_fold_replicas<T><<<(n + 255) / 256, 256, 0, st>>>(s, n, nvec, (size_t)NREP * n);
  return cudaGetLastError();
}

template <typename T>
cudaError_t launch_trial(const Dev<T> &d, cudaStream_t st) {
  if (d.nT > 0) k_trial<T><<<d.nT, TE, 0, st>>>(d);
  k_reduce<0><<<1, 1024, 0, st>>>(d.partials, d.nT, 2, d.scal + S_TRIAL_COST);
  return cudaGetLastError();
}

template <typename T>
cudaError_t launch_norms(const Dev<T> &d, cudaStream_t st) {
  k_norms<T><<<AUX_GRID, 256, 0, st>>>(d);
  k_reduce<0><<<1, 1024, 0, st>>>(d.partials, AUX_GRID, 4, d.scal + S_DX2_PT);
  return cudaGetLastError();
}

template <typename T>
cudaError_t launch_apply_step(const Dev<T> &d, cudaStream_t st) {
  k_apply_step<T><<<AUX_GRID, 256, 0, st>>>(d);
  return cudaGetLastError();
}

template <typename T>
cudaError_t launch_ginf(const Dev<T> &d, cudaStream_t st) {
  k_ginf<T><<<AUX_GRID, 256, 0, st>>>(d);
  k_reduce<1><<<1, 1024, 0, st>>>(d.partials, AUX_GRID, 1, d.scal + S_GINF);
  return cudaGetLastError();
}

template <typename T>
cudaError_t launch_eval_edges(const Dev<T> &d, T *r_slots, T *J_slots, cudaStream_t st) {
  if (d.nT > 0) k_eval_edges<T><<<d.nT, TE, 0, st>>>(d, r_slots, J_slots);
  return cudaGetLastError();
}

template <typename T>
cudaError_t launch_cam_matvec_sub(const Dev<T> &d, const T *p, const T *s, T *q, cudaStream_t st) {
  k_cam_matvec_sub<T><<<(9 * d.nC + 255) / 256, 256, 0, st>>>(d, p, s, q);
  return cudaGetLastError();
}

// dst[i] = op over r of src[r][i], in rank order (in-process communicator, ba_api.cu)
struct RankPtrs { const void *p[8]; };
template <typename U>
__global__ void __launch_bounds__(256) k_sum_ranks(U *dst, RankPtrs src, int nr, size_t n, int op) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    U a = reinterpret_cast<const volatile U *>(src.p[0])[i];
    for (int r = 1; r < nr; ++r) {
      const U b = reinterpret_cast<const volatile U *>(src.p[r])[i];
      a = op == 0 ? a + b : op == 1 ? (a > b ? a : b) : (a < b ? a : b);
    }
    dst[i] = a;
  }
}
template <typename U>
cudaError_t launch_sum_ranks(U *dst, const void *const *srcs, int nr, size_t n, int op, cudaStream_t st) {
  RankPtrs rp{};
  for (int r = 0; r < nr && r < 8; ++r) rp.p[r] = srcs[r];
  const int grid = (int)std::min<size_t>((n + 255) / 256, 592);
  if (n) k_sum_ranks<U><<<grid, 256, 0, st>>>(dst, rp, nr, n, op);
  return cudaGetLastError();
}
template cudaError_t launch_sum_ranks<double>(double *, const void *const *, int, size_t, int, cudaStream_t);
template cudaError_t launch_sum_ranks<float>(float *, const void *const *, int, size_t, int, cudaStream_t);

#define BA_INSTANTIATE(T)                                                                                   \
  template cudaError_t launch_linearize<T>(const Dev<T> &, cudaStream_t);                                   \
  template cudaError_t launch_lin_cm<T>(const Dev<T> &, cudaStream_t, int);                                      \
  template cudaError_t launch_schur_edge<T>(const Dev<T> &, int, const T *, T *, T *, int, cudaStream_t);   \
  template cudaError_t launch_prep<T>(const Dev<T> &, cudaStream_t);                                        \
  template cudaError_t launch_pcg_init<T>(const Dev<T> &, T *, int, double, double, int, const PeerExchange &, cudaStream_t); \
  template cudaError_t launch_pcg_step<T>(const Dev<T> &, T *, int, const PeerExchange &, cudaStream_t);    \
  template bool pcg_reg_kernel_available<T>(const Dev<T> &);                                                \
  template cudaError_t launch_fold_replicas<T>(const Dev<T> &, T *, int, cudaStream_t);                     \
  template cudaError_t launch_pad_cam<T>(const Dev<T> &, const T *, cudaStream_t);                          \
  template cudaError_t launch_trial<T>(const Dev<T> &, cudaStream_t);                                       \
  template cudaError_t launch_norms<T>(const Dev<T> &, cudaStream_t);                                       \
  template cudaError_t launch_apply_step<T>(const Dev<T> &, cudaStream_t);                                  \
  template cudaError_t launch_ginf<T>(const Dev<T> &, cudaStream_t);                                        \
  template cudaError_t launch_eval_edges<T>(const Dev<T> &, T *, T *, cudaStream_t);                        \
  template cudaError_t launch_cam_matvec_sub<T>(const Dev<T> &, const T *, const T *, T *, cudaStream_t);
BA_INSTANTIATE(double)
BA_INSTANTIATE(float)

}  // namespace ba
