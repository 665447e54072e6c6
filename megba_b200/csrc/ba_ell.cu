// ba_ell.cu -- sm_100a kernels of the ELL layout (thread per point, warp per slice; see ba_internal.h).
//
// Reference functions replaced (the same ones as ba_kernels.cu, different data layout):
//   k_cam_consts    per-camera rotation constants and the transform blocks of the factored operator; the
//                   reference recomputes sin/cos per OBSERVATION (src/geo/analytical_derivatives.cu:16-159).
//   k_lin_ell       <- AnalyticalDerivativesKernelGradKernel (analytical_derivatives.cu:162-285) fused with the
//                   point side of makeHSchur / makeHSchurWithInfo / makeHppHllSchur
//                   (src/edge/build_linear_system.cu:88-239, build_implicit_linear_system.cu:65-111) and
//                   computeResidualNorm (src/algo/lm_algo.cu:26-51): Hll, g_l and the operator stream; each
//                   thread owns one point, so Hll / g_l are plain register sums (no atomics, no shared memory).
//   k_schur_ell     <- one operator application of the PCG loop: the two cusparseSpMV calls + the batched
//                   Hll^-1 product (src/solver/schur_pcg_solver.cu:315-357), MakeV (:430-510), SolveW (:513-596)
//                   and the matrix-free implicitETMulx / implicitEMulx (implicit_schur_pcg_solver.cu:21-90).
//   k_trial_ell     <- updateDeltaXTwoVertices + forward + JdxpF + reductions (src/edge/update.cu:14-41,
//                   src/algo/lm_algo.cu:61-126).
//   k_prep_points_ell <- processDiag + the 3x3 cublasXmatinvBatched + w = Hll^-1 g_l
//                   (schur_LM_linear_system.cu:114-185, schur_pcg_solver.cu:61-97,467-476).
#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>

#include "ba_device_utils.cuh"
#include "ba_geometry.cuh"
#include "ba_internal.h"
#include "ba_program.cuh"

namespace cg = cooperative_groups;

namespace ba {

constexpr int ELL_THREADS = ELL_WARPS * 32;

struct WDesc { int row0, D, pt0, heavy, np, soff; };
__device__ __forceinline__ WDesc load_desc(const int4 *wdesc, int w) {
  const int4 v = __ldg(wdesc + w);
  return WDesc{v.x, v.y, v.z, v.w & 0xff, (v.w >> 8) & 0xff, (int)((unsigned)v.w >> 16)};
}

// one camera's constants (CAMC scalars, 16-byte aligned rows) with 16-byte loads
template <typename T>
__device__ __forceinline__ void load_camc(const T *__restrict__ camc, int c, T *cm, RotConst<T> &rc) {
  using V = typename CamPad<T>::vec;
  constexpr int NV = CAMC * sizeof(T) / 16;
  __align__(16) T v[CAMC];
  const V *src = reinterpret_cast<const V *>(camc + (size_t)CAMC * c);
#pragma unroll
  for (int k = 0; k < NV; ++k) reinterpret_cast<V *>(v)[k] = __ldg(src + k);
#pragma unroll
  for (int k = 0; k < 9; ++k) cm[k] = v[k];
  rc.c = v[9]; rc.a = v[10]; rc.b = v[11]; rc.ap = v[12]; rc.bp = v[13];
  rc.small = v[14] != T(0);
}

// the first 12 scalars of a row only (cam[9], c, a, b): all a kernel needs that runs when no camera is in the
// small-angle branch and evaluates no rotation derivative (6 instead of 8 sixteen-byte gathers per observation)
template <typename T>
__device__ __forceinline__ void load_camc_lite(const T *__restrict__ camc, int c, T *cm, RotConst<T> &rc) {
  using V = typename CamPad<T>::vec;
  constexpr int NV = 12 * sizeof(T) / 16;
  __align__(16) T v[12];
  const V *src = reinterpret_cast<const V *>(camc + (size_t)CAMC * c);
#pragma unroll
  for (int k = 0; k < NV; ++k) reinterpret_cast<V *>(v)[k] = __ldg(src + k);
#pragma unroll
  for (int k = 0; k < 9; ++k) cm[k] = v[k];
  rc.c = v[9]; rc.a = v[10]; rc.b = v[11]; rc.ap = T(0); rc.bp = T(0);
  rc.small = false;  // for a small-angle camera c = a = 1, b = 0 reproduce the first-order branch exactly
}

// the same through the texture path (the kernels that gather camera constants per observation are LSU-pipe bound too)
__device__ __forceinline__ void load_camc_lite_tex(cudaTextureObject_t tex, int c, double *cm, RotConst<double> &rc) {
  double v[12];
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const int4 t = tex1Dfetch<int4>(tex, 8 * c + k);
    v[2 * k] = __hiloint2double(t.y, t.x);
    v[2 * k + 1] = __hiloint2double(t.w, t.z);
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) cm[k] = v[k];
  rc.c = v[9]; rc.a = v[10]; rc.b = v[11]; rc.ap = 0.; rc.bp = 0.; rc.small = false;
}
__device__ __forceinline__ void load_camc_lite_tex(cudaTextureObject_t tex, int c, float *cm, RotConst<float> &rc) {
  float v[12];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float4 t = tex1Dfetch<float4>(tex, 4 * c + k);
    v[4 * k] = t.x; v[4 * k + 1] = t.y; v[4 * k + 2] = t.z; v[4 * k + 3] = t.w;
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) cm[k] = v[k];
  rc.c = v[9]; rc.a = v[10]; rc.b = v[11]; rc.ap = 0.f; rc.bp = 0.f; rc.small = false;
}
#ifndef BA_HP1_TEX
#define BA_HP1_TEX 1
#endif

template <typename T>
__global__ void __launch_bounds__(128) k_cam_consts(Dev<T> d) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d.nC) return;
  T cm[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) cm[k] = d.cam[9 * (size_t)c + k];
  const RotConst<T> rc = rot_const<T>(cm[0], cm[1], cm[2]);
  T *o = d.camc + (size_t)CAMC * c;
#pragma unroll
  for (int k = 0; k < 9; ++k) o[k] = cm[k];
  o[9] = rc.c; o[10] = rc.a; o[11] = rc.b; o[12] = rc.ap; o[13] = rc.bp;
  o[14] = rc.small ? T(1) : T(0);
  o[15] = T(0);
  T tr[CAMT];
  camera_transform<T>(cm, rc, tr);
  // a fixed camera has no columns: T_c = 0 removes it from both halves of the factored operator
  const bool fixed = d.cam_fixed && d.cam_fixed[c];
#pragma unroll
  for (int k = 0; k < CAMT; ++k) d.camT[(size_t)CAMT * c + k] = fixed ? T(0) : tr[k];
  if (rc.small) d.scal[S_ANY_SMALL] = 1.;  // the factored stream is not used while any camera is in the Taylor branch
}

// operator stream, row-blocked: pair j of slot (row, lane)
template <typename T, int K>
__device__ __forceinline__ void load_stream(const T *__restrict__ WJ, size_t row, int lane, T *v) {
  using V = typename Vec2<T>::type;
  const V *src = reinterpret_cast<const V *>(WJ) + row * (size_t)(K / 2 * 32) + lane;
#pragma unroll
  for (int j = 0; j < K / 2; ++j) {
    const V x = __ldg(src + j * 32);
    v[2 * j] = x.x;
    v[2 * j + 1] = x.y;
  }
}
template <typename T, int K>
__device__ __forceinline__ void store_stream(T *WJ, size_t row, int lane, const T *v) {
  using V = typename Vec2<T>::type;
  V *dst = reinterpret_cast<V *>(WJ) + row * (size_t)(K / 2 * 32) + lane;
#pragma unroll
  for (int j = 0; j < K / 2; ++j) {
    V x;
    x.x = v[2 * j];
    x.y = v[2 * j + 1];
    dst[j * 32] = x;
  }
}

template <typename T>
__device__ __forceinline__ void load_info(const T *__restrict__ r_info, size_t slot, T *om) {
  using V = typename Vec2<T>::type;
  const V a = __ldg(reinterpret_cast<const V *>(r_info) + 2 * slot), b = __ldg(reinterpret_cast<const V *>(r_info) + 2 * slot + 1);
  om[0] = a.x; om[1] = a.y; om[2] = b.x; om[3] = b.y;
}

// --------------------------------------------------------------------------------
// User-defined residuals: the camera-only part of the program, once per camera (ba_program_host.h).  trial: at
// cam + dx_c, values only (the trial cost needs no derivatives).
template <typename T, bool GRAD>
__global__ void __launch_bounds__(64) k_prog_cam(Dev<T> d, int trial) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d.nC) return;
  T in[PROG_NIN];
#pragma unroll
  for (int k = 0; k < 9; ++k) in[k] = d.cam[9 * (size_t)c + k] + (trial ? d.dxc[9 * (size_t)c + k] : T(0));
#pragma unroll
  for (int k = 9; k < PROG_NIN; ++k) in[k] = T(0);
  if (trial) {  // the trial camera, in 16-byte aligned rows for the generated prog_cost kernel
#pragma unroll
    for (int k = 0; k < 9; ++k) d.camc2[(size_t)CAMC * c + k] = in[k];
  }
  T *tab = (trial ? d.prog_tab2 : d.prog_tab) + (size_t)c * (PROG_FSTRIDE * d.prog_nf);
  prog_eval<T, GRAD>(d.prog_cam_code, d.prog_const, d.prog_cam_n, 0, 0, in, (T *)nullptr, (T *)nullptr, (const T *)nullptr, tab);
}
template <typename T>
void launch_prog_cam(const Dev<T> &d, int trial, cudaStream_t st) {
  if (!d.prog_code || d.nC == 0) return;
  if (trial) k_prog_cam<T, false><<<(d.nC + 63) / 64, 64, 0, st>>>(d, 1);
  else k_prog_cam<T, true><<<(d.nC + 63) / 64, 64, 0, st>>>(d, 0);
}

// --------------------------------------------------------------------------------
// HP1, point side.  DISPATCH: 0 run; 1 run unless a camera is in the small-angle branch; 2 run only then
// (the flag is produced on the device by k_cam_consts just before, so the host launches both flavours of a
// factored-store solver back to back and exactly one of them does the work).
template <typename T, int STORE, bool INFO, bool PROG = false>
__global__ void __launch_bounds__(ELL_THREADS) k_lin_ell(Dev<T> d, int dispatch) {
  constexpr int K = STORE == ST_FACT ? 8 : 24;
  if (dispatch) {
    const bool any_small = d.scal[S_ANY_SMALL] != 0.;
    if (any_small != (dispatch == 2)) return;
  }
  const int lane = threadIdx.x & 31, w = blockIdx.x * ELL_WARPS + (threadIdx.x >> 5);
  if (w >= d.nW) return;
  const WDesc wd = load_desc(d.wdesc, w);
  const int p = wd.heavy ? wd.pt0 : wd.pt0 + lane;
  const bool pvalid = wd.heavy || lane < wd.np;
  T X0 = T(0), X1 = T(0), X2 = T(0);
  if (pvalid) { X0 = d.pt[3 * (size_t)p]; X1 = d.pt[3 * (size_t)p + 1]; X2 = d.pt[3 * (size_t)p + 2]; }
  const bool pfixed = pvalid && d.pt_fixed && d.pt_fixed[p];
  T hl[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) hl[k] = T(0);
  double cost = 0.;
  for (int r = 0; r < wd.D; ++r) {
    const size_t row = (size_t)wd.row0 + r, slot = row * 32 + lane;
    const int c = __ldg(d.r_cam + slot);
    if (c < 0) continue;
    T cm[9];
    RotConst<T> rc;
    if (STORE == ST_FACT) {  // this flavour only runs without small-angle cameras
      if (BA_HP1_TEX && d.camc_tex) load_camc_lite_tex(d.camc_tex, c, cm, rc);
      else load_camc_lite<T>(d.camc, c, cm, rc);
    }
    else load_camc<T>(d.camc, c, cm, rc);
    const typename Vec2<T>::type ob = __ldg(reinterpret_cast<const typename Vec2<T>::type *>(d.r_obs) + slot);
    T r0, r1, Jp[6];
    if (STORE == ST_FACT) {
      T sv[8];
      edge_point_jacobian<T>(cm, rc, X0, X1, X2, ob.x, ob.y, r0, r1, sv, sv[6], sv[7]);
      if (pfixed) {  // a fixed point has no columns
#pragma unroll
        for (int k = 0; k < 6; ++k) sv[k] = T(0);
      }
      store_stream<T, K>(d.WJ, row, lane, sv);
#pragma unroll
      for (int k = 0; k < 6; ++k) Jp[k] = sv[k];
    } else {
      T J[24];
      if (PROG) {  // interpreter flavour (the generated prog_fill kernel replaces this whole kernel, ba_program_jit.h)
        prog_edge<T, true>(d, c, d.prog_tab, cm, X0, X1, X2, ob.x, ob.y, r0, r1, J);
        typename Vec2<T>::type rv;
        rv.x = r0; rv.y = r1;
        reinterpret_cast<typename Vec2<T>::type *>(d.r_slots)[slot] = rv;
      } else edge_residual_jacobian<T>(cm, rc, X0, X1, X2, ob.x, ob.y, r0, r1, J);
      if (pfixed) {
#pragma unroll
        for (int q = 0; q < 3; ++q) J[9 + q] = J[21 + q] = T(0);
      }
      if (d.cam_fixed && d.cam_fixed[c]) {
#pragma unroll
        for (int q = 0; q < 9; ++q) J[q] = J[12 + q] = T(0);
      }
      store_stream<T, K>(d.WJ, row, lane, J);
#pragma unroll
      for (int q = 0; q < 3; ++q) { Jp[q] = J[9 + q]; Jp[3 + q] = J[21 + q]; }
    }
    cost += (double)(r0 * r0) + (double)(r1 * r1);
    T L[6];  // Omega-weighted rows of Jp (JMulInfo, build_linear_system.cu:149-176); identity without information
    if (INFO) {
      T om[4];
      load_info<T>(d.r_info, slot, om);
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        L[q] = Jp[q] * om[0] + Jp[3 + q] * om[2];
        L[3 + q] = Jp[q] * om[1] + Jp[3 + q] * om[3];
      }
    } else {
#pragma unroll
      for (int q = 0; q < 6; ++q) L[q] = Jp[q];
    }
#pragma unroll
    for (int y = 0; y < 3; ++y) {
#pragma unroll
      for (int q = 0; q < 3; ++q) hl[3 * y + q] += L[y] * Jp[q] + L[3 + y] * Jp[3 + q];
      hl[9 + y] -= L[y] * r0 + L[3 + y] * r1;
    }
  }
  if (wd.heavy) {
#pragma unroll
    for (int k = 0; k < 12; ++k) hl[k] = warp_sum<T>(hl[k]);
  }
  if (pvalid && (!wd.heavy || lane == 0)) {
    if (pfixed) {  // identity block, zero gradient: dx_l = 0
#pragma unroll
      for (int k = 0; k < 9; ++k) hl[k] = (k % 4 == 0) ? T(1) : T(0);
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) d.Hll[9 * (size_t)p + k] = hl[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) d.gl[3 * (size_t)p + k] = hl[9 + k];
  }
  cost = warp_sum<double>(cost);
  if (lane == 0) d.partials[w] = cost;
}

// --------------------------------------------------------------------------------
// Solve preparation, point side: damped 3x3 inverse, w = Hll^-1 g_l, and the plane copies the edge kernel reads.
template <typename T>
__global__ void __launch_bounds__(256) k_prep_points_ell(Dev<T> d) {
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
  const size_t np = (size_t)d.nPlPad;
#pragma unroll
  for (int k = 0; k < 9; ++k) d.HllInvP[k * np + p] = inv[k];
  const T g0 = d.gl[3 * (size_t)p], g1 = d.gl[3 * (size_t)p + 1], g2 = d.gl[3 * (size_t)p + 2];
#pragma unroll
  for (int y = 0; y < 3; ++y) {
    d.wP[y * np + p] = inv[3 * y] * g0 + inv[3 * y + 1] * g1 + inv[3 * y + 2] * g2;
    d.XP[y * np + p] = d.pt[3 * (size_t)p + y];
  }
}

// --------------------------------------------------------------------------------
// Per-observation halves of the operator.
//   first half:  t += W^T x   second half: o = W u   (W = Jc^T Omega Jp, never formed)
// ST_FACT: s = {Jp[6], B[2]}, g = T_c x = (psi, xt, e); the second half produces o = (X x v, v, (B.z)(1,l2,l2^2)),
// to which the camera kernel applies T_c^T.
template <typename T, bool INFO>
__device__ __forceinline__ void fact_first(const T *s, const T *g, T X0, T X1, T X2, const T *om, T &t0, T &t1, T &t2) {
  const T q0 = g[1] * X2 - g[2] * X1 + g[3];
  const T q1 = g[2] * X0 - g[0] * X2 + g[4];
  const T q2 = g[0] * X1 - g[1] * X0 + g[5];
  const T l2 = s[6] * s[6] + s[7] * s[7];
  const T poly = g[6] + l2 * (g[7] + l2 * g[8]);
  T y0 = s[0] * q0 + s[1] * q1 + s[2] * q2 + s[6] * poly;
  T y1 = s[3] * q0 + s[4] * q1 + s[5] * q2 + s[7] * poly;
  if (INFO) {  // t = Jp^T Omega^T (Jc x)
    const T a = om[0] * y0 + om[2] * y1, b = om[1] * y0 + om[3] * y1;
    y0 = a; y1 = b;
  }
  t0 += s[0] * y0 + s[3] * y1;
  t1 += s[1] * y0 + s[4] * y1;
  t2 += s[2] * y0 + s[5] * y1;
}
template <typename T, bool INFO>
__device__ __forceinline__ void fact_second(const T *s, T u0, T u1, T u2, T X0, T X1, T X2, const T *om, T *o) {
  T z0 = s[0] * u0 + s[1] * u1 + s[2] * u2, z1 = s[3] * u0 + s[4] * u1 + s[5] * u2;
  if (INFO) {  // s = Jc^T Omega (Jp u)
    const T a = om[0] * z0 + om[1] * z1, b = om[2] * z0 + om[3] * z1;
    z0 = a; z1 = b;
  }
  const T v0 = s[0] * z0 + s[3] * z1, v1 = s[1] * z0 + s[4] * z1, v2 = s[2] * z0 + s[5] * z1;
  o[0] = X1 * v2 - X2 * v1;
  o[1] = X2 * v0 - X0 * v2;
  o[2] = X0 * v1 - X1 * v0;
  o[3] = v0; o[4] = v1; o[5] = v2;
  const T l2 = s[6] * s[6] + s[7] * s[7];
  const T bz = s[6] * z0 + s[7] * z1;
  o[6] = bz; o[7] = bz * l2; o[8] = bz * l2 * l2;
}
// ST_J: s = J[2][12], g = x (the plain camera vector)
template <typename T, bool INFO>
__device__ __forceinline__ void jac_first(const T *J, const T *x, const T *om, T &t0, T &t1, T &t2) {
  T a0 = T(0), a1 = T(0);
#pragma unroll
  for (int y = 0; y < 9; ++y) { a0 += J[y] * x[y]; a1 += J[12 + y] * x[y]; }
  if (INFO) {
    const T a = om[0] * a0 + om[2] * a1, b = om[1] * a0 + om[3] * a1;
    a0 = a; a1 = b;
  }
  t0 += J[9] * a0 + J[21] * a1;
  t1 += J[10] * a0 + J[22] * a1;
  t2 += J[11] * a0 + J[23] * a1;
}
template <typename T, bool INFO>
__device__ __forceinline__ void jac_second(const T *J, T u0, T u1, T u2, const T *om, T *o) {
  T b0 = J[9] * u0 + J[10] * u1 + J[11] * u2, b1 = J[21] * u0 + J[22] * u1 + J[23] * u2;
  if (INFO) {
    const T a = om[0] * b0 + om[1] * b1, b = om[2] * b0 + om[3] * b1;
    b0 = a; b1 = b;
  }
#pragma unroll
  for (int y = 0; y < 9; ++y) o[y] = J[y] * b0 + J[12 + y] * b1;
}

template <typename T>
__host__ __device__ constexpr size_t schur_ell_smem(int warps = ELL_WARPS) {
  return (size_t)warps * 32 * 9 * sizeof(T) + (size_t)warps * 32 * sizeof(int);
}
// SPLIT warps share one slice (rows r = part, part + SPLIT, ...): used when a shard has too few slices to fill the
// GPU (k_schur_ell below); the CTA then holds ell_cta_warps(SPLIT) / SPLIT slices
__host__ __device__ constexpr int ell_cta_warps(int split) { return split > ELL_WARPS ? split : ELL_WARPS; }

// HP2 edge kernel of the ELL layout.  One warp per slice (lane = point) or per heavy point (lane = observation).
//   phase 1: t_j = sum_e W_e^T x[cam(e)]     thread-local (slice) or one warp sum (heavy point)
//   u_j = Hll^-1 t_j | Hll^-1 t_j - w_j | (BACKSUB: dx_l = w_j - Hll^-1 t_j)
//   phase 2: s[cam(e)] += W_e u_j            RED.ADD, staged per warp so that one RED covers consecutive components
// RC rows of the factored stream are kept in registers between the phases; further rows (and the wider Jacobian
// stream) are read again in phase 2 (an L2 hit: the slice was streamed microseconds earlier).
// Resident warps per SM the register allocation aims at: 24 / 16 / 12 / 8 for RC = 0 / 2 / 4 / 6 rows of the factored
// stream kept in registers between the phases.  Measured on Venice fp64 (profiles/r02_edge_variants.txt): more resident
// warps beat fewer L2 re-reads (RC 0: 168.9 us with 4-warp CTAs and the Hll^-1 load issued after phase 1; RC 2: 170.8;
// RC 4: 185.3; RC 6: 207.4 us), keeping the stream out of L1 (ld.global.nc.L1::no_allocate) costs 6-10 us.
#ifndef BA_ELL_LATE_HI
#define BA_ELL_LATE_HI 1
#endif
#ifndef BA_ELL_U1
#define BA_ELL_U1 2
#endif
#ifndef BA_ELL_U2
#define BA_ELL_U2 1
#endif
constexpr int ELL_U1 = BA_ELL_U1, ELL_U2 = BA_ELL_U2;  // rows of phase 1 / phase 2 in flight per thread (unroll factors)
constexpr int ell_min_blocks(int rc) { return (rc == 0 ? 24 : rc <= 2 ? 16 : rc <= 4 ? 12 : 8) / ELL_WARPS > 0 ? (rc == 0 ? 24 : rc <= 2 ? 16 : rc <= 4 ? 12 : 8) / ELL_WARPS : 1; }

// The gather of the camera vector goes through the TEXTURE path (16-byte texels of xpad): the S*p pass is bound by the
// LSU data pipe of the SM (86 % busy, profiles/r02_ncu_final_summary.txt), and the texture pipe is a second way into the
// same L1 -- measured 157.4 us against 166.8 us with LSU gathers.  Moving the stream as well does not pay (phase-2
// re-read through TEX: 163.8 us, both stream passes: 173.0 us; profiles/r02_edge_variants.txt).  BA_ELL_TEX=0: LSU only.
#ifndef BA_ELL_TEX
#define BA_ELL_TEX 1
#endif
__device__ __forceinline__ void gather_cam_tex(cudaTextureObject_t tex, int c, double *x) {
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const int4 v = tex1Dfetch<int4>(tex, 5 * c + k);
    x[2 * k] = __hiloint2double(v.y, v.x);
    x[2 * k + 1] = __hiloint2double(v.w, v.z);
  }
}
__device__ __forceinline__ void gather_cam_tex(cudaTextureObject_t tex, int c, float *x) {
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float4 v = tex1Dfetch<float4>(tex, 3 * c + k);
    x[4 * k] = v.x; x[4 * k + 1] = v.y; x[4 * k + 2] = v.z; x[4 * k + 3] = v.w;
  }
}

// the factored stream (8 scalars per slot) through the texture path
__device__ __forceinline__ void load_stream_tex(cudaTextureObject_t tex, size_t row, int lane, double *v) {
  const int base = (int)(row * 128) + lane;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int4 t = tex1Dfetch<int4>(tex, base + j * 32);
    v[2 * j] = __hiloint2double(t.y, t.x);
    v[2 * j + 1] = __hiloint2double(t.w, t.z);
  }
}
__device__ __forceinline__ void load_stream_tex(cudaTextureObject_t tex, size_t row, int lane, float *v) {
  const int base = (int)(row * 128) + lane;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 t = tex1Dfetch<float2>(tex, base + j * 32);
    v[2 * j] = t.x;
    v[2 * j + 1] = t.y;
  }
}

// SPLIT > 1: a shard with about one wave of slices (an eighth of Venice: 3 900 slices for 3 552 warp slots) is bound by
// the latency chain of its tallest slices (long-scoreboard stalls, LSU pipe 34 % busy, 31 % of the warp slots active:
// profiles/r02_ncu_shard8_summary.txt).  SPLIT warps of one CTA then share a slice, each taking every SPLIT-th row;
// the partial sums t meet in shared memory between the phases (two barriers per pass).
template <typename T, int STORE, int MODE, bool INFO, int RC, int SPLIT = 1>
__global__ void __launch_bounds__(ell_cta_warps(SPLIT) * 32, ell_min_blocks(RC) * ELL_WARPS / ell_cta_warps(SPLIT))
k_schur_ell(Dev<T> d, const T *__restrict__ xin, T *sout, T *zero_set, int check_active) {
  static_assert(SPLIT == 1 || RC == 0, "register-cached rows and split slices are not combined");
  constexpr int K = STORE == ST_FACT ? 8 : 24;
  constexpr int NC = RC > 0 ? RC : 1;
  constexpr int NWC = ell_cta_warps(SPLIT), NT = NWC * 32, SPC = NWC / SPLIT;  // warps, threads, slices per CTA
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T *sv = reinterpret_cast<T *>(smem_raw);
  int *sc = reinterpret_cast<int *>(sv + NWC * 32 * 9);
  const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5, part = wib % SPLIT;
  const int w = blockIdx.x * SPC + wib / SPLIT;
  const bool wvalid = w < d.nW;
  WDesc wd{0, 0, 0, 0, 0, 0xffff};
  if (wvalid) wd = load_desc(d.wdesc, w);
  const int p = wd.heavy ? wd.pt0 : wd.pt0 + lane;
  const bool pvalid = wvalid && (wd.heavy || lane < wd.np);
  const size_t npl = (size_t)d.nPlPad;
  // ---- everything up to dependent_ready touches solve-constant data only (overlaps the preceding camera kernel)
  T X0 = T(0), X1 = T(0), X2 = T(0);
  if (STORE == ST_FACT && pvalid) { X0 = __ldg(d.XP + p); X1 = __ldg(d.XP + npl + p); X2 = __ldg(d.XP + 2 * npl + p); }
  T cs[NC][8];
  int cc[NC];
  if (RC > 0) {
#pragma unroll
    for (int r = 0; r < NC; ++r) {
      cc[r] = -1;
      if (r < wd.D) {
        const size_t row = (size_t)wd.row0 + r;
        cc[r] = __ldg(d.r_cam + row * 32 + lane);
        load_stream<T, 8>(d.WJ, row, lane, cs[r]);
      }
    }
  }
  T hi[9], wv[3] = {T(0), T(0), T(0)};
#pragma unroll
  for (int k = 0; k < 9; ++k) hi[k] = T(0);
  auto load_point_blocks = [&]() {
    if (pvalid) {
#pragma unroll
      for (int k = 0; k < 9; ++k) hi[k] = __ldg(d.HllInvP + k * npl + p);
      if (MODE != SM_FULL) {
#pragma unroll
        for (int k = 0; k < 3; ++k) wv[k] = __ldg(d.wP + k * npl + p);
      }
    }
  };
  if (!BA_ELL_LATE_HI) load_point_blocks();
  if (check_active) {
    asm volatile("griddepcontrol.wait;" ::: "memory");  // the preceding camera kernel (p, state, replica sets) is complete
    if (!d.st->active) return;
  }
  if (MODE == SM_FULL && zero_set) {  // clear the replica set the NEXT edge pass will accumulate into
    const unsigned tot = (unsigned)NREP * 9u * (unsigned)d.nC;
    for (unsigned i = blockIdx.x * NT + tid; i < tot; i += gridDim.x * NT) zero_set[i] = T(0);
  }
  if (SPLIT == 1 && !wvalid) return;  // whole warps only (split: an idle warp has D = 0 and still meets the barriers)
  T *dst = MODE != SM_BACKSUB ? sout + (size_t)(blockIdx.x % NREP) * (9 * (size_t)d.nC) : nullptr;
  // ---- phase 1
  T t0 = T(0), t1 = T(0), t2 = T(0);
  auto first = [&](const T *s, int c, size_t slot) {
    __align__(16) T g[CamPad<T>::N];
    if (BA_ELL_TEX) gather_cam_tex(d.xpad_tex, c, g);
    else gather_cam<T, true>(xin, c, g);
    T om[4];
    if (INFO) load_info<T>(d.r_info, slot, om);
    if (STORE == ST_FACT) fact_first<T, INFO>(s, g, X0, X1, X2, om, t0, t1, t2);
    else jac_first<T, INFO>(s, g, om, t0, t1, t2);
  };
  if (RC > 0) {
#pragma unroll
    for (int r = 0; r < NC; ++r)
      if (r < wd.D && cc[r] >= 0) first(cs[r], cc[r], ((size_t)wd.row0 + r) * 32 + lane);
  }
#pragma unroll ELL_U1
  for (int r = RC + part; r < wd.D; r += SPLIT) {
    const size_t row = (size_t)wd.row0 + r, slot = row * 32 + lane;
    const int c = __ldg(d.r_cam + slot);
    T s[K];
    if (BA_ELL_TEX >= 3 && STORE == ST_FACT) load_stream_tex(d.wj_tex, row, lane, s);
    else load_stream<T, K>(d.WJ, row, lane, s);
    if (c >= 0) first(s, c, slot);
  }
  if (BA_ELL_LATE_HI) load_point_blocks();
  if (wd.heavy) { t0 = warp_sum<T>(t0); t1 = warp_sum<T>(t1); t2 = warp_sum<T>(t2); }
  if (SPLIT > 1) {  // the parts' partial sums meet in the (not yet used) scatter staging area
    T *ex = sv + (size_t)wib * 32 * 9 + 3 * lane;
    ex[0] = t0; ex[1] = t1; ex[2] = t2;
    __syncthreads();
    t0 = t1 = t2 = T(0);
    const T *all = sv + (size_t)(wib - part) * 32 * 9 + 3 * lane;
#pragma unroll
    for (int q = 0; q < SPLIT; ++q) { t0 += all[q * 32 * 9]; t1 += all[q * 32 * 9 + 1]; t2 += all[q * 32 * 9 + 2]; }
    __syncthreads();
  }
  // ---- point step
  T u0 = hi[0] * t0 + hi[1] * t1 + hi[2] * t2;
  T u1 = hi[3] * t0 + hi[4] * t1 + hi[5] * t2;
  T u2 = hi[6] * t0 + hi[7] * t1 + hi[8] * t2;
  if (MODE == SM_INIT) { u0 -= wv[0]; u1 -= wv[1]; u2 -= wv[2]; }  // u = Hll^-1 Hlp x0 - w: one pass yields Hpl Hll^-1 Hlp x0 - Hpl w
  if (MODE == SM_BACKSUB) {
    const T l0 = wv[0] - u0, l1 = wv[1] - u1, l2 = wv[2] - u2;
    const bool owner = pvalid && part == 0 && (!wd.heavy || lane == 0);
    if (owner) {
      d.dxl[3 * (size_t)p] = l0;
      d.dxl[3 * (size_t)p + 1] = l1;
      d.dxl[3 * (size_t)p + 2] = l2;
    }
    if (d.model_partials) {
      // point part of  ||J dx + r||^2 - ||r||^2 = dx^T (J^T J) dx - 2 g^T dx  (computeRhoDenominator, lm_algo.cu:61-126,
      // without a pass over the observations): dx_l^T Hll dx_l + 2 (Hlp dx_c)^T dx_l - 2 g_l^T dx_l; t = Hlp dx_c is in hand
      double m = 0.;
      if (owner) {
        T h[9], gv[3];
#pragma unroll
        for (int k = 0; k < 9; ++k) h[k] = d.Hll[9 * (size_t)p + k];
#pragma unroll
        for (int k = 0; k < 3; ++k) gv[k] = d.gl[3 * (size_t)p + k];
        const T q0 = h[0] * l0 + h[1] * l1 + h[2] * l2, q1 = h[3] * l0 + h[4] * l1 + h[5] * l2, q2 = h[6] * l0 + h[7] * l1 + h[8] * l2;
        m = (double)(l0 * q0 + l1 * q1 + l2 * q2) + 2. * (double)(t0 * l0 + t1 * l1 + t2 * l2) -
            2. * (double)(gv[0] * l0 + gv[1] * l1 + gv[2] * l2);
      }
      m = warp_sum<double>(m);
      if (lane == 0 && part == 0 && wvalid) d.model_partials[w] = m;
    }
    return;
  }
  // ---- phase 2
  auto second = [&](const T *s, int c, size_t slot) {
    T o[9];
    if (c >= 0) {
      T om[4];
      if (INFO) load_info<T>(d.r_info, slot, om);
      if (STORE == ST_FACT) fact_second<T, INFO>(s, u0, u1, u2, X0, X1, X2, om, o);
      else jac_second<T, INFO>(s, u0, u1, u2, om, o);
    }
    warp_scatter9<T>(sv, sc, o, c, dst);
  };
  if (RC > 0) {
#pragma unroll
    for (int r = 0; r < NC; ++r)
      if (r < wd.D) second(cs[r], cc[r], ((size_t)wd.row0 + r) * 32 + lane);
  }
#pragma unroll ELL_U2
  for (int r = RC + part; r < wd.D; r += SPLIT) {
    const size_t row = (size_t)wd.row0 + r, slot = row * 32 + lane;
    const int c = __ldg(d.r_cam + slot);
    T s[K];
    if (BA_ELL_TEX >= 2 && STORE == ST_FACT) load_stream_tex(d.wj_tex, row, lane, s);
    else load_stream<T, K>(d.WJ, row, lane, s);
    second(s, c, slot);
  }
}

// --------------------------------------------------------------------------------
// TMA-staged flavour of the factored S*p pass.  A CTA owns up to ELL_WARPS consecutive slices whose rows fit its
// shared-memory stage (host-side grouping, Dev::cdesc).  Lane 0 of every warp fetches the warp's whole slice -- one
// contiguous run of D * 8 * 32 scalars plus D * 32 camera indices -- with two cp.async.bulk copies that complete on
// the warp's own mbarrier; both phases then read the stream from shared memory.  Compared with k_schur_ell this
// removes the second (L2) read of the stream, takes the stream's fills off the LSU data pipe (the limiter of the pass,
// profiles/README.md) and puts the whole slice in flight without landing registers, before griddepcontrol.wait.
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  unsigned ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  }
}

template <typename T>
__host__ __device__ constexpr size_t schur_tma_smem(int rows) {
  // stage rows [rows][8*32] T, camera rows [rows][32] int, one mbarrier per warp, scatter staging
  return (size_t)rows * (8 * 32 * sizeof(T) + 32 * sizeof(int)) + 128 + schur_ell_smem<T>();
}

#ifndef BA_TMA_MINB
#define BA_TMA_MINB 2
#endif
template <typename T, int MODE, bool INFO>
__global__ void __launch_bounds__(ELL_THREADS, BA_TMA_MINB)
k_schur_tma(Dev<T> d, const T *__restrict__ xin, T *sout, T *zero_set, int check_active) {
  using V = typename Vec2<T>::type;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int rows_cap = d.tma_rows;
  T *stage = reinterpret_cast<T *>(smem_raw);
  int *cstage = reinterpret_cast<int *>(stage + (size_t)rows_cap * 256);
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(cstage + (size_t)rows_cap * 32);
  T *sv = reinterpret_cast<T *>(bars + 16);
  int *sc = reinterpret_cast<int *>(sv + ELL_WARPS * 32 * 9);
  const int tid = threadIdx.x, lane = tid & 31, wi = tid >> 5;
  const int2 cd = __ldg(d.cdesc + blockIdx.x);
  const bool wvalid = wi < cd.y;
  WDesc wd{0, 0, 0, 0, 0, 0xffff};
  if (wvalid) wd = load_desc(d.wdesc, cd.x + wi);
  const bool staged = wvalid && wd.soff != 0xffff && wd.D > 0;
  const int p = wd.heavy ? wd.pt0 : wd.pt0 + lane;
  const bool pvalid = wvalid && (wd.heavy || lane < wd.np);
  const size_t npl = (size_t)d.nPlPad;
  // ---- solve-constant data: the bulk copies of this warp's slice go out first
  T *my_stage = stage + (size_t)(staged ? wd.soff : 0) * 256;
  int *my_cams = cstage + (size_t)(staged ? wd.soff : 0) * 32;
  if (staged) {
    if (lane == 0) {
      mbar_init(bars + wi, 1);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the async proxy must see the initialised barrier
      const unsigned b_stream = (unsigned)wd.D * 256u * (unsigned)sizeof(T), b_cam = (unsigned)wd.D * 128u;
      mbar_expect_tx(bars + wi, b_stream + b_cam);
      bulk_g2s(my_stage, d.WJ + (size_t)wd.row0 * 256, b_stream, bars + wi);
      bulk_g2s(my_cams, d.r_cam + (size_t)wd.row0 * 32, b_cam, bars + wi);
    }
    __syncwarp();
  }
  T X0 = T(0), X1 = T(0), X2 = T(0);
  T hi[9], wv[3] = {T(0), T(0), T(0)};
#pragma unroll
  for (int k = 0; k < 9; ++k) hi[k] = T(0);
  if (pvalid) {
    X0 = __ldg(d.XP + p); X1 = __ldg(d.XP + npl + p); X2 = __ldg(d.XP + 2 * npl + p);
#pragma unroll
    for (int k = 0; k < 9; ++k) hi[k] = __ldg(d.HllInvP + k * npl + p);
    if (MODE != SM_FULL) {
#pragma unroll
      for (int k = 0; k < 3; ++k) wv[k] = __ldg(d.wP + k * npl + p);
    }
  }
  if (check_active) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    // the bulk copies in flight target this CTA's shared memory: they must have landed before the CTA may exit
    if (!d.st->active) { if (staged) mbar_wait(bars + wi, 0); return; }
  }
  if (MODE == SM_FULL && zero_set) {
    const unsigned tot = (unsigned)NREP * 9u * (unsigned)d.nC;
    for (unsigned i = blockIdx.x * ELL_THREADS + tid; i < tot; i += gridDim.x * ELL_THREADS) zero_set[i] = T(0);
  }
  if (!wvalid) return;
  T *dst = MODE != SM_BACKSUB ? sout + (size_t)(blockIdx.x % NREP) * (9 * (size_t)d.nC) : nullptr;
  if (staged) mbar_wait(bars + wi, 0);
  // one row of the slice: camera index and the 8 stream values of this lane's slot
  auto row_cam = [&](int r) -> int {
    return staged ? my_cams[r * 32 + lane] : __ldg(d.r_cam + ((size_t)wd.row0 + r) * 32 + lane);
  };
  auto row_stream = [&](int r, T *sl) {
    if (staged) {
      const V *src = reinterpret_cast<const V *>(my_stage) + (size_t)r * 128 + lane;
#pragma unroll
      for (int j = 0; j < 4; ++j) { const V x = src[j * 32]; sl[2 * j] = x.x; sl[2 * j + 1] = x.y; }
    } else {
      load_stream<T, 8>(d.WJ, (size_t)wd.row0 + r, lane, sl);
    }
  };
  // ---- phase 1
  T t0 = T(0), t1 = T(0), t2 = T(0);
#pragma unroll 2
  for (int r = 0; r < wd.D; ++r) {
    const int c = row_cam(r);
    T sl[8];
    row_stream(r, sl);
    if (c >= 0) {
      __align__(16) T g[CamPad<T>::N];
      gather_cam<T, true>(xin, c, g);
      T om[4];
      if (INFO) load_info<T>(d.r_info, ((size_t)wd.row0 + r) * 32 + lane, om);
      fact_first<T, INFO>(sl, g, X0, X1, X2, om, t0, t1, t2);
    }
  }
  if (wd.heavy) { t0 = warp_sum<T>(t0); t1 = warp_sum<T>(t1); t2 = warp_sum<T>(t2); }
  T u0 = hi[0] * t0 + hi[1] * t1 + hi[2] * t2;
  T u1 = hi[3] * t0 + hi[4] * t1 + hi[5] * t2;
  T u2 = hi[6] * t0 + hi[7] * t1 + hi[8] * t2;
  if (MODE == SM_INIT) { u0 -= wv[0]; u1 -= wv[1]; u2 -= wv[2]; }
  if (MODE == SM_BACKSUB) {
    if (pvalid && (!wd.heavy || lane == 0)) {
      d.dxl[3 * (size_t)p] = wv[0] - u0;
      d.dxl[3 * (size_t)p + 1] = wv[1] - u1;
      d.dxl[3 * (size_t)p + 2] = wv[2] - u2;
    }
    return;
  }
  // ---- phase 2
  for (int r = 0; r < wd.D; ++r) {
    const int c = row_cam(r);
    T sl[8], o[9];
    row_stream(r, sl);
    if (c >= 0) {
      T om[4];
      if (INFO) load_info<T>(d.r_info, ((size_t)wd.row0 + r) * 32 + lane, om);
      fact_second<T, INFO>(sl, u0, u1, u2, X0, X1, X2, om, o);
    }
    warp_scatter9<T>(sv, sc, o, c, dst);
  }
}

// --------------------------------------------------------------------------------
// Persistent PCG kernel: the whole Schur-PCG solve (SchurPCGSolver::solve, src/solver/schur_pcg_solver.cu:599-679:
// MakeV fused with S*x0, the do-while :265-407, SolveW) in ONE cooperative launch, one CTA of 24 warps per SM.
//   edge phase   : warps take slices from a global counter (slices are sorted by row count, so this is longest-first
//                  scheduling); per slice exactly the two-phase pass of k_schur_ell<FACT>
//   camera phase : the 9 nC rows are spread over all CTAs (a CTA owns whole cameras); q = Hpp p - T^T Y, alpha, x / r
//                  update, z = Hpp^-1 r, rho, refuse / tol tests, p update, T p -> xpad; the two dot products go through
//                  per-CTA partials summed in CTA order by every CTA (deterministic, identical everywhere)
// separated by grid barriers (4 per iteration).  Compared with two launches per iteration this removes the launch /
// drain gaps and runs the camera step on every SM instead of one 16-CTA cluster.  With world_size > 1 the push exchange
// of k_pcg_cam_reg runs inside the camera phase (one more grid barrier orders the pushes before the flags go up).
constexpr int PK_THREADS = 768;
constexpr int PK_WARPS = PK_THREADS / 32;

struct PkArgs {
  int max_iter;
  double tol, refuse;
  int *counters;        // [4] slice counters of the edge phases (zero on entry)
  double *partials;     // [2][gridDim] per-CTA partial dot products
  double *timers;       // [4] device-timer sums: edge-phase ns, edge phases, camera-phase ns, camera phases
};

template <typename T>
__host__ __device__ constexpr size_t pk_smem(int rpt) {
  // scatter staging per warp + p / r / y exchange rows of the camera phase
  return (size_t)PK_WARPS * 32 * 9 * sizeof(T) + (size_t)PK_WARPS * 32 * sizeof(int) + 3 * (size_t)rpt * PK_THREADS * sizeof(T) + 64 * sizeof(double);
}

__device__ __forceinline__ unsigned long long pk_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// The gathered vector (xpad) is rewritten by the camera phase of the SAME kernel: it must be read through the coherent
// path (ld.global.ca is invalidated by the acquire side of the grid barrier; the non-coherent LDG.CONSTANT path of
// __ldg carries no such guarantee inside one kernel).
template <typename T>
__device__ __forceinline__ void gather_cam_coherent(const T *xpad, int c, T *x) {
  // plain (ld.global, default .ca) vector loads: no __ldg, no __restrict__, so the compiler cannot pick LDG.CONSTANT
  using V = typename CamPad<T>::vec;
  constexpr int NV = CamPad<T>::N * sizeof(T) / 16;
  const V *src = reinterpret_cast<const V *>(xpad + (size_t)CamPad<T>::N * c);
  V *xv = reinterpret_cast<V *>(x);
#pragma unroll
  for (int k = 0; k < NV; ++k) xv[k] = src[k];
}

// one slice (or heavy point), both phases; MODE as in k_schur_ell
template <typename T, int MODE, bool INFO>
__device__ __noinline__ void pk_slice(const Dev<T> &d, const WDesc &wd, int lane, const T *xin, T *dst, T *sv, int *sc) {
  const int p = wd.heavy ? wd.pt0 : wd.pt0 + lane;
  const bool pvalid = wd.heavy || lane < wd.np;
  const size_t npl = (size_t)d.nPlPad;
  T X0 = T(0), X1 = T(0), X2 = T(0);
  if (pvalid) { X0 = __ldg(d.XP + p); X1 = __ldg(d.XP + npl + p); X2 = __ldg(d.XP + 2 * npl + p); }
  T t0 = T(0), t1 = T(0), t2 = T(0);
#pragma unroll 2
  for (int r = 0; r < wd.D; ++r) {
    const size_t row = (size_t)wd.row0 + r, slot = row * 32 + lane;
    const int c = __ldg(d.r_cam + slot);
    T s[8];
    load_stream<T, 8>(d.WJ, row, lane, s);
    if (c >= 0) {
      __align__(16) T g[CamPad<T>::N];
      gather_cam_coherent(xin, c, g);
      T om[4];
      if (INFO) load_info<T>(d.r_info, slot, om);
      fact_first<T, INFO>(s, g, X0, X1, X2, om, t0, t1, t2);
    }
  }
  T hi[9], wv[3] = {T(0), T(0), T(0)};
#pragma unroll
  for (int k = 0; k < 9; ++k) hi[k] = T(0);
  if (pvalid) {
#pragma unroll
    for (int k = 0; k < 9; ++k) hi[k] = __ldg(d.HllInvP + k * npl + p);
    if (MODE != SM_FULL) {
#pragma unroll
      for (int k = 0; k < 3; ++k) wv[k] = __ldg(d.wP + k * npl + p);
    }
  }
  if (wd.heavy) { t0 = warp_sum<T>(t0); t1 = warp_sum<T>(t1); t2 = warp_sum<T>(t2); }
  T u0 = hi[0] * t0 + hi[1] * t1 + hi[2] * t2;
  T u1 = hi[3] * t0 + hi[4] * t1 + hi[5] * t2;
  T u2 = hi[6] * t0 + hi[7] * t1 + hi[8] * t2;
  if (MODE == SM_INIT) { u0 -= wv[0]; u1 -= wv[1]; u2 -= wv[2]; }
  if (MODE == SM_BACKSUB) {
    if (pvalid && (!wd.heavy || lane == 0)) {
      d.dxl[3 * (size_t)p] = wv[0] - u0;
      d.dxl[3 * (size_t)p + 1] = wv[1] - u1;
      d.dxl[3 * (size_t)p + 2] = wv[2] - u2;
    }
    return;
  }
  for (int r = 0; r < wd.D; ++r) {
    const size_t row = (size_t)wd.row0 + r, slot = row * 32 + lane;
    const int c = __ldg(d.r_cam + slot);
    T s[8], o[9];
    load_stream<T, 8>(d.WJ, row, lane, s);
    if (c >= 0) {
      T om[4];
      if (INFO) load_info<T>(d.r_info, slot, om);
      fact_second<T, INFO>(s, u0, u1, u2, X0, X1, X2, om, o);
    }
    warp_scatter9<T>(sv, sc, o, c, dst);
  }
}

template <typename T, int MODE, bool INFO>
__device__ __forceinline__ void pk_edge_phase(const Dev<T> &d, int *counter, const T *xin, T *sout, T *zero_set, T *sv, int *sc) {
  const int tid = threadIdx.x, lane = tid & 31;
  if (zero_set) {  // clear the replica set the NEXT edge phase accumulates into
    const unsigned tot = (unsigned)NREP * 9u * (unsigned)d.nC;
    for (unsigned i = blockIdx.x * PK_THREADS + tid; i < tot; i += gridDim.x * PK_THREADS) zero_set[i] = T(0);
  }
  T *dst = MODE != SM_BACKSUB ? sout + (size_t)(blockIdx.x % NREP) * (9 * (size_t)d.nC) : nullptr;
  auto grab = [&]() -> int {
    int v = 0;
    if (lane == 0) v = atomicAdd(counter, 1);
    return __shfl_sync(0xffffffffu, v, 0);
  };
  int sidx = grab();
  while (sidx < d.nW) {
    const int nxt = grab();  // requested before the slice is processed: its latency hides behind the slice
    const WDesc wd = load_desc(d.wdesc, sidx);
    pk_slice<T, MODE, INFO>(d, wd, lane, xin, dst, sv, sc);
    sidx = nxt;
  }
}

// sum of the per-CTA partials, the same fixed tree on every CTA; result valid in every thread
__device__ __forceinline__ double pk_total(const double *partials, int n, double *bcast) {
  const int lane = threadIdx.x & 31;
  if (threadIdx.x < 32) {
    double mine = 0.;
    for (int r = lane; r < n; r += 32) mine += __ldcg(partials + r);
    mine = warp_sum<double>(mine);
    if (lane == 0) *bcast = mine;
  }
  __syncthreads();
  const double v = *bcast;
  __syncthreads();
  return v;
}
__device__ __forceinline__ void pk_cta_partial(double v, double *red_warp, double *out) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = warp_sum<double>(v);
  if (lane == 0) red_warp[wid] = v;
  __syncthreads();
  if (wid == 0) {
    double x = lane < PK_WARPS ? red_warp[lane] : 0.;
    x = warp_sum<double>(x);
    if (lane == 0) *out = x;
  }
  __syncthreads();
}

template <typename T, bool INFO, int RPT>
__global__ void __launch_bounds__(PK_THREADS, 1) k_pcg_persistent(const __grid_constant__ Dev<T> d, const __grid_constant__ PkArgs a, const __grid_constant__ PeerExchange px) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T *sv = reinterpret_cast<T *>(smem_raw);
  int *sc = reinterpret_cast<int *>(sv + PK_WARPS * 32 * 9);
  T *p_s = reinterpret_cast<T *>(sc + PK_WARPS * 32);
  T *r_s = p_s + RPT * PK_THREADS;
  T *y_s = r_s + RPT * PK_THREADS;
  double *red_warp = reinterpret_cast<double *>(y_s + RPT * PK_THREADS);  // [32] + [2] broadcast slots
  double *bcast = red_warp + 32;
  const int tid = threadIdx.x, G = gridDim.x, b = blockIdx.x;
  const size_t rs = 9 * (size_t)d.nC, set_sz = (size_t)NREP * rs;
  // camera rows of this CTA: whole cameras
  const int per = (d.nC + G - 1) / G;
  const int c0 = min(d.nC, b * per), c1 = min(d.nC, c0 + per);
  const int i0 = 9 * c0, nrows = 9 * (c1 - c0);
  const bool timer = b == 0 && tid == 0 && a.timers;
  // The PCG state lives in shared memory between camera phases (every thread holds an identical copy while it works
  // on it): 20 registers that would otherwise stay live across the edge phases and push them into spills.
  __shared__ PcgState st_s;
  __shared__ unsigned long long epoch_s;
  if (tid == 0) {
    PcgState st0;
    st0.rho = 0.; st0.rho_prev = 0.; st0.rho_min = INFINITY; st0.alpha = 0.; st0.pq = 0.;
    st0.tol = a.tol; st0.refuse = a.refuse; st0.n = 0; st0.max_iter = a.max_iter > 1 ? a.max_iter : 1; st0.active = 1; st0.done = 0; st0.refused = 0; st0.pad = 0;
    st_s = st0;
    epoch_s = px.epoch;
  }
  __syncthreads();
  int pass = 0;

  // camera phase.  INIT: r = g_c - (Hpp x0 - T^T Y), z, rho, p = z.  Otherwise one PCG iteration's camera work.
  auto camera_phase = [&](const T *sB, bool init) -> bool {
    PcgState st = st_s;
    unsigned long long epoch = epoch_s;
    T hd[RPT][9], hinv[RPT][9], tin[RPT][3], tout[RPT][3], pj[RPT], rj[RPT], xj[RPT], xb[RPT], sj[RPT], qj[RPT], zj[RPT];
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      const int l = tid + j * PK_THREADS;
      pj[j] = rj[j] = xj[j] = xb[j] = sj[j] = qj[j] = zj[j] = T(0);
#pragma unroll
      for (int m = 0; m < 3; ++m) tin[j][m] = tout[j][m] = T(0);
#pragma unroll
      for (int m = 0; m < 9; ++m) hd[j][m] = hinv[j][m] = T(0);
      if (l < nrows) {
        const int i = i0 + l, c = i / 9, k = i - 9 * c, bk = k / 3, kk = k - 3 * bk;
#pragma unroll
        for (int m = 0; m < 9; ++m) {
          hd[j][m] = __ldg(d.HppD + 81 * (size_t)c + 9 * k + m);
          hinv[j][m] = __ldg(d.HppInv + 81 * (size_t)c + 9 * k + m);
        }
        const T *tb = d.camT + (size_t)CAMT * c + 9 * bk;
#pragma unroll
        for (int m = 0; m < 3; ++m) {
          const T rowv = __ldg(tb + 3 * kk + m), colv = __ldg(tb + 3 * m + kk);
          tin[j][m] = bk == 1 ? colv : rowv;
          tout[j][m] = bk == 1 ? rowv : colv;
        }
        xj[j] = d.dxc[i];
        T sbv[NREP];
#pragma unroll
        for (int rep = 0; rep < NREP; ++rep) sbv[rep] = __ldcg(sB + rep * rs + i);
        T sb = T(0);
#pragma unroll
        for (int rep = 0; rep < NREP; ++rep) sb += sbv[rep];
        sj[j] = sb;
        if (init) { pj[j] = xj[j]; rj[j] = d.gc[i]; }
        else { pj[j] = __ldcg(d.p + i); rj[j] = __ldcg(d.r + i); }
        p_s[l] = pj[j];
        if (px.ws > 1) {  // push this rank's partial into every rank's inbox
#pragma unroll
          for (int r = 0; r < 8; ++r)
            if (r < px.ws) (reinterpret_cast<T *>(px.data[r]) + ((size_t)((int)(epoch & 1) * px.ws + px.rank) * 2) * rs)[i] = sb;
        }
      }
    }
    if (px.ws > 1) {
      __threadfence_system();
      grid.sync();  // every CTA's pushes are ordered before the flags go up
      const int par = (int)(epoch & 1);
      if (b == 0 && tid < px.ws)
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(px.flag[tid] + par * 8 + px.rank), "l"(epoch) : "memory");
      if (nrows > 0) {
        if (tid < px.ws) {
          const unsigned long long *f = px.flag[px.rank] + par * 8 + tid;
          unsigned long long v = 0;
          const long long t0 = clock64();
          do {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
          } while (v < epoch && clock64() - t0 < 4000000000ll);
          if (v < epoch) d.scal[S_XCHG_ERR] = 1.;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
          const int l = tid + j * PK_THREADS;
          if (l < nrows) {
            const int i = i0 + l;
            T pb[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
              pb[r] = T(0);
              if (r < px.ws) pb[r] = __ldcv(reinterpret_cast<const T *>(px.data[px.rank]) + ((size_t)(par * px.ws + r) * 2) * rs + i);
            }
            T sb = T(0);
#pragma unroll
            for (int r = 0; r < 8; ++r) sb += pb[r];
            sj[j] = sb;
          }
        }
      }
      epoch += 1;
    }
    // s = T^T Y
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      const int l = tid + j * PK_THREADS;
      if (l < nrows) y_s[l] = sj[j];
    }
    __syncthreads();
    double acc = 0.;
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      const int l = tid + j * PK_THREADS;
      if (l < nrows) {
        const int k = (i0 + l) % 9, lb = l - k % 3, lc = l - k;
        const T sv_ = tout[j][0] * y_s[lb] + tout[j][1] * y_s[lb + 1] + tout[j][2] * y_s[lb + 2];
        T aq = T(0);
#pragma unroll
        for (int m = 0; m < 9; ++m) aq += hd[j][m] * p_s[lc + m];
        qj[j] = aq - sv_;                          // q = Hpp p - Hpl Hll^-1 Hlp p  (:359-366)
        if (init) rj[j] = rj[j] - qj[j];           // r = v - S x0                  (:182-258)
        else acc += (double)(pj[j] * qj[j]);
      }
    }
    if (!init) {
      pk_cta_partial(acc, red_warp, a.partials + b);
      grid.sync();
      const double pq = pk_total(a.partials, G, bcast);
      const T alpha = pq != 0. ? (T)(st.rho / pq) : T(0);   // :385
      st.alpha = (double)alpha; st.pq = pq;
#pragma unroll
      for (int j = 0; j < RPT; ++j) {
        xb[j] = xj[j];                             // :389
        xj[j] = xj[j] + alpha * pj[j];             // :391
        rj[j] = rj[j] + (-alpha) * qj[j];          // :400
      }
    }
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      const int l = tid + j * PK_THREADS;
      if (l < nrows) r_s[l] = rj[j];
    }
    __syncthreads();
    acc = 0.;
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      const int l = tid + j * PK_THREADS;
      if (l < nrows) {
        const int lc = l - (i0 + l) % 9;
        T az = T(0);
#pragma unroll
        for (int m = 0; m < 9; ++m) az += hinv[j][m] * r_s[lc + m];
        zj[j] = az;                                // z = M^-1 r                    (:273-275)
        acc += (double)(rj[j] * az);
      }
    }
    pk_cta_partial(acc, red_warp, a.partials + G + b);
    grid.sync();
    const double rho_new = (double)(T)pk_total(a.partials + G, G, bcast + 1);  // rhoN is a T in the reference
    bool cont = true, refused = false;
    if (!init) {
      st.n += 1;                                   // :405
      st.done = fabs(st.rho) < st.tol;             // :406
      cont = !st.done && st.n < st.max_iter;       // :407
    }
    T beta = T(0);
    if (cont) {
      if (rho_new > st.refuse * st.rho_min) {      // :288-295
        st.refused = 1; refused = true; cont = false;
      } else {
        st.rho_min = fmin(st.rho_min, rho_new);    // :296
        if (!init) beta = st.rho != 0. ? (T)(rho_new / st.rho) : T(0);  // :299
        st.rho_prev = st.rho;
        st.rho = rho_new;
      }
    }
    st.active = cont ? 1 : 0;
    __syncthreads();  // p_s is rewritten below
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      const int l = tid + j * PK_THREADS;
      if (l < nrows) {
        const int i = i0 + l;
        d.r[i] = rj[j];
        if (!init) {
          xj[j] = refused ? xb[j] : xj[j];
          d.dxc[i] = xj[j];
          d.xbk[i] = xb[j];
        }
        T pv = pj[j];
        if (cont) {
          pv = init ? zj[j] : zj[j] + beta * pj[j];  // :310 / :299-306
          d.p[i] = pv;
        }
        // what the next edge phase gathers: T p while the loop continues, T x for the back-substitution after it
        p_s[l] = cont ? pv : xj[j];
      }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      const int l = tid + j * PK_THREADS;
      if (l < nrows) {
        const int i = i0 + l, lb = l - (i % 9) % 3;
        d.xpad[(size_t)CamPad<T>::N * (i / 9) + i % 9] = tin[j][0] * p_s[lb] + tin[j][1] * p_s[lb + 1] + tin[j][2] * p_s[lb + 2];
      }
    }
    if (b == 0 && tid == 0) a.counters[(pass + 2) & 3] = 0;  // the counter of the edge phase after the next one
    __syncthreads();  // every thread has taken its copy of the state
    if (tid == 0) { st_s = st; epoch_s = epoch; }
    __syncthreads();
    return cont;
  };

  unsigned long long t_a = 0;
  auto tick = [&](int slot) {  // CTA 0, thread 0: device-timer bookkeeping between grid barriers
    if (timer) {
      const unsigned long long t = pk_now();
      if (slot >= 0) { a.timers[slot] += (double)(t - t_a); a.timers[slot + 1] += 1.; }
      t_a = t;
    }
  };

  // ---- MakeV + S*x0 in one pass (xpad = T x0 was prepared by the host), then the INIT camera phase
  tick(-1);
  pk_edge_phase<T, SM_INIT, INFO>(d, a.counters + (pass & 3), d.xpad, d.sB, nullptr, sv, sc);
  grid.sync();
  tick(0);
  bool cont = camera_phase(d.sB, true);
  pass += 1;
  grid.sync();
  tick(2);
  // ---- the do-while of schur_pcg_solver.cu:265-407
  int it = 0;
  while (cont) {
    T *cur = d.sB + (size_t)((it + 1) & 1) * set_sz, *other = d.sB + (size_t)(it & 1) * set_sz;
    pk_edge_phase<T, SM_FULL, INFO>(d, a.counters + (pass & 3), d.xpad, cur, other, sv, sc);
    grid.sync();
    tick(0);
    cont = camera_phase(cur, false);
    pass += 1;
    it += 1;
    grid.sync();
    tick(2);
  }
  // ---- back-substitution: dx_l = w - Hll^-1 Hlp dx_c  (xpad holds T dx_c)
  pk_edge_phase<T, SM_BACKSUB, INFO>(d, a.counters + (pass & 3), d.xpad, nullptr, nullptr, sv, sc);
  if (b == 0 && tid == 0) *d.st = st_s;
}

// --------------------------------------------------------------------------------
// LM scalars: sum (J dx + r)^2 at x and sum r^2 at x + dx in one pass.
template <typename T, bool PROG>
__global__ void __launch_bounds__(ELL_THREADS) k_trial_ell(Dev<T> d) {
  const int lane = threadIdx.x & 31, w = blockIdx.x * ELL_WARPS + (threadIdx.x >> 5);
  if (w >= d.nW) return;
  const WDesc wd = load_desc(d.wdesc, w);
  const int p = wd.heavy ? wd.pt0 : wd.pt0 + lane;
  const bool pvalid = wd.heavy || lane < wd.np;
  T X0 = T(0), X1 = T(0), X2 = T(0), D0 = T(0), D1 = T(0), D2 = T(0);
  if (pvalid) {
    X0 = d.pt[3 * (size_t)p]; X1 = d.pt[3 * (size_t)p + 1]; X2 = d.pt[3 * (size_t)p + 2];
    D0 = d.dxl[3 * (size_t)p]; D1 = d.dxl[3 * (size_t)p + 1]; D2 = d.dxl[3 * (size_t)p + 2];
  }
  double a_model = 0., a_cost = 0.;
  for (int r = 0; r < wd.D; ++r) {
    const size_t slot = ((size_t)wd.row0 + r) * 32 + lane;
    const int c = __ldg(d.r_cam + slot);
    if (c < 0) continue;
    T cm[9], dc[9];
    RotConst<T> rc;
    load_camc<T>(d.camc, c, cm, rc);
#pragma unroll
    for (int k = 0; k < 9; ++k) dc[k] = d.dxc[9 * (size_t)c + k];
    const typename Vec2<T>::type ob = __ldg(reinterpret_cast<const typename Vec2<T>::type *>(d.r_obs) + slot);
    T J[24], r0, r1;
    if (PROG) {  // the interpreter's Jacobian and residual of this linearisation are still in the stream
      load_stream<T, 24>(d.WJ, (size_t)wd.row0 + r, lane, J);
      const typename Vec2<T>::type rv = __ldg(reinterpret_cast<const typename Vec2<T>::type *>(d.r_slots) + slot);
      r0 = rv.x; r1 = rv.y;
    } else edge_residual_jacobian<T>(cm, rc, X0, X1, X2, ob.x, ob.y, r0, r1, J);
    T m0 = T(0), m1 = T(0);  // J dx  (JdxpF, lm_algo.cu:61-79)
#pragma unroll
    for (int k = 0; k < 9; ++k) { m0 += J[k] * dc[k]; m1 += J[12 + k] * dc[k]; }
    m0 += J[9] * D0 + J[10] * D1 + J[11] * D2;
    m1 += J[21] * D0 + J[22] * D1 + J[23] * D2;
    a_model += (double)((m0 + r0) * (m0 + r0)) + (double)((m1 + r1) * (m1 + r1));
#pragma unroll
    for (int k = 0; k < 9; ++k) cm[k] += dc[k];  // updateDeltaXTwoVertices, update.cu:14-41
    T s0 = T(0), s1 = T(0);
    if (PROG) {
      if (!d.prog_jit) prog_edge<T, false>(d, c, d.prog_tab2, cm, X0 + D0, X1 + D1, X2 + D2, ob.x, ob.y, s0, s1, (T *)nullptr);
    } else {
      rc = rot_const<T>(cm[0], cm[1], cm[2]);
      edge_residual<T>(cm, rc, X0 + D0, X1 + D1, X2 + D2, ob.x, ob.y, s0, s1);
    }
    a_cost += (double)(s0 * s0) + (double)(s1 * s1);
  }
  a_cost = warp_sum<double>(a_cost);
  a_model = warp_sum<double>(a_model);
  if (lane == 0) {
    if (!(PROG && d.prog_jit)) d.partials[2 * (size_t)w] = a_cost;  // else: written by the generated prog_cost kernel
    d.partials[2 * (size_t)w + 1] = a_model;
  }
}

// --------------------------------------------------------------------------------
// HP1 on the factored stream.  With J~ = [X x Jp_i | Jp_i | B_i (1, l2, l2^2)] (2 x 9) the camera block is Jc = J~ T_c,
// so   Hpp_c = T_c^T (sum_e J~^T Omega J~) T_c   and   g_c = -T_c^T sum_e J~^T Omega r:
// the camera-major pass needs r, Jp and B only (no rotation derivatives), and T_c is applied once per camera.
template <typename T, bool INFO>
__global__ void __launch_bounds__(CM_THREADS, 3) k_lin_cm_fact(Dev<T> d) {
  if (d.scal[S_ANY_SMALL] != 0.) return;  // a small-angle camera: k_lin_cm does the work (device-side dispatch)
  constexpr int NW = CM_THREADS / 32;
  __shared__ T sh[NW][NTRI + 9];
  const int4 ch = d.chunk[blockIdx.x];
  const int c = ch.x, beg = ch.y, end = ch.z;
  T cm[9];
  RotConst<T> rc;
  load_camc_lite<T>(d.camc, c, cm, rc);
  T acc[NTRI + 9];
#pragma unroll
  for (int k = 0; k < NTRI + 9; ++k) acc[k] = T(0);
  for (int k = beg + (int)threadIdx.x; k < end; k += CM_THREADS) {
    const int p = d.cm_pt[k];
    const T X0 = d.pt[3 * (size_t)p], X1 = d.pt[3 * (size_t)p + 1], X2 = d.pt[3 * (size_t)p + 2];
    const typename Vec2<T>::type ob = reinterpret_cast<const typename Vec2<T>::type *>(d.cm_obs)[k];
    T r0, r1, Jp[6], B0, B1;
    edge_point_jacobian<T>(cm, rc, X0, X1, X2, ob.x, ob.y, r0, r1, Jp, B0, B1);
    const T l2 = B0 * B0 + B1 * B1, l4 = l2 * l2;
    T Jt[18];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const T a0 = Jp[3 * i], a1 = Jp[3 * i + 1], a2 = Jp[3 * i + 2], Bi = i == 0 ? B0 : B1;
      Jt[9 * i] = X1 * a2 - X2 * a1;
      Jt[9 * i + 1] = X2 * a0 - X0 * a2;
      Jt[9 * i + 2] = X0 * a1 - X1 * a0;
      Jt[9 * i + 3] = a0; Jt[9 * i + 4] = a1; Jt[9 * i + 5] = a2;
      Jt[9 * i + 6] = Bi; Jt[9 * i + 7] = Bi * l2; Jt[9 * i + 8] = Bi * l4;
    }
    T Lb[18];
    const T *L = Jt;
    if (INFO) {
      T om[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) om[q] = d.cm_info[4 * (size_t)k + q];
#pragma unroll
      for (int y = 0; y < 9; ++y) {
        Lb[y] = Jt[y] * om[0] + Jt[9 + y] * om[2];
        Lb[9 + y] = Jt[y] * om[1] + Jt[9 + y] * om[3];
      }
      L = Lb;
    }
    int idx = 0;
#pragma unroll
    for (int y = 0; y < 9; ++y) {
#pragma unroll
      for (int i = y; i < 9; ++i) acc[idx++] += L[y] * Jt[i] + L[9 + y] * Jt[9 + i];
      acc[NTRI + y] -= L[y] * r0 + L[9 + y] * r1;
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

// element (i, a) of the block-diagonal T_c (blocks Jr, R^T, E; stored Jr, R, E row-major)
template <typename T>
__device__ __forceinline__ T camT_el(const T *tb, int i, int a) {
  const int bk = i / 3;
  if (bk != a / 3) return T(0);
  const int ii = i - 3 * bk, aa = a - 3 * bk;
  return bk == 1 ? tb[9 + 3 * aa + ii] : tb[9 * bk + 3 * ii + aa];
}

// chunk partials (fixed order) -> H~ (symmetric 9x9), g~ -> Hpp = T^T H~ T, g_c = T^T g~
template <typename T>
__global__ void __launch_bounds__(96) k_lin_cam_reduce_fact(Dev<T> d) {
  if (d.scal[S_ANY_SMALL] != 0.) return;
  __shared__ T Hs[81], gs[9], Ts[CAMT];
  const int c = blockIdx.x, k = threadIdx.x;
  if (k < NTRI + 9) {
    T v = T(0);
    for (int ch = d.cam_chunk_ptr[c]; ch < d.cam_chunk_ptr[c + 1]; ++ch) v += d.cm_part[(size_t)ch * (NTRI + 9) + k];
    if (k >= NTRI) {
      gs[k - NTRI] = v;
    } else {
      int y = 0, rem = k;
      while (rem >= 9 - y) { rem -= 9 - y; ++y; }
      const int i = y + rem;
      Hs[9 * y + i] = v;
      Hs[9 * i + y] = v;
    }
  }
  if (k < CAMT) Ts[k] = d.camT[(size_t)CAMT * c + k];
  __syncthreads();
  if (k < 81) {
    int a = k / 9, b = k - 9 * a;
    if (a > b) { const int t = a; a = b; b = t; }  // both mirror entries evaluate the same expression: exact symmetry
    const int ba = 3 * (a / 3), bb = 3 * (b / 3);
    T v = T(0);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      T w = T(0);
#pragma unroll
      for (int j = 0; j < 3; ++j) w += Hs[9 * (ba + i) + bb + j] * camT_el<T>(Ts, bb + j, b);
      v += camT_el<T>(Ts, ba + i, a) * w;
    }
    if (d.cam_fixed && d.cam_fixed[c]) v = (k / 9 == k % 9) ? T(1) : T(0);  // fixed camera: identity block
    d.Hpp[81 * (size_t)c + k] = v;
  } else if (k < 90) {
    const int a = k - 81, ba = 3 * (a / 3);
    T v = T(0);
#pragma unroll
    for (int i = 0; i < 3; ++i) v += camT_el<T>(Ts, ba + i, a) * gs[ba + i];
    d.gc[9 * (size_t)c + a] = v;  // zero for a fixed camera (T_c = 0)
  }
}

// rotation constants of the TRIAL cameras cam + dx_c (what k_trial_fact gathers)
template <typename T>
__global__ void __launch_bounds__(128) k_cam_consts_trial(Dev<T> d) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d.nC) return;
  T cm[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) cm[k] = d.cam[9 * (size_t)c + k] + d.dxc[9 * (size_t)c + k];  // updateDeltaXTwoVertices, update.cu:14-41
  const RotConst<T> rc = rot_const<T>(cm[0], cm[1], cm[2]);
  T *o = d.camc2 + (size_t)CAMC * c;
#pragma unroll
  for (int k = 0; k < 9; ++k) o[k] = cm[k];
  o[9] = rc.c; o[10] = rc.a; o[11] = rc.b; o[12] = rc.ap; o[13] = rc.bp;
  o[14] = rc.small ? T(1) : T(0);
  o[15] = T(0);
}

// LM scalars on the factored stream: J dx = Jp (psi x X + xt + D) + B poly(e) with (psi, xt, e) = T_c dx_c (xpad),
// r from the stored B; the trial residual from the per-camera trial constants.  No Jacobian, no sincos per observation.
template <typename T>
__global__ void __launch_bounds__(ELL_THREADS) k_trial_fact(Dev<T> d) {
  const int lane = threadIdx.x & 31, w = blockIdx.x * ELL_WARPS + (threadIdx.x >> 5);
  if (w >= d.nW) return;
  const WDesc wd = load_desc(d.wdesc, w);
  const int p = wd.heavy ? wd.pt0 : wd.pt0 + lane;
  const bool pvalid = wd.heavy || lane < wd.np;
  T X0 = T(0), X1 = T(0), X2 = T(0), D0 = T(0), D1 = T(0), D2 = T(0);
  if (pvalid) {
    X0 = d.pt[3 * (size_t)p]; X1 = d.pt[3 * (size_t)p + 1]; X2 = d.pt[3 * (size_t)p + 2];
    D0 = d.dxl[3 * (size_t)p]; D1 = d.dxl[3 * (size_t)p + 1]; D2 = d.dxl[3 * (size_t)p + 2];
  }
  double a_model = 0., a_cost = 0.;
#pragma unroll 2
  for (int r = 0; r < wd.D; ++r) {
    const size_t row = (size_t)wd.row0 + r, slot = row * 32 + lane;
    const int c = __ldg(d.r_cam + slot);
    T s[8];
    load_stream<T, 8>(d.WJ, row, lane, s);
    if (c < 0) continue;
    const typename Vec2<T>::type ob = __ldg(reinterpret_cast<const typename Vec2<T>::type *>(d.r_obs) + slot);
    __align__(16) T g[CamPad<T>::N];
    gather_cam<T, true>(d.xpad, c, g);
    const T f = __ldg(d.camc + (size_t)CAMC * c + 6), k1 = __ldg(d.camc + (size_t)CAMC * c + 7), k2 = __ldg(d.camc + (size_t)CAMC * c + 8);
    const T l2 = s[6] * s[6] + s[7] * s[7];
    const T fr = f * (T(1) + k1 * l2 + k2 * l2 * l2);
    const T r0 = fr * s[6] - ob.x, r1 = fr * s[7] - ob.y;  // the residual the linearisation saw (same arithmetic)
    const T q0 = g[1] * X2 - g[2] * X1 + g[3] + D0;
    const T q1 = g[2] * X0 - g[0] * X2 + g[4] + D1;
    const T q2 = g[0] * X1 - g[1] * X0 + g[5] + D2;
    const T poly = g[6] + l2 * (g[7] + l2 * g[8]);
    const T m0 = s[0] * q0 + s[1] * q1 + s[2] * q2 + s[6] * poly;
    const T m1 = s[3] * q0 + s[4] * q1 + s[5] * q2 + s[7] * poly;
    a_model += (double)((m0 + r0) * (m0 + r0)) + (double)((m1 + r1) * (m1 + r1));
    T cm[9];
    RotConst<T> rc;
    load_camc_lite<T>(d.camc2, c, cm, rc);
    T s0, s1;
    edge_residual<T>(cm, rc, X0 + D0, X1 + D1, X2 + D2, ob.x, ob.y, s0, s1);
    a_cost += (double)(s0 * s0) + (double)(s1 * s1);
  }
  a_cost = warp_sum<double>(a_cost);
  a_model = warp_sum<double>(a_model);
  if (lane == 0) {
    d.partials[2 * (size_t)w] = a_cost;
    d.partials[2 * (size_t)w + 1] = a_model;
  }
}

// camera part of the model decrease: dx_c^T Hpp dx_c - 2 g_c^T dx_c (one block, fixed order)
template <typename T>
__global__ void __launch_bounds__(1024) k_model_cam(Dev<T> d) {
  __shared__ double red[32];
  double a = 0.;
  for (int i = threadIdx.x; i < 9 * d.nC; i += blockDim.x) {
    const int c = i / 9, k = i - 9 * c;
    T q = T(0);
#pragma unroll
    for (int j = 0; j < 9; ++j) q += d.Hpp[81 * (size_t)c + 9 * k + j] * d.dxc[9 * c + j];
    a += (double)(d.dxc[i] * q) - 2. * (double)(d.gc[i] * d.dxc[i]);
  }
  a = block_sum<double>(a, red);
  if (threadIdx.x == 0) d.scal[S_MODEL_CAM] = a;
}

// trial cost only: sum r(x + dx)^2 from the per-camera trial constants
template <typename T>
__global__ void __launch_bounds__(ELL_THREADS) k_trial_cost(Dev<T> d) {
  const int lane = threadIdx.x & 31, w = blockIdx.x * ELL_WARPS + (threadIdx.x >> 5);
  if (w >= d.nW) return;
  const WDesc wd = load_desc(d.wdesc, w);
  const int p = wd.heavy ? wd.pt0 : wd.pt0 + lane;
  const bool pvalid = wd.heavy || lane < wd.np;
  T X0 = T(0), X1 = T(0), X2 = T(0);
  if (pvalid) {
    X0 = d.pt[3 * (size_t)p] + d.dxl[3 * (size_t)p];
    X1 = d.pt[3 * (size_t)p + 1] + d.dxl[3 * (size_t)p + 1];
    X2 = d.pt[3 * (size_t)p + 2] + d.dxl[3 * (size_t)p + 2];
  }
  double a_cost = 0.;
#pragma unroll 2
  for (int r = 0; r < wd.D; ++r) {
    const size_t slot = ((size_t)wd.row0 + r) * 32 + lane;
    const int c = __ldg(d.r_cam + slot);
    if (c < 0) continue;
    const typename Vec2<T>::type ob = __ldg(reinterpret_cast<const typename Vec2<T>::type *>(d.r_obs) + slot);
    T cm[9];
    RotConst<T> rc;
    if (BA_HP1_TEX && d.camc2_tex) load_camc_lite_tex(d.camc2_tex, c, cm, rc);
    else load_camc_lite<T>(d.camc2, c, cm, rc);
    T s0, s1;
    edge_residual<T>(cm, rc, X0, X1, X2, ob.x, ob.y, s0, s1);
    a_cost += (double)(s0 * s0) + (double)(s1 * s1);
  }
  a_cost = warp_sum<double>(a_cost);
  if (lane == 0) d.partials[w] = a_cost;
}

// debug / parity: r[slot][2], J[slot][24]
template <typename T, bool PROG>
__global__ void __launch_bounds__(ELL_THREADS) k_eval_edges_ell(Dev<T> d, T *r_out, T *J_out) {
  const int lane = threadIdx.x & 31, w = blockIdx.x * ELL_WARPS + (threadIdx.x >> 5);
  if (w >= d.nW) return;
  const WDesc wd = load_desc(d.wdesc, w);
  const int p = wd.heavy ? wd.pt0 : wd.pt0 + lane;
  const bool pvalid = wd.heavy || lane < wd.np;
  T X0 = T(0), X1 = T(0), X2 = T(0);
  if (pvalid) { X0 = d.pt[3 * (size_t)p]; X1 = d.pt[3 * (size_t)p + 1]; X2 = d.pt[3 * (size_t)p + 2]; }
  for (int r = 0; r < wd.D; ++r) {
    const size_t slot = ((size_t)wd.row0 + r) * 32 + lane;
    const int c = __ldg(d.r_cam + slot);
    if (c < 0) continue;
    T cm[9];
    RotConst<T> rc;
    load_camc<T>(d.camc, c, cm, rc);
    const T o0 = d.r_obs[2 * slot], o1 = d.r_obs[2 * slot + 1];
    T J[24], r0, r1;
    if (PROG) prog_edge<T, true>(d, c, d.prog_tab, cm, X0, X1, X2, o0, o1, r0, r1, J);
    else edge_residual_jacobian<T>(cm, rc, X0, X1, X2, o0, o1, r0, r1, J);
    r_out[2 * slot] = r0;
    r_out[2 * slot + 1] = r1;
    if (J_out) {
#pragma unroll
      for (int k = 0; k < 24; ++k) J_out[24 * slot + k] = J[k];
    }
  }
}

// xpad = T_c x (the vector the factored edge pass gathers)
template <typename T>
__global__ void __launch_bounds__(256) k_pad_cam_T(const T *__restrict__ src, const T *__restrict__ camT, T *dst, int nC) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 9 * nC) return;
  const int c = i / 9, k = i - 9 * c, bk = k / 3, kk = k - 3 * bk;
  const T *tb = camT + (size_t)CAMT * c + 9 * bk;
  const T *x = src + 9 * (size_t)c + 3 * bk;
  // block 1 of T_c is R^T (stored: R)
  const T v = bk == 1 ? tb[kk] * x[0] + tb[3 + kk] * x[1] + tb[6 + kk] * x[2]
                      : tb[3 * kk] * x[0] + tb[3 * kk + 1] * x[1] + tb[3 * kk + 2] * x[2];
  dst[(size_t)CamPad<T>::N * c + k] = v;
}

// dst[i] = src[perm[i]] (gather != 0) or dst[perm[i]] = src[i], three scalars per entry
template <typename T>
__global__ void __launch_bounds__(256) k_permute3(const T *__restrict__ src, T *dst, const int *__restrict__ perm, int n, int gather) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int j = perm[i];
  const size_t a = gather ? (size_t)j : (size_t)i, b = gather ? (size_t)i : (size_t)j;
  dst[3 * b] = src[3 * a];
  dst[3 * b + 1] = src[3 * a + 1];
  dst[3 * b + 2] = src[3 * a + 2];
}

// --------------------------------------------------------------------------------
// launchers
static int ell_grid(int nW) { return (nW + ELL_WARPS - 1) / ELL_WARPS; }

template <typename T>
cudaError_t launch_cam_consts(const Dev<T> &d, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(d.scal + S_ANY_SMALL, 0, sizeof(double), st);
  if (e != cudaSuccess) return e;
  k_cam_consts<T><<<(d.nC + 127) / 128, 128, 0, st>>>(d);
  return cudaGetLastError();
}

template <typename T, int STORE>
static void launch_lin_ell_s(const Dev<T> &d, int dispatch, cudaStream_t st) {
  if (d.r_info) k_lin_ell<T, STORE, true><<<ell_grid(d.nW), ELL_THREADS, 0, st>>>(d, dispatch);
  else k_lin_ell<T, STORE, false><<<ell_grid(d.nW), ELL_THREADS, 0, st>>>(d, dispatch);
}

// d.store is the CONFIGURED store here: ST_FACT launches both point-side flavours (device-side dispatch on the
// small-angle flag written by k_cam_consts), ST_J only the Jacobian flavour.
int ell_hp1_fact() {
  static int v = -1;
  if (v < 0) { const char *e = getenv("BA_HP1_FACT"); v = e ? atoi(e) : 1; }
  return v;
}

template <typename T>
cudaError_t launch_linearize_ell(const Dev<T> &d, cudaStream_t st) {
  if (d.nW > 0) {
    if (d.store == ST_FACT) {
      launch_lin_ell_s<T, ST_FACT>(d, 1, st);
      launch_lin_ell_s<T, ST_J>(d, 2, st);
    } else if (d.prog_code) {  // user-defined residual: the interpreter fills the Jacobian stream
      if (!d.prog_jit) {  // else: the generated prog_fill kernel has done the point side (ImplT::jit_fill_stream)
        launch_prog_cam<T>(d, 0, st);
        if (d.r_info) k_lin_ell<T, ST_J, true, true><<<ell_grid(d.nW), ELL_THREADS, 0, st>>>(d, 0);
        else k_lin_ell<T, ST_J, false, true><<<ell_grid(d.nW), ELL_THREADS, 0, st>>>(d, 0);
      }
    } else {
      launch_lin_ell_s<T, ST_J>(d, 0, st);
    }
  }
  if (d.store == ST_FACT && ell_hp1_fact()) {
    // both camera-major flavours are launched; the small-angle flag decides on the device which one works
    if (d.nChunks > 0) {
      if (d.cm_info) k_lin_cm_fact<T, true><<<d.nChunks, CM_THREADS, 0, st>>>(d);
      else k_lin_cm_fact<T, false><<<d.nChunks, CM_THREADS, 0, st>>>(d);
    }
    k_lin_cam_reduce_fact<T><<<d.nC, 96, 0, st>>>(d);
    cudaError_t e = launch_lin_cm<T>(d, st, 2);
    if (e != cudaSuccess) return e;
  } else {
    cudaError_t e = launch_lin_cm<T>(d, st, 0);
    if (e != cudaSuccess) return e;
  }
  k_reduce<0><<<1, 1024, 0, st>>>(d.partials, d.nW, 1, d.scal + S_COST);
  return cudaGetLastError();
}
int launches_of_linearize_ell() { return 7; }

int ell_rc() {
  static int v = -1;
  if (v < 0) { const char *e = getenv("BA_ELL_RC"); v = e ? atoi(e) : 0; }
  return v;
}

template <typename T, int STORE, int MODE, bool INFO, int RC, int SPLIT = 1>
static cudaError_t launch_schur_ell_v(const Dev<T> &d, const T *xin, T *sout, T *zero_set, int check_active, cudaStream_t st) {
  auto kern = k_schur_ell<T, STORE, MODE, INFO, RC, SPLIT>;
  constexpr int NWC = ell_cta_warps(SPLIT), SPC = NWC / SPLIT;
  constexpr size_t smem = schur_ell_smem<T>(NWC);
  const int grid = (d.nW + SPC - 1) / SPC;
  if (check_active && pdl_enabled() && !d.no_pdl) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(NWC * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, d, xin, sout, zero_set, check_active);
  }
  kern<<<grid, NWC * 32, smem, st>>>(d, xin, sout, zero_set, check_active);
  return cudaGetLastError();
}

template <typename T, int STORE, int MODE, bool INFO>
static cudaError_t launch_schur_ell_rc(const Dev<T> &d, const T *xin, T *sout, T *zero_set, int check_active, cudaStream_t st) {
  if (d.split == 2) return launch_schur_ell_v<T, STORE, MODE, INFO, 0, 2>(d, xin, sout, zero_set, check_active, st);
  if (d.split == 4) return launch_schur_ell_v<T, STORE, MODE, INFO, 0, 4>(d, xin, sout, zero_set, check_active, st);
  if constexpr (STORE != ST_FACT) {
    return launch_schur_ell_v<T, STORE, MODE, INFO, 0>(d, xin, sout, zero_set, check_active, st);
  } else if constexpr (INFO || MODE != SM_FULL) {
    return launch_schur_ell_v<T, STORE, MODE, INFO, 0>(d, xin, sout, zero_set, check_active, st);
  } else {
    switch (ell_rc()) {  // the hot flavour: how many rows stay in registers between the phases is a tuning knob (BA_ELL_RC)
      case 2: return launch_schur_ell_v<T, STORE, MODE, INFO, 2>(d, xin, sout, zero_set, check_active, st);
      case 4: return launch_schur_ell_v<T, STORE, MODE, INFO, 4>(d, xin, sout, zero_set, check_active, st);
      default: return launch_schur_ell_v<T, STORE, MODE, INFO, 0>(d, xin, sout, zero_set, check_active, st);
    }
  }
}

template <typename T, int STORE, int MODE>
static cudaError_t launch_schur_ell_i(const Dev<T> &d, const T *xin, T *sout, T *zero_set, int check_active, cudaStream_t st) {
  if (d.r_info) return launch_schur_ell_rc<T, STORE, MODE, true>(d, xin, sout, zero_set, check_active, st);
  return launch_schur_ell_rc<T, STORE, MODE, false>(d, xin, sout, zero_set, check_active, st);
}

template <typename T, int STORE>
static cudaError_t launch_schur_ell_m(const Dev<T> &d, int mode, const T *xin, T *sout, T *zero_set, int check_active, cudaStream_t st) {
  switch (mode) {
    case SM_FULL: return launch_schur_ell_i<T, STORE, SM_FULL>(d, xin, sout, zero_set, check_active, st);
    case SM_INIT: return launch_schur_ell_i<T, STORE, SM_INIT>(d, xin, sout, zero_set, check_active, st);
    case SM_BACKSUB: return launch_schur_ell_i<T, STORE, SM_BACKSUB>(d, xin, sout, zero_set, check_active, st);
    default: return cudaErrorInvalidValue;
  }
}

// ---- persistent PCG kernel: availability and launch
static int pk_grid(int *sms_out = nullptr) {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (sms_out) *sms_out = sms;
  return sms;
}
template <typename T, bool INFO, int RPT>
static cudaError_t launch_pk_v(const Dev<T> &d, const PkArgs &a, const PeerExchange &px, cudaStream_t st) {
  auto kern = k_pcg_persistent<T, INFO, RPT>;
  const size_t smem = pk_smem<T>(RPT);
  static bool prepared_dev[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!prepared_dev[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    prepared_dev[dev & 63] = true;
  }
  Dev<T> dd = d;
  PkArgs aa = a;
  PeerExchange pxx = px;
  void *args[] = {&dd, &aa, &pxx};
  return cudaLaunchCooperativeKernel((void *)kern, dim3(pk_grid()), dim3(PK_THREADS), args, smem, st);
}
template <typename T>
bool pcg_persistent_available(const Dev<T> &d) {
  if (!d.ell || d.store != ST_FACT || d.no_pdl) return false;  // no_pdl: ranks share the device, a cooperative grid would starve them
  int dev = 0, coop = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
  const int G = pk_grid(&sms);
  if (!coop || G <= 0) return false;
  const int per = (d.nC + G - 1) / G;
  return 9 * per <= 2 * PK_THREADS;
}
template <typename T>
cudaError_t launch_pcg_persistent(const Dev<T> &d, int max_iter, double tol, double refuse, const PeerExchange &px, cudaStream_t st) {
  PkArgs a{max_iter, tol, refuse, d.pk_counters, d.pk_partials, d.pk_timers};
  const int G = pk_grid();
  const int rows = 9 * ((d.nC + G - 1) / G);
  const bool two = rows > PK_THREADS;
  if (d.r_info) return two ? launch_pk_v<T, true, 2>(d, a, px, st) : launch_pk_v<T, true, 1>(d, a, px, st);
  return two ? launch_pk_v<T, false, 2>(d, a, px, st) : launch_pk_v<T, false, 1>(d, a, px, st);
}

int ell_tma() {
  static int v = -1;
  if (v < 0) { const char *e = getenv("BA_TMA"); v = e ? atoi(e) : 0; }
  return v;
}

template <typename T, int MODE, bool INFO>
static cudaError_t launch_schur_tma_v(const Dev<T> &d, const T *xin, T *sout, T *zero_set, int check_active, cudaStream_t st) {
  auto kern = k_schur_tma<T, MODE, INFO>;
  const size_t smem = schur_tma_smem<T>(d.tma_rows);
  static bool prepared_dev[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!prepared_dev[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)schur_tma_smem<T>(64));
    if (e != cudaSuccess) return e;
    prepared_dev[dev & 63] = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(d.nCta);
  cfg.blockDim = dim3(ELL_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (check_active && pdl_enabled() && !d.no_pdl) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, d, xin, sout, zero_set, check_active);
}

template <typename T, int MODE>
static cudaError_t launch_schur_tma_i(const Dev<T> &d, const T *xin, T *sout, T *zero_set, int check_active, cudaStream_t st) {
  if (d.r_info) return launch_schur_tma_v<T, MODE, true>(d, xin, sout, zero_set, check_active, st);
  return launch_schur_tma_v<T, MODE, false>(d, xin, sout, zero_set, check_active, st);
}

// d.store is the EFFECTIVE store of the current linearisation (ST_FACT or ST_J)
template <typename T>
cudaError_t launch_schur_ell(const Dev<T> &d, int mode, const T *xin, T *sout, T *zero_set, int check_active, cudaStream_t st) {
  if (d.nW == 0) return cudaSuccess;
  if (d.store == ST_FACT && ell_tma()) {
    switch (mode) {
      case SM_FULL: return launch_schur_tma_i<T, SM_FULL>(d, xin, sout, zero_set, check_active, st);
      case SM_INIT: return launch_schur_tma_i<T, SM_INIT>(d, xin, sout, zero_set, check_active, st);
      case SM_BACKSUB: return launch_schur_tma_i<T, SM_BACKSUB>(d, xin, sout, zero_set, check_active, st);
      default: return cudaErrorInvalidValue;
    }
  }
  if (d.store == ST_FACT) return launch_schur_ell_m<T, ST_FACT>(d, mode, xin, sout, zero_set, check_active, st);
  return launch_schur_ell_m<T, ST_J>(d, mode, xin, sout, zero_set, check_active, st);
}

template <typename T>
cudaError_t launch_prep_ell(const Dev<T> &d, cudaStream_t st) {
  if (d.nPl > 0) k_prep_points_ell<T><<<(d.nPl + 255) / 256, 256, 0, st>>>(d);
  return cudaGetLastError();
}

template <typename T>
cudaError_t launch_trial_ell(const Dev<T> &d, cudaStream_t st) {
  if (d.store == ST_FACT && ell_hp1_fact()) {  // d.store = the stream in effect; xpad must hold T_c dx_c (host: pad_cam)
    k_cam_consts_trial<T><<<(d.nC + 127) / 128, 128, 0, st>>>(d);
    if (d.nW > 0) k_trial_fact<T><<<ell_grid(d.nW), ELL_THREADS, 0, st>>>(d);
  } else if (d.nW > 0) {
    if (d.prog_code) {
      if (!d.prog_jit) launch_prog_cam<T>(d, 1, st);
      k_trial_ell<T, true><<<ell_grid(d.nW), ELL_THREADS, 0, st>>>(d);
    }
    else k_trial_ell<T, false><<<ell_grid(d.nW), ELL_THREADS, 0, st>>>(d);
  }
  k_reduce<0><<<1, 1024, 0, st>>>(d.partials, d.nW, 2, d.scal + S_TRIAL_COST);
  return cudaGetLastError();
}

// scal[scal_index .. + ncomp) = per-warp partials summed in warp order (the partials were left by a generated kernel)
template <typename T>
cudaError_t launch_reduce_partials(const Dev<T> &d, int ncomp, int scal_index, cudaStream_t st) {
  k_reduce<0><<<1, 1024, 0, st>>>(d.partials, d.nW, ncomp, d.scal + scal_index);
  return cudaGetLastError();
}

// after the back-substitution left the point partials: scal[S_MODEL_PT], scal[S_MODEL_CAM]
template <typename T>
cudaError_t launch_model_terms(const Dev<T> &d, cudaStream_t st) {
  k_reduce<0><<<1, 1024, 0, st>>>(d.model_partials, d.nW, 1, d.scal + S_MODEL_PT);
  k_model_cam<T><<<1, 1024, 0, st>>>(d);
  return cudaGetLastError();
}
template <typename T>
cudaError_t launch_trial_cost_ell(const Dev<T> &d, cudaStream_t st) {
  k_cam_consts_trial<T><<<(d.nC + 127) / 128, 128, 0, st>>>(d);
  if (d.nW > 0) k_trial_cost<T><<<ell_grid(d.nW), ELL_THREADS, 0, st>>>(d);
  k_reduce<0><<<1, 1024, 0, st>>>(d.partials, d.nW, 1, d.scal + S_TRIAL_COST);
  return cudaGetLastError();
}

template <typename T>
cudaError_t launch_eval_edges_ell(const Dev<T> &d, T *r_slots, T *J_slots, cudaStream_t st) {
  if (d.nW > 0) {
    if (d.prog_code) {
      launch_prog_cam<T>(d, 0, st);
      k_eval_edges_ell<T, true><<<ell_grid(d.nW), ELL_THREADS, 0, st>>>(d, r_slots, J_slots);
    }
    else k_eval_edges_ell<T, false><<<ell_grid(d.nW), ELL_THREADS, 0, st>>>(d, r_slots, J_slots);
  }
  return cudaGetLastError();
}

template <typename T>
cudaError_t launch_pad_cam_T(const Dev<T> &d, const T *src, cudaStream_t st) {
  k_pad_cam_T<T><<<(9 * d.nC + 255) / 256, 256, 0, st>>>(src, d.camT, d.xpad, d.nC);
  return cudaGetLastError();
}

template <typename T>
cudaError_t launch_permute3(const T *src, T *dst, const int *perm, int n, int gather, cudaStream_t st) {
  if (n > 0) k_permute3<T><<<(n + 255) / 256, 256, 0, st>>>(src, dst, perm, n, gather);
  return cudaGetLastError();
}

#define BA_INSTANTIATE_ELL(T)                                                                                 \
  template cudaError_t launch_cam_consts<T>(const Dev<T> &, cudaStream_t);                                    \
  template cudaError_t launch_linearize_ell<T>(const Dev<T> &, cudaStream_t);                                 \
  template cudaError_t launch_schur_ell<T>(const Dev<T> &, int, const T *, T *, T *, int, cudaStream_t);      \
  template cudaError_t launch_prep_ell<T>(const Dev<T> &, cudaStream_t);                                      \
  template cudaError_t launch_trial_ell<T>(const Dev<T> &, cudaStream_t);                                     \
  template void launch_prog_cam<T>(const Dev<T> &, int, cudaStream_t);                                        \
  template cudaError_t launch_reduce_partials<T>(const Dev<T> &, int, int, cudaStream_t);                     \
  template cudaError_t launch_trial_cost_ell<T>(const Dev<T> &, cudaStream_t);                                \
  template cudaError_t launch_model_terms<T>(const Dev<T> &, cudaStream_t);                                   \
  template cudaError_t launch_eval_edges_ell<T>(const Dev<T> &, T *, T *, cudaStream_t);                      \
  template cudaError_t launch_pad_cam_T<T>(const Dev<T> &, const T *, cudaStream_t);                          \
  template cudaError_t launch_permute3<T>(const T *, T *, const int *, int, int, cudaStream_t);                \
  template bool pcg_persistent_available<T>(const Dev<T> &);                                                  \
  template cudaError_t launch_pcg_persistent<T>(const Dev<T> &, int, double, double, const PeerExchange &, cudaStream_t);
BA_INSTANTIATE_ELL(double)
BA_INSTANTIATE_ELL(float)

}  // namespace ba
