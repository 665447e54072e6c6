// ba_device_utils.cuh -- small device helpers shared by ba_kernels.cu (tile layout, camera-space PCG kernel)
// and ba_ell.cu (ELL layout).
#pragma once
#include <cuda_runtime.h>

#include "ba_internal.h"

namespace ba {

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum over the block; result valid in every thread.  sm: >= 32 entries.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T *sm) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();  // protect sm from a previous use
  if (lane == 0) sm[wid] = v;
  __syncthreads();
  T x = lane < nw ? sm[lane] : T(0);
  x = warp_sum(x);
  return x;
}

template <typename T>
__device__ __forceinline__ size_t wj_base(int slot, int K) {
  return (size_t)(slot / TE) * (size_t)(K * TE) + (size_t)(slot % TE);
}

template <typename T> struct Vec2;
template <> struct Vec2<double> { using type = double2; };
template <> struct Vec2<float> { using type = float2; };

// Camera-side scatter of one warp's 32 x 9 values.  The values are staged through shared memory so
// that consecutive lanes add consecutive components of one camera (3 RED sectors per observation
// instead of 9), and the target vector exists in NREP replicas selected by CTA so the RED traffic is
// spread over all L2 slices (profiles/r01_microbench_red.txt: 0.47 ms -> 0.087 ms for Venice).
// fire-and-forget reduction: spelled in PTX because the compiler turns an unused-result atomicAdd into the returning
// ATOMG form in some contexts (seen inside the persistent cooperative kernel), which halves the scatter rate
__device__ __forceinline__ void red_add(double *p, double v) { asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory"); }
__device__ __forceinline__ void red_add(float *p, float v) { asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }

template <typename T>
__device__ __forceinline__ void warp_scatter9(T *sv, int *sc, const T *o, int c, T *dst) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  T *mine = sv + (size_t)w * 32 * 9;
  int *cams = sc + w * 32;
  cams[lane] = c;
#pragma unroll
  for (int k = 0; k < 9; ++k) mine[lane * 9 + k] = o[k];
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 9; ++j) {
    const int v = j * 32 + lane, el = v / 9, k = v - 9 * el;
    const int cc = cams[el];
    if (cc >= 0) red_add(dst + 9 * (size_t)cc + k, mine[v]);
  }
  __syncwarp();
}

// Camera-space vectors consumed by the edge kernel are kept in a padded copy (stride CamPad::N scalars,
// 16-byte aligned rows) so that one lane fetches its camera's 9 values with 16-byte loads.
template <typename T, bool VEC>
__device__ __forceinline__ void gather_cam(const T *__restrict__ xpad, int c, T *x) {
  const T *src = xpad + (size_t)CamPad<T>::N * c;
  if (VEC) {
    using V = typename CamPad<T>::vec;
    constexpr int NV = CamPad<T>::N * sizeof(T) / 16;
    V *xv = reinterpret_cast<V *>(x);
#pragma unroll
    for (int k = 0; k < NV; ++k) xv[k] = __ldg(reinterpret_cast<const V *>(src) + k);
  } else {
#pragma unroll
    for (int k = 0; k < 9; ++k) x[k] = __ldg(src + k);
  }
}

template <int N>
__device__ __forceinline__ void cp_async(void *smem_dst, const void *gsrc) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(s), "l"(gsrc), "n"(N) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Deterministic final reduction of per-CTA partials: out[j] = op_i part[i*ncomp + j].
template <int OP>  // 0 sum, 1 max
__global__ void __launch_bounds__(1024) k_reduce(const double *__restrict__ part, int n, int ncomp, double *out) {
  __shared__ double sm[32];
  for (int j = 0; j < ncomp; ++j) {
    double v = 0.;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const double x = part[(size_t)i * ncomp + j];
      v = OP == 0 ? v + x : fmax(v, x);
    }
    if (OP == 0) {
      v = block_sum<double>(v, sm);
    } else {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
      __syncthreads();
      if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
      __syncthreads();
      v = (threadIdx.x & 31) < (blockDim.x >> 5) ? sm[threadIdx.x & 31] : 0.;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    }
    if (threadIdx.x == 0) out[j] = v;
  }
}

}  // namespace ba
