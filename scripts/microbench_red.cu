// microbench_red.cu -- how fast can 5 M observations scatter-add their 9 camera-space values (fp64)?
// Decides the design of the camera-side scatter in k_schur_edge.  Build & run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/microbench_red scripts/microbench_red.cu && build/microbench_red
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

// PAT 0: lane <-> edge, 9 REDs per lane (v0 of k_schur_edge)
__global__ void __launch_bounds__(256) k_direct(const int* __restrict__ cam, double* s, int stride, int nrep, size_t rep_stride, int nE) {
  const int e = blockIdx.x * 256 + threadIdx.x;
  if (e >= nE) return;
  const int c = cam[e];
  double* dst = s + (size_t)(blockIdx.x % nrep) * rep_stride + (size_t)c * stride;
#pragma unroll
  for (int k = 0; k < 9; ++k) atomicAdd(dst + k, 1.0 + k);
}

// PAT 1: values staged through smem so that consecutive lanes add consecutive components of one camera
__global__ void __launch_bounds__(256) k_remap(const int* __restrict__ cam, double* s, int stride, int nrep, size_t rep_stride, int nE) {
  __shared__ double sv[256 * 9];
  __shared__ int sc[256];
  const int e = blockIdx.x * 256 + threadIdx.x;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  sc[threadIdx.x] = e < nE ? cam[e] : -1;
#pragma unroll
  for (int k = 0; k < 9; ++k) sv[(w * 32 + lane) * 9 + k] = 1.0 + k;
  __syncwarp();
  double* base = s + (size_t)(blockIdx.x % nrep) * rep_stride;
#pragma unroll
  for (int j = 0; j < 9; ++j) {
    const int v = j * 32 + lane, el = v / 9, k = v - 9 * el;
    const int c = sc[w * 32 + el];
    if (c >= 0) atomicAdd(base + (size_t)c * stride + k, sv[(w * 32) * 9 + v]);
  }
}

// PAT 2: whole camera vector privatised in shared memory (fp64 smem atomics = CAS loop), persistent CTAs
__global__ void __launch_bounds__(1024) k_smem(const int* __restrict__ cam, double* s, int nC9, int nE) {
  extern __shared__ double sh[];
  for (int i = threadIdx.x; i < nC9; i += blockDim.x) sh[i] = 0.;
  __syncthreads();
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nE; e += gridDim.x * blockDim.x) {
    const int c = cam[e];
#pragma unroll
    for (int k = 0; k < 9; ++k) atomicAdd(sh + 9 * c + k, 1.0 + k);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nC9; i += blockDim.x) atomicAdd(s + i, sh[i]);
}

// PAT 3: no atomics at all -- plain coalesced store of the 9 values (lower bound for "write c_e, reduce later")
__global__ void __launch_bounds__(256) k_store(const int* __restrict__ cam, double* out, int nE) {
  const int e = blockIdx.x * 256 + threadIdx.x;
  if (e >= nE) return;
  const int c = cam[e];
#pragma unroll
  for (int k = 0; k < 9; ++k) out[(size_t)k * nE + e] = 1.0 + k + c;
}

template <typename F> float timeit(F f, int reps = 5) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int i = 0; i < reps; ++i) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); best = std::min(best, ms); }
  return best;
}

int main() {
  const int nC = 1778, nE = 5001946;
  for (int dist = 0; dist < 2; ++dist) {
    std::mt19937_64 rng(42);
    std::vector<double> w(nC);
    std::lognormal_distribution<double> ln(0.0, 0.5);
    for (auto& x : w) x = dist == 0 ? ln(rng) : 1.0;
    std::discrete_distribution<int> dd(w.begin(), w.end());
    std::vector<int> cam(nE);
    for (auto& c : cam) c = dd(rng);
    int* d_cam; CK(cudaMalloc(&d_cam, sizeof(int) * nE));
    CK(cudaMemcpy(d_cam, cam.data(), sizeof(int) * nE, cudaMemcpyHostToDevice));
    const size_t maxbuf = (size_t)64 * nC * 16;
    double* d_s; CK(cudaMalloc(&d_s, sizeof(double) * maxbuf)); CK(cudaMemset(d_s, 0, sizeof(double) * maxbuf));
    double* d_out; CK(cudaMalloc(&d_out, sizeof(double) * 9 * (size_t)nE));
    const int grid = (nE + 255) / 256;
    printf("== camera popularity: %s   (nC=%d nE=%d; 45.0 M fp64 adds)\n", dist == 0 ? "lognormal(0,0.5)" : "uniform", nC, nE);
    for (int stride : {9, 12, 16}) for (int nrep : {1, 4, 16, 64}) {
      const size_t rs = (size_t)nC * stride;
      float t0 = timeit([&] { k_direct<<<grid, 256>>>(d_cam, d_s, stride, nrep, rs, nE); });
      float t1 = timeit([&] { k_remap<<<grid, 256>>>(d_cam, d_s, stride, nrep, rs, nE); });
      printf("stride %2d replicas %2d : direct %.3f ms   lane-remapped %.3f ms\n", stride, nrep, t0, t1);
    }
    CK(cudaFuncSetAttribute(k_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, nC * 9 * 8));
    for (int g : {148, 296}) {
      float t2 = timeit([&] { k_smem<<<g, 1024, nC * 9 * 8>>>(d_cam, d_s, nC * 9, nE); });
      printf("smem-privatised (grid %d x 1024, %d KB smem): %.3f ms\n", g, nC * 9 * 8 / 1024, t2);
    }
    float t3 = timeit([&] { k_store<<<grid, 256>>>(d_cam, d_out, nE); });
    printf("plain SoA store of 9 values/edge (360 MB): %.3f ms\n", t3);
    CK(cudaGetLastError());
    cudaFree(d_cam); cudaFree(d_s); cudaFree(d_out);
  }
  return 0;
}
