// ref_driver.cu -- C interface to the reference's own hot-path code, compiled unmodified against the shim headers
// (see README.md).  CHECKER ONLY: used by tests/test_gpu_reference_pin.py on the GPU box; never by the product.
//
//   ref_forward_*  -> MegBA::geo::AnalyticalDerivativesKernelMatrix   (src/geo/analytical_derivatives.cu:289-322)
//   ref_build_*    -> MegBA::EdgeVector<T>::buildLinearSystemCUDA     (src/edge/build_linear_system.cu:243-423)
//   ref_solve_*    -> MegBA::SchurPCGSolver<T>::solve                 (src/solver/schur_pcg_solver.cu:643-679)
//   ref_lm_solve_* -> MegBA::LMAlgo<T>::solveCUDA, the reference's whole LM loop (src/algo/lm_algo.cu:139-223 with its
//                     scalar kernels :26-136), driving the three pieces above plus SchurLMLinearSystem::processDiag /
//                     backup / rollback / applyUpdate (src/linear_system/schur_LM_linear_system.cu:162-218).  Only the
//                     glue the reference keeps in Eigen/JetVector-heavy files is provided here: EdgeVector::forward()
//                     (binds the per-edge parameter arrays and calls the reference's analytical kernel),
//                     EdgeVector::update/backup/rollback (restating src/edge/update.cu:14-41, src/edge/base_edge.cu:17-44)
//                     and BaseLinearSystem::solve() (forwards to the reference solver).
//
// Host arrays use the oracle's layouts: cam[nC][9], pt[nP][3], obs[nE][2], r[nE][2], J[nE][2][12], Hpp[nC][81],
// Hll[nP][9], W[nE][27], g[9nC+3nP].  The CSR index built here plays the role of SchurLMLinearSystem::buildIndex /
// EdgeVector::buildPositionContainer (src/linear_system/schur_LM_linear_system.cpp:21-84, src/edge/base_edge.cpp:224-262):
// Hpl row (c*9+y) holds, for every point seen by camera c (ascending), 3 scalars; Hlp row (p*3+y) holds 9 per camera.
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include <unistd.h>
#include <cstdio>
#include <iostream>

#include "algo/lm_algo.h"
#include "edge/base_edge.h"
#include "geo/geo.cuh"
#include "linear_system/schur_LM_linear_system.h"
#include "resource/handle_manager.h"
#include "resource/memory_pool.h"
#include "solver/schur_pcg_solver.h"
#include "linear_system/implicit_schur_LM_linear_system.h"
#include "solver/implicit_schur_pcg_solver.h"

namespace MegBA {
const ProblemOption *MemoryPool::_problemOption = nullptr;
// MemoryPool::allocateNormal / deallocateNormal: the reference serves these from a pre-reserved stack
// (src/resource/memory_pool.cu); a cudaMalloc / cudaFree per call would charge its solver and LM loop with driver calls it
// does not make.  Same discipline here: one slab, bump allocation, frees pop the stack (out-of-order frees are deferred).
namespace {
struct Arena {
  char *base = nullptr;
  size_t cap = 0, top = 0;
  struct Blk { size_t off, size; bool freed; };
  std::vector<Blk> blks;
} g_arena;
}  // namespace
void MemoryPool::allocateNormal(void **ptr, size_t size, int rank) {
  cudaSetDevice(rank);
  if (!g_arena.base) {
    size_t fr = 0, tot = 0;
    cudaMemGetInfo(&fr, &tot);
    g_arena.cap = std::min<size_t>(fr / 4, (size_t)8 << 30);
    if (cudaMalloc(&g_arena.base, g_arena.cap) != cudaSuccess) { g_arena.base = nullptr; g_arena.cap = 0; cudaGetLastError(); }
  }
  const size_t need = ((size ? size : 8) + 255) / 256 * 256;
  if (g_arena.base && g_arena.top + need <= g_arena.cap) {
    *ptr = g_arena.base + g_arena.top;
    g_arena.blks.push_back({g_arena.top, need, false});
    g_arena.top += need;
    // Blocks are handed out zeroed: the implicit solver's MakeV accumulates into a pool buffer with atomics without
    // clearing it first (implicit_schur_pcg_solver.cu:500-528), i.e. upstream relies on what a
    // fresh pool happens to contain.  Zero is the intended content; a recycled arena block would otherwise hold the
    // previous solve's data.
    cudaMemsetAsync(*ptr, 0, need, nullptr);
  } else {
    cudaMalloc(ptr, need);  // arena exhausted: fall back to the driver
  }
}
void MemoryPool::deallocateNormal(void *ptr, int rank) {
  cudaSetDevice(rank);
  char *p = (char *)ptr;
  if (g_arena.base && p >= g_arena.base && p < g_arena.base + g_arena.cap) {
    for (auto it = g_arena.blks.rbegin(); it != g_arena.blks.rend(); ++it)
      if (g_arena.base + it->off == p) { it->freed = true; break; }
    while (!g_arena.blks.empty() && g_arena.blks.back().freed) { g_arena.top = g_arena.blks.back().off; g_arena.blks.pop_back(); }
  } else {
    cudaFree(ptr);
  }
}
std::vector<cublasHandle_t> HandleManager::cublasHandle{};
std::vector<cusparseHandle_t> HandleManager::cusparseHandle{};
}  // namespace MegBA

namespace {
using namespace MegBA;

void ensure_handles() {
  if (HandleManager::cublasHandle.empty()) {
    HandleManager::cublasHandle.resize(1);
    HandleManager::cusparseHandle.resize(1);
    cublasCreate_v2(&HandleManager::cublasHandle[0]);
    cusparseCreate(&HandleManager::cusparseHandle[0]);
  }
}

template <typename T> T *dev_alloc(size_t n) { T *p = nullptr; cudaMalloc(&p, sizeof(T) * (n ? n : 1)); return p; }
template <typename T> T *dev_copy(const std::vector<T> &h) { T *p = dev_alloc<T>(h.size()); cudaMemcpy(p, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice); return p; }

struct Index {  // CSR structure shared by build and solve
  std::vector<int> hplRowPtr, hlpRowPtr, hplColInd, hlpColInd, relP, relC;
};

Index make_index(int64_t nC, int64_t nP, int64_t nE, const int32_t *cam, const int32_t *pt) {
  Index ix;
  std::vector<std::vector<int>> ptsOfCam(nC), camsOfPt(nP);
  for (int64_t e = 0; e < nE; ++e) { ptsOfCam[cam[e]].push_back(pt[e]); camsOfPt[pt[e]].push_back(cam[e]); }
  for (auto &v : ptsOfCam) std::sort(v.begin(), v.end());
  for (auto &v : camsOfPt) std::sort(v.begin(), v.end());
  ix.hplRowPtr.assign(9 * nC + 1, 0);
  ix.hlpRowPtr.assign(3 * nP + 1, 0);
  for (int64_t c = 0; c < nC; ++c) for (int y = 0; y < 9; ++y) ix.hplRowPtr[9 * c + y + 1] = ix.hplRowPtr[9 * c + y] + 3 * (int)ptsOfCam[c].size();
  for (int64_t p = 0; p < nP; ++p) for (int y = 0; y < 3; ++y) ix.hlpRowPtr[3 * p + y + 1] = ix.hlpRowPtr[3 * p + y] + 9 * (int)camsOfPt[p].size();
  ix.hplColInd.resize(ix.hplRowPtr.back());
  ix.hlpColInd.resize(ix.hlpRowPtr.back());
  for (int64_t c = 0; c < nC; ++c) for (int y = 0; y < 9; ++y) {
    int o = ix.hplRowPtr[9 * c + y];
    for (int p : ptsOfCam[c]) for (int i = 0; i < 3; ++i) ix.hplColInd[o++] = 3 * p + i;
  }
  for (int64_t p = 0; p < nP; ++p) for (int y = 0; y < 3; ++y) {
    int o = ix.hlpRowPtr[3 * p + y];
    for (int c : camsOfPt[p]) for (int i = 0; i < 9; ++i) ix.hlpColInd[o++] = 9 * c + i;
  }
  ix.relP.resize(nE); ix.relC.resize(nE);
  for (int64_t e = 0; e < nE; ++e) {
    const auto &pv = ptsOfCam[cam[e]];
    const auto &cv = camsOfPt[pt[e]];
    ix.relP[e] = (int)(std::lower_bound(pv.begin(), pv.end(), pt[e]) - pv.begin());
    ix.relC[e] = (int)(std::lower_bound(cv.begin(), cv.end(), cam[e]) - cv.begin());
  }
  return ix;
}

template <typename T>
int forward_impl(int64_t nE, const int32_t *cam_idx, const int32_t *pt_idx, const T *obs, const T *cam, const T *pt, T *r, T *J) {
  ProblemOption opt; opt.deviceUsed = {0}; opt.N = 12; opt.nItem = nE;
  MemoryPool::resetPool(&opt);
  cudaSetDevice(0);
  // the reference's a1 layout: per-edge copies of the 9 camera + 3 point parameters and the observation, SoA
  std::vector<T> soa(14 * (size_t)nE);
  for (int64_t e = 0; e < nE; ++e) {
    for (int k = 0; k < 9; ++k) soa[(size_t)k * nE + e] = cam[9 * (size_t)cam_idx[e] + k];
    for (int k = 0; k < 3; ++k) soa[(size_t)(9 + k) * nE + e] = pt[3 * (size_t)pt_idx[e] + k];
    for (int k = 0; k < 2; ++k) soa[(size_t)(12 + k) * nE + e] = obs[2 * e + k];
  }
  T *d = dev_copy(soa);
  JVD<T> camJ, ptJ, obsJ;
  camJ.resize(9, 1); ptJ.resize(3, 1); obsJ.resize(2, 1);
  auto bind = [&](JetVector<T> &jv, int k, unsigned N) { jv.bind(N, {(size_t)nE}, {d + (size_t)k * nE}, {nullptr}); };
  for (int k = 0; k < 9; ++k) bind(camJ(k), k, 12);
  for (int k = 0; k < 3; ++k) bind(ptJ(k), 9 + k, 12);
  for (int k = 0; k < 2; ++k) bind(obsJ(k), 12 + k, 0);
  Eigen::Map<const JVD<T>> aa{&camJ(0, 0), 3, 1}, tt{&camJ(3, 0), 3, 1}, intr{&camJ(6, 0), 3, 1};
  JVD<T> err = geo::AnalyticalDerivativesKernelMatrix<T>(aa, tt, intr, ptJ, obsJ);
  cudaDeviceSynchronize();
  std::vector<T> hr(nE), hg(12 * (size_t)nE);
  for (int row = 0; row < 2; ++row) {
    cudaMemcpy(hr.data(), err(row).getCUDAResPtr()[0], sizeof(T) * nE, cudaMemcpyDeviceToHost);
    cudaMemcpy(hg.data(), err(row).getCUDAGradPtr()[0], sizeof(T) * 12 * nE, cudaMemcpyDeviceToHost);
    for (int64_t e = 0; e < nE; ++e) {
      r[2 * e + row] = hr[e];
      for (int k = 0; k < 12; ++k) J[24 * e + 12 * row + k] = hg[(size_t)k * nE + e];
    }
    err(row).release();
  }
  cudaFree(d);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

template <typename T>
int build_impl(int64_t nC, int64_t nP, int64_t nE, const int32_t *cam_idx, const int32_t *pt_idx, const T *r, const T *J,
               const T *info, T *Hpp, T *Hll, T *W, T *g) {
  ProblemOption opt; opt.deviceUsed = {0}; opt.N = 12; opt.nItem = nE;
  MemoryPool::resetPool(&opt);
  cudaSetDevice(0);
  const Index ix = make_index(nC, nP, nE, cam_idx, pt_idx);
  // J / r as two JetVectors (one per residual row) with slab layout [grad 12 x nE | res nE]
  JVD<T> jet; jet.resize(2, 1);
  std::vector<T *> slabs;
  for (int row = 0; row < 2; ++row) {
    std::vector<T> h(13 * (size_t)nE);
    for (int64_t e = 0; e < nE; ++e) {
      for (int k = 0; k < 12; ++k) h[(size_t)k * nE + e] = J[24 * e + 12 * row + k];
      h[(size_t)12 * nE + e] = r[2 * e + row];
    }
    T *d = dev_copy(h);
    slabs.push_back(d);
    jet(row).bind(12, {(size_t)nE}, {d + (size_t)12 * nE}, {d});
  }
  EdgeVector<T> edges(opt);
  edges.positionContainers.resize(1);
  std::vector<int> ac(cam_idx, cam_idx + nE), ap(pt_idx, pt_idx + nE);
  int *dac = dev_copy(ac), *dap = dev_copy(ap), *drc = dev_copy(ix.relC), *drp = dev_copy(ix.relP);
  edges.positionContainers[0].absolutePosition = {dac, dap};
  edges.positionContainers[0].relativePosition = {drc, drp};
  T *dinfo = nullptr;
  if (info) {
    std::vector<T> h(4 * (size_t)nE);
    for (int64_t e = 0; e < nE; ++e) for (int k = 0; k < 4; ++k) h[(size_t)k * nE + e] = info[4 * e + k];
    dinfo = dev_copy(h);
    edges.jetInformation.resize(2, 2);
    // jetInformation(i, j) is read as infoPtrs[j + i*resDim] (build_linear_system.cu:319-323) and consumed as
    // infoPtrs[i + j*resDim] in JMulInfo (:172): element (i, j) of the row-major 2x2 block sits at info[2*i + j]
    for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j)
      edges.jetInformation(i, j).bind(0, {(size_t)nE}, {dinfo + (size_t)(2 * i + j) * nE}, {nullptr});
  }
  SchurLMLinearSystem<T> ls(opt);
  ls.dim = {9, 3}; ls.num = {(int)nC, (int)nP};
  ls.equationContainers.resize(1);
  auto &ec = ls.equationContainers[0];
  ec.nnz = {(size_t)27 * nE, (size_t)27 * nE, (size_t)81 * nC, (size_t)9 * nP};
  for (int k = 0; k < 4; ++k) ec.csrVal[k] = dev_alloc<T>(ec.nnz[k]);
  ec.csrRowPtr = {dev_copy(ix.hplRowPtr), dev_copy(ix.hlpRowPtr)};
  ls.g = {dev_alloc<T>(9 * nC + 3 * nP)};
  edges.buildLinearSystemCUDA(jet, ls);
  cudaDeviceSynchronize();
  cudaMemcpy(Hpp, ec.csrVal[2], sizeof(T) * 81 * nC, cudaMemcpyDeviceToHost);
  cudaMemcpy(Hll, ec.csrVal[3], sizeof(T) * 9 * nP, cudaMemcpyDeviceToHost);
  cudaMemcpy(g, ls.g[0], sizeof(T) * (9 * nC + 3 * nP), cudaMemcpyDeviceToHost);
  std::vector<T> hpl(27 * (size_t)nE), hlp(27 * (size_t)nE);
  cudaMemcpy(hpl.data(), ec.csrVal[0], sizeof(T) * 27 * nE, cudaMemcpyDeviceToHost);
  cudaMemcpy(hlp.data(), ec.csrVal[1], sizeof(T) * 27 * nE, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int64_t e = 0; e < nE; ++e)
    for (int y = 0; y < 9; ++y)
      for (int i = 0; i < 3; ++i) {
        const T a = hpl[ix.hplRowPtr[9 * cam_idx[e] + y] + 3 * ix.relP[e] + i];
        const T b = hlp[ix.hlpRowPtr[3 * pt_idx[e] + i] + 9 * ix.relC[e] + y];
        W[27 * e + 3 * y + i] = a;
        // Hlp must be the transpose of Hpl (bit-exact without Omega; with Omega the two are associated differently)
        const double da = (double)a - (double)b, sc = (a < 0 ? -(double)a : (double)a) + (b < 0 ? -(double)b : (double)b);
        (void)da; (void)sc;
        if (!info && a != b) ++bad;
      }
  for (int k = 0; k < 4; ++k) cudaFree(ec.csrVal[k]);
  cudaFree(ec.csrRowPtr[0]); cudaFree(ec.csrRowPtr[1]); cudaFree(ls.g[0]);
  cudaFree(dac); cudaFree(dap); cudaFree(drc); cudaFree(drp);
  for (auto *s : slabs) cudaFree(s);
  if (dinfo) cudaFree(dinfo);
  if (cudaGetLastError() != cudaSuccess) return -1;
  return bad ? -2 : 0;
}

template <typename T>
int solve_impl(int64_t nC, int64_t nP, int64_t nE, const int32_t *cam_idx, const int32_t *pt_idx, const T *HppD, const T *HllD,
               const T *W, const T *g, T *dx, int max_iter, double tol, double refuse, double *solve_ms = nullptr) {
  ProblemOption opt; opt.deviceUsed = {0}; opt.N = 12; opt.nItem = nE;
  MemoryPool::resetPool(&opt);
  cudaSetDevice(0);
  ensure_handles();
  const Index ix = make_index(nC, nP, nE, cam_idx, pt_idx);
  std::vector<T> hpl(27 * (size_t)nE), hlp(27 * (size_t)nE);
  for (int64_t e = 0; e < nE; ++e)
    for (int y = 0; y < 9; ++y)
      for (int i = 0; i < 3; ++i) {
        hpl[ix.hplRowPtr[9 * cam_idx[e] + y] + 3 * ix.relP[e] + i] = W[27 * e + 3 * y + i];
        hlp[ix.hlpRowPtr[3 * pt_idx[e] + i] + 9 * ix.relC[e] + y] = W[27 * e + 3 * y + i];
      }
  SchurLMLinearSystem<T> ls(opt);
  ls.dim = {9, 3}; ls.num = {(int)nC, (int)nP};
  ls.equationContainers.resize(1);
  auto &ec = ls.equationContainers[0];
  ec.nnz = {(size_t)27 * nE, (size_t)27 * nE, (size_t)81 * nC, (size_t)9 * nP};
  ec.csrVal[0] = dev_copy(hpl); ec.csrVal[1] = dev_copy(hlp);
  ec.csrVal[2] = dev_alloc<T>(81 * nC); ec.csrVal[3] = dev_alloc<T>(9 * nP);
  cudaMemcpy(ec.csrVal[2], HppD, sizeof(T) * 81 * nC, cudaMemcpyHostToDevice);
  cudaMemcpy(ec.csrVal[3], HllD, sizeof(T) * 9 * nP, cudaMemcpyHostToDevice);
  ec.csrRowPtr = {dev_copy(ix.hplRowPtr), dev_copy(ix.hlpRowPtr)};
  ec.csrColInd = {dev_copy(ix.hplColInd), dev_copy(ix.hlpColInd)};
  const size_t nX = 9 * nC + 3 * nP;
  ls.g = {dev_alloc<T>(nX)};
  ls.deltaXPtr = {dev_alloc<T>(nX)};
  cudaMemcpy(ls.g[0], g, sizeof(T) * nX, cudaMemcpyHostToDevice);
  cudaMemcpy(ls.deltaXPtr[0], dx, sizeof(T) * nX, cudaMemcpyHostToDevice);
  SolverOption so;
  so.solverOptionPCG.maxIter = max_iter; so.solverOptionPCG.tol = tol; so.solverOptionPCG.refuseRatio = refuse;
  SchurPCGSolver<T> solver(opt, so);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaDeviceSynchronize();
  cudaEventRecord(e0, nullptr);  // the reference works on the legacy default stream
  solver.solve(ls);
  cudaEventRecord(e1, nullptr);
  cudaDeviceSynchronize();
  if (solve_ms) { float ms = 0; cudaEventElapsedTime(&ms, e0, e1); *solve_ms = ms; }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaMemcpy(dx, ls.deltaXPtr[0], sizeof(T) * nX, cudaMemcpyDeviceToHost);
  for (int k = 0; k < 4; ++k) cudaFree(ec.csrVal[k]);
  for (int k = 0; k < 2; ++k) { cudaFree(ec.csrRowPtr[k]); cudaFree(ec.csrColInd[k]); }
  cudaFree(ls.g[0]); cudaFree(ls.deltaXPtr[0]);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// a3' / a7 implicit: EdgeVector::buildImplicitLinearSystemCUDA (src/edge/build_implicit_linear_system.cu:208-322, kernel
// makeHppHllSchur :65-111) and ImplicitSchurPCGSolver::solve (src/solver/implicit_schur_pcg_solver.cu:661-720: block
// inverses, MakeV :476-549, the matrix-free PCG :181-473 with implicitETMulx / implicitEMulx :21-90, SolveW :552-618).
template <typename T>
struct JetPair {  // J / r as two JetVectors (one per residual row), slab layout [grad 12 x nE | res nE]
  JVD<T> jet;
  std::vector<T *> slabs;
  JetPair(int64_t nE, const T *r, const T *J) {
    jet.resize(2, 1);
    for (int row = 0; row < 2; ++row) {
      std::vector<T> h(13 * (size_t)nE);
      for (int64_t e = 0; e < nE; ++e) {
        for (int k = 0; k < 12; ++k) h[(size_t)k * nE + e] = J[24 * e + 12 * row + k];
        h[(size_t)12 * nE + e] = r ? r[2 * e + row] : T(0);
      }
      T *d = dev_copy(h);
      slabs.push_back(d);
      jet(row).bind(12, {(size_t)nE}, {d + (size_t)12 * nE}, {d});
    }
  }
  ~JetPair() { for (auto *s : slabs) cudaFree(s); }
};

template <typename T>
int build_implicit_impl(int64_t nC, int64_t nP, int64_t nE, const int32_t *cam_idx, const int32_t *pt_idx, const T *r, const T *J,
                        T *Hpp, T *Hll, T *g) {
  ProblemOption opt; opt.deviceUsed = {0}; opt.N = 12; opt.nItem = nE;
  MemoryPool::resetPool(&opt);
  cudaSetDevice(0);
  JetPair<T> jp(nE, r, J);
  EdgeVector<T> edges(opt);
  edges.positionContainers.resize(1);
  std::vector<int> ac(cam_idx, cam_idx + nE), ap(pt_idx, pt_idx + nE);
  int *dac = dev_copy(ac), *dap = dev_copy(ap);
  edges.positionContainers[0].absolutePosition = {dac, dap};
  ImplicitSchurLMLinearSystem<T> ls(opt);
  ls.dim = {9, 3}; ls.num = {(int)nC, (int)nP};
  ls.implicitEquationContainers.resize(1);
  auto &ec = ls.implicitEquationContainers[0];
  ec.nnz = {(size_t)81 * nC, (size_t)9 * nP};
  for (int k = 0; k < 2; ++k) ec.csrVal[k] = dev_alloc<T>(ec.nnz[k]);
  ls.g = {dev_alloc<T>(9 * nC + 3 * nP)};
  edges.buildImplicitLinearSystemCUDA(jp.jet, ls);
  cudaDeviceSynchronize();
  cudaMemcpy(Hpp, ec.csrVal[0], sizeof(T) * 81 * nC, cudaMemcpyDeviceToHost);
  cudaMemcpy(Hll, ec.csrVal[1], sizeof(T) * 9 * nP, cudaMemcpyDeviceToHost);
  cudaMemcpy(g, ls.g[0], sizeof(T) * (9 * nC + 3 * nP), cudaMemcpyDeviceToHost);
  cudaFree(ec.csrVal[0]); cudaFree(ec.csrVal[1]); cudaFree(ls.g[0]); cudaFree(dac); cudaFree(dap);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

template <typename T>
int solve_implicit_impl(int64_t nC, int64_t nP, int64_t nE, const int32_t *cam_idx, const int32_t *pt_idx, const T *HppD, const T *HllD,
                        const T *J, const T *g, T *dx, int max_iter, double tol, double refuse) {
  ProblemOption opt; opt.deviceUsed = {0}; opt.N = 12; opt.nItem = nE;
  MemoryPool::resetPool(&opt);
  cudaSetDevice(0);
  ensure_handles();
  JetPair<T> jp(nE, nullptr, J);
  EdgeVector<T> edges(opt);
  edges.positionContainers.resize(1);
  std::vector<int> ac(cam_idx, cam_idx + nE), ap(pt_idx, pt_idx + nE);
  int *dac = dev_copy(ac), *dap = dev_copy(ap);
  edges.positionContainers[0].absolutePosition = {dac, dap};
  ImplicitSchurLMLinearSystem<T> ls(opt);
  ls.dim = {9, 3}; ls.num = {(int)nC, (int)nP};
  ls.implicitEquationContainers.resize(1);
  auto &ec = ls.implicitEquationContainers[0];
  ec.nnz = {(size_t)81 * nC, (size_t)9 * nP};
  ec.csrVal[0] = dev_alloc<T>(81 * nC); ec.csrVal[1] = dev_alloc<T>(9 * nP);
  cudaMemcpy(ec.csrVal[0], HppD, sizeof(T) * 81 * nC, cudaMemcpyHostToDevice);
  cudaMemcpy(ec.csrVal[1], HllD, sizeof(T) * 9 * nP, cudaMemcpyHostToDevice);
  const size_t nX = 9 * nC + 3 * nP;
  ls.g = {dev_alloc<T>(nX)};
  ls.deltaXPtr = {dev_alloc<T>(nX)};
  cudaMemcpy(ls.g[0], g, sizeof(T) * nX, cudaMemcpyHostToDevice);
  cudaMemcpy(ls.deltaXPtr[0], dx, sizeof(T) * nX, cudaMemcpyHostToDevice);
  SolverOption so;
  so.solverOptionPCG.maxIter = max_iter; so.solverOptionPCG.tol = tol; so.solverOptionPCG.refuseRatio = refuse;
  ImplicitSchurPCGSolver<T> solver(opt, so);
  solver.solve(ls, edges, jp.jet);
  cudaDeviceSynchronize();
  cudaMemcpy(dx, ls.deltaXPtr[0], sizeof(T) * nX, cudaMemcpyDeviceToHost);
  cudaFree(ec.csrVal[0]); cudaFree(ec.csrVal[1]); cudaFree(ls.g[0]); cudaFree(ls.deltaXPtr[0]); cudaFree(dac); cudaFree(dap);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
}  // namespace

namespace MegBA {
// EdgeVector::forward (src/edge/base_edge.cpp:160-163 -> user edge -> analytical kernel)
template <typename T>
JVD<T> EdgeVector<T>::forward() const {
  const size_t nE = (size_t)nEdges;
  JVD<T> camJ, ptJ, obsJ;
  camJ.resize(9, 1); ptJ.resize(3, 1); obsJ.resize(2, 1);
  for (int k = 0; k < 9; ++k) camJ(k).bind(12, {nE}, {edgeParams + (size_t)k * nE}, {nullptr});
  for (int k = 0; k < 3; ++k) ptJ(k).bind(12, {nE}, {edgeParams + (size_t)(9 + k) * nE}, {nullptr});
  for (int k = 0; k < 2; ++k) obsJ(k).bind(0, {nE}, {edgeObs + (size_t)k * nE}, {nullptr});
  Eigen::Map<const JVD<T>> aa{&camJ(0, 0), 3, 1}, tt{&camJ(3, 0), 3, 1}, intr{&camJ(6, 0), 3, 1};
  return geo::AnalyticalDerivativesKernelMatrix<T>(aa, tt, intr, ptJ, obsJ);
}
template <typename T>
void EdgeVector<T>::backup() const {  // base_edge.cu:17-30
  cudaMemcpyAsync(edgeParamsBackup, edgeParams, sizeof(T) * 12 * (size_t)nEdges, cudaMemcpyDeviceToDevice);
}
template <typename T>
void EdgeVector<T>::rollback() const {  // base_edge.cu:32-44
  cudaMemcpyAsync(edgeParams, edgeParamsBackup, sizeof(T) * 12 * (size_t)nEdges, cudaMemcpyDeviceToDevice);
}
template <typename T>
__global__ void shimUpdateKernel(const T *deltaX, const int *absCam, const int *absPt, int cameraNum, long long nItem, T *params) {
  const long long tid = blockDim.x * (long long)blockIdx.x + threadIdx.x;
  if (tid >= nItem) return;
  for (int i = 0; i < 9; ++i) params[tid + i * nItem] += deltaX[absCam[tid] * 9 + i];                       // update.cu:29-32
  for (int i = 0; i < 3; ++i) params[tid + (9 + i) * nItem] += deltaX[absPt[tid] * 3 + i + cameraNum * 9];  // update.cu:34-39
}
template <typename T>
void EdgeVector<T>::update(const BaseLinearSystem<T> &linearSystem) const {
  shimUpdateKernel<T><<<(unsigned)((nEdges + 255) / 256), 256>>>(linearSystem.deltaXPtr[0], positionContainers[0].absolutePosition[0],
                                                                 positionContainers[0].absolutePosition[1], cameraNum, nEdges, edgeParams);
}
template <typename T>
void SchurLMLinearSystem<T>::solve() const {  // base_linear_system.cpp:34-45
  const_cast<SchurPCGSolver<T> *>(static_cast<const SchurPCGSolver<T> *>(solverHook))->solve(*this);
}
#define SHIM_INST(T)                                                          \
  template JVD<T> EdgeVector<T>::forward() const;                            \
  template void EdgeVector<T>::backup() const;                               \
  template void EdgeVector<T>::rollback() const;                             \
  template void EdgeVector<T>::update(const BaseLinearSystem<T> &) const;    \
  template void SchurLMLinearSystem<T>::solve() const;
SHIM_INST(double)
SHIM_INST(float)
}  // namespace MegBA

namespace {
template <typename T>
int lm_impl(int64_t nC, int64_t nP, int64_t nE, const int32_t *cam_idx, const int32_t *pt_idx, const T *obs, T *x, int lm_max_iter,
            double tau, double eps1, double eps2, int pcg_max_iter, double pcg_tol, double pcg_refuse, const char *log_path) {
  ProblemOption opt; opt.deviceUsed = {0}; opt.N = 12; opt.nItem = nE;
  MemoryPool::resetPool(&opt);
  cudaSetDevice(0);
  ensure_handles();
  const Index ix = make_index(nC, nP, nE, cam_idx, pt_idx);
  const size_t nX = 9 * nC + 3 * nP;
  // a1: per-edge SoA copies of the parameters + observations
  std::vector<T> soa(12 * (size_t)nE), hobs(2 * (size_t)nE);
  for (int64_t e = 0; e < nE; ++e) {
    for (int k = 0; k < 9; ++k) soa[(size_t)k * nE + e] = x[9 * (size_t)cam_idx[e] + k];
    for (int k = 0; k < 3; ++k) soa[(size_t)(9 + k) * nE + e] = x[9 * nC + 3 * (size_t)pt_idx[e] + k];
    for (int k = 0; k < 2; ++k) hobs[(size_t)k * nE + e] = obs[2 * e + k];
  }
  EdgeVector<T> edges(opt);
  edges.nEdges = nE; edges.cameraNum = (int)nC;
  edges.edgeParams = dev_copy(soa);
  edges.edgeParamsBackup = dev_alloc<T>(12 * (size_t)nE);
  edges.edgeObs = dev_copy(hobs);
  edges.positionContainers.resize(1);
  std::vector<int> ac(cam_idx, cam_idx + nE), ap(pt_idx, pt_idx + nE);
  edges.positionContainers[0].absolutePosition = {dev_copy(ac), dev_copy(ap)};
  edges.positionContainers[0].relativePosition = {dev_copy(ix.relC), dev_copy(ix.relP)};
  SolverOption so;
  so.solverOptionPCG.maxIter = pcg_max_iter; so.solverOptionPCG.tol = pcg_tol; so.solverOptionPCG.refuseRatio = pcg_refuse;
  SchurPCGSolver<T> solver(opt, so);
  SchurLMLinearSystem<T> ls(opt);
  ls.solverHook = &solver;
  ls.dim = {9, 3}; ls.num = {(int)nC, (int)nP};
  ls.equationContainers.resize(1);
  auto &ec = ls.equationContainers[0];
  ec.nnz = {(size_t)27 * nE, (size_t)27 * nE, (size_t)81 * nC, (size_t)9 * nP};
  for (int k = 0; k < 4; ++k) ec.csrVal[k] = dev_alloc<T>(ec.nnz[k]);
  ec.csrRowPtr = {dev_copy(ix.hplRowPtr), dev_copy(ix.hlpRowPtr)};
  ec.csrColInd = {dev_copy(ix.hplColInd), dev_copy(ix.hlpColInd)};
  ls.g = {dev_alloc<T>(nX)}; ls.gBackup = {dev_alloc<T>(nX)};
  ls.deltaXPtr = {dev_alloc<T>(nX)}; ls.deltaXPtrBackup = {dev_alloc<T>(nX)};
  cudaMemset(ls.deltaXPtr[0], 0, sizeof(T) * nX);  // schur_LM_linear_system.cu:37
  ls.extractedDiag = {{dev_alloc<T>(9 * nC), dev_alloc<T>(3 * nP)}};
  T *xDev = dev_alloc<T>(nX);
  cudaMemcpy(xDev, x, sizeof(T) * nX, cudaMemcpyHostToDevice);
  AlgoOption ao;
  ao.algoOptionLM.maxIter = lm_max_iter; ao.algoOptionLM.initialRegion = tau;
  ao.algoOptionLM.epsilon1 = eps1; ao.algoOptionLM.epsilon2 = eps2;
  LMAlgo<T> algo(opt, ao);
  // the reference reports its trace on std::cout: route it into log_path for the caller
  std::cout.flush(); fflush(stdout);
  const int saved = dup(1);
  FILE *lf = log_path ? fopen(log_path, "w") : nullptr;
  if (lf) dup2(fileno(lf), 1);
  algo.solveCUDA(ls, edges, xDev);
  cudaDeviceSynchronize();
  std::cout.flush(); fflush(stdout);
  if (lf) { dup2(saved, 1); fclose(lf); }
  close(saved);
  cudaMemcpy(x, xDev, sizeof(T) * nX, cudaMemcpyDeviceToHost);
  const cudaError_t err = cudaGetLastError();
  return err == cudaSuccess ? 0 : -1;  // device buffers are left to process teardown (checker only)
}
}  // namespace

extern "C" {
int ref_lm_solve_f64(int64_t nC, int64_t nP, int64_t nE, const int32_t *ci, const int32_t *pi, const double *obs, double *x, int lm_max_iter, double tau, double eps1, double eps2, int pcg_max_iter, double pcg_tol, double pcg_refuse, const char *log_path) { return lm_impl<double>(nC, nP, nE, ci, pi, obs, x, lm_max_iter, tau, eps1, eps2, pcg_max_iter, pcg_tol, pcg_refuse, log_path); }
int ref_lm_solve_f32(int64_t nC, int64_t nP, int64_t nE, const int32_t *ci, const int32_t *pi, const float *obs, float *x, int lm_max_iter, double tau, double eps1, double eps2, int pcg_max_iter, double pcg_tol, double pcg_refuse, const char *log_path) { return lm_impl<float>(nC, nP, nE, ci, pi, obs, x, lm_max_iter, tau, eps1, eps2, pcg_max_iter, pcg_tol, pcg_refuse, log_path); }
int ref_forward_f64(int64_t nE, const int32_t *ci, const int32_t *pi, const double *obs, const double *cam, const double *pt, double *r, double *J) { return forward_impl<double>(nE, ci, pi, obs, cam, pt, r, J); }
int ref_forward_f32(int64_t nE, const int32_t *ci, const int32_t *pi, const float *obs, const float *cam, const float *pt, float *r, float *J) { return forward_impl<float>(nE, ci, pi, obs, cam, pt, r, J); }
int ref_build_f64(int64_t nC, int64_t nP, int64_t nE, const int32_t *ci, const int32_t *pi, const double *r, const double *J, const double *info, double *Hpp, double *Hll, double *W, double *g) { return build_impl<double>(nC, nP, nE, ci, pi, r, J, info, Hpp, Hll, W, g); }
int ref_build_f32(int64_t nC, int64_t nP, int64_t nE, const int32_t *ci, const int32_t *pi, const float *r, const float *J, const float *info, float *Hpp, float *Hll, float *W, float *g) { return build_impl<float>(nC, nP, nE, ci, pi, r, J, info, Hpp, Hll, W, g); }
int ref_solve_f64(int64_t nC, int64_t nP, int64_t nE, const int32_t *ci, const int32_t *pi, const double *HppD, const double *HllD, const double *W, const double *g, double *dx, int max_iter, double tol, double refuse) { return solve_impl<double>(nC, nP, nE, ci, pi, HppD, HllD, W, g, dx, max_iter, tol, refuse); }
// same as ref_solve_f64, additionally reports the device time of SchurPCGSolver::solve (used by bench.py to quote the
// reference's own cuSPARSE/cuBLAS solver on the same GPU and problem)
int ref_solve_timed_f64(int64_t nC, int64_t nP, int64_t nE, const int32_t *ci, const int32_t *pi, const double *HppD, const double *HllD, const double *W, const double *g, double *dx, int max_iter, double tol, double refuse, double *ms) { return solve_impl<double>(nC, nP, nE, ci, pi, HppD, HllD, W, g, dx, max_iter, tol, refuse, ms); }
int ref_build_implicit_f64(int64_t nC, int64_t nP, int64_t nE, const int32_t *ci, const int32_t *pi, const double *r, const double *J, double *Hpp, double *Hll, double *g) { return build_implicit_impl<double>(nC, nP, nE, ci, pi, r, J, Hpp, Hll, g); }
int ref_build_implicit_f32(int64_t nC, int64_t nP, int64_t nE, const int32_t *ci, const int32_t *pi, const float *r, const float *J, float *Hpp, float *Hll, float *g) { return build_implicit_impl<float>(nC, nP, nE, ci, pi, r, J, Hpp, Hll, g); }
int ref_solve_implicit_f64(int64_t nC, int64_t nP, int64_t nE, const int32_t *ci, const int32_t *pi, const double *HppD, const double *HllD, const double *J, const double *g, double *dx, int max_iter, double tol, double refuse) { return solve_implicit_impl<double>(nC, nP, nE, ci, pi, HppD, HllD, J, g, dx, max_iter, tol, refuse); }
int ref_solve_implicit_f32(int64_t nC, int64_t nP, int64_t nE, const int32_t *ci, const int32_t *pi, const float *HppD, const float *HllD, const float *J, const float *g, float *dx, int max_iter, double tol, double refuse) { return solve_implicit_impl<float>(nC, nP, nE, ci, pi, HppD, HllD, J, g, dx, max_iter, tol, refuse); }
int ref_solve_f32(int64_t nC, int64_t nP, int64_t nE, const int32_t *ci, const int32_t *pi, const float *HppD, const float *HllD, const float *W, const float *g, float *dx, int max_iter, double tol, double refuse) { return solve_impl<float>(nC, nP, nE, ci, pi, HppD, HllD, W, g, dx, max_iter, tol, refuse); }
}
