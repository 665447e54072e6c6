// shim of include/operator/jet_vector.h + the Eigen container it lives in: a JetVector here is just the pair of device
// pointers (value[nItem], grad[N][nItem]) per device that the hot-path kernels read and write; the slab layout
// [grad | value] follows src/resource/memory_pool.cu:89-92.
#pragma once
#include <cuda_runtime_api.h>
#include <cstddef>
#include <vector>
#include "common.h"
#include "resource/handle_manager.h"  // reached transitively in the reference too (jet_vector.h:18-19)
#include "resource/memory_pool.h"

namespace MegBA {
template <typename T>
class JetVector {
  unsigned N_{0};
  std::vector<std::size_t> nItem_{};
  std::vector<T *> res_{}, grad_{};

 public:
  void bind(unsigned N, const std::vector<std::size_t> &nItem, const std::vector<T *> &res, const std::vector<T *> &grad) {
    N_ = N; nItem_ = nItem; res_ = res; grad_ = grad;
  }
  void initAs(const JetVector<T> &o) {
    N_ = o.N_; nItem_ = o.nItem_;
    res_.assign(nItem_.size(), nullptr); grad_.assign(nItem_.size(), nullptr);
    for (std::size_t i = 0; i < nItem_.size(); ++i) {
      cudaSetDevice((int)i);
      T *slab = nullptr;
      cudaMalloc(&slab, sizeof(T) * (N_ + 1) * nItem_[i]);
      grad_[i] = slab;
      res_[i] = slab + (std::size_t)N_ * nItem_[i];
    }
  }
  void release() {  // frees slabs obtained through initAs
    for (auto *g : grad_) if (g) cudaFree(g);
    grad_.clear(); res_.clear();
  }
  unsigned getGradShape() const { return N_; }
  std::size_t getItemNum(int rank) const { return nItem_[rank]; }
  const std::vector<T *> &getCUDAResPtr() const { return res_; }
  const std::vector<T *> &getCUDAGradPtr() const { return grad_; }
};
}  // namespace MegBA

namespace Eigen {
constexpr int Dynamic = -1;
template <typename S, int R, int C>
class Matrix {
  long r_{0}, c_{0};
  std::vector<S> v_;
 public:
  void resize(long r, long c) { r_ = r; c_ = c; v_.assign((std::size_t)(r * c), S{}); }
  long rows() const { return r_; }
  long cols() const { return c_; }
  long size() const { return r_ * c_; }
  S &operator()(long i) { return v_[i]; }
  const S &operator()(long i) const { return v_[i]; }
  S &operator()(long i, long j) { return v_[(std::size_t)(j * r_ + i)]; }
  const S &operator()(long i, long j) const { return v_[(std::size_t)(j * r_ + i)]; }
};
template <typename M> class Map;
template <typename S, int R, int C>
class Map<const Matrix<S, R, C>> {
  const S *p_; long r_, c_;
 public:
  Map(const S *p, long r, long c) : p_(p), r_(r), c_(c) {}
  const S &operator()(long i) const { return p_[i]; }
  const S &operator()(long i, long j) const { return p_[j * r_ + i]; }
};
}  // namespace Eigen

namespace MegBA {
template <typename T>
using JVD = Eigen::Matrix<JetVector<T>, Eigen::Dynamic, Eigen::Dynamic>;
}
