// megba/dense.h -- the few dense-matrix operations the facade's public signatures need.
// The reference exposes Eigen types (TD<T> = Eigen::Matrix<T, Dynamic, Dynamic>, include/vertex/
// base_vertex.h:19-23); Eigen is an un-vendored submodule that is absent here, so the facade ships this
// minimal stand-in with the same spelling for what the BAL examples use: Matrix<T,R,C>, operator()(i),
// operator()(i,j), size/rows/cols, resize, Identity, Zero.  With real Eigen available, define
// MEGBA_USE_EIGEN before including the facade and these aliases map to Eigen instead.
#pragma once
#ifdef MEGBA_USE_EIGEN
#include <Eigen/Core>
namespace MegBA { namespace Dense = ::Eigen; }
#else
#include <cstddef>
#include <stdexcept>
#include <vector>

namespace MegBA {
namespace Dense {
constexpr int Dynamic = -1;

template <typename T, int R, int C>
class Matrix {
  int r_{R == Dynamic ? 0 : R}, c_{C == Dynamic ? 0 : C};
  std::vector<T> v_;  // column-major like Eigen's default

 public:
  Matrix() {  // (no T(0) for an empty matrix: a symbolic JetVector element cannot be created outside forward())
    const std::size_t n = static_cast<std::size_t>((R == Dynamic ? 0 : R) * (C == Dynamic ? 0 : C));
    if (n) v_.assign(n, T(0));
  }
  Matrix(int rows, int cols) { resize(rows, cols); }
  template <int R2, int C2>
  Matrix(const Matrix<T, R2, C2> &o) { *this = o; }  // NOLINT: implicit like Eigen
  template <int R2, int C2>
  Matrix &operator=(const Matrix<T, R2, C2> &o) {
    if ((R != Dynamic && R != o.rows()) || (C != Dynamic && C != o.cols())) throw std::runtime_error("Matrix: shape mismatch");
    r_ = o.rows(); c_ = o.cols();
    v_.assign(o.data(), o.data() + o.size());
    return *this;
  }
  void resize(int rows, int cols) {
    if ((R != Dynamic && R != rows) || (C != Dynamic && C != cols)) throw std::runtime_error("Matrix: cannot resize fixed dimension");
    r_ = rows; c_ = cols;
    v_.clear();
    if (rows * cols != 0) v_.assign(static_cast<std::size_t>(rows) * cols, T(0));
  }
  int rows() const { return r_; }
  int cols() const { return c_; }
  int size() const { return r_ * c_; }
  T *data() { return v_.data(); }
  const T *data() const { return v_.data(); }
  T &operator()(int i) { return v_[i]; }
  const T &operator()(int i) const { return v_[i]; }
  T &operator()(int i, int j) { return v_[static_cast<std::size_t>(j) * r_ + i]; }
  const T &operator()(int i, int j) const { return v_[static_cast<std::size_t>(j) * r_ + i]; }
  bool operator==(const Matrix &o) const { return r_ == o.r_ && c_ == o.c_ && v_ == o.v_; }
  // first n entries of a column vector (Eigen's head(n))
  Matrix<T, Dynamic, 1> head(int n) const {
    Matrix<T, Dynamic, 1> h(n, 1);
    for (int i = 0; i < n; ++i) h(i) = (*this)(i);
    return h;
  }
  static Matrix Zero() { return Matrix(); }
  static Matrix Identity() {
    Matrix m;
    for (int i = 0; i < m.rows() && i < m.cols(); ++i) m(i, i) = T(1);
    return m;
  }
};
// The arithmetic a residual written on JetVector matrices needs (examples/BAL_Double.cpp:24-32 of the reference):
// matrix * vector, +, -, unary -, vector / scalar, scalar * vector.  Results are dynamically sized; assigning them to a
// fixed-size Matrix checks the shape.
template <typename T, int R1, int C1, int R2, int C2>
Matrix<T, Dynamic, Dynamic> operator*(const Matrix<T, R1, C1> &a, const Matrix<T, R2, C2> &b) {
  if (a.cols() != b.rows()) throw std::runtime_error("Matrix: inner dimensions differ");
  Matrix<T, Dynamic, Dynamic> c(a.rows(), b.cols());
  for (int i = 0; i < a.rows(); ++i)
    for (int j = 0; j < b.cols(); ++j) {
      T acc = a(i, 0) * b(0, j);
      for (int k = 1; k < a.cols(); ++k) acc = acc + a(i, k) * b(k, j);
      c(i, j) = acc;
    }
  return c;
}
template <typename T, int R1, int C1, int R2, int C2>
Matrix<T, Dynamic, Dynamic> operator+(const Matrix<T, R1, C1> &a, const Matrix<T, R2, C2> &b) {
  if (a.rows() != b.rows() || a.cols() != b.cols()) throw std::runtime_error("Matrix: shapes differ");
  Matrix<T, Dynamic, Dynamic> c(a.rows(), a.cols());
  for (int i = 0; i < a.size(); ++i) c(i) = a(i) + b(i);
  return c;
}
template <typename T, int R1, int C1, int R2, int C2>
Matrix<T, Dynamic, Dynamic> operator-(const Matrix<T, R1, C1> &a, const Matrix<T, R2, C2> &b) {
  if (a.rows() != b.rows() || a.cols() != b.cols()) throw std::runtime_error("Matrix: shapes differ");
  Matrix<T, Dynamic, Dynamic> c(a.rows(), a.cols());
  for (int i = 0; i < a.size(); ++i) c(i) = a(i) - b(i);
  return c;
}
template <typename T, int R1, int C1>
Matrix<T, Dynamic, Dynamic> operator-(const Matrix<T, R1, C1> &a) {
  Matrix<T, Dynamic, Dynamic> c(a.rows(), a.cols());
  for (int i = 0; i < a.size(); ++i) c(i) = -a(i);
  return c;
}
template <typename T, int R1, int C1>
Matrix<T, Dynamic, Dynamic> operator/(const Matrix<T, R1, C1> &a, const T &s) {
  Matrix<T, Dynamic, Dynamic> c(a.rows(), a.cols());
  for (int i = 0; i < a.size(); ++i) c(i) = a(i) / s;
  return c;
}
template <typename T, int R1, int C1>
Matrix<T, Dynamic, Dynamic> operator*(const T &s, const Matrix<T, R1, C1> &a) {
  Matrix<T, Dynamic, Dynamic> c(a.rows(), a.cols());
  for (int i = 0; i < a.size(); ++i) c(i) = s * a(i);
  return c;
}
template <typename T, int R1, int C1>
Matrix<T, Dynamic, Dynamic> operator*(const Matrix<T, R1, C1> &a, const T &s) { return s * a; }

using Matrix2d = Matrix<double, 2, 2>;
using Matrix2f = Matrix<float, 2, 2>;
}  // namespace Dense
}  // namespace MegBA
namespace Eigen = MegBA::Dense;  // lets `Eigen::Matrix<T, 9, 1>` in user code resolve to the stand-in
#endif
