// megba/geo/geo.cuh -- the geometry entry points user residuals call inside forward()
// (reference: include/geo/geo.cuh:32-68).
//   AnalyticalDerivativesKernelMatrix  names the fused analytical BAL residual (src/geo/analytical_derivatives.cu:289-322);
//                                      the hand-written kernels evaluate it whenever the LM loop needs it.
//   AngleAxisToRotationKernelMatrix, RadialDistortion  (src/geo/angle_axis.cu, src/geo/distortion.cu) are recorded on
//                                      the residual tape as compositions of JetVector operations; the reference runs
//                                      one dedicated kernel each.
#pragma once
#include <limits>
#include <stdexcept>

#include "../edge/base_edge.h"

namespace MegBA {
namespace geo {
template <typename T>
using JVD = ::MegBA::JVD<T>;
template <typename T>
using JV3 = Dense::Matrix<JetVector<T>, 3, 1>;
template <typename T>
using JM33 = Dense::Matrix<JetVector<T>, 3, 3>;

// Eigen::Map<const JVD<T>>{&estimation(row, 0), rows, 1} in the reference's examples: a view of consecutive entries
template <typename T>
struct MapView {
  const JetVector<T> *p;
  int rows, cols;
  MapView(const JetVector<T> *ptr, int r, int c) : p(ptr), rows(r), cols(c) {}
  const JetVector<T> &operator()(int i) const { return p[i]; }
  int size() const { return rows * cols; }
  JVD<T> mat() const {
    JVD<T> m(rows, cols);
    for (int i = 0; i < rows * cols; ++i) m(i) = p[i];
    return m;
  }
};
template <typename T, int R, int C>
Dense::Matrix<JetVector<T>, Dense::Dynamic, Dense::Dynamic> operator+(const Dense::Matrix<JetVector<T>, R, C> &a, const MapView<T> &b) { return a + b.mat(); }
template <typename T, int R, int C>
Dense::Matrix<JetVector<T>, Dense::Dynamic, Dense::Dynamic> operator-(const Dense::Matrix<JetVector<T>, R, C> &a, const MapView<T> &b) { return a - b.mat(); }
template <typename T, int R, int C>
Dense::Matrix<JetVector<T>, Dense::Dynamic, Dense::Dynamic> operator*(const Dense::Matrix<JetVector<T>, R, C> &a, const MapView<T> &b) { return a * b.mat(); }

template <typename T>
JVD<T> AnalyticalDerivativesKernelMatrix(const MapView<T> &angleAxis, const MapView<T> &t, const MapView<T> &intrinsics,
                                         const JVD<T> &pointXyz, const JVD<T> &obsUv) {
  if (angleAxis.rows != 3 || t.rows != 3 || intrinsics.rows != 3 || pointXyz.rows() != 3 || obsUv.rows() != 2)
    throw std::runtime_error("AnalyticalDerivativesKernelMatrix: expected camera = [aa(3) t(3) f k1 k2], point(3), obs(2)");
  JVD<T> out(2, 1);
  out.program = ResidualProgram::BAL_ANALYTICAL;
  return out;
}

// Rotation matrix of an angle-axis vector, first-order branch for theta^2 <= eps (src/geo/angle_axis.cu:28-47,131-158)
template <typename T, typename Vec>
JM33<T> AngleAxisToRotationImpl(const Vec &aa) {
  using J = JetVector<T>;
  const J x = aa(0), y = aa(1), z = aa(2);
  const J theta2 = x * x + y * y + z * z;
  const J eps(std::numeric_limits<T>::epsilon());
  const J theta = math::sqrt(theta2);
  const J wx = x / theta, wy = y / theta, wz = z / theta;
  const J c = math::cos(theta), s = math::sin(theta);
  const J one(T(1)), omc = one - c;
  JM33<T> R;
  auto sel = [&](const J &big, const J &small) { return math::select_gt(theta2, eps, big, small); };
  R(0, 0) = sel(c + wx * wx * omc, one);
  R(1, 0) = sel(wz * s + wx * wy * omc, z);
  R(2, 0) = sel(wx * wz * omc - wy * s, -y);
  R(0, 1) = sel(wx * wy * omc - wz * s, -z);
  R(1, 1) = sel(c + wy * wy * omc, one);
  R(2, 1) = sel(wx * s + wy * wz * omc, x);
  R(0, 2) = sel(wy * s + wx * wz * omc, y);
  R(1, 2) = sel(wy * wz * omc - wx * s, -x);
  R(2, 2) = sel(c + wz * wz * omc, one);
  return R;
}
template <typename T>
JM33<T> AngleAxisToRotationKernelMatrix(const MapView<T> &aa) { return AngleAxisToRotationImpl<T>(aa); }
template <typename T>
JM33<T> AngleAxisToRotationKernelMatrix(const JV3<T> &aa) { return AngleAxisToRotationImpl<T>(aa); }

// f (1 + k1 rho^2 + k2 rho^4), rho^2 = x^2 + y^2 of the normalised image point (src/geo/distortion.cu)
template <typename T, typename Vec>
JetVector<T> RadialDistortionImpl(const JV3<T> &point, const Vec &intrinsic) {
  using J = JetVector<T>;
  const J rho2 = point(0) * point(0) + point(1) * point(1);
  return intrinsic(0) * (J(T(1)) + intrinsic(1) * rho2 + intrinsic(2) * rho2 * rho2);
}
template <typename T>
JetVector<T> RadialDistortion(const JV3<T> &point, const MapView<T> &intrinsic) { return RadialDistortionImpl<T>(point, intrinsic); }
template <typename T>
JetVector<T> RadialDistortion(const JV3<T> &point, const JV3<T> &intrinsic) { return RadialDistortionImpl<T>(point, intrinsic); }
}  // namespace geo
}  // namespace MegBA

// `Eigen::Map<const MegBA::geo::JVD<T>>` in user code (examples/BAL_Double_analytical.cpp:19-23 of the reference)
#ifndef MEGBA_USE_EIGEN
namespace MegBA { namespace Dense {
template <typename X> struct MapSelector;
template <typename T> struct MapSelector<const ::MegBA::JVD<T>> { using type = ::MegBA::geo::MapView<T>; };
template <typename X> using Map = typename MapSelector<X>::type;
} }
#endif
