// shim of include/geo/geo.cuh: only the declaration the analytical BAL edge uses (geo.cuh:61-66)
#pragma once
#include "operator/jet_vector.h"
namespace MegBA {
namespace geo {
template <typename T>
using JVD = Eigen::Matrix<JetVector<T>, Eigen::Dynamic, Eigen::Dynamic>;
template <typename T>
JVD<T> AnalyticalDerivativesKernelMatrix(const Eigen::Map<const JVD<T>> &AxisAngle, const Eigen::Map<const JVD<T>> &t,
                                         const Eigen::Map<const JVD<T>> &intrinsics, const JVD<T> &point_xyz,
                                         const JVD<T> &obs_uv);
}  // namespace geo
}  // namespace MegBA
