// megba/geo/geo.cuh -- geo::AnalyticalDerivativesKernelMatrix, the one residual this path implements.
// Reference: include/geo/geo.cuh:32-68, src/geo/analytical_derivatives.cu:289-322.  The reference returns
// evaluated JetVectors; here the call only *names* the residual, and the evaluation happens inside the fused
// linearisation / trial kernels (megba_b200/csrc/ba_kernels.cu) whenever the LM loop needs it.
#pragma once
#include <stdexcept>

#include "../edge/base_edge.h"

namespace MegBA {
namespace geo {
template <typename T>
using JVD = ::MegBA::JVD<T>;

// Eigen::Map<const JVD<T>>{&estimation(row, 0), rows, 1} in the reference's examples
template <typename T>
struct MapView {
  int rows;
  MapView(const EstimationView<T> *, int r, int /*cols*/) : rows(r) {}
};

template <typename T>
JVD<T> AnalyticalDerivativesKernelMatrix(const MapView<T> &angleAxis, const MapView<T> &t, const MapView<T> &intrinsics,
                                         const EstimationView<T> &pointXyz, const EstimationView<T> &obsUv) {
  if (angleAxis.rows != 3 || t.rows != 3 || intrinsics.rows != 3 || pointXyz.rows != 3 || obsUv.rows != 2)
    throw std::runtime_error("AnalyticalDerivativesKernelMatrix: expected camera = [aa(3) t(3) f k1 k2], point(3), obs(2)");
  JVD<T> out;
  out.program = ResidualProgram::BAL_ANALYTICAL;
  out.rows = 2;
  out.cols = 1;
  return out;
}
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
