// The two residuals of the BAL examples as a user of the facade writes them -- by construction the same few lines as
// in the reference's examples (examples/BAL_Double_analytical.cpp:15-31, examples/BAL_Double.cpp:15-33), since they
// are the API's usage, not its implementation:
//   BalEdgeAnalyticalDerivatives  names the fused analytical residual (hand-written kernels),
//   BAL_Edge                      writes the residual on JetVectors; the facade records it once and the library
//                                 evaluates and differentiates it per observation on the GPU.
#pragma once
#include "edge/base_edge.h"
#include "geo/geo.cuh"

template <typename T>
class BalEdgeAnalyticalDerivatives : public MegBA::BaseEdge<T> {
 public:
  MegBA::JVD<T> forward() override {
    using MappedJVD = Eigen::Map<const MegBA::geo::JVD<T>>;
    const auto &Vertices = this->getVertices();
    MappedJVD angle_axisd{&Vertices[0].getEstimation()(0, 0), 3, 1};
    MappedJVD t{&Vertices[0].getEstimation()(3, 0), 3, 1};
    MappedJVD intrinsics{&Vertices[0].getEstimation()(6, 0), 3, 1};
    const auto &point_xyz = Vertices[1].getEstimation();
    const auto &obs_uv = this->getMeasurement();
    MegBA::JVD<T> &&error = MegBA::geo::AnalyticalDerivativesKernelMatrix(angle_axisd, t, intrinsics, point_xyz, obs_uv);
    return error;
  }
};

template <typename T>
class BAL_Edge : public MegBA::BaseEdge<T> {
 public:
  MegBA::JVD<T> forward() override {
    using MappedJVD = Eigen::Map<const MegBA::geo::JVD<T>>;
    const auto &Vertices = this->getVertices();
    MappedJVD angle_axisd{&Vertices[0].getEstimation()(0, 0), 3, 1};
    MappedJVD t{&Vertices[0].getEstimation()(3, 0), 3, 1};
    MappedJVD intrinsics{&Vertices[0].getEstimation()(6, 0), 3, 1};
    const auto &point_xyz = Vertices[1].getEstimation();
    const auto &obs_uv = this->getMeasurement();
    auto R = MegBA::geo::AngleAxisToRotationKernelMatrix(angle_axisd);
    Eigen::Matrix<MegBA::JetVector<T>, 3, 1> re_projection = R * point_xyz + t;
    re_projection = -re_projection / re_projection(2);
    // f, k1, k2 = intrinsics
    auto fr = MegBA::geo::RadialDistortion(re_projection, intrinsics);
    MegBA::JVD<T> error = fr * re_projection.head(2) - obs_uv;
    return error;
  }
};

