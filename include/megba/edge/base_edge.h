// megba/edge/base_edge.h -- BaseEdge with the reference's residual plug-in seam
//   virtual JVD<T> forward() = 0            (include/edge/base_edge.h:43)
// In the reference forward() builds a JetVector expression that is evaluated for ALL edges at once
// (EdgeVector::forward, src/edge/base_edge.cpp:160-163).  Here forward() is called once, symbolically:
// the JVD it returns is a small tag naming the fused CUDA residual (see geo/geo.cuh).  Only the
// analytical BAL reprojection edge exists on this path; anything else is rejected loudly.
#pragma once
#include <stdexcept>
#include <utility>
#include <vector>

#include "../common.h"
#include "../vertex/base_vertex.h"

namespace MegBA {
enum EdgeKind { ONE, ONE_CAMERA_ONE_POINT, TWO_CAMERA, MULTI };

// What forward() returns: a handle on the residual program (the reference's JVD<T> is an Eigen matrix of
// JetVectors that owns device slabs; the fused kernels need neither).
enum class ResidualProgram { NONE, BAL_ANALYTICAL };
template <typename T>
struct JVD {
  ResidualProgram program{ResidualProgram::NONE};
  int rows{0}, cols{0};
};

// View of a vertex's estimation inside forward(): only identity/slicing is meaningful symbolically.
template <typename T>
struct EstimationView {
  int rows{0};
  const EstimationView &getEstimation() const { return *this; }
  const EstimationView *operator&() const { return this; }
  EstimationView operator()(int, int) const { return *this; }
};

template <typename T>
class EdgeWrapper {
  std::vector<EstimationView<T>> views_;
  EstimationView<T> measurement_;

 public:
  void bind(const std::vector<BaseVertex<T> *> &vs, int measRows) {
    views_.clear();
    for (auto *v : vs) views_.push_back(EstimationView<T>{v->getEstimation().rows()});
    measurement_.rows = measRows;
  }
  const EstimationView<T> &operator[](std::size_t n) const { return views_.at(n); }
  const EstimationView<T> &getMeasurement() const { return measurement_; }
};

template <typename T>
class BaseEdge {
  using PlainMatrix = Dense::Matrix<T, Dense::Dynamic, Dense::Dynamic>;
  std::vector<BaseVertex<T> *> m_data;
  EdgeWrapper<T> _edgeWrapper;
  PlainMatrix _measurement;
  PlainMatrix _information;

 public:
  virtual ~BaseEdge() = default;
  std::size_t size() { return m_data.size(); }
  BaseVertex<T> *&operator[](std::size_t n) { return m_data[n]; }
  BaseVertex<T> *const &operator[](std::size_t n) const { return m_data[n]; }

  void appendVertex(BaseVertex<T> *vertex) { m_data.push_back(vertex); }
  bool existVertex(const BaseVertex<T> &vertex) const {
    for (auto *v : m_data) if (v == &vertex) return true;
    return false;
  }

  virtual JVD<T> forward() = 0;

  template <typename M> void setMeasurement(M &&m) { _measurement = std::forward<M>(m); }
  template <typename M> void setInformation(M &&m) { _information = std::forward<M>(m); }
  const PlainMatrix &_getMeasurement() const { return _measurement; }
  const PlainMatrix &_getInformation() const { return _information; }

  // called by BaseProblem before the symbolic forward()
  void bindSymbolic() { _edgeWrapper.bind(m_data, _measurement.rows()); }

 protected:
  const EdgeWrapper<T> &getVertices() { return _edgeWrapper; }
  const EstimationView<T> &getMeasurement() const { return _edgeWrapper.getMeasurement(); }
};
}  // namespace MegBA
