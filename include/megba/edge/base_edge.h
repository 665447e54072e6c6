// megba/edge/base_edge.h -- BaseEdge with the reference's residual plug-in seam
//   virtual JVD<T> forward() = 0            (include/edge/base_edge.h:43)
// In the reference forward() builds a JetVector expression that is evaluated for ALL edges at once, one kernel launch
// per operator (EdgeVector::forward, src/edge/base_edge.cpp:160-163).  Here forward() is called ONCE, on the host, with
// symbolic JetVectors (operator/jet_vector.h): either it names the fused analytical BAL residual
// (geo::AnalyticalDerivativesKernelMatrix -> a tag, evaluated by the hand-written kernels), or its expression tape is
// compiled to a bytecode that the library evaluates per observation on the GPU with forward-mode automatic
// differentiation (ba_set_residual_program) -- the user-defined-residual path of the reference's BAL_Double.cpp.
#pragma once
#include <stdexcept>
#include <utility>
#include <vector>

#include "../common.h"
#include "../dense.h"
#include "../operator/jet_vector.h"
#include "../vertex/base_vertex.h"

namespace MegBA {
enum EdgeKind { ONE, ONE_CAMERA_ONE_POINT, TWO_CAMERA, MULTI };

enum class ResidualProgram { NONE, BAL_ANALYTICAL };

// What forward() works with and returns: a matrix of JetVectors, like the reference's
// JVD<T> = Eigen::Matrix<JetVector<T>, Dynamic, Dynamic>; `program` is set by the fused-residual entry points of geo.
template <typename T>
struct JVD : public Dense::Matrix<JetVector<T>, Dense::Dynamic, Dense::Dynamic> {
  using Base = Dense::Matrix<JetVector<T>, Dense::Dynamic, Dense::Dynamic>;
  ResidualProgram program{ResidualProgram::NONE};
  JVD() = default;
  JVD(int r, int c) : Base(r, c) {}
  template <int R, int C>
  JVD(const Dense::Matrix<JetVector<T>, R, C> &m) : Base(m) {}  // NOLINT: implicit like Eigen
};

// A vertex as forward() sees it: its estimation as symbolic inputs.
template <typename T>
struct VertexView {
  JVD<T> estimation;
  const JVD<T> &getEstimation() const { return estimation; }
};

template <typename T>
class EdgeWrapper {
  std::vector<VertexView<T>> views_;
  JVD<T> measurement_;

 public:
  // inputs of the residual program: vertex 0 (camera) = 0..8, vertex 1 (point) = 9..11, measurement = 12..13
  void bind(const std::vector<BaseVertex<T> *> &vs, int measRows) {
    views_.clear();
    int next = 0;
    for (auto *v : vs) {
      VertexView<T> vv;
      const int n = v->getEstimation().rows() * v->getEstimation().cols();
      vv.estimation = JVD<T>(n, 1);
      for (int i = 0; i < n; ++i) vv.estimation(i) = JetVector<T>::input(next++);
      views_.push_back(std::move(vv));
    }
    if (next != 12 || measRows != 2)
      throw std::runtime_error("edges of this path join a 9-parameter camera and a 3-parameter point with a 2-vector measurement");
    measurement_ = JVD<T>(measRows, 1);
    for (int i = 0; i < measRows; ++i) measurement_(i) = JetVector<T>::input(next++);
  }
  const VertexView<T> &operator[](std::size_t n) const { return views_.at(n); }
  const JVD<T> &getMeasurement() const { return measurement_; }
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
  const JVD<T> &getMeasurement() const { return _edgeWrapper.getMeasurement(); }
};
}  // namespace MegBA
