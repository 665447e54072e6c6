// megba/vertex/base_vertex.h -- BaseVertex / CameraVertex / PointVertex
// (API of the reference's include/vertex/base_vertex.h:25-77).
#pragma once
#include <utility>

#include "../common.h"
#include "../dense.h"

namespace MegBA {
template <typename T>
using TD = Dense::Matrix<T, Dense::Dynamic, Dense::Dynamic>;

enum VertexKind { CAMERA = 0, POINT = 1, NONE = 2 };

template <typename T>
struct BaseVertex {
  BaseVertex() = default;
  virtual ~BaseVertex() = default;

  template <typename Estimation>
  void setEstimation(Estimation &&estimation) { _estimation = std::forward<Estimation>(estimation); }
  template <typename Observation>
  void setObservation(Observation &&observation) { _observation = std::forward<Observation>(observation); }

  const TD<T> &getEstimation() const { return _estimation; }
  TD<T> &getEstimation() { return _estimation; }
  const TD<T> &getObservation() const { return _observation; }
  TD<T> &getObservation() { return _observation; }
  unsigned int getGradShape() const { return fixed ? 0 : _estimation.rows() * _estimation.cols(); }
  bool operator==(const BaseVertex<T> &v) const { return _estimation == v._estimation; }

  virtual VertexKind kind() { return NONE; }
  int absolutePosition{};
  bool fixed{false};

 private:
  TD<T> _estimation{};
  TD<T> _observation{};
};

template <typename T>
struct CameraVertex : public BaseVertex<T> {
  VertexKind kind() final { return CAMERA; }
};

template <typename T>
struct PointVertex : public BaseVertex<T> {
  VertexKind kind() final { return POINT; }
};
}  // namespace MegBA
