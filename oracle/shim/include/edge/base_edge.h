// shim of include/edge/base_edge.h: the members EdgeVector<T>::buildLinearSystemCUDA touches (base_edge.h:69-88,148-152)
#pragma once
#include <vector>
#include "common.h"
#include "operator/jet_vector.h"
namespace MegBA {
template <typename T> struct BaseLinearSystem;
template <typename T>
class EdgeVector {
 public:
  struct PositionContainer {
    std::vector<int *> relativePosition{};
    std::vector<int *> absolutePosition{};
  };
  explicit EdgeVector(const ProblemOption &o) : option(o) {}
  const ProblemOption &option;
  std::vector<PositionContainer> positionContainers;
  JVD<T> jetInformation;
  void buildLinearSystemCUDA(const JVD<T> &jetEstimation, const BaseLinearSystem<T> &linearSystem) const;
};
}  // namespace MegBA
