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
  void buildImplicitLinearSystemCUDA(const JVD<T> &jetEstimation, const BaseLinearSystem<T> &linearSystem) const;
  // what LMAlgo::solveCUDA calls (base_edge.h:99-152); forward/backup/rollback/update are provided by the driver
  const std::vector<PositionContainer> &getPositionContainers() const { return positionContainers; }
  JVD<T> forward() const;
  void buildLinearSystem(const JVD<T> &jetEstimation, const BaseLinearSystem<T> &linearSystem) const {
    buildLinearSystemCUDA(jetEstimation, linearSystem);
  }
  void backup() const;
  void rollback() const;
  void update(const BaseLinearSystem<T> &linearSystem) const;
  // a1: per-edge copies of the 9 camera + 3 point parameters, SoA [12][nE] (+ backup), and obs [2][nE]
  T *edgeParams{nullptr};
  T *edgeParamsBackup{nullptr};
  T *edgeObs{nullptr};
  long long nEdges{0};
  int cameraNum{0};
};
}  // namespace MegBA
