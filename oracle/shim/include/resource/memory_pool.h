// shim of include/resource/memory_pool.h: only the static members the three hot-path files call.
#pragma once
#include <cstddef>
#include "common.h"
namespace MegBA {
class MemoryPool {
 public:
  static const ProblemOption *_problemOption;
  static void resetPool(const ProblemOption *o) { _problemOption = o; }
  static void allocateNormal(void **ptr, size_t size, int rank = 0);
  static void deallocateNormal(void *ptr, int rank = 0);
  static void redistribute() {}
  static std::size_t getWorldSize() { return _problemOption->deviceUsed.size(); }
  // same partition rule as the reference (include/resource/memory_pool.h:48-63)
  static std::size_t getItemNum(int rank) { return getItemNum(rank, (std::size_t)_problemOption->nItem); }
  static std::size_t getItemNum(int rank, std::size_t nItem) {
    const auto ws = getWorldSize();
    return rank == (int)ws - 1 ? nItem - (nItem / ws + 1) * (ws - 1) : nItem / ws + 1;
  }
};
}  // namespace MegBA
