// shim of include/resource/handle_manager.h (cuBLAS / cuSPARSE handles, one per device)
#pragma once
#include <cublas_v2.h>
#include <cusparse_v2.h>
#include <vector>
namespace MegBA {
class HandleManager {
 public:
  static std::vector<cublasHandle_t> cublasHandle;
  static std::vector<cusparseHandle_t> cusparseHandle;
  static const std::vector<cublasHandle_t> &getCUBLASHandle() { return cublasHandle; }
  static const std::vector<cusparseHandle_t> &getCUSPARSEHandle() { return cusparseHandle; }
};
}  // namespace MegBA
