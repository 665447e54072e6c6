// shim of the reference's include/macro.h: release-mode behaviour (the debug macros compile to nothing, macro.h:96-107)
#pragma once
#include <cuda_runtime_api.h>
#include <stdexcept>
#include <string>
#define ASSERT_CUDA_NO_ERROR() do {} while (false)
#define PRINT_DMEMORY(...) do {} while (false)
#define PRINT_DMEMORY_SEGMENT(...) do {} while (false)
#define PRINT_DCSR(...) do {} while (false)
