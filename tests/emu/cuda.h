// TEST INFRASTRUCTURE ONLY -- stand-in for the driver API header <cuda.h>: the types fd2d_chain.cu needs to describe its
// input arrays to the TMA unit.  The emulated "tensor map" simply records what cuTensorMapEncodeTiled was told; the
// emulated cp.async.bulk.tensor (tests/emu/cuda_runtime.h, emu::tma_issue6) reads it back.
#pragma once
#include <cstdint>

typedef uint32_t cuuint32_t;
typedef uint64_t cuuint64_t;
typedef int CUresult;
constexpr CUresult CUDA_SUCCESS = 0;
enum CUtensorMapDataType { CU_TENSOR_MAP_DATA_TYPE_FLOAT32 = 7 };
enum CUtensorMapInterleave { CU_TENSOR_MAP_INTERLEAVE_NONE = 0 };
enum CUtensorMapSwizzle { CU_TENSOR_MAP_SWIZZLE_NONE = 0 };
enum CUtensorMapL2promotion { CU_TENSOR_MAP_L2_PROMOTION_NONE = 0, CU_TENSOR_MAP_L2_PROMOTION_L2_128B = 2 };
enum CUtensorMapFloatOOBfill { CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE = 0 };
struct alignas(64) CUtensorMap {
    const unsigned char *base;       // global address of element (0, 0)
    uint64_t dim[2];                 // elements: {columns, rows}
    uint64_t row_stride;             // bytes between rows
    uint32_t box[2];                 // elements: {columns, rows}
    uint32_t elem;                   // bytes per element
    unsigned char opaque[128 - 52];
};
static_assert(sizeof(CUtensorMap) == 128, "a tensor map is 128 bytes");
