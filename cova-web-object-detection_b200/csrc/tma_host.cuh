// Host-side TMA descriptor (CUtensorMap) construction.  cuTensorMapEncodeTiled is fetched through
// cudaGetDriverEntryPoint so libcova_b200.so links against cudart only (it must also load on a CPU-only box).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "common.cuh"

namespace cova {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// bf16 tensor, `rank` dims (innermost first), 128-byte swizzle, zero fill outside the tensor.
// dims/box in elements; strides in BYTES for dims 1..rank-1.
inline int make_tmap_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_b,
                          const uint32_t* box) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return COVA_ERR_CUDA;
  }
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_b[i];
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return COVA_ERR_CUDA;
  }
  return COVA_OK;
}

// fp32 matrix [rows][cols] (row pitch in bytes, a multiple of 16), box {box_cols, box_rows}, NO swizzle, zero fill outside:
// the raw activation tiles of linear_tc.cu (converted to split planes on the way from shared memory to shared memory).
inline int make_tmap_f32_2d(CUtensorMap* map, const void* base, uint64_t cols, uint64_t rows, uint64_t pitch_bytes,
                            uint32_t box_cols, uint32_t box_rows) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return COVA_ERR_CUDA;
  }
  cuuint64_t gd[2] = {cols, rows}, gs[1] = {pitch_bytes};
  cuuint32_t bx[2] = {box_cols, box_rows}, es[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(fp32) failed with CUresult %d", (int)r);
    return COVA_ERR_CUDA;
  }
  return COVA_OK;
}

}  // namespace cova
