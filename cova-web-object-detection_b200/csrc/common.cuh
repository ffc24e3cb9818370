// Shared helpers for libcova_b200.so (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/cova_b200.h"

namespace cova {

void set_error(const char* fmt, ...);

#define COVA_REQUIRE(cond, ...)               \
  do {                                        \
    if (!(cond)) {                            \
      cova::set_error(__VA_ARGS__);           \
      return COVA_ERR_ARG;                    \
    }                                         \
  } while (0)

#define COVA_CUDA_OK(expr)                                                                  \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      cova::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return COVA_ERR_CUDA;                                                                 \
    }                                                                                       \
  } while (0)

#define COVA_LAUNCH_OK() COVA_CUDA_OK(cudaGetLastError())

int sm_count();          // of the current device (cached)
int max_smem_optin();    // bytes

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// split-bf16: hi = bf16(x) (RNE), lo = bf16(x - hi)
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

__device__ __forceinline__ uint32_t pack_bf16x2(__nv_bfloat16 a, __nv_bfloat16 b) {
  return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}

__device__ __forceinline__ float bf16lo_to_f32(uint32_t packed) { return __uint_as_float(packed << 16); }
__device__ __forceinline__ float bf16hi_to_f32(uint32_t packed) { return __uint_as_float(packed & 0xffff0000u); }

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace cova
