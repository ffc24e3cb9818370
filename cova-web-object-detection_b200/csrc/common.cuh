// Shared helpers for libcova_b200.so (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
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
int knob(int id, int dflt);                       // cova_set_knob value, or `dflt` when unset (< 0)
unsigned long long* debug_words(long long need);  // cova_debug_buffer pointer if it holds >= need words, else NULL

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// image pixel fetch: fp32 as is, or uint8 -> v/255 with IEEE division (== torchvision ToTensor, datasets.py:41-45)
template <bool U8>
__device__ __forceinline__ float load_pixel(const void* base, size_t idx) {
  if (U8) return __fdiv_rn((float)__ldg(reinterpret_cast<const unsigned char*>(base) + idx), 255.f);
  return __ldg(reinterpret_cast<const float*>(base) + idx);
}

// split-bf16: hi = bf16(x) (RNE), lo = bf16(x - hi)
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

// two values at once: hi = {bf16(a), bf16(b)} (a in the low half), lo = {bf16(a - hi_a), bf16(b - hi_b)}.
// One packed F2FP per plane instead of four single F2F conversions (which run on the slow conversion pipe).
__device__ __forceinline__ void split_bf16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  hi = *reinterpret_cast<uint32_t*>(&h);
  const float ra = a - __uint_as_float(hi << 16);
  const float rb = b - __uint_as_float(hi & 0xffff0000u);
  __nv_bfloat162 l = __floats2bfloat162_rn(ra, rb);
  lo = *reinterpret_cast<uint32_t*>(&l);
}
__device__ __forceinline__ uint32_t pack2_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// fp16 pairs (a in the low half), for the single-plane fp16 mode (COVA_F16)
// fp16 saturates at +-65504 instead of overflowing to inf (an inf would poison the whole accumulator row downstream)
__device__ __forceinline__ float sat_f16(float x) { return x != x ? x : fminf(fmaxf(x, -65504.f), 65504.f); }   // NaN stays NaN
__device__ __forceinline__ uint32_t pack2_f16(float a, float b) {
  __half2 h = __floats2half2_rn(sat_f16(a), sat_f16(b));
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack2_f16(uint32_t packed) {
  return __half22float2(*reinterpret_cast<const __half2*>(&packed));
}

// split-fp16 ("fp16x3" mode): hi = f16(x), lo = f16(x - hi) carries 22 significand bits (split-bf16: 16), so the same
// three tensor-core products reproduce an fp32 convolution to ~1e-6 instead of ~1e-5.  Filters are pre-scaled by
// SPLIT_F16_WSCALE (a power of two, undone exactly in the epilogue) so that their lo plane stays a NORMAL fp16 number.
constexpr float SPLIT_F16_WSCALE = 256.f;
__device__ __forceinline__ void split_f16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
  __half2 h = __floats2half2_rn(sat_f16(a), sat_f16(b));
  hi = *reinterpret_cast<uint32_t*>(&h);
  const float2 hf = __half22float2(h);
  __half2 l = __floats2half2_rn(sat_f16(a - hf.x), sat_f16(b - hf.y));
  lo = *reinterpret_cast<uint32_t*>(&l);
}

// 256-bit global accesses (sm_100: LDG/STG.E.256): one full 32-byte sector per lane
__device__ __forceinline__ void st_global_v8(void* p, const uint32_t (&v)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
               "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void ld_global_nc_v8(const void* p, uint32_t (&v)[8]) {
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "l"(p));
}

__device__ __forceinline__ void ld_global_v8(const void* p, uint32_t (&v)[8]) {
  asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "l"(p));
}
__device__ __forceinline__ void ld_global_na_v8(const void* p, uint32_t (&v)[8]) {
  asm volatile("ld.global.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "l"(p));
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

// Column sums of a [32 lanes][32 values] register tile: lane l returns the sum over the warp's lanes of v[l] (v is clobbered).
// 31 shuffles (16 + 8 + 4 + 2 + 1: every step halves the values a lane is responsible for) instead of 32 x 5.  Used by the
// convolution epilogues (lane = pixel, value = output channel) to accumulate the BatchNorm batch statistics of the tile they
// are writing, so that no separate pass has to re-read the raw convolution output.
__device__ __forceinline__ float warp_transpose_sum32(float (&v)[32], int lane) {
#pragma unroll
  for (int step = 16; step >= 1; step >>= 1) {
    const bool up = (lane & step) != 0;
#pragma unroll
    for (int i = 0; i < step; ++i) {
      const float send = up ? v[i] : v[i + step];
      const float keep = up ? v[i + step] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, step);
    }
  }
  return v[0];
}
// The same for 16 values per lane: lanes l and l ^ 16 are added first, then the 16 x 16 transpose-reduce; lane l (and l + 16)
// returns the total of v[l & 15].
__device__ __forceinline__ float warp_transpose_sum16(float (&v)[16], int lane) {
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], 16);
#pragma unroll
  for (int step = 8; step >= 1; step >>= 1) {
    const bool up = (lane & step) != 0;
#pragma unroll
    for (int i = 0; i < step; ++i) {
      const float send = up ? v[i] : v[i + step];
      const float keep = up ? v[i + step] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, step);
    }
  }
  return v[0];
}
__device__ __forceinline__ float round_bf16(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

}  // namespace cova
