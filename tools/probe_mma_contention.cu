// Hardware probe #4 (development tool): what slows tcgen05.mma down INSIDE the stem kernel?  In isolation an
// M=128 x N=128 x K=16 MMA costs 64 clk (tools/probe_mma_rate.cu, any operand layout); in the stem the same instruction
// stream ran at ~110-150 clk per MMA.  One CTA per SM, 25 warps like the stem: warp 0 issues GROUPS x 14 MMAs on the
// stem's SWIZZLE_NONE sliding-window views (alternating accumulators, first MMA of a group overwrites); the other 24
// warps run one kind of background activity selected by MODE bits.  clock64 from first issue to the last commit.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_bf16.h>
#include "../cova-web-object-detection_b200/csrc/ptx.cuh"
using namespace cova;

__device__ __forceinline__ uint64_t desc_noswz(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

enum { M_COMMIT = 1, M_TMEM_LD = 2, M_SPIN = 4, M_STS = 8, M_ALU = 16, M_LDG = 32, M_LDS = 64, M_COMMIT_WAIT = 128, M_RANDOM = 256 };

template <int MODE>
__global__ void __launch_bounds__(800, 1) contention_kernel(int groups, long long* out, const float* gsrc, float* gdst) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar_done, bar_never, bar_grp[16], bar_grp2[16];
  __shared__ uint32_t tmem_base_s;
  __shared__ volatile int done;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 160 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (MODE & M_RANDOM) {   // operands with realistic bit activity: bf16 values in (-1, 1)
    for (int i = threadIdx.x; i < 160 * 1024 / 2; i += blockDim.x) {
      uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
      reinterpret_cast<__nv_bfloat16*>(smem)[i] = __float2bfloat16(((int)(h & 0xffff) - 32768) * (1.f / 32768.f));
    }
  }
  ptx::fence_proxy_async();
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar_done, 1); ptx::mbar_init(&bar_never, 1);
    for (int i = 0; i < 16; ++i) { ptx::mbar_init(&bar_grp[i], 1); ptx::mbar_init(&bar_grp2[i], 1); }
    done = 0;
    ptx::fence_barrier_init();
  }
  if (warp == 0) { ptx::tmem_alloc(&tmem_base_s, 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  long long t0 = 0, t1 = 0;
  float sink = 0.f;
  if (warp == 0) {
    constexpr uint32_t idesc = ptx::umma_idesc_bf16(128, 128);
    const uint64_t da0 = desc_noswz(ptx::smem_u32(smem), 16, 128);                 // ring: 16 rows x 2112 B
    const uint64_t db0 = desc_noswz(ptx::smem_u32(smem) + 64 * 1024, 2048, 128);   // filter: 28 chunks x 2048 B
    t0 = clock64();
    for (int g = 0; g < groups; ++g) {
      if (MODE & M_COMMIT_WAIT) { if (g >= 2) ptx::mbar_wait(&bar_grp[(g - 2) & 15], ((g - 2) >> 4) & 1); }
      if (ptx::elect_one()) {
        const uint32_t d = tmem + (g & 1) * 128;
#pragma unroll
        for (int i = 0; i < 14; ++i) {
          const uint32_t ao = (((i >> 1) * 2112 + (i & 1) * 32) >> 4);
          const uint32_t bo = ((i * 4096) >> 4);
          ptx::umma_bf16(d, da0 + ao, db0 + bo, idesc, i != 0);
        }
        if (MODE & (M_COMMIT | M_COMMIT_WAIT)) { ptx::umma_commit(&bar_grp[g & 15]); ptx::umma_commit(&bar_grp2[g & 15]); }
      }
      __syncwarp();
    }
    if (ptx::elect_one()) ptx::umma_commit(&bar_done);
    __syncwarp();
    ptx::mbar_wait(&bar_done, 0);
    t1 = clock64();
    if (lane == 0) { done = 1; ptx::mbar_arrive(&bar_never); }
  } else if (warp <= 16) {
    if (MODE & M_TMEM_LD) {
      uint32_t raw[16];
      const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + ((warp - 1) >> 2) * 16 + 256;   // columns the MMA does not touch
      while (!done) {
#pragma unroll 1
        for (int k = 0; k < 8; ++k) { ptx::tmem_ld16(taddr, raw); ptx::tmem_ld_wait(); sink += __uint_as_float(raw[k]); }
      }
    } else if (MODE & M_SPIN) {
      while (!ptx::mbar_try_wait(&bar_never, 0)) {}
    } else if (MODE & M_ALU) {
      float a = lane, b2 = 1.0001f;
      while (!done) {
#pragma unroll
        for (int k = 0; k < 64; ++k) a = fmaf(a, b2, 0.5f);
      }
      sink += a;
    } else if (MODE & M_LDS) {
      const float4* src = reinterpret_cast<const float4*>(smem + 128 * 1024) + lane;   // conflict-free 512-byte reads
      while (!done) {
#pragma unroll
        for (int k = 0; k < 16; ++k) { const float4 v = src[k * 32]; sink += v.x; }
      }
    }
  } else {
    if (MODE & M_STS) {
      uint4* dst = reinterpret_cast<uint4*>(smem + 144 * 1024) + lane;                 // conflict-free 512-byte stores
      uint32_t k = 0;
      while (!done) {
#pragma unroll
        for (int u = 0; u < 16; ++u) dst[((k + u) & 15) * 32] = make_uint4(k, k, k, k);
        ptx::fence_proxy_async();
        ++k;
      }
    } else if (MODE & M_LDG) {
      size_t off = ((size_t)blockIdx.x * 8 + (warp - 17)) * 1024 * 1024 / 4 + lane * 4;
      while (!done) {
#pragma unroll
        for (int u = 0; u < 8; ++u) { const float4 v = __ldg(reinterpret_cast<const float4*>(gsrc + off + u * 128)); sink += v.x; }
        off = (off + 1024) % (size_t)(148 * 8 * 1024 * 1024 / 4 - 2048);
      }
    }
  }
  if (sink == 123.456f) gdst[threadIdx.x] = sink;
  if (threadIdx.x == 0) { out[blockIdx.x * 2] = t0; out[blockIdx.x * 2 + 1] = t1; }
  ptx::tc_fence_before(); __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, 512);
}

template <int MODE>
static void run(const char* what, long long* d, const float* gsrc, float* gdst) {
  const int groups = 400, grid = 148;
  cudaFuncSetAttribute(contention_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 170 * 1024);
  contention_kernel<MODE><<<grid, 800, 170 * 1024>>>(groups, d, gsrc, gdst);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: %s\n", what, cudaGetErrorString(e)); exit(1); }
  std::vector<long long> h(grid * 2);
  cudaMemcpy(h.data(), d, grid * 2 * sizeof(long long), cudaMemcpyDeviceToHost);
  double worst = 0;
  for (int b = 0; b < grid; ++b) worst = std::max(worst, (double)(h[2 * b + 1] - h[2 * b]) / (groups * 14));
  printf("%-78s : %6.1f clk/MMA (%5.0f clk per 14-MMA tile)\n", what, worst, worst * 14);
  fflush(stdout);
}

int main() {
  long long* d; cudaMalloc(&d, 148 * 2 * sizeof(long long));
  float *gsrc, *gdst;
  cudaMalloc(&gsrc, (size_t)148 * 8 * 1024 * 1024); cudaMemset(gsrc, 0, (size_t)148 * 8 * 1024 * 1024);
  cudaMalloc(&gdst, 4096);
  run<0>("14 x (M128 N128 K16) per tile, idle background warps", d, gsrc, gdst);
  run<M_COMMIT>("+ 2 tcgen05.commit per tile", d, gsrc, gdst);
  run<M_COMMIT_WAIT>("+ 2 commits per tile and the issuer waits for tile t-2 (accumulator hand-back)", d, gsrc, gdst);
  run<M_COMMIT | M_TMEM_LD>("+ commits, 16 warps looping tcgen05.ld", d, gsrc, gdst);
  run<M_COMMIT | M_SPIN>("+ commits, 16 warps spinning on mbarrier.try_wait", d, gsrc, gdst);
  run<M_COMMIT | M_ALU>("+ commits, 16 warps of dependent FFMA (issue pressure)", d, gsrc, gdst);
  run<M_COMMIT | M_LDS>("+ commits, 16 warps of conflict-free LDS.128", d, gsrc, gdst);
  run<M_COMMIT | M_STS>("+ commits, 8 warps of conflict-free STS.128 + fence.proxy.async", d, gsrc, gdst);
  run<M_COMMIT | M_LDG>("+ commits, 8 warps streaming LDG.128 from HBM", d, gsrc, gdst);
  run<M_COMMIT | M_TMEM_LD | M_STS>("+ commits, tcgen05.ld warps and STS warps", d, gsrc, gdst);
  run<M_COMMIT | M_ALU | M_LDG>("+ commits, FFMA warps and LDG warps", d, gsrc, gdst);
  run<M_RANDOM>("random bf16 operands (not zeros), idle background warps", d, gsrc, gdst);
  run<M_RANDOM | M_COMMIT_WAIT>("random operands + commits + accumulator hand-back wait", d, gsrc, gdst);
  run<M_RANDOM | M_COMMIT_WAIT | M_TMEM_LD | M_STS>("random operands + commits + hand-back + tcgen05.ld warps + STS warps", d, gsrc, gdst);
  run<M_RANDOM | M_COMMIT_WAIT | M_ALU | M_LDG>("random operands + commits + hand-back + FFMA warps + LDG warps", d, gsrc, gdst);
  return 0;
}
