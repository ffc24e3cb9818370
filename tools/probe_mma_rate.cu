// Hardware probe #3 (development tool): tcgen05.mma throughput per instruction shape, operands in shared memory.
// One CTA per SM; one elected thread issues REPS x 32 fully unrolled M=128 x N x K=16 bf16 MMAs (descriptor =
// uniform base + compile-time offset, i.e. the cheapest possible issue sequence), one accumulator chain, then a
// commit; clock64 from first issue to commit arrival.  Answers: what does an N = 64 MMA really cost?
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../cova-web-object-detection_b200/csrc/ptx.cuh"
using namespace cova;

// SWIZZLE_NONE K-major descriptor: lbo = stride between the two 8-element K chunks, sbo = stride between 8-row groups
__device__ __forceinline__ uint64_t desc_noswz(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// AM / BM: operand layouts.  0 = SWIZZLE_128B K-major (the 3x3 conv kernels); 1 = SWIZZLE_NONE as the stem uses it (A: the
// sliding-window view, rows 16 B apart, K chunks 16 B apart, i.e. overlapping core matrices; B: [chunk][cout][8], LBO 2048);
// 2 = SWIZZLE_NONE canonical interleaved (128-byte core matrices, LBO 128, SBO 256)
template <int N, bool VARY, int AM = 0, int BM = 0>
__global__ void __launch_bounds__(128, 1) rate_kernel(int reps, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 192 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  ptx::fence_proxy_async();
  if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); }
  if (warp == 0) { ptx::tmem_alloc(&tmem_base_s, 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  long long t0 = 0, t1 = 0;
  if (warp == 0) {
    constexpr uint32_t idesc = ptx::umma_idesc_bf16(128, N);
    const uint64_t da0 = AM == 0 ? ptx::umma_desc_sw128(ptx::smem_u32(smem), 1280)
                       : AM == 1 ? desc_noswz(ptx::smem_u32(smem), 16, 128) : desc_noswz(ptx::smem_u32(smem), 128, 256);
    const uint64_t db0 = BM == 0 ? ptx::umma_desc_sw128(ptx::smem_u32(smem) + 64 * 1024, 1024)
                       : BM == 1 ? desc_noswz(ptx::smem_u32(smem) + 64 * 1024, 2048, 128)
                                 : desc_noswz(ptx::smem_u32(smem) + 64 * 1024, 128, 256);
    if (ptx::elect_one()) {
      t0 = clock64();
      for (int rep = 0; rep < reps; ++rep) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          // stem-like views: 14 (row, half) steps: ring rows 2112 B apart, halves 32 B apart; filter chunks 4096 B apart
          const uint32_t ao = !VARY ? 0 : AM == 1 ? ((((i % 14) >> 1) * 2112 + (i & 1) * 32) >> 4) : AM == 2 ? (((i & 7) * 4096) >> 4)
                                                  : (((i & 3) * 32 + (i >> 2) * 128) >> 4);
          const uint32_t bo = !VARY ? 0 : BM != 0 ? (((i % 14) * 4096) >> 4) : (((i & 3) * 32 + (i >> 3) * (N * 128)) >> 4);
          ptx::umma_bf16(tmem, da0 + ao, db0 + bo, idesc, 1);
        }
      }
      ptx::umma_commit(&bar);
    }
    __syncwarp();
  }
  ptx::mbar_wait(&bar, 0);
  t1 = clock64();
  if (threadIdx.x == 0) { out[blockIdx.x * 2] = t0; out[blockIdx.x * 2 + 1] = t1; }
  ptx::tc_fence_before(); __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, 512);
}

template <int N, bool VARY, int AM = 0, int BM = 0>
static void run(int grid, long long* d) {
  const int reps = 64;
  cudaFuncSetAttribute(rate_kernel<N, VARY, AM, BM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  rate_kernel<N, VARY, AM, BM><<<grid, 128, 200 * 1024>>>(reps, d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("N=%d: %s\n", N, cudaGetErrorString(e)); exit(1); }
  std::vector<long long> h(grid * 2);
  cudaMemcpy(h.data(), d, grid * 2 * sizeof(long long), cudaMemcpyDeviceToHost);
  double worst = 0;
  for (int b = 0; b < grid; ++b) worst = std::max(worst, (double)(h[2 * b + 1] - h[2 * b]) / (reps * 32));
  static const char* lay[3] = {"SW128", "NOSWZ stem view", "NOSWZ interleaved"};
  printf("A %-17s B %-17s ", lay[AM], lay[BM]);
  printf("grid %3d  M=128 N=%3d K=16 %s : %6.1f clk/MMA  (%5.0f MAC/clk/SM = %4.1f%% of 4096, smem operands %4.0f B/clk)\n", grid, N,
         VARY ? "varying operand views" : "one fixed operand    ", worst, 128.0 * N * 16 / worst, 100.0 * 128 * N * 16 / worst / 4096,
         (4096.0 + N * 32) / worst);
}

int main() {
  long long* d; cudaMalloc(&d, 148 * 2 * sizeof(long long));
  for (int grid : {1, 148}) {
    run<8, false>(grid, d);  run<16, true>(grid, d);  run<32, true>(grid, d);
    run<64, false>(grid, d); run<64, true>(grid, d);
    run<128, false>(grid, d); run<128, true>(grid, d);
    run<192, true>(grid, d); run<256, false>(grid, d); run<256, true>(grid, d);
  }
  // operand layouts (round 2): what does the stem's SWIZZLE_NONE sliding-window operand cost per MMA?
  run<128, true, 1, 1>(148, d); run<128, false, 1, 1>(148, d); run<64, true, 1, 1>(148, d);
  run<128, true, 1, 0>(148, d); run<128, true, 0, 1>(148, d);
  run<128, true, 2, 2>(148, d); run<64, true, 2, 2>(148, d); run<128, true, 2, 0>(148, d); run<128, true, 0, 2>(148, d);
  run<256, true, 1, 1>(148, d); run<256, true, 1, 0>(148, d);
  return 0;
}
