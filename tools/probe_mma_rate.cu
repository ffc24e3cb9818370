// Hardware probe #3 (development tool): tcgen05.mma throughput per instruction shape, operands in shared memory.
// One CTA per SM; one elected thread issues REPS x 32 fully unrolled M=128 x N x K=16 bf16 MMAs (descriptor =
// uniform base + compile-time offset, i.e. the cheapest possible issue sequence), one accumulator chain, then a
// commit; clock64 from first issue to commit arrival.  Answers: what does an N = 64 MMA really cost?
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../cova-web-object-detection_b200/csrc/ptx.cuh"
using namespace cova;

template <int N, bool VARY>
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
    const uint64_t da0 = ptx::umma_desc_sw128(ptx::smem_u32(smem), 1280);
    const uint64_t db0 = ptx::umma_desc_sw128(ptx::smem_u32(smem) + 64 * 1024, 1024);
    if (ptx::elect_one()) {
      t0 = clock64();
      for (int rep = 0; rep < reps; ++rep) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const uint32_t ao = VARY ? (((i & 3) * 32 + (i >> 2) * 128) >> 4) : 0;
          const uint32_t bo = VARY ? (((i & 3) * 32 + (i >> 3) * (N * 128)) >> 4) : 0;
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

template <int N, bool VARY>
static void run(int grid, long long* d) {
  const int reps = 64;
  cudaFuncSetAttribute(rate_kernel<N, VARY>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  rate_kernel<N, VARY><<<grid, 128, 200 * 1024>>>(reps, d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("N=%d: %s\n", N, cudaGetErrorString(e)); exit(1); }
  std::vector<long long> h(grid * 2);
  cudaMemcpy(h.data(), d, grid * 2 * sizeof(long long), cudaMemcpyDeviceToHost);
  double worst = 0;
  for (int b = 0; b < grid; ++b) worst = std::max(worst, (double)(h[2 * b + 1] - h[2 * b]) / (reps * 32));
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
  return 0;
}
