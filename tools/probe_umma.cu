// Hardware probe (development tool, not part of the product library): which tcgen05 shared-memory descriptor
// forms address a K-major SWIZZLE_128B operand whose start is NOT 1024-byte aligned, and whose 8-row groups are
// not 1024 B apart?  Decides whether a conv tile can load ONE halo patch and take all 9 taps as shifted views.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bin/probe_umma tools/probe_umma.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../cova-web-object-detection_b200/csrc/ptx.cuh"
#include "../cova-web-object-detection_b200/csrc/tma_host.cuh"

namespace cova {
void set_error(const char* fmt, ...) { fprintf(stderr, "error: %s\n", fmt); }
}
using namespace cova;

struct Case {
  int pitch_rows;    // smem rows between consecutive 8-row groups of the operand (8 = dense)
  int shift_rows;    // operand start, in 128-B rows from the (1024-aligned) buffer base
  int base_offset;   // descriptor base_offset field [49,52)
};

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, int a_rows, Case cs,
             float* __restrict__ D) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  unsigned char* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  unsigned char* sa = smem;                 // a_rows x 128 B
  unsigned char* sb = smem + 32 * 1024;     // 64 x 128 B
  __shared__ uint64_t bar_load, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar_load, 1);
    ptx::mbar_init(&bar_mma, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc(&tmem_base_s, 64);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (threadIdx.x == 0) {
    ptx::mbar_arrive_expect_tx(&bar_load, (uint32_t)(a_rows + 64) * 128u);
    ptx::tma_load_2d(sa, &tm_a, &bar_load, 0, 0);
    ptx::tma_load_2d(sb, &tm_b, &bar_load, 0, 0);
    ptx::mbar_wait(&bar_load, 0);
    ptx::tc_fence_after();
    const uint32_t idesc = ptx::umma_idesc_bf16(128, 64);
    for (int kk = 0; kk < 4; ++kk) {
      uint64_t da = ptx::umma_desc_sw128(ptx::smem_u32(sa) + cs.shift_rows * 128 + kk * 32, cs.pitch_rows * 128);
      da |= (uint64_t)(cs.base_offset & 7) << 49;
      const uint64_t db = ptx::umma_desc_sw128(ptx::smem_u32(sb) + kk * 32, 1024);
      ptx::umma_bf16(tmem, da, db, idesc, kk > 0);
    }
    ptx::umma_commit(&bar_mma);
  }
  ptx::mbar_wait(&bar_mma, 0);
  ptx::tc_fence_after();
  uint32_t v[4][16];
  for (int q = 0; q < 4; ++q) ptx::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + q * 16, v[q]);
  ptx::tmem_ld_wait();
  const int m = warp * 32 + lane;
  for (int q = 0; q < 4; ++q)
    for (int j = 0; j < 16; ++j) D[m * 64 + q * 16 + j] = __uint_as_float(v[q][j]);
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, 64);
}

static uint16_t f2bf(float f) { uint32_t u; memcpy(&u, &f, 4); return (uint16_t)(u >> 16); }

int main() {
  const int A_ROWS = 200;
  std::vector<uint16_t> hA(A_ROWS * 64), hB(64 * 64);
  std::vector<float> fA(A_ROWS * 64), fB(64 * 64);
  srand(7);
  for (size_t i = 0; i < hA.size(); ++i) { fA[i] = (float)(rand() % 5 - 2); hA[i] = f2bf(fA[i]); }
  for (size_t i = 0; i < hB.size(); ++i) { fB[i] = (float)(rand() % 5 - 2); hB[i] = f2bf(fB[i]); }
  void *dA, *dB; float* dD;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, 128 * 64 * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap ta, tb;
  const uint64_t ad[2] = {64, (uint64_t)A_ROWS}, bd[2] = {64, 64}, st[1] = {128};
  const uint32_t abox[2] = {64, (uint32_t)A_ROWS}, bbox[2] = {64, 64};
  if (make_tmap_bf16(&ta, dA, 2, ad, st, abox) || make_tmap_bf16(&tb, dB, 2, bd, st, bbox)) return 1;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
  const int pitches[] = {8, 10, 9};
  for (int pitch : pitches)
    for (int shift = 0; shift <= 11; ++shift)
      for (int bo_mode = 0; bo_mode < 2; ++bo_mode) {
        Case cs{pitch, shift, bo_mode ? (shift & 7) : 0};
        if (bo_mode && (shift & 7) == 0) continue;
        cudaMemset(dD, 0xff, 128 * 64 * 4);
        probe_kernel<<<1, 128, 48 * 1024>>>(ta, tb, A_ROWS, cs, dD);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("pitch %d shift %d bo %d: CUDA error %s\n", pitch, shift, cs.base_offset, cudaGetErrorString(e)); return 2; }
        std::vector<float> hD(128 * 64);
        cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
        // expected: operand row m = buffer row shift + (m/8)*pitch + m%8
        int bad = 0; double maxerr = 0;
        for (int m = 0; m < 128; ++m) {
          const int row = shift + (m / 8) * pitch + (m % 8);
          for (int n = 0; n < 64; ++n) {
            float ref = 0;
            for (int k = 0; k < 64; ++k) ref += fA[row * 64 + k] * fB[n * 64 + k];
            const double err = fabs((double)hD[m * 64 + n] - ref);
            if (err > 1e-3) ++bad;
            if (err > maxerr) maxerr = err;
          }
        }
        printf("pitch_rows %2d shift_rows %2d base_offset %d : %s (mismatches %d / 8192, max err %.1f)\n", pitch, shift,
               cs.base_offset, bad == 0 ? "OK" : "WRONG", bad, maxerr);
      }
  return 0;
}
