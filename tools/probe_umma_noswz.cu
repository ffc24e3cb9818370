// Hardware probe #2 (development tool): SWIZZLE_NONE K-major tcgen05 operands written by ordinary threads.
//  (1) canonical interleaved layout: element (m,k) at (m%8)*16 + (m/8)*SBO + (k/8)*LBO + (k%8)*2 bytes -
//      which descriptor field is the M-group stride and which the K-chunk stride?
//  (2) "Toeplitz view": A is NOT materialised; it is a sliding window over one flat row buffer
//      (LBO = 16 B, SBO = 128 B, rows 16 B apart, so K-chunk 1 of row j aliases K-chunk 0 of row j+1).
//      This is the stride-2 7x7 stem conv read straight from an NHWC4 bf16 image row, with no im2col copy.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bin/probe_umma_noswz tools/probe_umma_noswz.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../cova-web-object-detection_b200/csrc/ptx.cuh"
using namespace cova;

__device__ __forceinline__ uint64_t desc_noswz(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;   // layout_type 0 = SWIZZLE_NONE
}

// a_bytes / b_bytes: raw smem images copied from global; descriptors given by (start offset, lbo, sbo)
__global__ void __launch_bounds__(128, 1)
probe_kernel(const uint4* __restrict__ a_img, int a_bytes, const uint4* __restrict__ b_img, int b_bytes, int a_off,
             int a_lbo, int a_sbo, int b_lbo, int b_sbo, int nk, int a_kstep, int b_kstep, float* __restrict__ D) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* sa = smem;
  unsigned char* sb = smem + 64 * 1024;
  __shared__ uint64_t bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < a_bytes / 16; i += 128) reinterpret_cast<uint4*>(sa)[i] = a_img[i];
  for (int i = threadIdx.x; i < b_bytes / 16; i += 128) reinterpret_cast<uint4*>(sb)[i] = b_img[i];
  ptx::fence_proxy_async();      // generic-proxy smem writes -> visible to the tensor core (async proxy)
  if (threadIdx.x == 0) {
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
    const uint32_t idesc = ptx::umma_idesc_bf16(128, 64);
    for (int kk = 0; kk < nk; ++kk) {
      const uint64_t da = desc_noswz(ptx::smem_u32(sa) + a_off + kk * a_kstep, a_lbo, a_sbo);
      const uint64_t db = desc_noswz(ptx::smem_u32(sb) + kk * b_kstep, b_lbo, b_sbo);
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

static int run(const char* name, const std::vector<uint16_t>& a_img, const std::vector<uint16_t>& b_img, int a_off,
               int a_lbo, int a_sbo, int b_lbo, int b_sbo, int nk, int a_kstep, int b_kstep,
               const std::vector<float>& ref) {
  void *dA, *dB; float* dD;
  cudaMalloc(&dA, a_img.size() * 2); cudaMalloc(&dB, b_img.size() * 2); cudaMalloc(&dD, 128 * 64 * 4);
  cudaMemcpy(dA, a_img.data(), a_img.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, b_img.data(), b_img.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0xff, 128 * 64 * 4);
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
  probe_kernel<<<1, 128, 128 * 1024>>>((const uint4*)dA, (int)a_img.size() * 2, (const uint4*)dB, (int)b_img.size() * 2,
                                      a_off, a_lbo, a_sbo, b_lbo, b_sbo, nk, a_kstep, b_kstep, dD);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: CUDA error %s\n", name, cudaGetErrorString(e)); exit(2); }
  std::vector<float> hD(128 * 64);
  cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int i = 0; i < 128 * 64; ++i) bad += fabs(hD[i] - ref[i]) > 1e-3;
  printf("%-70s : %s (mismatches %d / 8192)\n", name, bad == 0 ? "OK" : "WRONG", bad);
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
  return bad;
}

int main() {
  srand(11);
  const int K = 32;   // two K=16 MMAs
  // ---------- (1) canonical interleaved: A [K/8][128][8], B [K/8][64][8]
  std::vector<float> A(128 * K), B(64 * K), ref(128 * 64, 0.f);
  for (auto& v : A) v = (float)(rand() % 5 - 2);
  for (auto& v : B) v = (float)(rand() % 5 - 2);
  for (int m = 0; m < 128; ++m) for (int n = 0; n < 64; ++n) { float s = 0; for (int k = 0; k < K; ++k) s += A[m * K + k] * B[n * K + k]; ref[m * 64 + n] = s; }
  std::vector<uint16_t> ai(K / 8 * 128 * 8), bi(K / 8 * 64 * 8);
  for (int m = 0; m < 128; ++m) for (int k = 0; k < K; ++k) ai[(k / 8) * 1024 + m * 8 + k % 8] = f2bf(A[m * K + k]);
  for (int n = 0; n < 64; ++n) for (int k = 0; k < K; ++k) bi[(k / 8) * 512 + n * 8 + k % 8] = f2bf(B[n * K + k]);
  // per MMA (K=16 = 2 chunks): A chunk stride 2048 B, group stride 128 B; B chunk stride 1024 B, group stride 128 B
  run("interleaved: LBO = K-chunk stride, SBO = 8-row-group stride", ai, bi, 0, 2048, 128, 1024, 128, 2, 4096, 2048, ref);
  run("interleaved: LBO = 8-row-group stride, SBO = K-chunk stride (swapped)", ai, bi, 0, 128, 2048, 128, 1024, 2, 4096, 2048, ref);

  // ---------- (2) Toeplitz view: flat row of 4-channel pixels (8 B each); output px m, K index = s*4+c,
  //            s in [0,8): element = row[(2m + s)*4 + c]; two MMAs: s 0..3 (offset 0) and s 4..7 (offset 32 B)
  const int NPX = 2 * 128 + 8;
  std::vector<float> rowf(NPX * 4);
  for (auto& v : rowf) v = (float)(rand() % 5 - 2);
  std::vector<uint16_t> rowi(NPX * 4 + 64, 0);
  for (int i = 0; i < NPX * 4; ++i) rowi[i] = f2bf(rowf[i]);
  for (int m = 0; m < 128; ++m) for (int n = 0; n < 64; ++n) {
    float s = 0;
    for (int k = 0; k < 32; ++k) s += rowf[(2 * m + k / 4) * 4 + k % 4] * B[n * K + k];
    ref[m * 64 + n] = s;
  }
  for (int off = 0; off <= 48; off += 16) {
    // start offset `off` bytes = shift by off/8 input pixels: reference shifts accordingly
    std::vector<float> r2(128 * 64);
    for (int m = 0; m < 128; ++m) for (int n = 0; n < 64; ++n) {
      float s = 0;
      for (int k = 0; k < 32; ++k) s += rowf[(2 * m + off / 8 + k / 4) * 4 + k % 4] * B[n * K + k];
      r2[m * 64 + n] = s;
    }
    char nm[128];
    snprintf(nm, sizeof nm, "toeplitz view: A LBO=16 SBO=128 start +%d B, kstep 32 B", off);
    run(nm, rowi, bi, off, 16, 128, 1024, 128, 2, 32, 2048, r2);
  }
  return 0;
}
