// Row-major linear layer, CUDA-core fp32 engine:  Y[M,N] = act((X[M,K] @ W[N,K]^T + bias) * scale + shift).
// Replaces `nn.Linear` (+ folded `nn.BatchNorm1d`, + ReLU) of `/root/reference/models.py:160-164` (GAT W_i/W_j),
// `:85-87` (decoder.1 + decoder.2 + ReLU) and `:89` (decoder.5).  Exact-fp32 engine / on-device check for the
// tcgen05 GEMM; 64x64 tile, 16-deep K slices through shared memory, 4x4 register tile per thread.
#include "common.cuh"

namespace cova {

constexpr int LS_BM = 64, LS_BN = 64, LS_BK = 16, LS_THREADS = 256, LS_LD = LS_BM + 4;

template <bool VEC>
__global__ void __launch_bounds__(LS_THREADS)
linear_simt_kernel(const float* __restrict__ X, int64_t ldx, int M, int K, const float* __restrict__ Wt, int N,
                   const float* __restrict__ bias, const float* __restrict__ scale, const float* __restrict__ shift,
                   const float* __restrict__ res, int64_t ldr, int relu, float* __restrict__ Y, int64_t ldy) {
  __shared__ __align__(16) float As[LS_BK][LS_LD];
  __shared__ __align__(16) float Bs[LS_BK][LS_LD];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * LS_BM, n0 = blockIdx.x * LS_BN;
  const int lr = tid >> 2, lk = (tid & 3) * 4;   // this thread stages row lr, k-quad lk of both tiles
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += LS_BK) {
    float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
    const int gm = m0 + lr, gn = n0 + lr, gk = k0 + lk;
    if (VEC) {
      if (gm < M && gk < K) {
        float4 v = __ldg(reinterpret_cast<const float4*>(X + (size_t)gm * ldx + gk));
        a[0] = v.x; a[1] = v.y; a[2] = v.z; a[3] = v.w;
      }
      if (gn < N && gk < K) {
        float4 v = __ldg(reinterpret_cast<const float4*>(Wt + (size_t)gn * K + gk));
        b[0] = v.x; b[1] = v.y; b[2] = v.z; b[3] = v.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (gm < M && gk + j < K) a[j] = __ldg(X + (size_t)gm * ldx + gk + j);
        if (gn < N && gk + j < K) b[j] = __ldg(Wt + (size_t)gn * K + gk + j);
      }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      As[lk + j][lr] = a[j];
      Bs[lk + j][lr] = b[j];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < LS_BK; ++k) {
      const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float ar[4] = {av.x, av.y, av.z, av.w}, br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (bias) v += bias[n];
      if (scale) v = fmaf(v, scale[n], shift[n]);
      if (res) v += res[(size_t)m * ldr + n];
      if (relu) v = fmaxf(v, 0.f);
      Y[(size_t)m * ldy + n] = v;
    }
  }
}

// Skinny outputs (N <= 8: decoder.5 = Linear(992, 4), `models.py:89`): one warp per row of X, every lane 4 consecutive k of
// each 128-wide slice, N dot products per lane, warp-shuffle reduction.  Exact fp32 FMAs on both engines - the tcgen05 engine
// passes its packed split-bf16 filter [2][N][K], summed back to hi + lo here.  A 128 x 32 tensor-core tile for four output
// columns spent ~15 us on pipeline latency for 5.7 MFLOP (GPU-side, CUDA-graph replay); this kernel 8.7 us.
constexpr int LSK_NMAX = 8;
template <bool PACKED>
__global__ void __launch_bounds__(256)
linear_skinny_kernel(const float* __restrict__ X, int64_t ldx, int M, int K, const void* __restrict__ Wv, int N,
                     const float* __restrict__ bias, const float* __restrict__ scale, const float* __restrict__ shift,
                     const float* __restrict__ res, int64_t ldr, int relu, float* __restrict__ Y, int64_t ldy) {
  const int m = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (m >= M) return;
  float acc[LSK_NMAX];
#pragma unroll
  for (int n = 0; n < LSK_NMAX; ++n) acc[n] = 0.f;
  const float* xr = X + (size_t)m * ldx;
  for (int k = lane * 4; k < K; k += 128) {   // (unrolling this loop by 4 measured slower: 11.6 vs 8.7 us at M = 1440, K = 992)
    const float4 xv = __ldg(reinterpret_cast<const float4*>(xr + k));
#pragma unroll
    for (int n = 0; n < LSK_NMAX; ++n) {
      if (n < N) {
        float4 wv;
        if (PACKED) {
          const __nv_bfloat16* wh = reinterpret_cast<const __nv_bfloat16*>(Wv) + (size_t)n * K + k;
          const uint2 h = __ldg(reinterpret_cast<const uint2*>(wh)), l = __ldg(reinterpret_cast<const uint2*>(wh + (size_t)N * K));
          wv = make_float4(bf16lo_to_f32(h.x) + bf16lo_to_f32(l.x), bf16hi_to_f32(h.x) + bf16hi_to_f32(l.x),
                           bf16lo_to_f32(h.y) + bf16lo_to_f32(l.y), bf16hi_to_f32(h.y) + bf16hi_to_f32(l.y));
        } else {
          wv = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(Wv) + (size_t)n * K + k));
        }
        acc[n] = fmaf(xv.x, wv.x, fmaf(xv.y, wv.y, fmaf(xv.z, wv.z, fmaf(xv.w, wv.w, acc[n]))));
      }
    }
  }
#pragma unroll
  for (int n = 0; n < LSK_NMAX; ++n) {
    if (n < N) {
      float v = acc[n];
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == n) {
        if (bias) v += bias[n];
        if (scale) v = v * scale[n] + shift[n];
        if (res) v += res[(size_t)m * ldr + n];
        if (relu) v = fmaxf(v, 0.f);
        Y[(size_t)m * ldy + n] = v;
      }
    }
  }
}

bool linear_skinny_supported(const float* x, int64_t ld_x, int K, int N, const void* w) {
  return N <= LSK_NMAX && (K % 4 == 0) && (ld_x % 4 == 0) && (((uintptr_t)x & 15) == 0) && (((uintptr_t)w & 15) == 0);
}

int linear_skinny(const float* x, int64_t ld_x, int M, int K, const void* w, bool packed, int N, const float* bias,
                  const float* scale, const float* shift, const float* res, int64_t ld_res, int relu, float* y,
                  int64_t ld_y, cudaStream_t st) {
  if (packed)
    linear_skinny_kernel<true><<<ceil_div(M, 8), 256, 0, st>>>(x, ld_x, M, K, w, N, bias, scale, shift, res, ld_res, relu, y, ld_y);
  else
    linear_skinny_kernel<false><<<ceil_div(M, 8), 256, 0, st>>>(x, ld_x, M, K, w, N, bias, scale, shift, res, ld_res, relu, y, ld_y);
  COVA_LAUNCH_OK();
  return COVA_OK;
}

int linear_simt(const float* x, int64_t ld_x, int M, int K, const float* w, int N, const float* bias,
                const float* scale, const float* shift, const float* res, int64_t ld_res, int relu, float* y,
                int64_t ld_y, cudaStream_t st) {
  dim3 grid(ceil_div(N, LS_BN), ceil_div(M, LS_BM));
  const bool vec = (K % 4 == 0) && (ld_x % 4 == 0) && (((uintptr_t)x & 15) == 0) && (((uintptr_t)w & 15) == 0);
  if (vec)
    linear_simt_kernel<true><<<grid, LS_THREADS, 0, st>>>(x, ld_x, M, K, w, N, bias, scale, shift, res, ld_res, relu, y, ld_y);
  else
    linear_simt_kernel<false><<<grid, LS_THREADS, 0, st>>>(x, ld_x, M, K, w, N, bias, scale, shift, res, ld_res, relu, y, ld_y);
  COVA_LAUNCH_OK();
  return COVA_OK;
}

}  // namespace cova
