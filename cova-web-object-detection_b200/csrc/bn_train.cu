// A2 / A9, training mode: BatchNorm2d with BATCH statistics (+ residual) (+ ReLU), forward and backward, and the
// stem's maxpool 3x3 s2 p1 forward / backward, on NHWC fp32 maps.  Replaces, on the autograd path
// (`/root/reference/train.py:45-60`), what torchvision's `BasicBlock.forward` / `Bottleneck.forward` run through
// `nn.BatchNorm2d` in `model.train()` (`models.py:49-51`, `train.py:27`): biased batch variance over (B,H,W) per
// channel, eps 1e-5, running statistics updated with momentum 0.1 (unbiased variance) - SURVEY.md row A2.
// Profiled motivation (tools/prof_train.py): cuDNN's NCHW spatial-BN kernels were 29 ms of a 61 ms train step and
// `max_pool_backward_nchw` another 4.4 ms; these are plain HBM-bound passes:
//   forward : reduce (sum x, sum x^2)  ->  finalize (mean, 1/sqrt(var+eps), running stats)  ->  apply (+res)(+ReLU)
//   backward: reduce (sum g, sum g*xhat), g = dy * [y > 0]  ->  apply dx = gamma*inv*(g - S1/M - xhat*S2/M), dres = g
// Layout: x [M = B*H*W pixels][C channels], C a power of two in [4, 1024]; a thread owns one float4 channel group and
// strides over pixels, so every warp load is 512 contiguous bytes.  Partial sums are fp32 per thread (a few hundred
// values), combined in double (smem tree + one double atomic per channel and CTA): var = E[x^2] - mean^2 is evaluated
// in double.
#include <float.h>

#include "common.cuh"

namespace cova {

constexpr int BN_THREADS = 256;

__device__ __forceinline__ float bn_val(float x, float mean, float inv, float g, float b) {
  return (x - mean) * inv * g + b;     // torch: (x - mean) * invstd * weight + bias
}

// fp32 float4 -> split planes (bf16 hi/lo, or fp16 hi/lo when f16 != 0); both formats are 16 bits per element
__device__ __forceinline__ void split4(const float4 v, __nv_bfloat16* hi, __nv_bfloat16* lo, size_t idx, int f16 = 0) {
  uint32_t h01, l01, h23, l23;
  if (f16) {
    split_f16x2(v.x, v.y, h01, l01);
    split_f16x2(v.z, v.w, h23, l23);
  } else {
    split_bf16x2(v.x, v.y, h01, l01);
    split_bf16x2(v.z, v.w, h23, l23);
  }
  *reinterpret_cast<uint2*>(hi + idx) = make_uint2(h01, h23);
  *reinterpret_cast<uint2*>(lo + idx) = make_uint2(l01, l23);
}

// Storage-type accessors: the training maps are fp32 (fp32-parity mode) or bf16 (the bf16 training mode, BASELINE config 3:
// half the bytes of every pass); 4 consecutive channels per access, arithmetic always in fp32.
template <typename T> __device__ __forceinline__ float4 ld4(const T* p, int64_t e);
template <> __device__ __forceinline__ float4 ld4<float>(const float* p, int64_t e) {
  return __ldg(reinterpret_cast<const float4*>(p + e));
}
template <> __device__ __forceinline__ float4 ld4<__nv_bfloat16>(const __nv_bfloat16* p, int64_t e) {
  const uint2 u = __ldg(reinterpret_cast<const uint2*>(p + e));
  return make_float4(bf16lo_to_f32(u.x), bf16hi_to_f32(u.x), bf16lo_to_f32(u.y), bf16hi_to_f32(u.y));
}
__device__ __forceinline__ void st4(float* p, int64_t e, const float4 v) { *reinterpret_cast<float4*>(p + e) = v; }
__device__ __forceinline__ void st4(__nv_bfloat16* p, int64_t e, const float4 v) {
  *reinterpret_cast<uint2*>(p + e) = make_uint2(pack2_bf16(v.x, v.y), pack2_bf16(v.z, v.w));
}

// power-of-two scale that brings max|x| into [2^target, 2^(target+1)); 1 for a zero / non-finite maximum
__device__ __forceinline__ float pow2_scale(unsigned int amax_bits, int target_log2) {
  const int e = (int)(amax_bits >> 23) & 0xff;
  if (e == 0 || e == 0xff) return 1.f;
  int se = 127 + target_log2 - (e - 127);
  se = min(max(se, 1), 254);
  return __uint_as_float((unsigned int)se << 23);
}

// Vector accessors: V channels per thread = one 16-byte access (4 fp32 / 8 bf16); a map with C % V != 0 falls back to the
// host-side check.  Arithmetic on float[V].
template <typename T> struct Vec { static constexpr int V = 16 / sizeof(T); };
template <typename T> __device__ __forceinline__ void ldv(const T* p, int64_t e, float (&v)[Vec<T>::V]);
template <> __device__ __forceinline__ void ldv<float>(const float* p, int64_t e, float (&v)[4]) {
  const float4 t = __ldg(reinterpret_cast<const float4*>(p + e));
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <> __device__ __forceinline__ void ldv<__nv_bfloat16>(const __nv_bfloat16* p, int64_t e, float (&v)[8]) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(p + e));
  v[0] = bf16lo_to_f32(u.x); v[1] = bf16hi_to_f32(u.x); v[2] = bf16lo_to_f32(u.y); v[3] = bf16hi_to_f32(u.y);
  v[4] = bf16lo_to_f32(u.z); v[5] = bf16hi_to_f32(u.z); v[6] = bf16lo_to_f32(u.w); v[7] = bf16hi_to_f32(u.w);
}
__device__ __forceinline__ void stv(float* p, int64_t e, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p + e) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void stv(__nv_bfloat16* p, int64_t e, const float (&v)[8]) {
  *reinterpret_cast<uint4*>(p + e) = make_uint4(pack2_bf16(v[0], v[1]), pack2_bf16(v[2], v[3]), pack2_bf16(v[4], v[5]), pack2_bf16(v[6], v[7]));
}
// V fp32 values -> the storage type TO (V may be 4 or 8: a bf16-storage kernel writing an fp32 map, or the reverse)
template <int V> __device__ __forceinline__ void stv_any(float* p, int64_t e, const float (&v)[V]) {
#pragma unroll
  for (int k = 0; k < V; k += 4) *reinterpret_cast<float4*>(p + e + k) = make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]);
}
template <int V> __device__ __forceinline__ void stv_any(__nv_bfloat16* p, int64_t e, const float (&v)[V]) {
#pragma unroll
  for (int k = 0; k < V; k += 4) *reinterpret_cast<uint2*>(p + e + k) = make_uint2(pack2_bf16(v[k], v[k + 1]), pack2_bf16(v[k + 2], v[k + 3]));
}
// V values of a map of type TI starting at element e, whatever V is (TI fp32 with V = 8: two 16-byte loads)
template <int V> __device__ __forceinline__ void ldv_any(const float* p, int64_t e, float (&v)[V]) {
#pragma unroll
  for (int k = 0; k < V; k += 4) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p + e + k));
    v[k] = t.x; v[k + 1] = t.y; v[k + 2] = t.z; v[k + 3] = t.w;
  }
}
template <int V> __device__ __forceinline__ void ldv_any(const __nv_bfloat16* p, int64_t e, float (&v)[V]) {
#pragma unroll
  for (int k = 0; k < V; k += 4) {
    const uint2 u = __ldg(reinterpret_cast<const uint2*>(p + e + k));
    v[k] = bf16lo_to_f32(u.x); v[k + 1] = bf16hi_to_f32(u.x); v[k + 2] = bf16lo_to_f32(u.y); v[k + 3] = bf16hi_to_f32(u.y);
  }
}

// BatchNorm as ONE fma per element: y = x * A + B with A = invstd * gamma, B = beta - mean * A.  Forward, the backward's mask
// recomputation and its reduction all use THIS expression (so the ReLU mask of the backward is bit-identical to the forward's);
// these passes sit at the SM's issue limit, not only at the HBM limit (~30 lane-instructions per element against 128 per
// clock and SM: 3e12 bf16 elements/s at 6 TB/s would need 2.5x the issue rate), so instructions per element are what counts.
__device__ __forceinline__ void bn_affine(float mean, float inv, float g, float b, float& A, float& B) {
  A = inv * g;
  B = fmaf(-mean, A, b);
}

// V per-channel parameters (global or shared memory, 16-byte aligned at c0) with 128-bit loads
template <int V> __device__ __forceinline__ void ldp(const float* p, int c0, float (&v)[V]) {
#pragma unroll
  for (int k = 0; k < V; k += 4) {
    const float4 t = *reinterpret_cast<const float4*>(p + c0 + k);
    v[k] = t.x; v[k + 1] = t.y; v[k + 2] = t.z; v[k + 3] = t.w;
  }
}

// MODE 0: a = sum x, b = sum x^2.   MODE 1: a = sum g, b = sum g * xhat with g = dy * [y > 0] (y recomputed).
// MODE 2: MODE 1 + wmax[c] = max |x - mean| per channel and wmax[C] = max |g| (bit patterns of non-negative floats): what
// bounds |dx| before dx exists, so that the apply pass can emit SCALED split-fp16 planes directly (bn_act_bwd_kernel<true>).
// TS = storage type of x / res, TG = type of dy; a thread owns V = 16 B / sizeof(TS) channels and strides over pixels.
template <int MODE, typename TS = float, typename TG = float, bool MASK = false>
__global__ void __launch_bounds__(BN_THREADS, 4)
bn_reduce_kernel(const TS* __restrict__ x, const TG* __restrict__ dy, const TS* __restrict__ res, int64_t M,
                 int C, const float* __restrict__ mean, const float* __restrict__ invstd,
                 const float* __restrict__ gamma, const float* __restrict__ beta, int relu, double* __restrict__ ws,
                 unsigned int* __restrict__ wmax = nullptr, const unsigned char* __restrict__ mask = nullptr) {
  constexpr int V = Vec<TS>::V;
  extern __shared__ __align__(16) float red[];                       // [2][rows][C]
  const int cvn = C / V, rows = BN_THREADS / cvn;
  const int cg = threadIdx.x % cvn, prow = threadIdx.x / cvn;
  const int c0 = V * cg;
  float a[V], b[V], mu[V], iv[V], ga[V], be[V], mx[V];
  float gm = 0.f;
#pragma unroll
  for (int k = 0; k < V; ++k) {
    a[k] = b[k] = mx[k] = 0.f;
    mu[k] = iv[k] = ga[k] = be[k] = 0.f;
  }
  if (MODE >= 1) {
    ldp<V>(mean, c0, mu); ldp<V>(invstd, c0, iv); ldp<V>(gamma, c0, ga); ldp<V>(beta, c0, be);
#pragma unroll
    for (int k = 0; k < V; ++k) {        // ga <- A, be <- B
      float A, B;
      bn_affine(mu[k], iv[k], ga[k], be[k], A, B);
      ga[k] = A; be[k] = B;
    }
  }
  const int64_t stride = (int64_t)gridDim.x * rows;
  int64_t p = (int64_t)blockIdx.x * rows + prow;
  if (prow < rows) {
    if (MODE == 0) {   // statistics pass: four independent loads in flight per thread (a plain loop ran at 4 TB/s)
      for (; p + 3 * stride < M; p += 4 * stride) {
        float v[4][V];
#pragma unroll
        for (int u = 0; u < 4; ++u) ldv<TS>(x, (p + u * stride) * C + c0, v[u]);
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int k = 0; k < V; ++k) { a[k] += v[u][k]; b[k] = fmaf(v[u][k], v[u][k], b[k]); }
      }
      for (; p < M; p += stride) {
        float v[V];
        ldv<TS>(x, p * C + c0, v);
#pragma unroll
        for (int k = 0; k < V; ++k) { a[k] += v[k]; b[k] = fmaf(v[k], v[k], b[k]); }
      }
    } else {
      for (; p < M; p += stride) {
        float xv[V], g[V], r[V];
        ldv<TS>(x, p * C + c0, xv);
        ldv_any<V>(dy, p * C + c0, g);
        // ReLU decision: the forward's bit mask (one byte per thread and pixel) when it was kept - the residual map is then
        // not read at all - otherwise y is recomputed from x (and res)
        const unsigned bits = (MASK && relu) ? (unsigned)__ldg(mask + p * cvn + cg) : 0u;
        if (!MASK && relu && res != nullptr) ldv<TS>(res, p * C + c0, r);
#pragma unroll
        for (int k = 0; k < V; ++k) {
          // bf16 storage: accumulate sum g*x and form sum g*xhat = inv (sum g*x - mean sum g) once per thread below (two
          // instructions and 16 registers less per element; the cancellation costs |mean|/std ulps of fp32, irrelevant next to
          // the bf16 maps).  fp32 storage keeps the centred product.
          const float xh = sizeof(TS) == 2 ? xv[k] : (xv[k] - mu[k]) * iv[k];
          if (MASK) {
            if (relu) g[k] = ((bits >> k) & 1u) ? g[k] : 0.f;
          } else if (relu) {
            float y = fmaf(xv[k], ga[k], be[k]);
            if (res != nullptr) y += r[k];
            g[k] = y > 0.f ? g[k] : 0.f;
          }
          a[k] += g[k];
          b[k] = fmaf(g[k], xh, b[k]);
          if (MODE == 2) {
            mx[k] = fmaxf(mx[k], fabsf(xv[k] - mu[k]));
            gm = fmaxf(gm, fabsf(g[k]));
          }
        }
      }
    }
  }
  float* ra = red + (size_t)prow * C + c0;
  float* rb = red + (size_t)(rows + prow) * C + c0;
  if (MODE >= 1 && sizeof(TS) == 2) {
#pragma unroll
    for (int k = 0; k < V; ++k) b[k] = invstd[c0 + k] * fmaf(-mean[c0 + k], a[k], b[k]);
  }
#pragma unroll
  for (int k = 0; k < V; ++k) { ra[k] = a[k]; rb[k] = b[k]; }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * C; c += BN_THREADS) {
    const int which = c / C, ch = c % C;
    double s = 0.0;
    for (int r = 0; r < rows; ++r) s += (double)red[(size_t)(which * rows + r) * C + ch];
    atomicAdd(ws + c, s);
  }
  if (MODE == 2) {
    __syncthreads();
#pragma unroll
    for (int k = 0; k < V; ++k) ra[k] = mx[k];
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += BN_THREADS) {
      float m = 0.f;
      for (int r = 0; r < rows; ++r) m = fmaxf(m, red[(size_t)r * C + c]);
      atomicMax(wmax + c, __float_as_uint(m));
    }
    gm = warp_max(gm);
    if ((threadIdx.x & 31) == 0) atomicMax(wmax + C, __float_as_uint(gm));
  }
}

// mean / invstd from the sums, running statistics like torch (momentum on the UNBIASED variance)
__global__ void bn_finalize_kernel(const double* __restrict__ ws, int64_t M, int C, float eps, float momentum,
                                   float* __restrict__ mean, float* __restrict__ invstd, float* __restrict__ rmean,
                                   float* __restrict__ rvar) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double m = ws[c] / (double)M;
  double var = ws[C + c] / (double)M - m * m;
  if (var < 0.0) var = 0.0;
  mean[c] = (float)m;
  invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  if (rmean != nullptr) {
    const double unb = M > 1 ? var * (double)M / (double)(M - 1) : var;
    rmean[c] = (1.f - momentum) * rmean[c] + momentum * (float)m;
    rvar[c] = (1.f - momentum) * rvar[c] + momentum * (float)unb;
  }
}

// TS = storage type of x / res, TY = type of y; V = 16 B / sizeof(TS) channels per thread.  In the bf16 training mode the
// backward recomputes the ReLU mask exactly as it is computed here: bn_val in fp32, + res, compare with 0, THEN the store rounds
// (a positive value never rounds to <= 0).
template <typename TS, typename TY>
__global__ void __launch_bounds__(BN_THREADS)
bn_act_fwd_kernel(const TS* __restrict__ x, int64_t nv, int C, const float* __restrict__ mean,
                  const float* __restrict__ invstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                  const TS* __restrict__ res, int relu, TY* __restrict__ y, __nv_bfloat16* __restrict__ y_hi,
                  __nv_bfloat16* __restrict__ y_lo, int f16, unsigned char* __restrict__ mask = nullptr) {
  constexpr int V = Vec<TS>::V;
  extern __shared__ __align__(16) float tab[];                       // [2][C]: A, B
  const int cvn = C / V;
  for (int c = threadIdx.x; c < C; c += BN_THREADS) bn_affine(mean[c], invstd[c], gamma[c], beta[c], tab[c], tab[C + c]);
  __syncthreads();
  // descending order: the statistics pass that ran just before this one read the map ascending, so its tail is what L2 still holds
  for (int64_t i = nv - 1 - ((int64_t)blockIdx.x * BN_THREADS + threadIdx.x); i >= 0; i -= (int64_t)gridDim.x * BN_THREADS) {
    const int c0 = V * (int)(i % cvn);
    float xv[V], r[V], o[V], A[V], B[V];
    ldv<TS>(x, i * V, xv);
    if (res != nullptr) ldv<TS>(res, i * V, r);
    ldp<V>(tab, c0, A); ldp<V>(tab + C, c0, B);
#pragma unroll
    for (int k = 0; k < V; ++k) {
      o[k] = fmaf(xv[k], A[k], B[k]);
      if (res != nullptr) o[k] += r[k];
      if (relu) o[k] = fmaxf(o[k], 0.f);
    }
    if (mask != nullptr) {   // bit k = [y_k > 0]: what the backward passes need of the residual map (1 bit instead of 16 / 32)
      unsigned bits = 0u;
#pragma unroll
      for (int k = 0; k < V; ++k) bits |= (o[k] > 0.f ? 1u : 0u) << k;
      mask[i] = (unsigned char)bits;
    }
    if (y != nullptr) stv_any<V>(y, i * V, o);
    if (y_hi != nullptr) {
#pragma unroll
      for (int k = 0; k < V; k += 4) split4(make_float4(o[k], o[k + 1], o[k + 2], o[k + 3]), y_hi, y_lo, (size_t)i * V + k, f16);
    }
  }
}

// PLANES = false: dx in the storage type.  PLANES = true: dx as SCALED split planes (hi, lo) of dx * s, s = the power of two that
// brings an upper bound of max|dx| - computed here from the reduce pass's maxima: |dx| <= |gamma inv| (max|g| + |S1/M| +
// max|xhat| |S2/M|) - into [2^target, 2^(target+1)); block 0 publishes 256 copies of 1/s for the consuming convolution
// kernels (dgrad epilogue scale / wgrad finalize).  The fp32 gradient map is never written nor re-read for the split.
template <bool PLANES, typename TS = float, typename TG = float, bool MASK = false>
__global__ void __launch_bounds__(BN_THREADS)
bn_act_bwd_kernel(const TG* __restrict__ dy, const TS* __restrict__ x, const TS* __restrict__ res, int64_t nv,
                  int64_t M, int C, const float* __restrict__ mean, const float* __restrict__ invstd,
                  const float* __restrict__ gamma, const float* __restrict__ beta, int relu,
                  const double* __restrict__ ws, TS* __restrict__ dx, TS* __restrict__ dres,
                  float* __restrict__ dgamma, float* __restrict__ dbeta, const unsigned int* __restrict__ wmax,
                  __nv_bfloat16* __restrict__ dx_hi, __nv_bfloat16* __restrict__ dx_lo, int f16, int target_log2,
                  float* __restrict__ inv_vec, const unsigned char* __restrict__ mask = nullptr) {
  constexpr int V = Vec<TS>::V;
  extern __shared__ __align__(16) float sums[];                      // [2][C]: S1/M, S2/M, then [4][C]: A, B, c1, c2
  float* tab = sums + 2 * C;
  __shared__ float s_red[BN_THREADS / 32];
  __shared__ float s_scale;
  const int cvn = C / V;
  if (blockIdx.x == 0) {                               // parameter gradients: dbeta = S1, dgamma = S2
    for (int c = threadIdx.x; c < C; c += BN_THREADS) {
      if (dbeta != nullptr) dbeta[c] = (float)ws[c];
      if (dgamma != nullptr) dgamma[c] = (float)ws[C + c];
    }
  }
  for (int c = threadIdx.x; c < 2 * C; c += BN_THREADS) sums[c] = (float)(ws[c] / (double)M);
  __syncthreads();
  // dx = A (g - s1 - xhat s2) = A g + c2 x + c1 with c2 = -A inv s2, c1 = -A s1 - c2 mean: two fma per element
  for (int c = threadIdx.x; c < C; c += BN_THREADS) {
    float A, B;
    bn_affine(mean[c], invstd[c], gamma[c], beta[c], A, B);
    const float c2 = -A * invstd[c] * sums[C + c];
    tab[c] = A; tab[C + c] = B; tab[2 * C + c] = fmaf(-c2, mean[c], -A * sums[c]); tab[3 * C + c] = c2;
  }
  __syncthreads();
  float scale = 1.f;
  if (PLANES) {
    const float gmax = __uint_as_float(wmax[C]);
    float bound = 0.f;
    for (int c = threadIdx.x; c < C; c += BN_THREADS) {
      const float iv = invstd[c];
      bound = fmaxf(bound, fabsf(gamma[c] * iv) * (gmax + fabsf(sums[c]) + __uint_as_float(wmax[c]) * iv * fabsf(sums[C + c])));
    }
    bound = warp_max(bound);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = bound;
    __syncthreads();
    if (threadIdx.x == 0) {
      float m = 0.f;
      for (int w = 0; w < BN_THREADS / 32; ++w) m = fmaxf(m, s_red[w]);
      s_scale = pow2_scale(__float_as_uint(m), target_log2);
    }
    __syncthreads();
    scale = s_scale;
    if (blockIdx.x == 0) inv_vec[threadIdx.x] = 1.f / scale;
  }
  // descending order: the statistics pass that ran just before this one read the map ascending, so its tail is what L2 still holds
  for (int64_t i = nv - 1 - ((int64_t)blockIdx.x * BN_THREADS + threadIdx.x); i >= 0; i -= (int64_t)gridDim.x * BN_THREADS) {
    const int c0 = V * (int)(i % cvn);
    float xv[V], g[V], r[V], o[V], A[V], B[V], c1[V], c2[V];
    ldv<TS>(x, i * V, xv);
    ldv_any<V>(dy, i * V, g);
    const unsigned bits = (MASK && relu) ? (unsigned)__ldg(mask + i) : 0u;
    if (!MASK && relu && res != nullptr) ldv<TS>(res, i * V, r);
    ldp<V>(tab, c0, A); ldp<V>(tab + 2 * C, c0, c1); ldp<V>(tab + 3 * C, c0, c2);
    if (!MASK && relu) ldp<V>(tab + C, c0, B);
#pragma unroll
    for (int k = 0; k < V; ++k) {
      if (MASK) {
        if (relu) g[k] = ((bits >> k) & 1u) ? g[k] : 0.f;
      } else if (relu) {
        float y = fmaf(xv[k], A[k], B[k]);
        if (res != nullptr) y += r[k];
        g[k] = y > 0.f ? g[k] : 0.f;
      }
      o[k] = fmaf(A[k], g[k], fmaf(c2[k], xv[k], c1[k]));
      if (PLANES) o[k] *= scale;
    }
    if (dres != nullptr) stv(dres, i * V, g);
    if (PLANES) {
#pragma unroll
      for (int k = 0; k < V; k += 4) split4(make_float4(o[k], o[k + 1], o[k + 2], o[k + 3]), dx_hi, dx_lo, (size_t)i * V + k, f16);
    } else {
      stv(dx, i * V, o);
    }
  }
}

// ---------------------------------------------------------------- maxpool 3x3 s2 p1, NHWC
// window of output (oh, ow): rows 2oh-1 .. 2oh+1, cols 2ow-1 .. 2ow+1 clipped to the map; scan order rows then
// columns, first maximum wins (strict >) - torch's max_pool2d_with_indices rule, which decides where the gradient
// goes.  The forward stores the winner's position inside the (unclipped) window as one byte per output element
// (code = r*3 + s); the backward is a gather over the <= 2 x 2 windows that contain an input pixel: compare codes,
// no rescans, no atomics.
// IdxT = uint32_t whenever the element count fits (64-bit div / mod per element made these kernels index-math bound:
// the backward ran at 1.7 TB/s)
// One thread = 8 channels of one pixel; blockIdx.y / z = image row / page, so no per-element division (the first version
// did 64-bit div / mod per float4 and ran the backward at 1.1-1.7 TB/s); cvn = C / 8 is a power of two (shift / mask).
template <typename TS>
__global__ void __launch_bounds__(256)
maxpool_fwd_kernel(const TS* __restrict__ x, int H, int W, int C, int Ho, int Wo, int cshift, TS* __restrict__ y,
                   unsigned char* __restrict__ code, __nv_bfloat16* __restrict__ y_hi, __nv_bfloat16* __restrict__ y_lo,
                   int f16) {
  const int t = blockIdx.x * 256 + threadIdx.x;
  const int cvn = 1 << cshift;
  if (t >= Wo * cvn) return;
  const int cg = t & (cvn - 1), ow = t >> cshift, oh = blockIdx.y, b = blockIdx.z;
  float m[8];
  int mi[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { m[k] = -FLT_MAX; mi[k] = -1; }
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int h = 2 * oh - 1 + r;
    if (h < 0 || h >= H) continue;
#pragma unroll
    for (int s2 = 0; s2 < 3; ++s2) {
      const int w = 2 * ow - 1 + s2;
      if (w < 0 || w >= W) continue;
      float v[8];
      ldv_any<8>(x, (int64_t)((((size_t)b * H + h) * W + w) * C) + 8 * cg, v);
      const int kk = r * 3 + s2;
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (v[k] > m[k] || mi[k] < 0) { m[k] = v[k]; mi[k] = kk; }
    }
  }
  const size_t o = ((((size_t)b * Ho + oh) * Wo + ow) * C) + 8 * cg;
  stv_any<8>(y, (int64_t)o, m);
  if (code != nullptr)
    *reinterpret_cast<uint2*>(code + o) = make_uint2((uint32_t)mi[0] | ((uint32_t)mi[1] << 8) | ((uint32_t)mi[2] << 16) | ((uint32_t)mi[3] << 24),
                                                    (uint32_t)mi[4] | ((uint32_t)mi[5] << 8) | ((uint32_t)mi[6] << 16) | ((uint32_t)mi[7] << 24));
  if (y_hi != nullptr) {
    split4(make_float4(m[0], m[1], m[2], m[3]), y_hi, y_lo, o, f16);
    split4(make_float4(m[4], m[5], m[6], m[7]), y_hi, y_lo, o + 4, f16);
  }
}

// Backward: a 2 x 2 block of input pixels per thread (rows 2i, 2i+1, columns 2j, 2j+1, 8 channels).  The block is covered by the
// four windows (i | i+1, j | j+1), loaded ONCE (a thread per input pixel loads nine window records for the same four pixels:
// 0.66 -> 0.39 ms on the 640^2 fp32 map at B=16), and each window can only have put its maximum on fixed positions of the block:
//   (2i, 2j): (i,j) code 4      (2i, 2j+1): (i,j) 5, (i,j+1) 3      (2i+1, 2j): (i,j) 7, (i+1,j) 1
//   (2i+1, 2j+1): (i,j) 8, (i,j+1) 6, (i+1,j) 2, (i+1,j+1) 0          - summed in row-major window order.
template <typename TS>
__global__ void __launch_bounds__(256)
maxpool_bwd2x2_kernel(const unsigned char* __restrict__ code, const TS* __restrict__ dy, int H, int W, int C, int Ho, int Wo,
                      int cshift, TS* __restrict__ dx) {
  const int t = blockIdx.x * 256 + threadIdx.x;
  const int cvn = 1 << cshift, Wb = (W + 1) >> 1;
  if (t >= Wb * cvn) return;
  const int cg = t & (cvn - 1), j = t >> cshift, i = blockIdx.y, b = blockIdx.z;
  float g[4][8];
  uint2 kc[4];
#pragma unroll
  for (int wq = 0; wq < 4; ++wq) {
    const int oh = i + (wq >> 1), ow = j + (wq & 1);
    kc[wq] = make_uint2(0xffffffffu, 0xffffffffu);          // code 255: matches nothing
#pragma unroll
    for (int k = 0; k < 8; ++k) g[wq][k] = 0.f;
    if (oh < Ho && ow < Wo) {
      const size_t o = ((((size_t)b * Ho + oh) * Wo + ow) * C) + 8 * cg;
      kc[wq] = __ldg(reinterpret_cast<const uint2*>(code + o));
      ldv_any<8>(dy, (int64_t)o, g[wq]);
    }
  }
  float a00[8], a01[8], a10[8], a11[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int sh = 8 * (k & 3);
    const uint32_t c0 = ((k < 4 ? kc[0].x : kc[0].y) >> sh) & 255u, c1 = ((k < 4 ? kc[1].x : kc[1].y) >> sh) & 255u,
                   c2 = ((k < 4 ? kc[2].x : kc[2].y) >> sh) & 255u, c3 = ((k < 4 ? kc[3].x : kc[3].y) >> sh) & 255u;
    float v;
    v = 0.f; if (c0 == 4u) v += g[0][k]; a00[k] = v;
    v = 0.f; if (c0 == 5u) v += g[0][k]; if (c1 == 3u) v += g[1][k]; a01[k] = v;
    v = 0.f; if (c0 == 7u) v += g[0][k]; if (c2 == 1u) v += g[2][k]; a10[k] = v;
    v = 0.f; if (c0 == 8u) v += g[0][k]; if (c1 == 6u) v += g[1][k]; if (c2 == 2u) v += g[2][k]; if (c3 == 0u) v += g[3][k]; a11[k] = v;
  }
  const int h0 = 2 * i, w0 = 2 * j;
  const size_t base = ((((size_t)b * H + h0) * W + w0) * C) + 8 * cg;
  stv_any<8>(dx, (int64_t)base, a00);
  if (w0 + 1 < W) stv_any<8>(dx, (int64_t)(base + C), a01);
  if (h0 + 1 < H) {
    stv_any<8>(dx, (int64_t)(base + (size_t)W * C), a10);
    if (w0 + 1 < W) stv_any<8>(dx, (int64_t)(base + (size_t)W * C + C), a11);
  }
}

// fp32 -> split-bf16 planes (hi = bf16(x), lo = bf16(x - hi)): the operand format of the tensor-core convolutions
__global__ void __launch_bounds__(256)
split_planes_kernel(const float* __restrict__ x, int64_t n4, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                    int f16) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x)
    split4(__ldg(reinterpret_cast<const float4*>(x) + i), hi, lo, (size_t)i * 4, f16);
}

// max |x| of a map, as the bit pattern of a non-negative float (monotonic in the value): one atomicMax per CTA
__global__ void __launch_bounds__(256)
absmax_kernel(const float* __restrict__ x, int64_t n4, unsigned int* __restrict__ out_bits) {
  float m = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));   // fmaxf drops NaNs
  }
  m = warp_max(m);
  __shared__ float sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 8) {
    m = sm[threadIdx.x];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffu, m, o));
    if (threadIdx.x == 0) atomicMax(out_bits, __float_as_uint(m));
  }
}

__global__ void __launch_bounds__(256)
split_planes_scaled_kernel(const float* __restrict__ x, int64_t n4, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                           int f16, const unsigned int* __restrict__ amax_bits, int target_log2, float* __restrict__ inv_vec) {
  const float s = pow2_scale(__ldg(amax_bits), target_log2);
  if (blockIdx.x == 0) inv_vec[threadIdx.x] = 1.f / s;               // 256 copies of 1/s (exact: s is a power of two)
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    v.x *= s; v.y *= s; v.z *= s; v.w *= s;
    split4(v, hi, lo, (size_t)i * 4, f16);
  }
}

// ---------------------------------------------------------------- stem: BatchNorm(batch stats) + ReLU + maxpool, fused
// The normalised [B,H,W,C] map between bn1 and the maxpool (1.7 GB at B=16) is never written: the forward pools
// relu(bn(x)) on the fly from the raw conv1 output, the backward re-derives each input pixel's gradient from the
// pooled gradient and the winner codes inside the BatchNorm reduction / apply passes (and emits dx in the format its
// consumer - conv1's wgrad - reads).  One thread = 8 channels; rows / pages come from the grid (forward) or from a
// per-CTA row loop (backward): no per-element division (the first version of these kernels was index-math bound).
template <typename TS, typename TY>
__global__ void __launch_bounds__(256)
bn_relu_pool_fwd_kernel(const TS* __restrict__ x, int H, int W, int C, int Ho, int Wo, int cshift,
                        const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ gamma,
                        const float* __restrict__ beta, TY* __restrict__ y, unsigned char* __restrict__ code,
                        __nv_bfloat16* __restrict__ y_hi, __nv_bfloat16* __restrict__ y_lo, int f16) {
  extern __shared__ __align__(16) float tab[];                       // [2][C]: A, B
  for (int c = threadIdx.x; c < C; c += 256) bn_affine(mean[c], invstd[c], gamma[c], beta[c], tab[c], tab[C + c]);
  __syncthreads();
  const int t = blockIdx.x * 256 + threadIdx.x;
  const int cvn = 1 << cshift;
  if (t >= Wo * cvn) return;
  const int cg = t & (cvn - 1), ow = t >> cshift, oh = blockIdx.y, b = blockIdx.z;
  float A[8], Bc[8], m[8];
  int mi[8];
  ldp<8>(tab, 8 * cg, A);
  ldp<8>(tab + C, 8 * cg, Bc);
#pragma unroll
  for (int k = 0; k < 8; ++k) { m[k] = -FLT_MAX; mi[k] = -1; }
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int h = 2 * oh - 1 + r;
    if (h < 0 || h >= H) continue;
#pragma unroll
    for (int s2 = 0; s2 < 3; ++s2) {
      const int w = 2 * ow - 1 + s2;
      if (w < 0 || w >= W) continue;
      float v[8];
      ldv_any<8>(x, (int64_t)((((size_t)b * H + h) * W + w) * C) + 8 * cg, v);
      const int kk = r * 3 + s2;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float yv = fmaxf(fmaf(v[k], A[k], Bc[k]), 0.f);
        if (yv > m[k] || mi[k] < 0) { m[k] = yv; mi[k] = kk; }
      }
    }
  }
  const size_t o = ((((size_t)b * Ho + oh) * Wo + ow) * C) + 8 * cg;
  if (y != nullptr) stv_any<8>(y, (int64_t)o, m);
  *reinterpret_cast<uint2*>(code + o) = make_uint2((uint32_t)mi[0] | ((uint32_t)mi[1] << 8) | ((uint32_t)mi[2] << 16) | ((uint32_t)mi[3] << 24),
                                                  (uint32_t)mi[4] | ((uint32_t)mi[5] << 8) | ((uint32_t)mi[6] << 16) | ((uint32_t)mi[7] << 24));
  if (y_hi != nullptr) {
    split4(make_float4(m[0], m[1], m[2], m[3]), y_hi, y_lo, o, f16);
    split4(make_float4(m[4], m[5], m[6], m[7]), y_hi, y_lo, o + 4, f16);
  }
}

// gradient that reaches 8 channels of input pixel (b, h0, w0) of the pooled map through the <= 2 x 2 windows it won
template <typename TG>
__device__ __forceinline__ void pool_gather8(const unsigned char* __restrict__ code, const TG* __restrict__ dyp, int b, int h0, int w0,
                                             int cg, int C, int Ho, int Wo, float (&acc)[8]) {
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  const int oh_hi = (h0 + 1) >> 1, ow_hi = (w0 + 1) >> 1;
  for (int oh = h0 >> 1; oh <= oh_hi; ++oh) {
    if (oh >= Ho) continue;
    for (int ow = w0 >> 1; ow <= ow_hi; ++ow) {
      if (ow >= Wo) continue;
      const uint32_t me = (uint32_t)((h0 - (2 * oh - 1)) * 3 + (w0 - (2 * ow - 1)));
      const size_t o = ((((size_t)b * Ho + oh) * Wo + ow) * C) + 8 * cg;
      const uint2 kc = __ldg(reinterpret_cast<const uint2*>(code + o));
      float g[8];
      ldv_any<8>(dyp, (int64_t)o, g);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (((kc.x >> (8 * k)) & 255u) == me) acc[k] += g[k];
        if (((kc.y >> (8 * k)) & 255u) == me) acc[4 + k] += g[4 + k];
      }
    }
  }
}

// PASS 0: ws = [sum g | sum g*xhat] (+ wmax: per-channel max|x - mean|, max|g|, for the PLANES scale bound)
// PASS 1: dx = A g + c2 x + c1 (bn_act_bwd_kernel's form), stored in the map's type or as scaled split planes; dgamma / dbeta
// CTA = a strided set of image rows (b, h); thread = 8 channels (fixed) of every (256 / cvn)-th pixel of the row.
template <int PASS, bool PLANES, typename TS, typename TG>
__global__ void __launch_bounds__(256)
bn_relu_pool_bwd_kernel(const TS* __restrict__ x, const unsigned char* __restrict__ code, const TG* __restrict__ dyp, int B, int H,
                        int W, int C, int Ho, int Wo, int cshift, const float* __restrict__ mean, const float* __restrict__ invstd,
                        const float* __restrict__ gamma, const float* __restrict__ beta, double* __restrict__ ws,
                        unsigned int* __restrict__ wmax, TS* __restrict__ dx, __nv_bfloat16* __restrict__ dx_hi,
                        __nv_bfloat16* __restrict__ dx_lo, int f16, int target_log2, float* __restrict__ inv_vec,
                        float* __restrict__ dgamma, float* __restrict__ dbeta) {
  extern __shared__ __align__(16) float sm[];           // PASS 0: [2][ppi][C] partial sums; PASS 1: [2][C] S1/M, S2/M
  __shared__ float s_red[8];
  __shared__ float s_scale;
  const int cvn = 1 << cshift, ppi = 256 >> cshift;     // channel groups, pixels per iteration
  const int cg = threadIdx.x & (cvn - 1), wl = threadIdx.x >> cshift, c0 = 8 * cg;
  const int64_t M = (int64_t)B * H * W;
  float A[8], Bc[8], p2[8], p3[8];                      // PASS 0: p2 = mean, p3 = invstd;  PASS 1: p2 = c1, p3 = c2
  {
    float mu[8], iv[8], ga[8], be[8];
    ldp<8>(mean, c0, mu); ldp<8>(invstd, c0, iv); ldp<8>(gamma, c0, ga); ldp<8>(beta, c0, be);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      bn_affine(mu[k], iv[k], ga[k], be[k], A[k], Bc[k]);
      p2[k] = mu[k]; p3[k] = iv[k];
    }
  }
  float scale = 1.f;
  if (PASS == 1) {
    if (blockIdx.x == 0)
      for (int c = threadIdx.x; c < C; c += 256) { dbeta[c] = (float)ws[c]; dgamma[c] = (float)ws[C + c]; }
    for (int c = threadIdx.x; c < 2 * C; c += 256) sm[c] = (float)(ws[c] / (double)M);
    __syncthreads();
    if (PLANES) {
      const float gmax = __uint_as_float(wmax[C]);
      float bound = 0.f;
      for (int c = threadIdx.x; c < C; c += 256) {
        const float iv = invstd[c];
        bound = fmaxf(bound, fabsf(gamma[c] * iv) * (gmax + fabsf(sm[c]) + __uint_as_float(wmax[c]) * iv * fabsf(sm[C + c])));
      }
      bound = warp_max(bound);
      if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = bound;
      __syncthreads();
      if (threadIdx.x == 0) {
        float m = 0.f;
        for (int w = 0; w < 8; ++w) m = fmaxf(m, s_red[w]);
        s_scale = pow2_scale(__float_as_uint(m), target_log2);
      }
      __syncthreads();
      scale = s_scale;
      if (blockIdx.x == 0) inv_vec[threadIdx.x] = 1.f / scale;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float mu = p2[k], iv = p3[k];
      const float c2 = -A[k] * iv * sm[C + c0 + k];
      p2[k] = fmaf(-c2, mu, -A[k] * sm[c0 + k]);        // c1
      p3[k] = c2;
    }
  }
  float a[8], bs[8], mx[8];
  float gm = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] = bs[k] = mx[k] = 0.f;
  for (int row = blockIdx.x; row < B * H; row += gridDim.x) {
    const int b = row / H, h0 = row - b * H;
    for (int w0 = wl; w0 < W; w0 += ppi) {
      const size_t e = (((size_t)row * W + w0) * C) + c0;
      float xv[8], g[8];
      ldv_any<8>(x, (int64_t)e, xv);
      pool_gather8<TG>(code, dyp, b, h0, w0, cg, C, Ho, Wo, g);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        g[k] = fmaf(xv[k], A[k], Bc[k]) > 0.f ? g[k] : 0.f;
        if (PASS == 0) {
          const float xc = xv[k] - p2[k];
          a[k] += g[k];
          bs[k] = fmaf(g[k], xc * p3[k], bs[k]);
          if (PLANES) { mx[k] = fmaxf(mx[k], fabsf(xc)); gm = fmaxf(gm, fabsf(g[k])); }
        } else {
          g[k] = fmaf(A[k], g[k], fmaf(p3[k], xv[k], p2[k]));
          if (PLANES) g[k] *= scale;
        }
      }
      if (PASS == 1) {
        if (PLANES) {
          split4(make_float4(g[0], g[1], g[2], g[3]), dx_hi, dx_lo, e, f16);
          split4(make_float4(g[4], g[5], g[6], g[7]), dx_hi, dx_lo, e + 4, f16);
        } else {
          stv_any<8>(dx, (int64_t)e, g);
        }
      }
    }
  }
  if (PASS == 0) {
    float* ra = sm + (size_t)wl * C + c0;
    float* rb = sm + (size_t)(ppi + wl) * C + c0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { ra[k] = a[k]; rb[k] = bs[k]; }
    __syncthreads();
    for (int c = threadIdx.x; c < 2 * C; c += 256) {
      const int which = c / C, ch = c % C;
      double t = 0.0;
      for (int r = 0; r < ppi; ++r) t += (double)sm[(size_t)(which * ppi + r) * C + ch];
      atomicAdd(ws + c, t);
    }
    if (PLANES) {
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 8; ++k) ra[k] = mx[k];
      __syncthreads();
      for (int c = threadIdx.x; c < C; c += 256) {
        float m = 0.f;
        for (int r = 0; r < ppi; ++r) m = fmaxf(m, sm[(size_t)r * C + c]);
        atomicMax(wmax + c, __float_as_uint(m));
      }
      gm = warp_max(gm);
      if ((threadIdx.x & 31) == 0) atomicMax(wmax + C, __float_as_uint(gm));
    }
  }
}

static bool bn_c_ok(int C) { return C >= 4 && C <= 1024 && (C & (C - 1)) == 0; }
static bool pool_c_ok(int C) { return C >= 8 && (C & (C - 1)) == 0; }
static int ilog2(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }
static int ew_grid(int64_t n, int threads) {
  int64_t b = (n + threads - 1) / threads;
  const int64_t cap = (int64_t)sm_count() * 16;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

}  // namespace cova

using namespace cova;

extern "C" int cova_bn_train_stats(const float* x, int64_t M, int C, double* ws, void* stream) {
  COVA_REQUIRE(x && ws && M > 0, "cova_bn_train_stats: bad arguments");
  COVA_REQUIRE(bn_c_ok(C), "cova_bn_train_stats: C=%d must be a power of two in [4, 1024]", C);
  COVA_REQUIRE(((uintptr_t)x & 15) == 0, "cova_bn_train_stats: x must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  COVA_CUDA_OK(cudaMemsetAsync(ws, 0, 2 * C * sizeof(double), st));
  const int rows = BN_THREADS / (C / 4);
  const size_t smem = (size_t)2 * rows * C * sizeof(float);
  int64_t grid = (M + rows - 1) / rows;
  if (grid > (int64_t)sm_count() * 8) grid = (int64_t)sm_count() * 8;
  bn_reduce_kernel<0, float, float><<<(int)grid, BN_THREADS, smem, st>>>(x, nullptr, nullptr, M, C, nullptr, nullptr, nullptr, nullptr, 0,
                                                           ws);
  COVA_LAUNCH_OK();
  return COVA_OK;
}

extern "C" int cova_bn_train_finalize(const double* ws, int64_t M, int C, float eps, float momentum, float* mean,
                                      float* invstd, float* running_mean, float* running_var, void* stream) {
  COVA_REQUIRE(ws && mean && invstd && M > 0 && C > 0, "cova_bn_train_finalize: bad arguments");
  COVA_REQUIRE((running_mean == nullptr) == (running_var == nullptr), "cova_bn_train_finalize: running stats come together");
  bn_finalize_kernel<<<ceil_div(C, 128), 128, 0, (cudaStream_t)stream>>>(ws, M, C, eps, momentum, mean, invstd, running_mean,
                                                                        running_var);
  COVA_LAUNCH_OK();
  return COVA_OK;
}

extern "C" int cova_bn_act_fwd(const float* x, int64_t M, int C, const float* mean, const float* invstd,
                               const float* gamma, const float* beta, const float* res, int relu, float* y,
                               void* y_hi, void* y_lo, int planes_dtype, unsigned char* relu_mask, void* stream) {
  COVA_REQUIRE(planes_dtype == COVA_BF16X2 || planes_dtype == COVA_F16X2, "cova_bn_act_fwd: planes are split-bf16 or split-fp16");
  COVA_REQUIRE(!relu_mask || relu, "cova_bn_act_fwd: the ReLU bit mask comes with relu");
  COVA_REQUIRE(x && (y || y_hi) && mean && invstd && gamma && beta && M > 0, "cova_bn_act_fwd: bad arguments");
  COVA_REQUIRE(bn_c_ok(C), "cova_bn_act_fwd: C=%d must be a power of two in [4, 1024]", C);
  COVA_REQUIRE((((uintptr_t)x | (uintptr_t)y | (uintptr_t)res | (uintptr_t)y_hi | (uintptr_t)y_lo) & 15) == 0,
               "cova_bn_act_fwd: 16-byte alignment");
  COVA_REQUIRE((y_hi == nullptr) == (y_lo == nullptr), "cova_bn_act_fwd: the split planes come together");
  const int64_t n4 = M * (C / 4);
  bn_act_fwd_kernel<float, float><<<ew_grid(n4, BN_THREADS), BN_THREADS, 2 * C * sizeof(float), (cudaStream_t)stream>>>(x, n4, C, mean, invstd, gamma, beta,
                                                                                     res, relu, y, (__nv_bfloat16*)y_hi,
                                                                                     (__nv_bfloat16*)y_lo,
                                                                                     planes_dtype == COVA_F16X2, relu_mask);
  COVA_LAUNCH_OK();
  return COVA_OK;
}

extern "C" int cova_bn_act_bwd(const float* dy, const float* x, const float* res, int64_t M, int C, const float* mean,
                               const float* invstd, const float* gamma, const float* beta, int relu, double* ws,
                               float* dx, float* dres, float* dgamma, float* dbeta, void* stream) {
  COVA_REQUIRE(dy && x && dx && ws && mean && invstd && gamma && beta && M > 0, "cova_bn_act_bwd: bad arguments");
  COVA_REQUIRE(bn_c_ok(C), "cova_bn_act_bwd: C=%d must be a power of two in [4, 1024]", C);
  COVA_REQUIRE((((uintptr_t)dy | (uintptr_t)x | (uintptr_t)res | (uintptr_t)dx | (uintptr_t)dres) & 15) == 0,
               "cova_bn_act_bwd: 16-byte alignment");
  cudaStream_t st = (cudaStream_t)stream;
  COVA_CUDA_OK(cudaMemsetAsync(ws, 0, 2 * C * sizeof(double), st));
  const int rows = BN_THREADS / (C / 4);
  const size_t smem = (size_t)2 * rows * C * sizeof(float);
  int64_t grid = (M + rows - 1) / rows;
  if (grid > (int64_t)sm_count() * 8) grid = (int64_t)sm_count() * 8;
  bn_reduce_kernel<1, float, float><<<(int)grid, BN_THREADS, smem, st>>>(x, dy, res, M, C, mean, invstd, gamma, beta, relu, ws);
  COVA_LAUNCH_OK();
  const int64_t n4 = M * (C / 4);
  bn_act_bwd_kernel<false, float, float><<<ew_grid(n4, BN_THREADS), BN_THREADS, 6 * C * sizeof(float), st>>>(
      dy, x, res, n4, M, C, mean, invstd, gamma, beta, relu, ws, dx, dres, dgamma, dbeta, nullptr, nullptr, nullptr, 0, 0, nullptr);
  COVA_LAUNCH_OK();
  return COVA_OK;
}

extern "C" int cova_bn_act_bwd_planes(const float* dy, const float* x, const float* res, int64_t M, int C, const float* mean,
                                      const float* invstd, const float* gamma, const float* beta, int relu, double* ws,
                                      unsigned int* ws_max, void* dx_hi, void* dx_lo, int planes_dtype, int target_log2,
                                      float* inv_scale_vec, float* dres, float* dgamma, float* dbeta, const unsigned char* relu_mask,
                                      void* stream) {
  COVA_REQUIRE(!relu_mask || relu, "cova_bn_act_bwd_planes: the ReLU bit mask comes with relu");
  COVA_REQUIRE(dy && x && dx_hi && dx_lo && ws && ws_max && inv_scale_vec && mean && invstd && gamma && beta && M > 0,
               "cova_bn_act_bwd_planes: bad arguments");
  COVA_REQUIRE(bn_c_ok(C), "cova_bn_act_bwd_planes: C=%d must be a power of two in [4, 1024]", C);
  COVA_REQUIRE(planes_dtype == COVA_BF16X2 || planes_dtype == COVA_F16X2, "cova_bn_act_bwd_planes: planes are split-bf16 or split-fp16");
  COVA_REQUIRE(target_log2 >= -14 && target_log2 <= 14, "cova_bn_act_bwd_planes: target_log2 out of range");
  COVA_REQUIRE((((uintptr_t)dy | (uintptr_t)x | (uintptr_t)res | (uintptr_t)dres) & 15) == 0 &&
                   (((uintptr_t)dx_hi | (uintptr_t)dx_lo) & 7) == 0, "cova_bn_act_bwd_planes: alignment");
  cudaStream_t st = (cudaStream_t)stream;
  COVA_CUDA_OK(cudaMemsetAsync(ws, 0, 2 * C * sizeof(double), st));
  COVA_CUDA_OK(cudaMemsetAsync(ws_max, 0, (C + 1) * sizeof(unsigned int), st));
  const int rows = BN_THREADS / (C / 4);
  const size_t smem = (size_t)2 * rows * C * sizeof(float);
  int64_t grid = (M + rows - 1) / rows;
  if (grid > (int64_t)sm_count() * 8) grid = (int64_t)sm_count() * 8;
  const int64_t n4 = M * (C / 4);
  if (relu_mask) {   // the forward's ReLU decisions (4 bits of a byte per 4 channels): the residual map is not read
    bn_reduce_kernel<2, float, float, true><<<(int)grid, BN_THREADS, smem, st>>>(x, dy, res, M, C, mean, invstd, gamma, beta, relu, ws,
                                                                               ws_max, relu_mask);
    COVA_LAUNCH_OK();
    bn_act_bwd_kernel<true, float, float, true><<<ew_grid(n4, BN_THREADS), BN_THREADS, 6 * C * sizeof(float), st>>>(
        dy, x, res, n4, M, C, mean, invstd, gamma, beta, relu, ws, nullptr, dres, dgamma, dbeta, ws_max, (__nv_bfloat16*)dx_hi,
        (__nv_bfloat16*)dx_lo, planes_dtype == COVA_F16X2, target_log2, inv_scale_vec, relu_mask);
    COVA_LAUNCH_OK();
    return COVA_OK;
  }
  bn_reduce_kernel<2, float, float><<<(int)grid, BN_THREADS, smem, st>>>(x, dy, res, M, C, mean, invstd, gamma, beta, relu, ws, ws_max);
  COVA_LAUNCH_OK();
  bn_act_bwd_kernel<true, float, float><<<ew_grid(n4, BN_THREADS), BN_THREADS, 6 * C * sizeof(float), st>>>(
      dy, x, res, n4, M, C, mean, invstd, gamma, beta, relu, ws, nullptr, dres, dgamma, dbeta, ws_max, (__nv_bfloat16*)dx_hi,
      (__nv_bfloat16*)dx_lo, planes_dtype == COVA_F16X2, target_log2, inv_scale_vec);
  COVA_LAUNCH_OK();
  return COVA_OK;
}

extern "C" int cova_maxpool3x3s2_fwd(const float* x, int B, int H, int W, int C, float* y, unsigned char* code, void* y_hi,
                                     void* y_lo, int planes_dtype, void* stream) {
  COVA_REQUIRE(planes_dtype == COVA_BF16X2 || planes_dtype == COVA_F16X2, "cova_maxpool3x3s2_fwd: planes are split-bf16 or split-fp16");
  COVA_REQUIRE(x && y && B > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0, "cova_maxpool3x3s2_fwd: bad arguments");
  COVA_REQUIRE((((uintptr_t)x | (uintptr_t)y | (uintptr_t)y_hi | (uintptr_t)y_lo) & 15) == 0 && ((uintptr_t)code & 7) == 0,
               "cova_maxpool3x3s2_fwd: alignment");
  COVA_REQUIRE((y_hi == nullptr) == (y_lo == nullptr), "cova_maxpool3x3s2_fwd: the split planes come together");
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  COVA_REQUIRE(pool_c_ok(C) && Ho <= 65535 && B <= 65535, "cova_maxpool3x3s2_fwd: C / 8 must be a power of two, H/2 and B <= 65535");
  const dim3 grid(ceil_div(Wo * (C / 8), 256), Ho, B);
  maxpool_fwd_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(x, H, W, C, Ho, Wo, ilog2(C / 8), y, code, (__nv_bfloat16*)y_hi,
                                                                   (__nv_bfloat16*)y_lo, planes_dtype == COVA_F16X2);
  COVA_LAUNCH_OK();
  return COVA_OK;
}

extern "C" int cova_maxpool3x3s2_bwd(const unsigned char* code, const float* dy, int B, int H, int W, int C, float* dx,
                                     void* stream) {
  COVA_REQUIRE(code && dy && dx && B > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0, "cova_maxpool3x3s2_bwd: bad arguments");
  COVA_REQUIRE((((uintptr_t)dy | (uintptr_t)dx) & 15) == 0 && ((uintptr_t)code & 7) == 0, "cova_maxpool3x3s2_bwd: alignment");
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  COVA_REQUIRE(pool_c_ok(C) && H <= 65535 && B <= 65535, "cova_maxpool3x3s2_bwd: C / 8 must be a power of two, H and B <= 65535");
  const dim3 grid(ceil_div(((W + 1) / 2) * (C / 8), 256), (H + 1) / 2, B);
  maxpool_bwd2x2_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(code, dy, H, W, C, Ho, Wo, ilog2(C / 8), dx);
  COVA_LAUNCH_OK();
  return COVA_OK;
}

extern "C" int cova_split_planes(const float* x, int64_t n, void* hi, void* lo, int planes_dtype, void* stream) {
  COVA_REQUIRE(planes_dtype == COVA_BF16X2 || planes_dtype == COVA_F16X2, "cova_split_planes: planes are split-bf16 or split-fp16");
  COVA_REQUIRE(x && hi && lo && n > 0 && n % 4 == 0, "cova_split_planes: n must be a positive multiple of 4");
  COVA_REQUIRE((((uintptr_t)x & 15) | ((uintptr_t)hi & 7) | ((uintptr_t)lo & 7)) == 0, "cova_split_planes: alignment");
  split_planes_kernel<<<ew_grid(n / 4, 256), 256, 0, (cudaStream_t)stream>>>(x, n / 4, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo,
                                                                            planes_dtype == COVA_F16X2);
  COVA_LAUNCH_OK();
  return COVA_OK;
}

extern "C" int cova_split_planes_scaled(const float* x, int64_t n, void* hi, void* lo, int planes_dtype, int target_log2,
                                        unsigned int* ws, float* inv_scale_vec, void* stream) {
  COVA_REQUIRE(planes_dtype == COVA_BF16X2 || planes_dtype == COVA_F16X2, "cova_split_planes_scaled: planes are split-bf16 or split-fp16");
  COVA_REQUIRE(x && hi && lo && ws && inv_scale_vec && n > 0 && n % 4 == 0, "cova_split_planes_scaled: n must be a positive multiple of 4");
  COVA_REQUIRE((((uintptr_t)x & 15) | ((uintptr_t)hi & 7) | ((uintptr_t)lo & 7)) == 0, "cova_split_planes_scaled: alignment");
  COVA_REQUIRE(target_log2 >= -14 && target_log2 <= 14, "cova_split_planes_scaled: target_log2 out of range");
  cudaStream_t st = (cudaStream_t)stream;
  COVA_CUDA_OK(cudaMemsetAsync(ws, 0, sizeof(unsigned int), st));
  absmax_kernel<<<ew_grid(n / 4, 256), 256, 0, st>>>(x, n / 4, ws);
  COVA_LAUNCH_OK();
  split_planes_scaled_kernel<<<ew_grid(n / 4, 256), 256, 0, st>>>(x, n / 4, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo,
                                                                  planes_dtype == COVA_F16X2, ws, target_log2, inv_scale_vec);
  COVA_LAUNCH_OK();
  return COVA_OK;
}

typedef __nv_bfloat16 bf16_t;
static bool dt_ok(int d) { return d == COVA_F32 || d == COVA_BF16; }

extern "C" int cova_bn_relu_pool_fwd_t(const void* x, int s_dtype, int B, int H, int W, int C, const float* mean, const float* invstd,
                                       const float* gamma, const float* beta, void* y, int y_dtype, unsigned char* code,
                                       void* y_hi, void* y_lo, int planes_dtype, void* stream) {
  COVA_REQUIRE(x && code && (y || y_hi) && mean && invstd && gamma && beta && B > 0 && H > 0 && W > 0, "cova_bn_relu_pool_fwd_t: bad arguments");
  COVA_REQUIRE(dt_ok(s_dtype) && dt_ok(y_dtype) && s_dtype == y_dtype, "cova_bn_relu_pool_fwd_t: x and y are both fp32 or both bf16");
  COVA_REQUIRE(bn_c_ok(C) && C >= 8, "cova_bn_relu_pool_fwd_t: C=%d must be a power of two in [8, 1024]", C);
  COVA_REQUIRE((y_hi == nullptr) == (y_lo == nullptr), "cova_bn_relu_pool_fwd_t: the split planes come together");
  COVA_REQUIRE(planes_dtype == COVA_BF16X2 || planes_dtype == COVA_F16X2, "cova_bn_relu_pool_fwd_t: planes are split-bf16 or split-fp16");
  COVA_REQUIRE((((uintptr_t)x | (uintptr_t)y | (uintptr_t)y_hi | (uintptr_t)y_lo) & 15) == 0 && ((uintptr_t)code & 7) == 0,
               "cova_bn_relu_pool_fwd_t: alignment");
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  COVA_REQUIRE(Ho <= 65535 && B <= 65535, "cova_bn_relu_pool_fwd_t: H/2 and B <= 65535");
  const dim3 grid(ceil_div(Wo * (C / 8), 256), Ho, B);
  const size_t smem = (size_t)2 * C * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  if (s_dtype == COVA_F32)
    bn_relu_pool_fwd_kernel<float, float><<<grid, 256, smem, st>>>((const float*)x, H, W, C, Ho, Wo, ilog2(C / 8), mean, invstd, gamma, beta,
                                                                  (float*)y, code, (__nv_bfloat16*)y_hi, (__nv_bfloat16*)y_lo,
                                                                  planes_dtype == COVA_F16X2);
  else
    bn_relu_pool_fwd_kernel<bf16_t, bf16_t><<<grid, 256, smem, st>>>((const bf16_t*)x, H, W, C, Ho, Wo, ilog2(C / 8), mean, invstd, gamma,
                                                                    beta, (bf16_t*)y, code, (__nv_bfloat16*)y_hi, (__nv_bfloat16*)y_lo,
                                                                    planes_dtype == COVA_F16X2);
  COVA_LAUNCH_OK();
  return COVA_OK;
}

extern "C" int cova_bn_relu_pool_bwd_t(const void* x, int s_dtype, const unsigned char* code, const void* dy_pooled, int dy_dtype,
                                       int B, int H, int W, int C, const float* mean, const float* invstd, const float* gamma,
                                       const float* beta, double* ws, unsigned int* ws_max, void* dx, void* dx_hi, void* dx_lo,
                                       int planes_dtype, int target_log2, float* inv_scale_vec, float* dgamma, float* dbeta,
                                       void* stream) {
  const bool planes = dx_hi != nullptr;
  COVA_REQUIRE(x && code && dy_pooled && ws && (dx || planes) && dgamma && dbeta && mean && invstd && gamma && beta && B > 0 && H > 0 && W > 0,
               "cova_bn_relu_pool_bwd_t: bad arguments");
  COVA_REQUIRE(dt_ok(s_dtype) && dt_ok(dy_dtype) && s_dtype == dy_dtype, "cova_bn_relu_pool_bwd_t: x and dy are both fp32 or both bf16");
  COVA_REQUIRE(!planes || (s_dtype == COVA_F32 && dx_lo && ws_max && inv_scale_vec), "cova_bn_relu_pool_bwd_t: plane output goes with fp32 maps");
  COVA_REQUIRE(bn_c_ok(C) && C >= 8, "cova_bn_relu_pool_bwd_t: C=%d must be a power of two in [8, 1024]", C);
  COVA_REQUIRE(planes_dtype == COVA_BF16X2 || planes_dtype == COVA_F16X2, "cova_bn_relu_pool_bwd_t: planes are split-bf16 or split-fp16");
  COVA_REQUIRE((((uintptr_t)x | (uintptr_t)dy_pooled | (uintptr_t)dx) & 15) == 0 && ((uintptr_t)code & 7) == 0 &&
                   (((uintptr_t)dx_hi | (uintptr_t)dx_lo) & 7) == 0, "cova_bn_relu_pool_bwd_t: alignment");
  cudaStream_t st = (cudaStream_t)stream;
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  COVA_CUDA_OK(cudaMemsetAsync(ws, 0, 2 * C * sizeof(double), st));
  if (planes) COVA_CUDA_OK(cudaMemsetAsync(ws_max, 0, (C + 1) * sizeof(unsigned int), st));
  const int cs = ilog2(C / 8), ppi = 256 >> cs;
  int grid = B * H < sm_count() * 8 ? B * H : sm_count() * 8;
  const size_t sm0 = (size_t)2 * ppi * C * sizeof(float), sm1 = (size_t)2 * C * sizeof(float);
  const int f16 = planes_dtype == COVA_F16X2;
#define COVA_POOL_BWD(PL, TS, TG)                                                                                              \
  do {                                                                                                                         \
    bn_relu_pool_bwd_kernel<0, PL, TS, TG><<<grid, 256, sm0, st>>>((const TS*)x, code, (const TG*)dy_pooled, B, H, W, C, Ho, Wo, cs, mean, \
        invstd, gamma, beta, ws, ws_max, nullptr, nullptr, nullptr, f16, target_log2, nullptr, nullptr, nullptr);              \
    COVA_LAUNCH_OK();                                                                                                          \
    bn_relu_pool_bwd_kernel<1, PL, TS, TG><<<grid, 256, sm1, st>>>((const TS*)x, code, (const TG*)dy_pooled, B, H, W, C, Ho, Wo, cs, mean, \
        invstd, gamma, beta, ws, ws_max, (TS*)dx, (__nv_bfloat16*)dx_hi, (__nv_bfloat16*)dx_lo, f16, target_log2, inv_scale_vec,   \
        dgamma, dbeta);                                                                                                        \
  } while (0)
  if (s_dtype == COVA_BF16) COVA_POOL_BWD(false, bf16_t, bf16_t);
  else if (planes) COVA_POOL_BWD(true, float, float);
  else COVA_POOL_BWD(false, float, float);
#undef COVA_POOL_BWD
  COVA_LAUNCH_OK();
  return COVA_OK;
}

extern "C" int cova_bn_relu_pool_fwd(const float* x, int B, int H, int W, int C, const float* mean, const float* invstd,
                                     const float* gamma, const float* beta, float* y, unsigned char* code, void* y_hi,
                                     void* y_lo, int planes_dtype, void* stream) {
  return cova_bn_relu_pool_fwd_t(x, COVA_F32, B, H, W, C, mean, invstd, gamma, beta, y, COVA_F32, code, y_hi, y_lo, planes_dtype, stream);
}

extern "C" int cova_bn_relu_pool_bwd(const float* x, const unsigned char* code, const float* dy_pooled, int B, int H, int W,
                                     int C, const float* mean, const float* invstd, const float* gamma, const float* beta,
                                     double* ws, float* dx, float* dgamma, float* dbeta, void* stream) {
  return cova_bn_relu_pool_bwd_t(x, COVA_F32, code, dy_pooled, COVA_F32, B, H, W, C, mean, invstd, gamma, beta, ws, nullptr, dx, nullptr,
                                 nullptr, COVA_F16X2, 10, nullptr, dgamma, dbeta, stream);
}

// ---------------------------------------------------------------- typed entry points (bf16 training mode)

extern "C" int cova_bn_train_stats_t(const void* x, int x_dtype, int64_t M, int C, double* ws, void* stream) {
  if (x_dtype == COVA_F32) return cova_bn_train_stats((const float*)x, M, C, ws, stream);
  COVA_REQUIRE(x_dtype == COVA_BF16, "cova_bn_train_stats_t: x is fp32 or bf16");
  COVA_REQUIRE(x && ws && M > 0, "cova_bn_train_stats_t: bad arguments");
  COVA_REQUIRE(bn_c_ok(C) && C >= 8, "cova_bn_train_stats_t: C=%d must be a power of two in [8, 1024]", C);
  COVA_REQUIRE(((uintptr_t)x & 15) == 0, "cova_bn_train_stats_t: x must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  COVA_CUDA_OK(cudaMemsetAsync(ws, 0, 2 * C * sizeof(double), st));
  const int rows = BN_THREADS / (C / 8);
  const size_t smem = (size_t)2 * rows * C * sizeof(float);
  int64_t grid = (M + rows - 1) / rows;
  if (grid > (int64_t)sm_count() * 8) grid = (int64_t)sm_count() * 8;
  bn_reduce_kernel<0, bf16_t, bf16_t><<<(int)grid, BN_THREADS, smem, st>>>((const bf16_t*)x, nullptr, nullptr, M, C, nullptr, nullptr,
                                                                         nullptr, nullptr, 0, ws);
  COVA_LAUNCH_OK();
  return COVA_OK;
}

extern "C" int cova_bn_act_fwd_t(const void* x, int s_dtype, int64_t M, int C, const float* mean, const float* invstd,
                                 const float* gamma, const float* beta, const void* res, int relu, void* y, int y_dtype,
                                 unsigned char* relu_mask, void* stream) {
  COVA_REQUIRE(dt_ok(s_dtype) && dt_ok(y_dtype), "cova_bn_act_fwd_t: dtypes are fp32 or bf16");
  COVA_REQUIRE(!relu_mask || (s_dtype == COVA_BF16 && relu), "cova_bn_act_fwd_t: the ReLU bit mask comes with bf16 storage and relu");
  unsigned char* mk = relu_mask;
  COVA_REQUIRE(x && y && mean && invstd && gamma && beta && M > 0, "cova_bn_act_fwd_t: bad arguments");
  COVA_REQUIRE(bn_c_ok(C) && (s_dtype == COVA_F32 || C >= 8), "cova_bn_act_fwd_t: C=%d must be a power of two in [4 (fp32) / 8 (bf16), 1024]", C);
  COVA_REQUIRE((((uintptr_t)x | (uintptr_t)y | (uintptr_t)res) & 15) == 0, "cova_bn_act_fwd_t: 16-byte alignment");
  const int V = s_dtype == COVA_F32 ? 4 : 8;
  const int64_t nv = M * (C / V);
  const int grid = ew_grid(nv, BN_THREADS);
  cudaStream_t st = (cudaStream_t)stream;
  if (s_dtype == COVA_F32 && y_dtype == COVA_F32)
    bn_act_fwd_kernel<float, float><<<grid, BN_THREADS, 2 * C * sizeof(float), st>>>((const float*)x, nv, C, mean, invstd, gamma, beta, (const float*)res, relu, (float*)y, nullptr, nullptr, 0);
  else if (s_dtype == COVA_BF16 && y_dtype == COVA_BF16)
    bn_act_fwd_kernel<bf16_t, bf16_t><<<grid, BN_THREADS, 2 * C * sizeof(float), st>>>((const bf16_t*)x, nv, C, mean, invstd, gamma, beta, (const bf16_t*)res, relu, (bf16_t*)y, nullptr, nullptr, 0, mk);
  else if (s_dtype == COVA_BF16)
    bn_act_fwd_kernel<bf16_t, float><<<grid, BN_THREADS, 2 * C * sizeof(float), st>>>((const bf16_t*)x, nv, C, mean, invstd, gamma, beta, (const bf16_t*)res, relu, (float*)y, nullptr, nullptr, 0, mk);
  else
    COVA_REQUIRE(false, "cova_bn_act_fwd_t: fp32 storage with a bf16 output is not built");
  COVA_LAUNCH_OK();
  return COVA_OK;
}

extern "C" int cova_bn_act_bwd_t(const void* dy, int dy_dtype, const void* x, const void* res, int s_dtype, int64_t M, int C,
                                 const float* mean, const float* invstd, const float* gamma, const float* beta, int relu,
                                 double* ws, void* dx, void* dres, float* dgamma, float* dbeta, const unsigned char* relu_mask,
                                 void* stream) {
  COVA_REQUIRE(dt_ok(s_dtype) && dt_ok(dy_dtype), "cova_bn_act_bwd_t: dtypes are fp32 or bf16");
  COVA_REQUIRE(!relu_mask || (s_dtype == COVA_BF16 && relu), "cova_bn_act_bwd_t: the ReLU bit mask comes with bf16 storage and relu");
  const unsigned char* mk = relu_mask;
  if (s_dtype == COVA_F32 && dy_dtype == COVA_F32)
    return cova_bn_act_bwd((const float*)dy, (const float*)x, (const float*)res, M, C, mean, invstd, gamma, beta, relu, ws,
                           (float*)dx, (float*)dres, dgamma, dbeta, stream);
  COVA_REQUIRE(s_dtype == COVA_BF16, "cova_bn_act_bwd_t: fp32 storage with a bf16 gradient is not built");
  COVA_REQUIRE(dy && x && dx && ws && mean && invstd && gamma && beta && M > 0, "cova_bn_act_bwd_t: bad arguments");
  COVA_REQUIRE(bn_c_ok(C) && C >= 8, "cova_bn_act_bwd_t: C=%d must be a power of two in [8, 1024]", C);
  COVA_REQUIRE((((uintptr_t)x | (uintptr_t)res | (uintptr_t)dx | (uintptr_t)dres | (uintptr_t)dy) & 15) == 0,
               "cova_bn_act_bwd_t: 16-byte alignment");
  cudaStream_t st = (cudaStream_t)stream;
  COVA_CUDA_OK(cudaMemsetAsync(ws, 0, 2 * C * sizeof(double), st));
  const int rows = BN_THREADS / (C / 8);
  const size_t smem = (size_t)2 * rows * C * sizeof(float);
  int64_t grid = (M + rows - 1) / rows;
  if (grid > (int64_t)sm_count() * 8) grid = (int64_t)sm_count() * 8;
  const int64_t nv = M * (C / 8);
  const int g2 = ew_grid(nv, BN_THREADS);
  const bf16_t *xb = (const bf16_t*)x, *rb = (const bf16_t*)res;
#define COVA_BN_BWD_T(TGT, MK)                                                                                                  \
  do {                                                                                                                         \
    bn_reduce_kernel<1, bf16_t, TGT, MK><<<(int)grid, BN_THREADS, smem, st>>>(xb, (const TGT*)dy, rb, M, C, mean, invstd, gamma, \
                                                                              beta, relu, ws, nullptr, mk);                     \
    COVA_LAUNCH_OK();                                                                                                          \
    bn_act_bwd_kernel<false, bf16_t, TGT, MK><<<g2, BN_THREADS, 6 * C * sizeof(float), st>>>(                                   \
        (const TGT*)dy, xb, rb, nv, M, C, mean, invstd, gamma, beta, relu, ws, (bf16_t*)dx, (bf16_t*)dres, dgamma, dbeta,      \
        nullptr, nullptr, nullptr, 0, 0, nullptr, mk);                                                                         \
  } while (0)
  if (dy_dtype == COVA_BF16) {
    if (mk) COVA_BN_BWD_T(bf16_t, true); else COVA_BN_BWD_T(bf16_t, false);
  } else {
    if (mk) COVA_BN_BWD_T(float, true); else COVA_BN_BWD_T(float, false);
  }
#undef COVA_BN_BWD_T
  COVA_LAUNCH_OK();
  return COVA_OK;
}

extern "C" int cova_maxpool3x3s2_fwd_t(const void* x, int dtype, int B, int H, int W, int C, void* y, unsigned char* code,
                                       void* stream) {
  if (dtype == COVA_F32) return cova_maxpool3x3s2_fwd((const float*)x, B, H, W, C, (float*)y, code, nullptr, nullptr, COVA_BF16X2, stream);
  COVA_REQUIRE(dtype == COVA_BF16, "cova_maxpool3x3s2_fwd_t: dtype is fp32 or bf16");
  COVA_REQUIRE(x && y && B > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0, "cova_maxpool3x3s2_fwd_t: bad arguments");
  COVA_REQUIRE((((uintptr_t)x | (uintptr_t)y) & 7) == 0 && ((uintptr_t)code & 7) == 0, "cova_maxpool3x3s2_fwd_t: alignment");
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  COVA_REQUIRE(pool_c_ok(C) && Ho <= 65535 && B <= 65535, "cova_maxpool3x3s2_fwd_t: C / 8 must be a power of two, H/2 and B <= 65535");
  const dim3 grid(ceil_div(Wo * (C / 8), 256), Ho, B);
  maxpool_fwd_kernel<bf16_t><<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16_t*)x, H, W, C, Ho, Wo, ilog2(C / 8), (bf16_t*)y, code,
                                                                    nullptr, nullptr, 0);
  COVA_LAUNCH_OK();
  return COVA_OK;
}

extern "C" int cova_maxpool3x3s2_bwd_t(const unsigned char* code, const void* dy, int dtype, int B, int H, int W, int C, void* dx,
                                       void* stream) {
  if (dtype == COVA_F32) return cova_maxpool3x3s2_bwd(code, (const float*)dy, B, H, W, C, (float*)dx, stream);
  COVA_REQUIRE(dtype == COVA_BF16, "cova_maxpool3x3s2_bwd_t: dtype is fp32 or bf16");
  COVA_REQUIRE(code && dy && dx && B > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0, "cova_maxpool3x3s2_bwd_t: bad arguments");
  COVA_REQUIRE((((uintptr_t)dy | (uintptr_t)dx) & 7) == 0 && ((uintptr_t)code & 7) == 0, "cova_maxpool3x3s2_bwd_t: alignment");
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  COVA_REQUIRE(pool_c_ok(C) && H <= 65535 && B <= 65535, "cova_maxpool3x3s2_bwd_t: C / 8 must be a power of two, H and B <= 65535");
  const dim3 grid(ceil_div(((W + 1) / 2) * (C / 8), 256), (H + 1) / 2, B);
  maxpool_bwd2x2_kernel<bf16_t><<<grid, 256, 0, (cudaStream_t)stream>>>(code, (const bf16_t*)dy, H, W, C, Ho, Wo, ilog2(C / 8), (bf16_t*)dx);
  COVA_LAUNCH_OK();
  return COVA_OK;
}
