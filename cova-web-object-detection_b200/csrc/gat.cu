// A6: fused graph-attention gather.  Replaces `GraphAttentionLayer.forward` lines 180-208 of
// `/root/reference/models.py` after the once-per-node projections (whj = W_j h, s = a_i.W_i h, t = a_j.W_j h;
// SURVEY.md row A6 - algebraically identical to the as-written layer, which re-projects every neighbour K
// times and materialises [T,K,F], [T,K,H] x2 and [T,K,2H] intermediates).
//
// One CTA = GAT_E consecutive elements, one warp per element:
//   1. lanes read the element's K neighbour ids; block-reduce the CTA's [min,max] id range
//   2. if that range fits the staging buffer, ONE elected thread TMA-bulk-copies rows [min,max] of whj into
//      shared memory (neighbour windows of consecutive elements overlap almost entirely: reuse E*K/(E+K));
//      otherwise rows are gathered straight from L2
//   3. logits e_k = LeakyReLU(s_i + t_c(k) + b), mask, softmax over K with warp shuffles
//   4. out_i = sum_k alpha_k * whj[c(k)], float4 per lane across the hidden dim
// HBM-bound; algorithmic bytes = T*K*Hd*4 read + T*Hd*4 written; real DRAM traffic is
// ~T*Hd*4*(1+K/E) because of the staging (and whj sits in L2 straight out of the projection GEMM).
#include <float.h>
#include <limits.h>

#include "common.cuh"
#include "ptx.cuh"

namespace cova {

constexpr int GAT_E = 8;                 // elements (= warps) per CTA
constexpr int GAT_THREADS = GAT_E * 32;
constexpr int GAT_KMAX = 128;            // neighbours per element handled by one warp (4 per lane)
constexpr int GAT_STAGE_BYTES = 96 * 1024;
constexpr int GAT_MAX_HEADS = 8;
struct GatBias { float b[GAT_MAX_HEADS]; };

__global__ void __launch_bounds__(GAT_THREADS)
gat_fwd_kernel(const float* __restrict__ whj, int64_t ld_whj, const float* __restrict__ s_vec,
               const float* __restrict__ t_vec, int64_t ld_st, GatBias att_b_h, float alpha, const int64_t* __restrict__ ctx, int T,
               int K, int Hd, float* __restrict__ out, int64_t ld_out, float* __restrict__ attn, int stage_ok, int out_vec,
               int n_heads) {
  // All attention heads in ONE CTA: the neighbour ids are read and the neighbour rows staged once for
  // every head (a grid.y-per-head layout re-read the ids and issued twice the bulk copies: 22 vs 15 us at config 2).
  // Head h reads whj columns [h*Hd, (h+1)*Hd), s/t at +2h, writes out columns [h*Hd, (h+1)*Hd) and attn [h][T][K].
  const int Wst = n_heads * Hd;            // staged row width (floats)
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* stage = reinterpret_cast<float*>(smem_raw);
  __shared__ uint64_t bar;
  __shared__ int s_min[GAT_E], s_max[GAT_E];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * GAT_E + warp;
  const bool live = i < T;

  // neighbour ids (4 per lane, k = lane + 32*j) and this warp's id range
  int cid[GAT_KMAX / 32];
  int lo = INT_MAX, hi = -1;
#pragma unroll
  for (int j = 0; j < GAT_KMAX / 32; ++j) {
    const int k = lane + 32 * j;
    int c = -1;
    if (live && k < K) {
      const int64_t c64 = ctx[(size_t)i * K + k];
      c = (c64 >= 0 && c64 < T) ? (int)c64 : -1;     // an id outside [0, T) is treated as padding, not dereferenced
    }
    cid[j] = c;
    if (c >= 0) { lo = min(lo, c); hi = max(hi, c); }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if (lane == 0) { s_min[warp] = lo; s_max[warp] = hi; }
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar, 1);
    ptx::fence_barrier_init();
  }
  __syncthreads();
  int cmin = INT_MAX, cmax = -1;
#pragma unroll
  for (int w = 0; w < GAT_E; ++w) { cmin = min(cmin, s_min[w]); cmax = max(cmax, s_max[w]); }
  const int nrows = cmax >= cmin ? cmax - cmin + 1 : 0;
  const bool staged = stage_ok && nrows > 0 && (size_t)nrows * Wst * 4 <= (size_t)GAT_STAGE_BYTES;
  if (staged && threadIdx.x == 0) {
    const uint32_t row_bytes = (uint32_t)Wst * 4u;
    ptx::mbar_arrive_expect_tx(&bar, row_bytes * (uint32_t)nrows);
    if (ld_whj == Wst) {
      ptx::tma_bulk_g2s(stage, whj + (size_t)cmin * ld_whj, row_bytes * (uint32_t)nrows, &bar);
    } else {
      for (int r = 0; r < nrows; ++r)
        ptx::tma_bulk_g2s(stage + (size_t)r * Wst, whj + (size_t)(cmin + r) * ld_whj, row_bytes, &bar);
    }
  }

  for (int head = 0; head < n_heads; ++head) {
  const float att_b = att_b_h.b[head];
  const float* whj_h = whj + (size_t)head * Hd;
  float* out_h = out + (size_t)head * Hd;
  float* attn_h = attn != nullptr ? attn + (size_t)head * T * K : nullptr;
  // attention logits + softmax over K (head 0: overlaps the bulk copy)
  const float si = live ? s_vec[(size_t)i * ld_st + 2 * head] : 0.f;
  float e[GAT_KMAX / 32];
  float mx = -FLT_MAX;
#pragma unroll
  for (int j = 0; j < GAT_KMAX / 32; ++j) {
    const int k = lane + 32 * j;
    float v = -FLT_MAX;                    // k >= K: not part of the softmax
    if (live && k < K) {
      if (cid[j] >= 0) {
        v = si + __ldg(t_vec + (size_t)cid[j] * ld_st + 2 * head) + att_b;
        v = v > 0.f ? v : alpha * v;       // LeakyReLU (models.py:200)
      } else {
        v = -9e15f;                        // models.py:202-203
      }
    }
    e[j] = v;
    mx = fmaxf(mx, v);
  }
  mx = warp_max(mx);
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < GAT_KMAX / 32; ++j) {
    const int k = lane + 32 * j;
    e[j] = (live && k < K) ? expf(e[j] - mx) : 0.f;
    sum += e[j];
  }
  sum = warp_sum(sum);
  const float inv = 1.f / sum;
#pragma unroll
  for (int j = 0; j < GAT_KMAX / 32; ++j) {
    e[j] *= inv;
    const int k = lane + 32 * j;
    if (attn_h != nullptr && live && k < K) attn_h[(size_t)i * K + k] = e[j];
  }

  if (staged && head == 0) ptx::mbar_wait(&bar, 0);
  if (!live) continue;

  // weighted sum of neighbour rows; lane covers float4 columns lane, lane+32, ...
  const int nvec = Hd >> 2;
  for (int v0 = 0; v0 < nvec; v0 += 128) {     // up to 4 float4 per lane per pass
    float4 acc[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < GAT_KMAX / 32; ++j) {
      if (32 * j >= K) break;
      const int kend = min(32, K - 32 * j);
      for (int kk = 0; kk < kend; ++kk) {
        const int c = __shfl_sync(0xffffffffu, cid[j], kk);
        const float a = __shfl_sync(0xffffffffu, e[j], kk);
        if (c < 0) continue;                  // zero row of models.py:180-184: contributes exactly 0
        const float4* row = staged ? reinterpret_cast<const float4*>(stage + (size_t)(c - cmin) * Wst + (size_t)head * Hd)
                                   : reinterpret_cast<const float4*>(whj_h + (size_t)c * ld_whj);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int v = v0 + lane + 32 * q;
          if (v < nvec) {
            const float4 x = staged ? row[v] : __ldg(row + v);
            acc[q].x = fmaf(a, x.x, acc[q].x); acc[q].y = fmaf(a, x.y, acc[q].y);
            acc[q].z = fmaf(a, x.z, acc[q].z); acc[q].w = fmaf(a, x.w, acc[q].w);
          }
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int v = v0 + lane + 32 * q;
      if (v >= nvec) continue;
      float* o = out_h + (size_t)i * ld_out + 4 * v;
      if (out_vec) {
        *reinterpret_cast<float4*>(o) = acc[q];
      } else {   // output columns not 16-byte aligned (e.g. ctx lands at column n_feat of an odd-width row)
        o[0] = acc[q].x; o[1] = acc[q].y; o[2] = acc[q].z; o[3] = acc[q].w;
      }
    }
  }
  }   // head
}

// Backward of the gather (A9, `train.py:59`).  One warp per element i, same lane <-> neighbour mapping as the forward.
// With G = dL/dout_i:  dalpha_k = G . whj[c_k];  d whj[c_k] += alpha_k G (atomics: a node is the neighbour of ~K
// elements);  de_k = alpha_k (dalpha_k - sum_k' alpha_k' dalpha_k');  dz_k = de_k * LeakyReLU'(s_i + t_c + b)
// (0 for padded neighbours);  ds_i = sum_k dz_k;  dt[c_k] += dz_k;  db += sum_k dz_k.
// The projections' own backward (GEMMs against W_i / W_j / a) stays with autograd.
__global__ void __launch_bounds__(GAT_THREADS)
gat_bwd_kernel(const float* __restrict__ gout, int64_t ld_go, const float* __restrict__ whj, int64_t ld_whj,
               const float* __restrict__ s_vec, const float* __restrict__ t_vec, int64_t ld_st, float att_b, float alpha,
               const int64_t* __restrict__ ctx, const float* __restrict__ attn, int T, int K, int Hd,
               float* __restrict__ d_whj, int64_t ld_dw, float* __restrict__ d_s, float* __restrict__ d_t, int64_t ld_dst,
               float* __restrict__ d_b) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * GAT_E + warp;
  if (i >= T) return;
  const int nvec = Hd >> 2;
  float4 g[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int v = lane + 32 * q;
    g[q] = v < nvec ? __ldg(reinterpret_cast<const float4*>(gout + (size_t)i * ld_go) + v) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  int cid[GAT_KMAX / 32];
  float a[GAT_KMAX / 32], da[GAT_KMAX / 32];
#pragma unroll
  for (int j = 0; j < GAT_KMAX / 32; ++j) {
    const int k = lane + 32 * j;
    const int64_t c64 = k < K ? ctx[(size_t)i * K + k] : -1;
    cid[j] = (c64 >= 0 && c64 < T) ? (int)c64 : -1;
    a[j] = k < K ? attn[(size_t)i * K + k] : 0.f;
    da[j] = 0.f;
  }
#pragma unroll
  for (int j = 0; j < GAT_KMAX / 32; ++j) {
    if (32 * j >= K) break;
    const int kend = min(32, K - 32 * j);
    for (int kk = 0; kk < kend; ++kk) {
      const int c = __shfl_sync(0xffffffffu, cid[j], kk);
      const float ak = __shfl_sync(0xffffffffu, a[j], kk);
      if (c < 0) continue;                                  // zero row: no gradient anywhere
      const float4* row = reinterpret_cast<const float4*>(whj + (size_t)c * ld_whj);
      float* drow = d_whj + (size_t)c * ld_dw;
      float dot = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int v = lane + 32 * q;
        if (v < nvec) {
          const float4 x = __ldg(row + v);
          dot += g[q].x * x.x + g[q].y * x.y + g[q].z * x.z + g[q].w * x.w;
          atomicAdd(drow + 4 * v + 0, ak * g[q].x);
          atomicAdd(drow + 4 * v + 1, ak * g[q].y);
          atomicAdd(drow + 4 * v + 2, ak * g[q].z);
          atomicAdd(drow + 4 * v + 3, ak * g[q].w);
        }
      }
      dot = warp_sum(dot);
      if (lane == kk) da[j] = dot;
    }
  }
  float sdot = 0.f;
#pragma unroll
  for (int j = 0; j < GAT_KMAX / 32; ++j) sdot += a[j] * da[j];
  sdot = warp_sum(sdot);
  const float si = s_vec[(size_t)i * ld_st];
  float dsum = 0.f;
#pragma unroll
  for (int j = 0; j < GAT_KMAX / 32; ++j) {
    const int k = lane + 32 * j;
    if (k < K && cid[j] >= 0) {
      const float z = si + __ldg(t_vec + (size_t)cid[j] * ld_st) + att_b;
      const float dz = a[j] * (da[j] - sdot) * (z > 0.f ? 1.f : alpha);
      dsum += dz;
      atomicAdd(d_t + (size_t)cid[j] * ld_dst, dz);
    }
  }
  dsum = warp_sum(dsum);
  if (lane == 0) {
    d_s[(size_t)i * ld_dst] = dsum;     // s_i only feeds element i: plain store
    atomicAdd(d_b, dsum);
  }
}

}  // namespace cova

extern "C" int cova_gat_bwd(const float* grad_out, int64_t ld_go, const float* whj, int64_t ld_whj, const float* s,
                            const float* t, int64_t ld_st, float att_b, float alpha, const int64_t* ctx_idx,
                            const float* attn, int T, int K, int Hd, float* d_whj, int64_t ld_dw, float* d_s, float* d_t,
                            int64_t ld_dst, float* d_b, void* stream) {
  using namespace cova;
  COVA_REQUIRE(T >= 0 && K >= 1 && K <= GAT_KMAX && Hd > 0 && Hd % 4 == 0 && Hd <= 512, "cova_gat_bwd: bad dims");
  if (T == 0) return COVA_OK;
  COVA_REQUIRE(grad_out && whj && s && t && ctx_idx && attn && d_whj && d_s && d_t && d_b, "cova_gat_bwd: null pointer");
  COVA_REQUIRE(ld_go % 4 == 0 && ld_whj % 4 == 0 && ((uintptr_t)grad_out & 15) == 0 && ((uintptr_t)whj & 15) == 0,
               "cova_gat_bwd: grad_out / whj rows must be 16-byte aligned");
  gat_bwd_kernel<<<ceil_div(T, GAT_E), GAT_THREADS, 0, (cudaStream_t)stream>>>(
      grad_out, ld_go, whj, ld_whj, s, t, ld_st, att_b, alpha, ctx_idx, attn, T, K, Hd, d_whj, ld_dw, d_s, d_t, ld_dst, d_b);
  COVA_LAUNCH_OK();
  return COVA_OK;
}

extern "C" int cova_gat_multihead_fwd(const float* ext, int64_t ld_ext, int Hd, int n_heads, const float* h_att_b,
                                      float alpha, const int64_t* ctx_idx, int T, int K, float* out, int64_t ld_out,
                                      float* attn, void* stream) {
  using namespace cova;
  COVA_REQUIRE(T >= 0 && K >= 0 && Hd > 0 && n_heads >= 1 && n_heads <= GAT_MAX_HEADS, "cova_gat_multihead_fwd: bad dims");
  if (T == 0) return COVA_OK;
  COVA_REQUIRE(ext && h_att_b && out && ctx_idx, "cova_gat_multihead_fwd: null pointer");
  COVA_REQUIRE(K >= 1 && K <= GAT_KMAX, "cova_gat_multihead_fwd: K=%d outside [1,%d]", K, GAT_KMAX);
  COVA_REQUIRE(Hd % 4 == 0 && ld_ext % 4 == 0 && ld_ext >= (int64_t)n_heads * (Hd + 2) && ld_out >= (int64_t)n_heads * Hd,
               "cova_gat_multihead_fwd: Hd and ld_ext must be multiples of 4; ext = [whj_0..whj_H-1 | s_0 t_0 .. | pad]");
  COVA_REQUIRE(((uintptr_t)ext & 15) == 0, "cova_gat_multihead_fwd: ext must be 16-byte aligned");
  const int out_vec = (ld_out % 4 == 0) && (((uintptr_t)out & 15) == 0);
  GatBias bias;
  for (int h = 0; h < GAT_MAX_HEADS; ++h) bias.b[h] = h < n_heads ? h_att_b[h] : 0.f;
  const float* st = ext + (size_t)n_heads * Hd;
  COVA_CUDA_OK(cudaFuncSetAttribute(gat_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GAT_STAGE_BYTES));
  gat_fwd_kernel<<<ceil_div(T, GAT_E), GAT_THREADS, GAT_STAGE_BYTES, (cudaStream_t)stream>>>(
      ext, ld_ext, st, st + 1, ld_ext, bias, alpha, ctx_idx, T, K, Hd, out, ld_out, attn, 1, out_vec, n_heads);
  COVA_LAUNCH_OK();
  return COVA_OK;
}

extern "C" int cova_gat_fwd(const float* whj, int64_t ld_whj, const float* s, const float* t, int64_t ld_st, float att_b,
                            float alpha, const int64_t* ctx_idx, int T, int K, int Hd, float* out, int64_t ld_out,
                            float* attn, void* stream) {
  using namespace cova;
  COVA_REQUIRE(T >= 0 && K >= 0 && Hd > 0, "cova_gat_fwd: bad dims");
  if (T == 0) return COVA_OK;
  COVA_REQUIRE(whj && s && t && out && (ctx_idx || K == 0), "cova_gat_fwd: null pointer");
  COVA_REQUIRE(ld_st >= 1, "cova_gat_fwd: ld_st must be >= 1");
  COVA_REQUIRE(K >= 1 && K <= GAT_KMAX, "cova_gat_fwd: K=%d outside [1,%d]", K, GAT_KMAX);
  COVA_REQUIRE(Hd % 4 == 0 && ld_whj % 4 == 0 && ld_whj >= Hd && ld_out >= Hd,
               "cova_gat_fwd: Hd and ld_whj must be multiples of 4 (float4 rows), ld >= Hd");
  COVA_REQUIRE(((uintptr_t)whj & 15) == 0, "cova_gat_fwd: whj must be 16-byte aligned");
  const int out_vec = (ld_out % 4 == 0) && (((uintptr_t)out & 15) == 0);
  COVA_CUDA_OK(cudaFuncSetAttribute(gat_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GAT_STAGE_BYTES));
  GatBias bias;
  for (int h = 0; h < GAT_MAX_HEADS; ++h) bias.b[h] = att_b;
  gat_fwd_kernel<<<ceil_div(T, GAT_E), GAT_THREADS, GAT_STAGE_BYTES, (cudaStream_t)stream>>>(
      whj, ld_whj, s, t, ld_st, bias, alpha, ctx_idx, T, K, Hd, out, ld_out, attn, 1, out_vec, 1);
  COVA_LAUNCH_OK();
  return COVA_OK;
}
