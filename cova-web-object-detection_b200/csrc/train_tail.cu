// The callers either side of `CoVA.forward`, as small HBM-/latency-bound kernels:
//   * cova_ce_sum_fwd_bwd   `nn.CrossEntropyLoss(reduction="sum")` (`/root/reference/main.py:139`, applied
//                           `train.py:56`) forward + gradient w.r.t. the logits in ONE pass, plus the
//                           `(output.argmax(1) == labels).sum()` of `train.py:53-54`
//   * cova_adam_step        `torch.optim.Adam(lr, weight_decay)` (`main.py:133-135`, stepped `train.py:60`) over ONE
//                           flat fp32 bucket - the same bucket the NCCL gradient all-reduce uses
//   * cova_topk_hits        the per-page / per-class top-k test of `evaluate_model` (`train.py:131-154`)
//   * cova_build_batch      `WebDataset.__getitem__` context window (`datasets.py:117-128`) + `custom_collate_fn`'s
//                           batch-index column and batch-global ids (`datasets.py:170-178`), built on the device
//                           from per-page box counts
#include <float.h>
#include <limits.h>
#include <math.h>

#include "common.cuh"

namespace cova {

constexpr int CE_THREADS = 256;
constexpr int CE_MAX_CLS = 32;

// One thread per row (n_cls is tiny: 4); block-level deterministic tree reduction, then ONE atomicAdd per block
// (double, so the order of the <= a-few-dozen block partials does not show in the fp32 result).
__global__ void __launch_bounds__(CE_THREADS)
ce_sum_kernel(const float* __restrict__ logits, int64_t ld, const int64_t* __restrict__ labels, int T, int C,
              int64_t ignore_index, double* __restrict__ loss_acc, float* __restrict__ dlogits, int64_t ld_d,
              int* __restrict__ n_correct) {
  __shared__ double s_loss[CE_THREADS / 32];
  __shared__ int s_corr[CE_THREADS / 32];
  const int i = blockIdx.x * CE_THREADS + threadIdx.x;
  double li = 0.0;
  int corr = 0;
  if (i < T) {
    float x[CE_MAX_CLS];
    float mx = -FLT_MAX;
    int am = 0;
#pragma unroll
    for (int c = 0; c < CE_MAX_CLS; ++c) {
      if (c < C) {
        x[c] = logits[(size_t)i * ld + c];
        if (x[c] > mx) { mx = x[c]; am = c; }     // first maximum wins, like torch.argmax
      }
    }
    float se = 0.f;
#pragma unroll
    for (int c = 0; c < CE_MAX_CLS; ++c)
      if (c < C) se += expf(x[c] - mx);
    const float lse = mx + logf(se);
    const int64_t y = labels[i];
    const bool live = y != ignore_index && y >= 0 && y < C;
    if (live) {
      li = (double)(lse - x[(int)y]);
      corr = (am == (int)y);
    }
    if (dlogits != nullptr) {
#pragma unroll
      for (int c = 0; c < CE_MAX_CLS; ++c)
        if (c < C) dlogits[(size_t)i * ld_d + c] = live ? expf(x[c] - lse) - (c == (int)y ? 1.f : 0.f) : 0.f;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    li += __shfl_xor_sync(0xffffffffu, li, o);
    corr += __shfl_xor_sync(0xffffffffu, corr, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { s_loss[warp] = li; s_corr[warp] = corr; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    int k = 0;
    for (int w = 0; w < CE_THREADS / 32; ++w) { t += s_loss[w]; k += s_corr[w]; }
    atomicAdd(loss_acc, t);
    if (n_correct != nullptr) atomicAdd(n_correct, k);
  }
}

__global__ void ce_finish_kernel(const double* __restrict__ acc, float* __restrict__ loss) { *loss = (float)*acc; }

// torch.optim.Adam (amsgrad=False, maximize=False), single-tensor formulation of torch/optim/adam.py:
//   g += wd*p;  m.lerp_(g, 1-b1);  v = b2*v + (1-b2)*g*g;  p -= (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps)
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n,
            float omb1, float b2, float omb2, float eps, float wd, float step_size, float sqrt_bc2, float grad_scale) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 4 <= n && (((uintptr_t)(p + i) | (uintptr_t)(g + i) | (uintptr_t)(m + i) | (uintptr_t)(v + i)) & 15) == 0) {
      float4 pp = *reinterpret_cast<float4*>(p + i), gg = *reinterpret_cast<const float4*>(g + i);
      float4 mm = *reinterpret_cast<float4*>(m + i), vv = *reinterpret_cast<float4*>(v + i);
      float* pa = &pp.x; float* ga = &gg.x; float* ma = &mm.x; float* va = &vv.x;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float gr = ga[e] * grad_scale;
        gr = fmaf(wd, pa[e], gr);
        ma[e] = fmaf(omb1, gr - ma[e], ma[e]);
        va[e] = fmaf(omb2, gr * gr, b2 * va[e]);
        const float denom = sqrtf(va[e]) / sqrt_bc2 + eps;
        pa[e] = pa[e] - step_size * (ma[e] / denom);
      }
      *reinterpret_cast<float4*>(p + i) = pp;
      *reinterpret_cast<float4*>(m + i) = mm;
      *reinterpret_cast<float4*>(v + i) = vv;
    } else {
      for (int64_t j = i; j < n && j < i + 4; ++j) {
        float gr = g[j] * grad_scale;
        gr = fmaf(wd, p[j], gr);
        const float mj = fmaf(omb1, gr - m[j], m[j]);
        const float vj = fmaf(omb2, gr * gr, b2 * v[j]);
        m[j] = mj; v[j] = vj;
        p[j] = p[j] - step_size * (mj / (sqrtf(vj) / sqrt_bc2 + eps));
      }
    }
  }
}

// One warp per (page, class c >= 1).  true row = FIRST row of the page whose label is c (train.py:146); it is a hit
// iff fewer than k rows of the page rank above it in `torch.argsort(output_img, dim=0)[n-k:]` (train.py:141-149) taken
// as a STABLE ascending sort: row j outranks row i iff v_j > v_i or (v_j == v_i and j > i).  No row with label c: -1
// (the reference raises IndexError there).
__global__ void topk_hits_kernel(const float* __restrict__ logits, int64_t ld, const int64_t* __restrict__ labels,
                                 const int* __restrict__ page_off, int B, int C, int k, int* __restrict__ hits) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= B * (C - 1)) return;
  const int b = w / (C - 1), c = 1 + w % (C - 1);
  const int r0 = page_off[b], r1 = page_off[b + 1];
  int first = INT_MAX;
  for (int r = r0 + lane; r < r1; r += 32)
    if (labels[r] == c) { first = r; break; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
  if (first == INT_MAX) {
    if (lane == 0) hits[b * C + c] = -1;
    return;
  }
  const float vt = logits[(size_t)first * ld + c];
  int above = 0;
  for (int r = r0 + lane; r < r1; r += 32) {
    const float v = logits[(size_t)r * ld + c];
    above += (v > vt) || (v == vt && r > first);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) above += __shfl_xor_sync(0xffffffffu, above, o);
  if (lane == 0) hits[b * C + c] = above < k ? 1 : 0;
}

// Thread per (row, slot).  Row i of a page with n rows: context = [max(0,i-cs) .. i-1] ++ [i+1 .. min(n,i+cs+1)-1], right
// padded with -1 (datasets.py:120-127), then shifted by the page's first row (datasets.py:175).  Slot 0 also writes the
// collated box [page, x1, y1, x1+w, y1+h] (datasets.py:114-115, :172-174) when raw [x,y,w,h] boxes are given.
__global__ void build_batch_kernel(const int* __restrict__ page_off, int B, int cs, const float* __restrict__ xywh,
                                   float* __restrict__ bboxes, int64_t* __restrict__ ctx, int T) {
  const int K = 2 * cs;
  const int slots = K > 0 ? K : 1;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)T * slots) return;
  const int row = (int)(idx / slots), k = (int)(idx % slots);
  int lo = 0, hi = B;                       // page of `row`: largest b with page_off[b] <= row
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (page_off[mid] <= row) lo = mid; else hi = mid;
  }
  const int r0 = page_off[lo], n = page_off[lo + 1] - r0, i = row - r0;
  if (K > 0) {
    const int left0 = max(0, i - cs), nleft = i - left0;
    const int nright = min(n, i + cs + 1) - (i + 1);
    int64_t c = -1;
    if (k < nleft) c = r0 + left0 + k;
    else if (k < nleft + nright) c = r0 + i + 1 + (k - nleft);
    ctx[(size_t)row * K + k] = c;
  }
  if (k == 0 && bboxes != nullptr) {
    const float x = xywh[row * 4 + 0], y = xywh[row * 4 + 1], w = xywh[row * 4 + 2], h = xywh[row * 4 + 3];
    float* o = bboxes + (size_t)row * 5;
    o[0] = (float)lo; o[1] = x; o[2] = y; o[3] = x + w; o[4] = y + h;
  }
}

}  // namespace cova

using namespace cova;

extern "C" int cova_ce_sum_fwd_bwd(const float* logits, int64_t ld, const int64_t* labels, int T, int n_cls,
                                   int64_t ignore_index, double* ws_acc, float* loss, float* dlogits, int64_t ld_d,
                                   int* n_correct, void* stream) {
  COVA_REQUIRE(T >= 0 && n_cls >= 1 && n_cls <= CE_MAX_CLS, "cova_ce_sum_fwd_bwd: n_cls=%d outside [1,%d]", n_cls, CE_MAX_CLS);
  COVA_REQUIRE(ws_acc && loss && (T == 0 || (logits && labels)), "cova_ce_sum_fwd_bwd: null pointer");
  COVA_REQUIRE(ld >= n_cls && (!dlogits || ld_d >= n_cls), "cova_ce_sum_fwd_bwd: row stride smaller than n_cls");
  cudaStream_t st = (cudaStream_t)stream;
  COVA_CUDA_OK(cudaMemsetAsync(ws_acc, 0, sizeof(double), st));
  if (n_correct) COVA_CUDA_OK(cudaMemsetAsync(n_correct, 0, sizeof(int), st));
  if (T > 0) {
    ce_sum_kernel<<<ceil_div(T, CE_THREADS), CE_THREADS, 0, st>>>(logits, ld, labels, T, n_cls, ignore_index, ws_acc,
                                                                   dlogits, ld_d, n_correct);
    COVA_LAUNCH_OK();
  }
  ce_finish_kernel<<<1, 1, 0, st>>>(ws_acc, loss);
  COVA_LAUNCH_OK();
  return COVA_OK;
}

extern "C" int cova_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, double lr,
                              double beta1, double beta2, double eps, double weight_decay, int step, double grad_scale,
                              void* stream) {
  COVA_REQUIRE(n >= 0 && step >= 1, "cova_adam_step: bad n / step");
  if (n == 0) return COVA_OK;
  COVA_REQUIRE(param && grad && exp_avg && exp_avg_sq, "cova_adam_step: null pointer");
  // bias corrections in double on the host, exactly the scalars torch/optim/adam.py computes in Python floats
  const double bc1 = 1.0 - pow(beta1, step), bc2 = 1.0 - pow(beta2, step);
  const float step_size = (float)(lr / bc1), sqrt_bc2 = (float)sqrt(bc2);
  int64_t blocks = (n + 256 * 4 - 1) / (256 * 4);
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  adam_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(
      param, grad, exp_avg, exp_avg_sq, n, (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2), (float)eps,
      (float)weight_decay, step_size, sqrt_bc2, (float)grad_scale);
  COVA_LAUNCH_OK();
  return COVA_OK;
}

extern "C" int cova_topk_hits(const float* logits, int64_t ld, const int64_t* labels, const int* page_offsets, int B,
                              int n_cls, int k, int* hits, void* stream) {
  COVA_REQUIRE(B >= 0 && n_cls >= 1 && k >= 1, "cova_topk_hits: bad dims");
  if (B == 0 || n_cls == 1) return COVA_OK;
  COVA_REQUIRE(logits && labels && page_offsets && hits && ld >= n_cls, "cova_topk_hits: bad arguments");
  const int warps = B * (n_cls - 1);
  topk_hits_kernel<<<ceil_div(warps * 32, 128), 128, 0, (cudaStream_t)stream>>>(logits, ld, labels, page_offsets, B, n_cls,
                                                                               k, hits);
  COVA_LAUNCH_OK();
  return COVA_OK;
}

extern "C" int cova_build_batch(const int* page_offsets, int B, int T, int context_size, const float* boxes_xywh,
                                float* bboxes, int64_t* context_indices, void* stream) {
  COVA_REQUIRE(B >= 0 && T >= 0 && context_size >= 0, "cova_build_batch: bad dims");
  if (B == 0 || T == 0) return COVA_OK;
  COVA_REQUIRE(page_offsets && (context_size == 0 || context_indices), "cova_build_batch: null pointer");
  COVA_REQUIRE((boxes_xywh == nullptr) == (bboxes == nullptr), "cova_build_batch: boxes_xywh and bboxes come together");
  const int64_t total = (int64_t)T * (context_size > 0 ? 2 * context_size : 1);
  build_batch_kernel<<<(int)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(page_offsets, B, context_size,
                                                                                  boxes_xywh, bboxes, context_indices, T);
  COVA_LAUNCH_OK();
  return COVA_OK;
}
