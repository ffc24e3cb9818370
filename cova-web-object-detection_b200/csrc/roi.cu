// A4 / A4' / A5: RoIPool (bit-exact), RoIAlign and the positional encoder, writing straight into the
// `own` feature row so the concat of `/root/reference/models.py:110` costs no pass of its own.
//
// Replaces `torchvision.ops.RoIPool(P, scale)` (`models.py:58`, applied `:125-127`), the D1 variant
// `RoIAlign(P, scale, sampling_ratio, aligned=False)`, and `_get_bbox_features` + `bbox_feat_encoder`
// (`models.py:129-148`, `:65-70`).  Feature map is NHWC fp32: a pixel's 64 channels are 256 contiguous
// bytes, a crop row is one contiguous run, so every warp load is a full 512 B (two pixels x 16 lanes x 16 B).
//
// HBM-bound: algorithmic bytes per box = crop_h*crop_w*C*4 read + C*P*P*4 written.
#include <float.h>
#include <math.h>

#include "common.cuh"

namespace cova {

constexpr int ROI_THREADS = 256;
constexpr int ROI_WARPS = ROI_THREADS / 32;
constexpr int ROI_CB = 64;   // channels per CTA (grid.y walks channel blocks)

struct RoiGeom {
  int b, sw, sh, rw, rh;
  float bin_h, bin_w;
};

// box edge quantisation of torchvision's roi_pool kernel: C round() of the fp32 product
__device__ __forceinline__ RoiGeom roi_geom(const float* __restrict__ roi, float scale, int PH, int PW) {
  RoiGeom g;
  g.b = (int)roi[0];
  g.sw = (int)roundf(roi[1] * scale);
  g.sh = (int)roundf(roi[2] * scale);
  const int ew = (int)roundf(roi[3] * scale), eh = (int)roundf(roi[4] * scale);
  g.rw = max(ew - g.sw + 1, 1);
  g.rh = max(eh - g.sh + 1, 1);
  g.bin_h = (float)g.rh / (float)PH;
  g.bin_w = (float)g.rw / (float)PW;
  return g;
}
__device__ __forceinline__ int bin_lo(int p, float bin, int start, int limit) {
  return min(max((int)floorf((float)p * bin) + start, 0), limit);
}
__device__ __forceinline__ int bin_hi(int p, float bin, int start, int limit) {
  return min(max((int)ceilf((float)(p + 1) * bin) + start, 0), limit);
}

// One CTA = one box x one 64-channel block.  Warps take crop rows round-robin; per row a warp walks the
// PW column segments, two pixels per load (half-warp each, float4 per lane), and folds each segment's
// max into its private smem accumulators for the (<=2) row bins the row belongs to.
// ROWSPLIT: one CTA per (box, row bin) instead of per box - crop areas vary 1000x (8 px ... full-width boxes), and a
// third of a box per CTA (adjacent row bins share at most one crop row) evens out the tail; ROI_SPLIT_WARPS warps.
constexpr int ROI_SPLIT_WARPS = 4;
template <bool WITH_ARGMAX, bool ROWSPLIT>
__global__ void __launch_bounds__(ROI_THREADS)
roi_pool_kernel(const float* __restrict__ fm, int Hf, int Wf, int C, const float* __restrict__ rois, int PH,
                int PW, float scale, float* __restrict__ out, int64_t ld_out, int32_t* __restrict__ argmax, int B) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int NW = ROWSPLIT ? ROI_SPLIT_WARPS : ROI_WARPS;
  const int nbins = PH * PW;                       // bins of the box
  const int abins = ROWSPLIT ? PW : nbins;         // bins this CTA accumulates
  float* acc = reinterpret_cast<float*>(smem_raw);                 // [warp][bin][64]
  int* acc_i = reinterpret_cast<int*>(acc + NW * abins * ROI_CB);  // same shape, only WITH_ARGMAX
  const int t = ROWSPLIT ? blockIdx.x / PH : blockIdx.x, cb = blockIdx.y * ROI_CB;
  const int ph_own = ROWSPLIT ? blockIdx.x % PH : 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int half = lane >> 4, c4 = (lane & 15) * 4;
  const RoiGeom g = roi_geom(rois + (size_t)t * 5, scale, PH, PW);

  for (int i = threadIdx.x; i < NW * abins * ROI_CB; i += NW * 32) {
    acc[i] = -FLT_MAX;
    if (WITH_ARGMAX) acc_i[i] = -1;
  }
  __syncthreads();

  const int h_lo = bin_lo(ROWSPLIT ? ph_own : 0, g.bin_h, g.sh, Hf);
  const int h_hi = bin_hi(ROWSPLIT ? ph_own : PH - 1, g.bin_h, g.sh, Hf);
  const int bq = min(max(g.b, 0), B - 1);
  const float* base = fm + (size_t)bq * Hf * Wf * C + cb;
  float* wacc = acc + warp * abins * ROI_CB;
  int* wacc_i = acc_i + warp * abins * ROI_CB;

  for (int h = h_lo + warp; h < h_hi; h += NW) {
    const float* row = base + (size_t)h * Wf * C;
    for (int pw = 0; pw < PW; ++pw) {
      const int ws = bin_lo(pw, g.bin_w, g.sw, Wf), we = bin_hi(pw, g.bin_w, g.sw, Wf);
      if (we <= ws) continue;
      float4 m = make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX);
      int4 mi = make_int4(-1, -1, -1, -1);
      int w = ws + half;
#pragma unroll 4
      for (; w < we; w += 2) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(row + (size_t)w * C + c4));
        if (WITH_ARGMAX) {
          const int idx = h * Wf + w;   // ascending per lane, so strict '>' keeps the first maximum
          if (v.x > m.x) { m.x = v.x; mi.x = idx; }
          if (v.y > m.y) { m.y = v.y; mi.y = idx; }
          if (v.z > m.z) { m.z = v.z; mi.z = idx; }
          if (v.w > m.w) { m.w = v.w; mi.w = idx; }
        } else {
          m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
        }
      }
      // merge the two half-warps (odd/even pixels)
      float4 o;
      o.x = __shfl_xor_sync(0xffffffffu, m.x, 16); o.y = __shfl_xor_sync(0xffffffffu, m.y, 16);
      o.z = __shfl_xor_sync(0xffffffffu, m.z, 16); o.w = __shfl_xor_sync(0xffffffffu, m.w, 16);
      if (WITH_ARGMAX) {
        int4 oi;
        oi.x = __shfl_xor_sync(0xffffffffu, mi.x, 16); oi.y = __shfl_xor_sync(0xffffffffu, mi.y, 16);
        oi.z = __shfl_xor_sync(0xffffffffu, mi.z, 16); oi.w = __shfl_xor_sync(0xffffffffu, mi.w, 16);
#define MERGE(f)                                                                  \
  if (oi.f >= 0 && (mi.f < 0 || o.f > m.f || (o.f == m.f && oi.f < mi.f))) {      \
    m.f = o.f;                                                                    \
    mi.f = oi.f;                                                                  \
  }
        MERGE(x) MERGE(y) MERGE(z) MERGE(w)
#undef MERGE
      } else {
        m.x = fmaxf(m.x, o.x); m.y = fmaxf(m.y, o.y); m.z = fmaxf(m.z, o.z); m.w = fmaxf(m.w, o.w);
      }
      if (half == 0) {
        for (int ph = 0; ph < (ROWSPLIT ? 1 : PH); ++ph) {   // rows belong to at most two adjacent row bins
          if (!ROWSPLIT && (h < bin_lo(ph, g.bin_h, g.sh, Hf) || h >= bin_hi(ph, g.bin_h, g.sh, Hf))) continue;
          float4* a = reinterpret_cast<float4*>(wacc + (ph * PW + pw) * ROI_CB + c4);
          float4 cur = *a;
          if (WITH_ARGMAX) {
            int4* ai = reinterpret_cast<int4*>(wacc_i + (ph * PW + pw) * ROI_CB + c4);
            int4 ci = *ai;
#define FOLD(f)                                                                       \
  if (mi.f >= 0 && (ci.f < 0 || m.f > cur.f || (m.f == cur.f && mi.f < ci.f))) {      \
    cur.f = m.f;                                                                      \
    ci.f = mi.f;                                                                      \
  }
            FOLD(x) FOLD(y) FOLD(z) FOLD(w)
#undef FOLD
            *ai = ci;
          } else {
            cur.x = fmaxf(cur.x, m.x); cur.y = fmaxf(cur.y, m.y); cur.z = fmaxf(cur.z, m.z); cur.w = fmaxf(cur.w, m.w);
          }
          *a = cur;
        }
      }
    }
  }
  __syncthreads();

  // cross-warp reduce; output index = c*PH*PW + bin  (the `.view(T, C*P*P)` order of models.py:125-127)
  for (int o = threadIdx.x; o < ROI_CB * abins; o += NW * 32) {
    const int c = o / abins, abin = o % abins;
    const int bin = ROWSPLIT ? ph_own * PW + abin : abin;
    const int ph = bin / PW, pw = bin % PW;
    const bool empty = bin_hi(ph, g.bin_h, g.sh, Hf) <= bin_lo(ph, g.bin_h, g.sh, Hf) ||
                       bin_hi(pw, g.bin_w, g.sw, Wf) <= bin_lo(pw, g.bin_w, g.sw, Wf);
    float m = -FLT_MAX;
    int mi = -1;
    for (int w = 0; w < NW; ++w) {
      const float v = acc[(w * abins + abin) * ROI_CB + c];
      if (WITH_ARGMAX) {
        const int vi = acc_i[(w * abins + abin) * ROI_CB + c];
        if (vi >= 0 && (mi < 0 || v > m || (v == m && vi < mi))) { m = v; mi = vi; }
      } else {
        m = fmaxf(m, v);
      }
    }
    if (empty) { m = 0.f; mi = -1; }
    out[(size_t)t * ld_out + (size_t)(cb + c) * nbins + bin] = m;
    if (WITH_ARGMAX) argmax[((size_t)t * C + cb + c) * nbins + bin] = mi;
  }
  // A batch index outside [0, B) (never produced by the loader, `datasets.py:170-177`) selects no page: the box was pooled
  // from a clamped (in-bounds) page above; its outputs are overwritten here with 0 / arg-max -1 (no gradient) by the same
  // threads that wrote them.  Kept OUT of the loop above on purpose: folding the test into `empty` cost the row-split
  // kernel 30 % (136 -> 178 us at config 2, tools/bench_kernels.py) through the compiler's register allocation.
  if ((unsigned)(int)__ldg(rois + (size_t)t * 5) >= (unsigned)B) {
    for (int o = threadIdx.x; o < ROI_CB * abins; o += NW * 32) {
      const int c = o / abins, abin = o % abins;
      const int bin = ROWSPLIT ? ph_own * PW + abin : abin;
      out[(size_t)t * ld_out + (size_t)(cb + c) * nbins + bin] = 0.f;
      if (WITH_ARGMAX) argmax[((size_t)t * C + cb + c) * nbins + bin] = -1;
    }
  }
}

// torchvision roi_align bilinear sample (clamp-to-edge, zero outside [-1, size])
__device__ __forceinline__ float4 bilinear4(const float* __restrict__ base, int Hf, int Wf, int C, float y, float x) {
  if (y < -1.0f || y > (float)Hf || x < -1.0f || x > (float)Wf) return make_float4(0.f, 0.f, 0.f, 0.f);
  y = fmaxf(y, 0.f);
  x = fmaxf(x, 0.f);
  int yl = (int)y, xl = (int)x, yh, xh;
  if (yl >= Hf - 1) { yh = yl = Hf - 1; y = (float)yl; } else { yh = yl + 1; }
  if (xl >= Wf - 1) { xh = xl = Wf - 1; x = (float)xl; } else { xh = xl + 1; }
  const float ly = y - (float)yl, lx = x - (float)xl, hy = 1.f - ly, hx = 1.f - lx;
  const float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
  const float4 a = __ldg(reinterpret_cast<const float4*>(base + ((size_t)yl * Wf + xl) * C));
  const float4 b = __ldg(reinterpret_cast<const float4*>(base + ((size_t)yl * Wf + xh) * C));
  const float4 c = __ldg(reinterpret_cast<const float4*>(base + ((size_t)yh * Wf + xl) * C));
  const float4 d = __ldg(reinterpret_cast<const float4*>(base + ((size_t)yh * Wf + xh) * C));
  float4 r;   // same association order as torchvision: w1*v1 + w2*v2 + w3*v3 + w4*v4
  r.x = w1 * a.x + w2 * b.x + w3 * c.x + w4 * d.x;
  r.y = w1 * a.y + w2 * b.y + w3 * c.y + w4 * d.y;
  r.z = w1 * a.z + w2 * b.z + w3 * c.z + w4 * d.z;
  r.w = w1 * a.w + w2 * b.w + w3 * c.w + w4 * d.w;
  return r;
}

// Branch-free form of the same sample: 4 tap offsets (pixels) + 4 weights; a sample outside [-1, size] gets zero
// weights on a valid address, so all taps of all samples can be loaded back to back.
struct BilinearTaps {
  int o[4];
  float w[4];
};
__device__ __forceinline__ BilinearTaps bilinear_taps(int Hf, int Wf, float y, float x) {
  BilinearTaps t;
  const bool valid = !(y < -1.0f || y > (float)Hf || x < -1.0f || x > (float)Wf);
  y = fminf(fmaxf(y, 0.f), (float)Hf);
  x = fminf(fmaxf(x, 0.f), (float)Wf);
  int yl = (int)y, xl = (int)x, yh, xh;
  if (yl >= Hf - 1) { yh = yl = Hf - 1; y = (float)yl; } else { yh = yl + 1; }
  if (xl >= Wf - 1) { xh = xl = Wf - 1; x = (float)xl; } else { xh = xl + 1; }
  const float ly = y - (float)yl, lx = x - (float)xl, hy = 1.f - ly, hx = 1.f - lx;
  const float m = valid ? 1.f : 0.f;
  t.w[0] = hy * hx * m; t.w[1] = hy * lx * m; t.w[2] = ly * hx * m; t.w[3] = ly * lx * m;
  t.o[0] = yl * Wf + xl; t.o[1] = yl * Wf + xh; t.o[2] = yh * Wf + xl; t.o[3] = yh * Wf + xh;
  return t;
}

// 16 lanes x float4 = one 64-channel block of one (box, bin); the T*P*P*(C/64) items are spread evenly over all
// half-warps of the grid (a CTA-per-box layout left 7 of 16 half-warps idle for 3x3 bins and was latency-bound).
template <int G>   // G = sampling_ratio when it is a compile-time 2 (all 16 bilinear taps in flight at once), 0 = runtime
__global__ void __launch_bounds__(ROI_THREADS)
roi_align_kernel(const float* __restrict__ fm, int B, int Hf, int Wf, int C, const float* __restrict__ rois, int T, int PH,
                 int PW, float scale, int sampling_ratio, float* __restrict__ out, int64_t ld_out) {
  const int nbins = PH * PW, ncb = C / ROI_CB;
  const int64_t n_items = (int64_t)T * nbins * ncb;
  const int64_t item = (int64_t)blockIdx.x * (ROI_THREADS / 16) + (threadIdx.x >> 4);
  if (item >= n_items) return;
  const int c4 = (threadIdx.x & 15) * 4;
  const int t = (int)(item / (nbins * ncb));
  const int rem = (int)(item % (nbins * ncb));
  const int bin = rem % nbins, cb = (rem / nbins) * ROI_CB;
  const int ph = bin / PW, pw = bin % PW;
  const float* roi = rois + (size_t)t * 5;
  const int b = (int)roi[0];
  const bool page_ok = b >= 0 && b < B;              // a batch index outside [0, B): zeros, not an out-of-bounds read
  const float x1 = roi[1] * scale, y1 = roi[2] * scale, x2 = roi[3] * scale, y2 = roi[4] * scale;
  const float rw = fmaxf(x2 - x1, 1.f), rh = fmaxf(y2 - y1, 1.f);
  const float bh = rh / (float)PH, bw = rw / (float)PW;
  const int gh = G > 0 ? G : (sampling_ratio > 0 ? sampling_ratio : (int)ceilf(rh / (float)PH));
  const int gw = G > 0 ? G : (sampling_ratio > 0 ? sampling_ratio : (int)ceilf(rw / (float)PW));
  const float cnt = (float)max(gh * gw, 1);
  const float* base = fm + (size_t)(page_ok ? b : 0) * Hf * Wf * C + cb + c4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (!page_ok) {
  } else if (G > 0) {
    constexpr int NS = G > 0 ? G * G : 1;
    BilinearTaps tp[NS];
    float4 v[NS][4];
#pragma unroll
    for (int i = 0; i < NS; ++i) {
      const float y = y1 + (float)ph * bh + ((float)(i / (G > 0 ? G : 1)) + 0.5f) * bh / (float)G;
      const float x = x1 + (float)pw * bw + ((float)(i % (G > 0 ? G : 1)) + 0.5f) * bw / (float)G;
      tp[i] = bilinear_taps(Hf, Wf, y, x);
    }
#pragma unroll
    for (int i = 0; i < NS; ++i)
#pragma unroll
      for (int k = 0; k < 4; ++k) v[i][k] = __ldg(reinterpret_cast<const float4*>(base + (size_t)tp[i].o[k] * C));
#pragma unroll
    for (int i = 0; i < NS; ++i) {   // per sample w1*v1 + w2*v2 + w3*v3 + w4*v4, samples summed in order (as torchvision)
      acc.x += tp[i].w[0] * v[i][0].x + tp[i].w[1] * v[i][1].x + tp[i].w[2] * v[i][2].x + tp[i].w[3] * v[i][3].x;
      acc.y += tp[i].w[0] * v[i][0].y + tp[i].w[1] * v[i][1].y + tp[i].w[2] * v[i][2].y + tp[i].w[3] * v[i][3].y;
      acc.z += tp[i].w[0] * v[i][0].z + tp[i].w[1] * v[i][1].z + tp[i].w[2] * v[i][2].z + tp[i].w[3] * v[i][3].z;
      acc.w += tp[i].w[0] * v[i][0].w + tp[i].w[1] * v[i][1].w + tp[i].w[2] * v[i][2].w + tp[i].w[3] * v[i][3].w;
    }
  } else {
    for (int iy = 0; iy < gh; ++iy) {
      const float y = y1 + (float)ph * bh + ((float)iy + 0.5f) * bh / (float)gh;
      for (int ix = 0; ix < gw; ++ix) {
        const float x = x1 + (float)pw * bw + ((float)ix + 0.5f) * bw / (float)gw;
        const float4 v = bilinear4(base, Hf, Wf, C, y, x);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    }
  }
  float* o = out + (size_t)t * ld_out + (size_t)(cb + c4) * nbins + bin;
  o[0] = acc.x / cnt;
  o[nbins] = acc.y / cnt;
  o[2 * nbins] = acc.z / cnt;
  o[3 * nbins] = acc.w / cnt;
}

// RoIAlign backward (torchvision roi_align backward): every output bin spreads grad / (gh*gw) over the 4 bilinear taps
// of each of its samples.  Same work split as the forward (16 lanes x float4 = one 64-channel block of one (box, bin));
// boxes and samples overlap freely -> vector fp32 atomics (RED.ADD.F32x4) into the zero-initialised NHWC gradient map.
__global__ void __launch_bounds__(ROI_THREADS)
roi_align_bwd_kernel(const float* __restrict__ grad_out, int64_t ld_go, const float* __restrict__ rois, int T, int C, int PH,
                     int PW, float scale, int sampling_ratio, int B, int Hf, int Wf, float* __restrict__ grad_fm) {
  const int nbins = PH * PW, ncb = C / ROI_CB;
  const int64_t n_items = (int64_t)T * nbins * ncb;
  const int64_t item = (int64_t)blockIdx.x * (ROI_THREADS / 16) + (threadIdx.x >> 4);
  if (item >= n_items) return;
  const int c4 = (threadIdx.x & 15) * 4;
  const int t = (int)(item / (nbins * ncb));
  const int rem = (int)(item % (nbins * ncb));
  const int bin = rem % nbins, cb = (rem / nbins) * ROI_CB;
  const int ph = bin / PW, pw = bin % PW;
  const float* roi = rois + (size_t)t * 5;
  const int b = (int)roi[0];
  if (b < 0 || b >= B) return;
  const float x1 = roi[1] * scale, y1 = roi[2] * scale, x2 = roi[3] * scale, y2 = roi[4] * scale;
  const float rw = fmaxf(x2 - x1, 1.f), rh = fmaxf(y2 - y1, 1.f);
  const float bh = rh / (float)PH, bw = rw / (float)PW;
  const int gh = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(rh / (float)PH);
  const int gw = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(rw / (float)PW);
  const float cnt = (float)max(gh * gw, 1);
  const float* go = grad_out + (size_t)t * ld_go + (size_t)(cb + c4) * nbins + bin;
  const float4 g = make_float4(go[0] / cnt, go[nbins] / cnt, go[2 * nbins] / cnt, go[3 * nbins] / cnt);
  float* base = grad_fm + (size_t)b * Hf * Wf * C + cb + c4;
  for (int iy = 0; iy < gh; ++iy) {
    const float y = y1 + (float)ph * bh + ((float)iy + 0.5f) * bh / (float)gh;
    for (int ix = 0; ix < gw; ++ix) {
      const float x = x1 + (float)pw * bw + ((float)ix + 0.5f) * bw / (float)gw;
      const BilinearTaps tp = bilinear_taps(Hf, Wf, y, x);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float w = tp.w[k];
        if (w != 0.f)
          atomicAdd(reinterpret_cast<float4*>(base + (size_t)tp.o[k] * C), make_float4(w * g.x, w * g.y, w * g.z, w * g.w));
      }
    }
  }
}

// RoIPool backward (torchvision roi_pool backward): every pooled output sends its gradient to its arg-max pixel.
// Bins of one box overlap by up to a pixel and boxes overlap freely, so several outputs hit the same pixel:
// fp32 atomicAdd (RED.ADD.F32) into the zero-initialised NHWC gradient map.
__global__ void roi_pool_bwd_kernel(const float* __restrict__ grad_out, int64_t ld_go, const int32_t* __restrict__ argmax,
                                    const float* __restrict__ rois, int T, int C, int nbins, int B, int Hf, int Wf,
                                    float* __restrict__ grad_fm) {
  const int64_t n = (int64_t)T * C * nbins;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int idx = argmax[i];
    if (idx < 0) continue;                         // empty bin: no gradient
    const int t = (int)(i / ((int64_t)C * nbins));
    const int r = (int)(i % ((int64_t)C * nbins));
    const int c = r / nbins;
    const int b = (int)rois[(size_t)t * 5];
    if (b < 0 || b >= B) continue;
    atomicAdd(grad_fm + ((size_t)b * Hf * Wf + idx) * C + c, grad_out[(size_t)t * ld_go + r]);
  }
}

// [x1,y1,w,h,w/h] -> Linear(5,D) -> folded BN1d -> ReLU; one thread per (box, d)
__global__ void bbox_enc_kernel(const float* __restrict__ rois, int T, const float* __restrict__ w,
                                const float* __restrict__ bias, const float* __restrict__ scale,
                                const float* __restrict__ shift, int D, float* __restrict__ out, int64_t ld_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T * D) return;
  const int t = i / D, d = i % D;
  const float* r = rois + (size_t)t * 5;
  const float x = r[1], y = r[2], bw = r[3] - r[1], bh = r[4] - r[2];
  const float f[5] = {x, y, bw, bh, bw / bh};   // models.py:134-142 (h == 0 -> inf/NaN, as the reference)
  float acc = 0.f;                              // fp32 dot in index order, then + bias (nn.Linear)
#pragma unroll
  for (int k = 0; k < 5; ++k) acc = fmaf(f[k], w[d * 5 + k], acc);
  acc += bias[d];
  if (scale) acc = fmaf(acc, scale[d], shift[d]);
  out[(size_t)t * ld_out + d] = fmaxf(acc, 0.f);
}

__global__ void affine_cols_kernel(const float* __restrict__ x, int T, int D, int64_t ld_x,
                                   const float* __restrict__ scale, const float* __restrict__ shift,
                                   float* __restrict__ out, int64_t ld_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T * D) return;
  const int t = i / D, d = i % D;
  float v = x[(size_t)t * ld_x + d];
  if (scale) v = fmaf(v, scale[d], shift[d]);
  out[(size_t)t * ld_out + d] = v;
}

}  // namespace cova

extern "C" int cova_roi_fwd(const float* fm, int B, int Hf, int Wf, int C, const float* rois, int T, int PH, int PW,
                            float spatial_scale, int mode, int sampling_ratio, float* out, int64_t ld_out,
                            int32_t* argmax, void* stream) {
  using namespace cova;
  COVA_REQUIRE(B > 0 && Hf > 0 && Wf > 0 && PH > 0 && PW > 0 && T >= 0, "cova_roi_fwd: bad dims");
  if (T == 0) return COVA_OK;   // a batch without boxes: empty (possibly null) rois / out are fine
  COVA_REQUIRE(fm && rois && out, "cova_roi_fwd: null pointer");
  COVA_REQUIRE(C % ROI_CB == 0, "cova_roi_fwd: C=%d must be a multiple of %d", C, ROI_CB);
  COVA_REQUIRE(ld_out >= (int64_t)C * PH * PW, "cova_roi_fwd: ld_out too small");
  COVA_REQUIRE(mode == 0 || mode == 1, "cova_roi_fwd: mode must be 0 (pool) or 1 (align)");
  cudaStream_t st = (cudaStream_t)stream;
  if (mode == 1) {
    COVA_REQUIRE(argmax == nullptr, "cova_roi_fwd: RoIAlign has no argmax");
    const int64_t n_items = (int64_t)T * PH * PW * (C / ROI_CB);
    const unsigned grid = (unsigned)((n_items + ROI_THREADS / 16 - 1) / (ROI_THREADS / 16));
    if (sampling_ratio == 2)
      roi_align_kernel<2><<<grid, ROI_THREADS, 0, st>>>(fm, B, Hf, Wf, C, rois, T, PH, PW, spatial_scale, 2, out, ld_out);
    else
      roi_align_kernel<0><<<grid, ROI_THREADS, 0, st>>>(fm, B, Hf, Wf, C, rois, T, PH, PW, spatial_scale, sampling_ratio,
                                                       out, ld_out);
  } else {
    const bool rowsplit = knob(COVA_KNOB_ROI_ROWSPLIT, 1) != 0 && (int64_t)T * PH < (1LL << 31);
    const size_t one = rowsplit ? (size_t)ROI_SPLIT_WARPS * PW * ROI_CB * 4 : (size_t)ROI_WARPS * PH * PW * ROI_CB * 4;
    const size_t smem = argmax ? 2 * one : one;
    COVA_REQUIRE(smem <= (size_t)max_smem_optin(), "cova_roi_fwd: P=%dx%d needs %zu B of shared memory", PH, PW, smem);
    dim3 grid(rowsplit ? T * PH : T, C / ROI_CB);
    const int threads = rowsplit ? ROI_SPLIT_WARPS * 32 : ROI_THREADS;
#define ROI_GO(AM, RS)                                                                                                   \
  do {                                                                                                                   \
    COVA_CUDA_OK(cudaFuncSetAttribute(roi_pool_kernel<AM, RS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    roi_pool_kernel<AM, RS><<<grid, threads, smem, st>>>(fm, Hf, Wf, C, rois, PH, PW, spatial_scale, out, ld_out, argmax, B); \
  } while (0)
    if (argmax) { if (rowsplit) ROI_GO(true, true); else ROI_GO(true, false); }
    else        { if (rowsplit) ROI_GO(false, true); else ROI_GO(false, false); }
#undef ROI_GO
  }
  COVA_LAUNCH_OK();
  return COVA_OK;
}

extern "C" int cova_roi_pool_bwd(const float* grad_out, int64_t ld_go, const int32_t* argmax, const float* rois, int T,
                                 int C, int PH, int PW, int B, int Hf, int Wf, float* grad_fm, void* stream) {
  COVA_REQUIRE(T >= 0 && C > 0 && PH > 0 && PW > 0 && B > 0 && Hf > 0 && Wf > 0, "cova_roi_pool_bwd: bad dims");
  if (T == 0) return COVA_OK;
  COVA_REQUIRE(grad_out && argmax && rois && grad_fm && ld_go >= (int64_t)C * PH * PW, "cova_roi_pool_bwd: bad arguments");
  const int64_t n = (int64_t)T * C * PH * PW;
  const int blocks = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
  cova::roi_pool_bwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(grad_out, ld_go, argmax, rois, T, C, PH * PW, B, Hf,
                                                                     Wf, grad_fm);
  COVA_LAUNCH_OK();
  return COVA_OK;
}

extern "C" int cova_roi_align_bwd(const float* grad_out, int64_t ld_go, const float* rois, int T, int C, int PH, int PW,
                                  float spatial_scale, int sampling_ratio, int B, int Hf, int Wf, float* grad_fm,
                                  void* stream) {
  using namespace cova;
  COVA_REQUIRE(T >= 0 && C > 0 && PH > 0 && PW > 0 && B > 0 && Hf > 0 && Wf > 0, "cova_roi_align_bwd: bad dims");
  if (T == 0) return COVA_OK;
  COVA_REQUIRE(grad_out && rois && grad_fm && ld_go >= (int64_t)C * PH * PW, "cova_roi_align_bwd: bad arguments");
  COVA_REQUIRE(C % ROI_CB == 0 && ((uintptr_t)grad_fm & 15) == 0, "cova_roi_align_bwd: C must be a multiple of %d, grad_fm 16-B aligned", ROI_CB);
  const int64_t n_items = (int64_t)T * PH * PW * (C / ROI_CB);
  const unsigned grid = (unsigned)((n_items + ROI_THREADS / 16 - 1) / (ROI_THREADS / 16));
  roi_align_bwd_kernel<<<grid, ROI_THREADS, 0, (cudaStream_t)stream>>>(grad_out, ld_go, rois, T, C, PH, PW, spatial_scale,
                                                                      sampling_ratio, B, Hf, Wf, grad_fm);
  COVA_LAUNCH_OK();
  return COVA_OK;
}

extern "C" int cova_bbox_enc_fwd(const float* rois, int T, const float* w, const float* b, const float* bn_scale,
                                 const float* bn_shift, int D, float* out, int64_t ld_out, void* stream) {
  COVA_REQUIRE(D > 0 && T >= 0, "cova_bbox_enc_fwd: bad dims");
  if (T == 0) return COVA_OK;
  COVA_REQUIRE(rois && w && b && out && ld_out >= D, "cova_bbox_enc_fwd: bad arguments");
  COVA_REQUIRE((bn_scale == nullptr) == (bn_shift == nullptr), "cova_bbox_enc_fwd: scale/shift must come together");
  cova::bbox_enc_kernel<<<cova::ceil_div(T * D, 256), 256, 0, (cudaStream_t)stream>>>(rois, T, w, b, bn_scale, bn_shift,
                                                                                   D, out, ld_out);
  COVA_LAUNCH_OK();
  return COVA_OK;
}

extern "C" int cova_affine_cols_fwd(const float* x, int T, int D, int64_t ld_x, const float* scale, const float* shift,
                                    float* out, int64_t ld_out, void* stream) {
  COVA_REQUIRE(T >= 0 && D >= 0, "cova_affine_cols_fwd: bad dims");
  if (T == 0 || D == 0) return COVA_OK;
  COVA_REQUIRE(x && out && ld_x >= D && ld_out >= D, "cova_affine_cols_fwd: bad arguments");
  COVA_REQUIRE((scale == nullptr) == (shift == nullptr), "cova_affine_cols_fwd: scale/shift must come together");
  cova::affine_cols_kernel<<<cova::ceil_div(T * D, 256), 256, 0, (cudaStream_t)stream>>>(x, T, D, ld_x, scale, shift, out,
                                                                                     ld_out);
  COVA_LAUNCH_OK();
  return COVA_OK;
}
