// A9: weight gradient of conv1 (7x7 s2 p3, 3->64; `loss.backward()`, /root/reference/train.py:59, through
// `convnet[0]` of /root/reference/models.py:49-51) on the 5th-gen tensor cores.  conv1 has no dgrad (the image needs no
// gradient), so this is its whole backward:
//     dW[co][c][r][s] = sum over output pixels (b,oy,ox) of  dY[b,oy,ox,co] * X[b,c,2oy+r-3,2ox+s-3]      (zero padding)
// The contraction runs over PIXELS (K = 16 output pixels per MMA), so both operands are "MN-major" for tcgen05:
//   A = dY   [K = pixel][M = 128 = {hi plane, lo plane} x 64 cout]: NHWC split planes, one TMA box {64 c, 128 px} per plane
//            with the 128-byte swizzle is already the canonical MN-major SWIZZLE_128B image (rows = pixels, 8-row groups
//            1024 B apart, the second 64-element atom = the lo plane, LBO = plane distance) - as in wgrad_tc.cu.
//   B = X    [K = pixel][N = 32 = (s, c) of ONE filter row r]: never materialised.  Converter warps stage raw image rows
//            as 4-channel 16-bit pixels (8 B: c0,c1,c2,0), exactly like the forward stem (stem_tc.cu); the 8 input pixels
//            2ox-3 .. 2ox+4 under output pixel ox are 64 contiguous bytes, 16 B further along per output pixel.  That is
//            the MN-major SWIZZLE_NONE canonical layout ((1,n),(8,k)):((X,SBO),(1,LBO)) in 16-byte units with
//            K rows 16 B apart (the core matrix), N-chunk stride SBO = 16 B (chunk j+1 of pixel k aliases chunk j of
//            pixel k+1 - a sliding window) and 8-pixel-group stride LBO = 128 B.
// One descriptor has ONE N-chunk stride, so the seven filter rows (seven different ring rows) are seven N = 32 MMAs into
// seven 32-column blocks of one accumulator [128 lanes][224 columns]; x = hi + lo doubles that (all four hi/lo products are
// accumulated: D[0:64] = dYhi * (Xhi + Xlo), D[64:128] = dYlo * (Xhi + Xlo), summed when drained).  Columns s = 7 and
// c = 3 of a block are padding (finite garbage / zeros), dropped by the finalize kernel.
// Cost: 14 MMAs x ~40 clk (an N = 32 MMA is bound by re-reading the 4 KB A operand) per 16 pixels.
//
// CTA = (page, band of conv rows), strip by strip (128 conv columns); roles: warp 0 MMA issue, warp 1 TMA producer of the
// dY tiles (3-stage ring), warps 2-5 drain (TMEM lane group = warp % 4), 8 converter warps (16-slot ring of image rows,
// the forward stem's barrier scheme).  The accumulator is double buffered in TMEM (2 x 256 columns): every `drain_every`
// row tiles the issuer switches buffers and the drain warps add the finished one to the global fp32 sum [64][224] with
// 128-bit reductions (short accumulation chains keep the tensor core's fp32 accumulator rounding out of the result).
//
// Algorithmic work: 2*64*147 FLOP per conv pixel (= the forward's 7.707 GFLOP per 1280x1280 page).
#include "common.cuh"
#include "ptx.cuh"
#include "tma_host.cuh"

namespace cova {

constexpr int SW_TM = 128;                 // conv pixels per row tile
constexpr int SW_NPX = 264;                // staged input pixels per ring row (2*127 + 8 = 262, padded)
constexpr int SW_ROW_BYTES = SW_NPX * 8;   // 2112
constexpr int SW_R = 16;                   // ring slots
constexpr int SW_ND = 16;                  // row-tile completion barriers
constexpr int SW_NS = 3;                   // dY stages
constexpr int SW_DY_PLANE = SW_TM * 128;   // 16,384 B: one plane of a dY tile
constexpr int SW_DY_STAGE = 2 * SW_DY_PLANE;
constexpr int SW_NCV = 8;                  // converter warps
constexpr int SW_THREADS = 32 * (3 + 4 + SW_NCV);   // + a second MMA issuer warp (the last one)
constexpr int SW_ISSUER_B = 2 + 4 + SW_NCV;
constexpr int SW_ACC_COLS = 256;           // 7 x 32 used
constexpr int SW_TMEM_COLS = 512;
constexpr int SW_WS = 64 * 224;            // floats of the global partial sum

struct StemWgradSmem {
  alignas(1024) unsigned char dy[SW_NS][SW_DY_STAGE];
  alignas(128) unsigned char ring[2][SW_R][SW_ROW_BYTES];
  uint64_t in_full[SW_R], pair_full[SW_ND], mma_done[SW_ND], dy_full[SW_NS], dy_empty[SW_NS], acc_full[2], acc_empty[2];
  uint32_t tmem_base;
};

struct StemWgradParams {
  const void* img;
  int B, H, W, Hc, Wc;
  int bands_per_page, rows_per_band, drain_every;
  int dy_single;                           // bf16 training mode: dY is one bf16 plane - the lo slots are zero-filled once
  float* ws;                               // [64 co][224 = r*32 + s*4 + c] fp32, zeroed
};

// MN-major SWIZZLE_NONE: 8-pixel-group stride LBO, N-chunk (8 elements) stride SBO, pixels 16 B apart
__device__ __forceinline__ uint64_t sw_desc_noswz(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// MN-major SWIZZLE_128B (see wgrad_tc.cu): rows = pixels (128 B), 8-row groups `sbo` apart, second 64-element atom `lbo` on
__device__ __forceinline__ uint64_t sw_desc_mn_sw128(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t sw_idesc_mn(int M, int N, bool half) {
  return (1u << 4) | (half ? 0u : ((1u << 7) | (1u << 10))) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

// The MMAs of filter rows R0 .. R1-1 for one tile of 128 conv pixels (8 K-steps of 16 pixels).  `first0` = 0: the tile opens
// its accumulator buffer (the first K-step overwrites).
template <int R0, int R1, bool HALF>
__device__ __forceinline__ void sw_issue_rows(uint32_t d_tmem, uint32_t dy_addr, uint32_t ring0, uint32_t g0, uint32_t first0,
                                              bool dy_single) {
  constexpr uint32_t idesc = sw_idesc_mn(128, 32, HALF);
#pragma unroll 1
  for (int kk = 0; kk < 8; ++kk) {                         // K = 16 conv pixels
    const uint64_t da = sw_desc_mn_sw128(dy_addr + kk * 2048, SW_DY_PLANE, 1024);      // [dYhi ; dYlo]
    const uint32_t first = (first0 == 0u && kk == 0) ? 0u : 1u;
#pragma unroll
    for (int r = R0; r < R1; ++r) {
      const uint32_t row = ring0 + ((g0 + r) % SW_R) * SW_ROW_BYTES + kk * 256;        // 16 pixels x 16 B
      ptx::umma_bf16(d_tmem + r * 32, da, sw_desc_noswz(row, 128, 16), idesc, first);
      // bf16 training mode (one dY plane): the forward convolved bf16(image), so the gradient of what was computed
      // takes the hi plane of the image alone - half the MMAs
      if (!dy_single) ptx::umma_bf16(d_tmem + r * 32, da, sw_desc_noswz(row + SW_R * SW_ROW_BYTES, 128, 16), idesc, 1u);
    }
  }
}

template <bool U8, bool HALF>
__global__ void __launch_bounds__(SW_THREADS, 1)
stem_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tm_dy_hi, const __grid_constant__ CUtensorMap tm_dy_lo,
                     const StemWgradParams p) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  StemWgradSmem& sm = *reinterpret_cast<StemWgradSmem*>(smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int b = blockIdx.x / p.bands_per_page, band = blockIdx.x % p.bands_per_page;
  const int oy_begin = band * p.rows_per_band;
  const int oy_end = min(oy_begin + p.rows_per_band, p.Hc);
  const int n_conv = oy_end - oy_begin;
  const int NQ = 2 * (n_conv - 1) + 7;               // input rows per strip
  const int y_first = 2 * oy_begin - 3;
  const int n_strips = (p.Wc + SW_TM - 1) / SW_TM;
  const int n_tiles = n_strips * n_conv;

  if (threadIdx.x == 0) {
    for (int i = 0; i < SW_R; ++i) ptx::mbar_init(&sm.in_full[i], 1);
    for (int i = 0; i < SW_ND; ++i) {
      ptx::mbar_init(&sm.mma_done[i], 2);          // one commit per issuer warp
      ptx::mbar_init(&sm.pair_full[i], 2);
    }
    for (int i = 0; i < SW_NS; ++i) {
      ptx::mbar_init(&sm.dy_full[i], 1);
      ptx::mbar_init(&sm.dy_empty[i], 2);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&sm.acc_full[i], 2);
      ptx::mbar_init(&sm.acc_empty[i], 128);
    }
    ptx::fence_barrier_init();
    ptx::prefetch_tensormap(&tm_dy_hi);
    ptx::prefetch_tensormap(&tm_dy_lo);
  }
  if (warp == 0) {
    ptx::tmem_alloc(&sm.tmem_base, SW_TMEM_COLS);
    ptx::tmem_relinquish();
  }
  if (p.dy_single) {
    for (int st = 0; st < SW_NS; ++st) {
      uint4* lo = reinterpret_cast<uint4*>(sm.dy[st] + SW_DY_PLANE);
      for (int i = threadIdx.x; i < SW_DY_PLANE / 16; i += SW_THREADS) lo[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    ptx::fence_proxy_async();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = sm.tmem_base;
  if (n_conv <= 0) goto teardown;

  if (warp == 0 || warp == SW_ISSUER_B) {
    // ======================= MMA issuers =======================
    // Two warps share every tile: filter rows 0..3 / 4..6 accumulate in disjoint TMEM columns (r * 32), so the two issue streams
    // never touch the same accumulator.  Fourteen N = 32 MMAs per 16 pixels from one thread ran at ~53 clk each against the
    // pipe's ~40-clk floor for this shape (stem_tc.cu: the issuing thread's latency).
    const uint32_t ring0 = ptx::smem_u32(&sm.ring[0][0][0]);
    const uint32_t dy0 = ptx::smem_u32(&sm.dy[0][0]);
    uint32_t t = 0, buf = 0, n_in_buf = 0, uses[2] = {0u, 0u};
    for (int strip = 0; strip < n_strips; ++strip) {
      for (int i = 0; i < n_conv; ++i, ++t) {
        const uint32_t g0 = (uint32_t)strip * NQ + 2 * i;
        if (i == 0) {
          for (int r = 0; r < 5; ++r) {
            const uint32_t g = g0 + r;
            ptx::mbar_wait(&sm.in_full[g % SW_R], (g / SW_R) & 1);
          }
        }
        ptx::mbar_wait(&sm.pair_full[t % SW_ND], (t / SW_ND) & 1);
        const uint32_t stage = t % SW_NS;
        ptx::mbar_wait(&sm.dy_full[stage], (t / SW_NS) & 1);
        if (n_in_buf == 0 && uses[buf] > 0) ptx::mbar_wait(&sm.acc_empty[buf], (uses[buf] - 1) & 1);
        ptx::tc_fence_after();
        const bool last = (int)t + 1 == n_tiles;
        if (ptx::elect_one()) {
          const uint32_t d_tmem = tmem_base + buf * SW_ACC_COLS;
          const uint32_t dy_addr = dy0 + stage * SW_DY_STAGE;
          const uint32_t first0 = n_in_buf == 0 ? 0u : 1u;
          if (warp == 0) sw_issue_rows<0, 4, HALF>(d_tmem, dy_addr, ring0, g0, first0, p.dy_single != 0);
          else           sw_issue_rows<4, 7, HALF>(d_tmem, dy_addr, ring0, g0, first0, p.dy_single != 0);
          ptx::umma_commit(&sm.mma_done[t % SW_ND]);
          ptx::umma_commit(&sm.dy_empty[stage]);
          if (n_in_buf + 1 == (uint32_t)p.drain_every || last) ptx::umma_commit(&sm.acc_full[buf]);
        }
        __syncwarp();
        if (++n_in_buf == (uint32_t)p.drain_every || last) {
          ++uses[buf];
          buf ^= 1u;
          n_in_buf = 0;
        }
      }
    }
  } else if (warp == 1) {
    // ======================= TMA producer: dY row tiles {64 c, 128 px} x {hi, lo} =======================
    uint32_t t = 0;
    for (int strip = 0; strip < n_strips; ++strip) {
      for (int i = 0; i < n_conv; ++i, ++t) {
        const uint32_t stage = t % SW_NS;
        ptx::mbar_wait(&sm.dy_empty[stage], ((t / SW_NS) & 1) ^ 1);
        if (ptx::elect_one()) {
          ptx::mbar_arrive_expect_tx(&sm.dy_full[stage], p.dy_single ? SW_DY_PLANE : SW_DY_STAGE);
          ptx::tma_load_4d(sm.dy[stage], &tm_dy_hi, &sm.dy_full[stage], 0, strip * SW_TM, oy_begin + i, b);
          if (!p.dy_single) ptx::tma_load_4d(sm.dy[stage] + SW_DY_PLANE, &tm_dy_lo, &sm.dy_full[stage], 0, strip * SW_TM, oy_begin + i, b);
        }
        __syncwarp();
      }
    }
  } else if (warp >= 2 && warp < 6) {
    // ======================= drain: TMEM -> RED.ADD.F32x4 into ws[co][224] =======================
    const int lg = warp & 3;
    const int co = (lg * 32 + lane) & 63;          // lanes 0-63: dYhi rows, 64-127: dYlo rows of the same cout
    const int n_drains = (n_tiles + p.drain_every - 1) / p.drain_every;
    float* dst = p.ws + (size_t)co * 224;
    for (int d = 0; d < n_drains; ++d) {
      const uint32_t bf = d & 1;
      ptx::mbar_wait(&sm.acc_full[bf], (d >> 1) & 1);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + bf * SW_ACC_COLS;
#pragma unroll 1
      for (int q = 0; q < 14; ++q) {
        uint32_t v[16];
        ptx::tmem_ld16(taddr + q * 16, v);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float4 a;
          a.x = __uint_as_float(v[4 * j]); a.y = __uint_as_float(v[4 * j + 1]);
          a.z = __uint_as_float(v[4 * j + 2]); a.w = __uint_as_float(v[4 * j + 3]);
          atomicAdd(reinterpret_cast<float4*>(dst + q * 16 + 4 * j), a);
        }
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(&sm.acc_empty[bf]);
    }
  } else if (warp < SW_ISSUER_B) {
    // ======================= converters: NCHW image rows -> ring of 4-channel 16-bit pixels (hi, lo) =======================
    const int cw = warp - 6;
    const size_t plane = (size_t)p.H * p.W;
    const size_t img_b = (size_t)b * 3 * plane;
    const uint32_t n_rows_total = (uint32_t)n_strips * NQ;
    for (uint32_t g = cw; g < n_rows_total; g += SW_NCV) {
      const int strip = g / NQ, q = g % NQ;
      const int y = y_first + q;
      const int x0 = 2 * strip * SW_TM - 3;
      if (g >= SW_R) {   // the previous occupant of this slot must have been consumed by its last conv row
        const uint32_t gp = g - SW_R;
        const int sp = gp / NQ, qp = gp % NQ;
        const uint32_t t_last = (uint32_t)sp * n_conv + min(qp >> 1, n_conv - 1);
        ptx::mbar_wait(&sm.mma_done[t_last % SW_ND], (t_last / SW_ND) & 1);
      }
      const bool row_ok = y >= 0 && y < p.H;
      unsigned char* dst_hi = sm.ring[0][g % SW_R];
      unsigned char* dst_lo = sm.ring[1][g % SW_R];
      if (!U8 && (p.W & 3) == 0 && (reinterpret_cast<uintptr_t>(p.img) & 15) == 0) {
        // x0 - 1 is a multiple of 4: the row segment is 67 aligned float4 per channel
        const float* imgf = reinterpret_cast<const float*>(p.img) + img_b;
        float4 fv[3][3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const int wi = lane + 32 * j, xw = x0 - 1 + 4 * wi;
          const bool ok = row_ok && wi < 67 && xw >= 0 && xw < p.W;
#pragma unroll
          for (int c = 0; c < 3; ++c)
            fv[j][c] = ok ? __ldg(reinterpret_cast<const float4*>(imgf + c * plane + (size_t)y * p.W + xw))
                          : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const int wi = lane + 32 * j;
          if (wi < 67) {
            const float* c0 = &fv[j][0].x;
            const float* c1 = &fv[j][1].x;
            const float* c2 = &fv[j][2].x;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int i = 4 * wi + k - 1;
              if (i >= 0 && i < SW_NPX) {
                uint32_t h01, l01, h2, l2;
                if (HALF) {
                  split_f16x2(c0[k], c1[k], h01, l01);
                  split_f16x2(c2[k], 0.f, h2, l2);
                } else {
                  split_bf16x2(c0[k], c1[k], h01, l01);
                  split_bf16x2(c2[k], 0.f, h2, l2);
                }
                *reinterpret_cast<uint2*>(dst_hi + i * 8) = make_uint2(h01, h2);
                *reinterpret_cast<uint2*>(dst_lo + i * 8) = make_uint2(l01, l2);
              }
            }
          }
        }
      } else {
        float f[9][3];
#pragma unroll
        for (int j = 0; j < 9; ++j) {
          const int i = lane + 32 * j, x = x0 + i;
          const bool ok = row_ok && i < SW_NPX && x >= 0 && x < p.W;
#pragma unroll
          for (int c = 0; c < 3; ++c) f[j][c] = ok ? load_pixel<U8>(p.img, img_b + c * plane + (size_t)y * p.W + x) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < 9; ++j) {
          const int i = lane + 32 * j;
          if (i < SW_NPX) {
            uint32_t h01, l01, h2, l2;
            if (HALF) {
              split_f16x2(f[j][0], f[j][1], h01, l01);
              split_f16x2(f[j][2], 0.f, h2, l2);
            } else {
              split_bf16x2(f[j][0], f[j][1], h01, l01);
              split_bf16x2(f[j][2], 0.f, h2, l2);
            }
            *reinterpret_cast<uint2*>(dst_hi + i * 8) = make_uint2(h01, h2);
            *reinterpret_cast<uint2*>(dst_lo + i * 8) = make_uint2(l01, l2);
          }
        }
      }
      ptx::fence_proxy_async();      // generic-proxy writes -> visible to tcgen05 (async proxy) reads
      __syncwarp();
      if (lane == 0) {
        ptx::mbar_arrive(&sm.in_full[g % SW_R]);
        if (q >= 5) {
          const uint32_t tq = (uint32_t)strip * n_conv + ((q - 5) >> 1);
          ptx::mbar_arrive(&sm.pair_full[tq % SW_ND]);
        }
      }
    }
  }

teardown:
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, SW_TMEM_COLS);
  }
}

// ws [64 co][r*32 + s*4 + c] -> OIHW [64][3][7][7], scaled by *inv_scale (device scalar; NULL = 1)
__global__ void stem_wgrad_finalize_kernel(const float* __restrict__ ws, const float* __restrict__ inv_scale,
                                           float* __restrict__ dw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 64 * 147) return;
  const int s = i % 7, r = (i / 7) % 7, c = (i / 49) % 3, co = i / 147;
  dw[i] = ws[co * 224 + r * 32 + s * 4 + c] * (inv_scale ? __ldg(inv_scale) : 1.f);
}

}  // namespace cova

extern "C" int cova_stem_wgrad(const void* images, int img_dtype, int B, int H, int W, const void* dy_hi, const void* dy_lo,
                               int planes_dtype, const float* inv_scale, float* ws, float* dw_oihw, void* stream) {
  using namespace cova;
  COVA_REQUIRE(images && dy_hi && ws && dw_oihw, "cova_stem_wgrad: null pointer");
  COVA_REQUIRE(planes_dtype == COVA_BF16 || dy_lo, "cova_stem_wgrad: split planes need their lo plane");
  COVA_REQUIRE(B > 0 && H >= 7 && W >= 7, "cova_stem_wgrad: bad dims");
  COVA_REQUIRE(img_dtype == COVA_F32 || img_dtype == COVA_U8, "cova_stem_wgrad: images must be fp32 or uint8");
  COVA_REQUIRE(planes_dtype == COVA_F16X2 || planes_dtype == COVA_BF16X2 || planes_dtype == COVA_BF16,
               "cova_stem_wgrad: planes are split-fp16, split-bf16 or one bf16 plane (dy_lo NULL)");
  COVA_REQUIRE((((uintptr_t)dy_hi | (uintptr_t)dy_lo | (uintptr_t)ws) & 15) == 0, "cova_stem_wgrad: 16-byte alignment");
  cudaStream_t st = (cudaStream_t)stream;
  StemWgradParams p;
  p.img = images; p.B = B; p.H = H; p.W = W;
  p.Hc = (H + 6 - 7) / 2 + 1; p.Wc = (W + 6 - 7) / 2 + 1;
  // bands per page: fill whole waves of SMs, keep bands tall enough that the 5-row halo / pipeline fill stays small
  const int sms = sm_count();
  int best = 1;
  double best_score = -1.0;
  for (int bands = 1; bands <= p.Hc && bands <= 4 * sms; ++bands) {
    const int rows = ceil_div(p.Hc, bands), ctas = B * ceil_div(p.Hc, rows);
    const double eff = (double)ctas / ((double)ceil_div(ctas, sms) * sms);
    const double score = eff * rows / (rows + 3.0);
    if (score > best_score + 1e-9) { best_score = score; best = bands; }
  }
  p.rows_per_band = ceil_div(p.Hc, best);
  p.bands_per_page = ceil_div(p.Hc, p.rows_per_band);
  p.drain_every = knob(COVA_KNOB_WGRAD_DRAIN, 16);
  if (p.drain_every < 1) p.drain_every = 1;
  p.ws = ws;
  p.dy_single = dy_lo == nullptr;
  CUtensorMap td_hi, td_lo;
  const uint64_t xd[4] = {64, (uint64_t)p.Wc, (uint64_t)p.Hc, (uint64_t)B};
  const uint64_t xs[3] = {128, (uint64_t)p.Wc * 128, (uint64_t)p.Hc * p.Wc * 128};
  const uint32_t db[4] = {64, SW_TM, 1, 1};
  int rc;
  if ((rc = make_tmap_bf16(&td_hi, dy_hi, 4, xd, xs, db))) return rc;
  if ((rc = make_tmap_bf16(&td_lo, dy_lo ? dy_lo : dy_hi, 4, xd, xs, db))) return rc;
  COVA_CUDA_OK(cudaMemsetAsync(ws, 0, (size_t)SW_WS * sizeof(float), st));
  const int smem = (int)sizeof(StemWgradSmem) + 1024;
  const int grid = B * p.bands_per_page;
  const bool u8 = img_dtype == COVA_U8, half = planes_dtype == COVA_F16X2;
#define COVA_SW_GO(U, HF)                                                                              \
  do {                                                                                                 \
    auto kern = stem_wgrad_tc_kernel<U, HF>;                                                           \
    COVA_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));       \
    kern<<<grid, SW_THREADS, smem, st>>>(td_hi, td_lo, p);                                             \
  } while (0)
  if (u8 && half) COVA_SW_GO(true, true);
  else if (u8) COVA_SW_GO(true, false);
  else if (half) COVA_SW_GO(false, true);
  else COVA_SW_GO(false, false);
#undef COVA_SW_GO
  COVA_LAUNCH_OK();
  stem_wgrad_finalize_kernel<<<ceil_div(64 * 147, 256), 256, 0, st>>>(ws, inv_scale, dw_oihw);
  COVA_LAUNCH_OK();
  return COVA_OK;
}
