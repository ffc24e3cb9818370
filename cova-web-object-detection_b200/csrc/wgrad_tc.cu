// A9: weight gradient of the 3x3 s1 p1 64->64 convolutions (`loss.backward()`, /root/reference/train.py:59, through
// torchvision BasicBlock / Bottleneck conv2) on the 5th-gen tensor cores - the contraction runs over PIXELS:
//     dW[tap][ci][co] = sum over pixels p of  X[p + tap][ci] * dY[p][co]
// NHWC activations make both operands "MN-major" for tcgen05 (the contraction index - the pixel - is the slow axis, the
// 64 channels of a pixel are one contiguous 128-byte row), which is exactly what a TMA box {64 ch, pixels} with the
// 128-byte swizzle lands in shared memory: operand rows = K (pixel) index, 8-row groups 1024 B apart.  As in the
// forward kernel (conv_tc.cu) the nine taps are nine VIEWS of one staged halo patch {64 c, 10 w, 18 h}: a K = 16 step
// is two tile rows of 8 pixels, i.e. two 8-row groups `10 patch rows` (1280 B) apart, starting at patch row
// (2 kk + r) * 10 + s.
//
// UMMA shape: A = X (M side), B = dY (N side).  Cin = 64 would leave half of the 128 accumulator lanes idle, so ONE MMA
// covers TWO taps: the MN-major descriptor addresses its second 64-element atom `LBO` bytes after the first, and a
// second tap of the same patch is just another byte offset ((r'*10+s') - (r*10+s)) * 128.  Five tap pairs
// (0,1) (2,3) (4,5) (6,7) (8,-); the partner of tap 8 reads one patch row further (finite garbage, rows ignored).
//
// fp32 parity from 16-bit operands (split planes x = hi + lo, the same three products as the forward):
//   pairs 0-2:  Xhi x [dYhi ; dYlo]  as one N = 128 MMA (the two dY planes are the two N atoms, LBO = plane distance)
//               Xlo x dYhi           as an N = 64 MMA into the first 64 columns            -> 128 TMEM columns per pair
//   pairs 3-4:  Xhi x dYhi, Xhi x dYlo, Xlo x dYhi as three N = 64 MMAs into the same 64 columns -> 64 columns per pair
// = 3 x 128 + 2 x 64 = 512 TMEM columns, the whole tensor memory of the SM: all nine taps of a pixel tile are
// accumulated from ONE pass over the operands (a taps-over-several-passes scheme would re-stream 78 KB per tile and pass).
// Cost per 128-pixel tile: 8 K-steps x (3 x (64 + 48) + 2 x 3 x 48) = 4992 tensor clocks (forward: 4032).
//
// Persistent CTAs (grid = #SMs), warp-specialised: warp 0 TMA producer (2-stage ring of {Xhi, Xlo patch, dYhi, dYlo tile}
// = 77 KB per stage), warp 1 single-thread MMA issue, warps 2-5 drain: every `drain_every` tiles (and at the end) the
// accumulators are read back with tcgen05.ld and added to the global fp32 gradient [9][64 ci][64 co] with 128-bit
// reductions (RED.ADD.F32x4) - the short accumulation chains keep the tensor core's fp32 accumulator rounding out of
// the result.  A second tiny kernel scales by 1/s (the power-of-two scale of the dY planes) and transposes to OIHW.
//
// Algorithmic work: 2*9*64*64 = 73,728 FLOP per pixel (the same as the forward conv); operands 4 planes x 128 B per pixel.
#include "common.cuh"
#include "ptx.cuh"
#include "tma_host.cuh"

namespace cova {

constexpr int WG_C = 64;
constexpr int WG_TW = 8, WG_TH = 16;
constexpr int WG_HW = WG_TW + 2, WG_HH = WG_TH + 2;
constexpr int WG_PATCH_BYTES = WG_HW * WG_HH * 128;      // 23,040
constexpr int WG_PATCH_SLOT = 23 * 1024;                 // 1024-B aligned sub-buffer
constexpr int WG_DY_BYTES = WG_TW * WG_TH * 128;         // 16,384
constexpr int WG_STAGE_BYTES = 2 * WG_PATCH_SLOT + 2 * WG_DY_BYTES;   // 79,872 = 78 KB
constexpr int WG_STAGE_TX = 2 * WG_PATCH_BYTES + 2 * WG_DY_BYTES;     // bytes the four TMA loads deliver
constexpr int WG_NSTAGE = 2;
constexpr int WG_THREADS = 192;                          // warp 0 TMA, warp 1 MMA, warps 2-5 drain (TMEM lane group = warp % 4)
constexpr int WG_SMEM_BYTES = WG_NSTAGE * WG_STAGE_BYTES + 1024 /*tail*/ + 1024 /*align slack*/;
constexpr int WG_TMEM_COLS = 512;

struct WgradTail {
  uint64_t full[WG_NSTAGE], empty[WG_NSTAGE], acc_full, acc_empty;
  uint32_t tmem_base;
};

struct WgradParams {
  int B, H, W, tiles_w, tiles_h, n_tiles, drain_every;
  int x_single, dy_single;   // bf16 training mode: an operand is ONE plane - its lo slots are zero-filled once and never loaded
  float* ws;   // [9][64 ci][64 co] fp32, zeroed
};

// MN-major operand, 128-byte swizzle: rows = K index (128 B each = 64 MN elements), 8-row groups `sbo` bytes apart,
// the second 64-element MN atom `lbo` bytes after the first (cute::UMMA canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO))).
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16 instruction descriptor with both operands MN-major (a_major = bit 15, b_major = bit 16)
__host__ __device__ constexpr uint32_t umma_idesc_mn(int M, int N, bool half) {
  return (1u << 4) | (half ? 0u : ((1u << 7) | (1u << 10))) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

__host__ __device__ constexpr int wg_tap_off(int tap) { return ((tap / 3) * WG_HW + (tap % 3)) * 128; }
// accumulator column of tap pair p (pairs 0-2: 128 columns, pairs 3-4: 64 columns)
__host__ __device__ constexpr int wg_pair_col(int p) { return p < 3 ? p * 128 : 384 + (p - 3) * 64; }

template <bool HALF>
__global__ void __launch_bounds__(WG_THREADS, 1)
conv3x3_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tm_x_hi, const __grid_constant__ CUtensorMap tm_x_lo,
                        const __grid_constant__ CUtensorMap tm_dy_hi, const __grid_constant__ CUtensorMap tm_dy_lo,
                        const WgradParams p) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  unsigned char* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  WgradTail& tail = *reinterpret_cast<WgradTail*>(smem + WG_NSTAGE * WG_STAGE_BYTES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int my_tiles = p.n_tiles > (int)blockIdx.x ? (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < WG_NSTAGE; ++i) {
      ptx::mbar_init(&tail.full[i], 1);
      ptx::mbar_init(&tail.empty[i], 1);
    }
    ptx::mbar_init(&tail.acc_full, 1);
    ptx::mbar_init(&tail.acc_empty, 128);
    ptx::fence_barrier_init();
    ptx::prefetch_tensormap(&tm_x_hi);
    ptx::prefetch_tensormap(&tm_x_lo);
    ptx::prefetch_tensormap(&tm_dy_hi);
    ptx::prefetch_tensormap(&tm_dy_lo);
  }
  if (warp == 1) {
    ptx::tmem_alloc(&tail.tmem_base, WG_TMEM_COLS);
    ptx::tmem_relinquish();
  }
  if (p.x_single || p.dy_single) {
    for (int st = 0; st < WG_NSTAGE; ++st) {
      uint4* xl = reinterpret_cast<uint4*>(smem + st * WG_STAGE_BYTES + WG_PATCH_SLOT);
      uint4* dl = reinterpret_cast<uint4*>(smem + st * WG_STAGE_BYTES + 2 * WG_PATCH_SLOT + WG_DY_BYTES);
      if (p.x_single) for (int i = threadIdx.x; i < WG_PATCH_SLOT / 16; i += WG_THREADS) xl[i] = make_uint4(0u, 0u, 0u, 0u);
      if (p.dy_single) for (int i = threadIdx.x; i < WG_DY_BYTES / 16; i += WG_THREADS) dl[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    ptx::fence_proxy_async();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tail.tmem_base;

  if (warp == 0) {
    // ======================= TMA producer =======================
    uint32_t stage = 0, phase = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      const int b = tile / (p.tiles_h * p.tiles_w);
      const int th = (tile / p.tiles_w) % p.tiles_h, tw = tile % p.tiles_w;
      const int h0 = th * WG_TH, w0 = tw * WG_TW;
      ptx::mbar_wait(&tail.empty[stage], phase ^ 1);
      if (ptx::elect_one()) {
        unsigned char* s = smem + stage * WG_STAGE_BYTES;
        ptx::mbar_arrive_expect_tx(&tail.full[stage], (p.x_single ? 1 : 2) * WG_PATCH_BYTES + (p.dy_single ? 1 : 2) * WG_DY_BYTES);
        ptx::tma_load_4d(s, &tm_x_hi, &tail.full[stage], 0, w0 - 1, h0 - 1, b);
        if (!p.x_single) ptx::tma_load_4d(s + WG_PATCH_SLOT, &tm_x_lo, &tail.full[stage], 0, w0 - 1, h0 - 1, b);
        ptx::tma_load_4d(s + 2 * WG_PATCH_SLOT, &tm_dy_hi, &tail.full[stage], 0, w0, h0, b);
        if (!p.dy_single) ptx::tma_load_4d(s + 2 * WG_PATCH_SLOT + WG_DY_BYTES, &tm_dy_lo, &tail.full[stage], 0, w0, h0, b);
      }
      __syncwarp();
      if (++stage == WG_NSTAGE) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer (one elected lane) =======================
    constexpr uint32_t idesc128 = umma_idesc_mn(128, 128, HALF), idesc64 = umma_idesc_mn(128, 64, HALF);
    uint32_t stage = 0, phase = 0;
    int since_drain = 0, n_drain = 0;
    for (int i = 0; i < my_tiles; ++i) {
      if (since_drain == 0 && n_drain > 0) {          // the drain warps have emptied the accumulators
        ptx::mbar_wait(&tail.acc_empty, (n_drain - 1) & 1);
        ptx::tc_fence_after();
      }
      ptx::mbar_wait(&tail.full[stage], phase);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        const uint32_t s_addr = ptx::smem_u32(smem + stage * WG_STAGE_BYTES);
        const uint32_t x_hi = s_addr, x_lo = s_addr + WG_PATCH_SLOT, dy_hi = s_addr + 2 * WG_PATCH_SLOT;
#pragma unroll 1
        for (int kk = 0; kk < 8; ++kk) {              // K = 16 pixels: tile rows 2 kk and 2 kk + 1
          const uint32_t acc0 = (since_drain > 0 || kk > 0) ? 1u : 0u;
          const uint32_t xrow = kk * 2 * WG_HW * 128;
          const uint64_t db_hl = umma_desc_mn_sw128(dy_hi + kk * 2048, WG_DY_BYTES, 1024);               // [dYhi ; dYlo]
          const uint64_t db_lo = umma_desc_mn_sw128(dy_hi + WG_DY_BYTES + kk * 2048, WG_DY_BYTES, 1024); // dYlo alone
#pragma unroll
          for (int pr = 0; pr < 5; ++pr) {
            const int ta = 2 * pr, tb = pr < 4 ? 2 * pr + 1 : 9;   // "tap 9" = one patch row past tap 8 (ignored rows)
            const uint32_t lbo = (tb < 9 ? wg_tap_off(tb) : wg_tap_off(8) + 128) - wg_tap_off(ta);
            const uint64_t da_hi = umma_desc_mn_sw128(x_hi + xrow + wg_tap_off(ta), lbo, WG_HW * 128);
            const uint64_t da_lo = umma_desc_mn_sw128(x_lo + xrow + wg_tap_off(ta), lbo, WG_HW * 128);
            const uint32_t d = tmem_base + wg_pair_col(pr);
            if (pr < 3) {
              ptx::umma_bf16(d, da_hi, db_hl, idesc128, acc0);   // Xhi x [dYhi ; dYlo]
              ptx::umma_bf16(d, da_lo, db_hl, idesc64, 1u);      // Xlo x dYhi
            } else {
              ptx::umma_bf16(d, da_hi, db_hl, idesc64, acc0);    // Xhi x dYhi
              ptx::umma_bf16(d, da_hi, db_lo, idesc64, 1u);      // Xhi x dYlo
              ptx::umma_bf16(d, da_lo, db_hl, idesc64, 1u);      // Xlo x dYhi
            }
          }
        }
        ptx::umma_commit(&tail.empty[stage]);
        if (since_drain + 1 == p.drain_every || i + 1 == my_tiles) ptx::umma_commit(&tail.acc_full);
      }
      __syncwarp();
      if (++since_drain == p.drain_every || i + 1 == my_tiles) { since_drain = 0; ++n_drain; }
      if (++stage == WG_NSTAGE) { stage = 0; phase ^= 1; }
    }
  } else {
    // ======================= drain warps: TMEM -> registers -> RED.ADD.F32x4 into ws =======================
    const int lg = warp & 3;                       // TMEM lane group of this warp
    const int row = lg * 32 + lane;                // accumulator row = (tap half, ci)
    const int half_sel = row >> 6, ci = row & 63;
    const int n_drains = (my_tiles + p.drain_every - 1) / p.drain_every;
    for (int dr = 0; dr < n_drains; ++dr) {
      ptx::mbar_wait(&tail.acc_full, dr & 1);
      ptx::tc_fence_after();
#pragma unroll 1
      for (int pr = 0; pr < 5; ++pr) {
        const int tap = 2 * pr + half_sel;
        const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + wg_pair_col(pr);
        float* dst = p.ws + ((size_t)min(tap, 8) * WG_C + ci) * WG_C;
#pragma unroll 1
        for (int q = 0; q < 4; ++q) {              // 16 output channels per step
          uint32_t v[16], u[16];
          ptx::tmem_ld16(taddr + q * 16, v);
          if (pr < 3) ptx::tmem_ld16(taddr + 64 + q * 16, u);
          ptx::tmem_ld_wait();
          if (tap < 9) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float4 a;
              a.x = __uint_as_float(v[4 * j]); a.y = __uint_as_float(v[4 * j + 1]);
              a.z = __uint_as_float(v[4 * j + 2]); a.w = __uint_as_float(v[4 * j + 3]);
              if (pr < 3) {
                a.x += __uint_as_float(u[4 * j]); a.y += __uint_as_float(u[4 * j + 1]);
                a.z += __uint_as_float(u[4 * j + 2]); a.w += __uint_as_float(u[4 * j + 3]);
              }
              atomicAdd(reinterpret_cast<float4*>(dst + q * 16 + 4 * j), a);
            }
          }
        }
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(&tail.acc_empty);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, WG_TMEM_COLS);
  }
}

// ws [9][ci][co] -> OIHW [co][ci][3][3], scaled by *inv_scale (device scalar; NULL = 1)
__global__ void wgrad_finalize_kernel(const float* __restrict__ ws, const float* __restrict__ inv_scale, float mult, int taps,
                                      int Cin, int Cout, float* __restrict__ dw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= taps * Cin * Cout) return;
  const int t = i % taps, ci = (i / taps) % Cin, co = i / (taps * Cin);
  const float s = (inv_scale ? __ldg(inv_scale) : 1.f) * mult;
  dw[i] = ws[((size_t)t * Cin + ci) * Cout + co] * s;
}

int launch_wgrad_finalize(const float* ws, const float* inv_scale, float mult, int taps, int Cin, int Cout, float* dw,
                          cudaStream_t st) {
  wgrad_finalize_kernel<<<ceil_div(taps * Cin * Cout, 256), 256, 0, st>>>(ws, inv_scale, mult, taps, Cin, Cout, dw);
  COVA_LAUNCH_OK();
  return COVA_OK;
}

}  // namespace cova

extern "C" int cova_conv3x3_wgrad(const void* x_hi, const void* x_lo, const void* dy_hi, const void* dy_lo, int B, int H, int W,
                                  int planes_dtype, const float* inv_scale, float* ws, float* dw_oihw, void* stream) {
  using namespace cova;
  COVA_REQUIRE(x_hi && dy_hi && ws && dw_oihw, "cova_conv3x3_wgrad: null pointer");
  COVA_REQUIRE(B > 0 && H > 0 && W > 0, "cova_conv3x3_wgrad: bad dims");
  COVA_REQUIRE(planes_dtype == COVA_F16X2 || planes_dtype == COVA_BF16X2 || planes_dtype == COVA_BF16,
               "cova_conv3x3_wgrad: planes are split-fp16, split-bf16 or single bf16 planes (x_lo / dy_lo NULL)");
  COVA_REQUIRE(planes_dtype == COVA_BF16 || (x_lo && dy_lo), "cova_conv3x3_wgrad: split planes need their lo plane");
  COVA_REQUIRE((((uintptr_t)x_hi | (uintptr_t)x_lo | (uintptr_t)dy_hi | (uintptr_t)dy_lo | (uintptr_t)ws) & 15) == 0,
               "cova_conv3x3_wgrad: 16-byte alignment");
  cudaStream_t st = (cudaStream_t)stream;
  CUtensorMap tx_hi, tx_lo, td_hi, td_lo;
  const uint64_t xd[4] = {(uint64_t)WG_C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
  const uint64_t xs[3] = {(uint64_t)WG_C * 2, (uint64_t)W * WG_C * 2, (uint64_t)H * W * WG_C * 2};
  const uint32_t xb[4] = {WG_C, WG_HW, WG_HH, 1};
  const uint32_t db[4] = {WG_C, WG_TW, WG_TH, 1};
  int rc;
  if ((rc = make_tmap_bf16(&tx_hi, x_hi, 4, xd, xs, xb))) return rc;
  if ((rc = make_tmap_bf16(&tx_lo, x_lo ? x_lo : x_hi, 4, xd, xs, xb))) return rc;
  if ((rc = make_tmap_bf16(&td_hi, dy_hi, 4, xd, xs, db))) return rc;
  if ((rc = make_tmap_bf16(&td_lo, dy_lo ? dy_lo : dy_hi, 4, xd, xs, db))) return rc;
  WgradParams p;
  p.B = B; p.H = H; p.W = W;
  p.tiles_w = ceil_div(W, WG_TW);
  p.tiles_h = ceil_div(H, WG_TH);
  p.n_tiles = B * p.tiles_w * p.tiles_h;
  p.drain_every = knob(COVA_KNOB_WGRAD_DRAIN, 16);
  if (p.drain_every < 1) p.drain_every = 1;
  p.ws = ws;
  p.x_single = x_lo == nullptr; p.dy_single = dy_lo == nullptr;
  COVA_CUDA_OK(cudaMemsetAsync(ws, 0, (size_t)9 * WG_C * WG_C * sizeof(float), st));
  static_assert(sizeof(WgradTail) <= 1024, "tail too large");
  const int grid = p.n_tiles < sm_count() ? p.n_tiles : sm_count();
  if (planes_dtype == COVA_F16X2) {
    auto kern = conv3x3_wgrad_tc_kernel<true>;
    COVA_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM_BYTES));
    kern<<<grid, WG_THREADS, WG_SMEM_BYTES, st>>>(tx_hi, tx_lo, td_hi, td_lo, p);
  } else {
    auto kern = conv3x3_wgrad_tc_kernel<false>;
    COVA_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM_BYTES));
    kern<<<grid, WG_THREADS, WG_SMEM_BYTES, st>>>(tx_hi, tx_lo, td_hi, td_lo, p);
  }
  COVA_LAUNCH_OK();
  return launch_wgrad_finalize(ws, inv_scale, 1.f, 9, WG_C, WG_C, dw_oihw, st);
}
