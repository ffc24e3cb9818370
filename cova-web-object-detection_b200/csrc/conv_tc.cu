// A2 BasicBlock half on the 5th-gen tensor cores: 3x3 s1 p1 conv (64->64) as an implicit GEMM with
// tcgen05.mma (accumulators in TMEM), operands staged by TMA, + folded BN (+ residual) (+ ReLU) epilogue.
// Replaces `convnet.4.{b}.conv{1,2}` + `bn{1,2}` (torchvision BasicBlock.forward via
// `/root/reference/models.py:49-51`).
//
// GEMM view per output tile (8 wide x 16 high = 128 pixels = UMMA M):
//     D[128 px, 64 cout] = sum over 9 taps (r,s):  A_rs[128 px, 64 cin] * W_rs[64 cin, 64 cout]
// Activations are NHWC bf16 planes, so a pixel's 64 input channels are exactly one 128-byte swizzle row and
// "im2col" is nothing but a shifted VIEW of one staged halo patch: ONE 4-D TMA box {64 c, 10 w, 18 h} at
// (w0-1, h0-1) lands in shared memory as 180 rows x 128 B (row = hh*10 + ww), and tap (r,s) is the UMMA
// descriptor {start = patch + (r*10+s)*128 B, 8-row-group stride = 10 rows = 1280 B}: M row m = py*8+px reads
// patch row (py+r)*10 + (px+s).  The hardware applies the 128-B swizzle on absolute shared-memory address bits,
// so a start that is only 128-B aligned and a non-1024 group stride address the TMA-written data correctly
// (measured with tools/probe_umma.cu -> profiles/r01a_probe_umma_descriptor.txt).  Conv zero padding = TMA
// out-of-bounds zero fill.  L2->smem traffic is 180/128 = 1.41x the tile (a per-tap load would be 9x); HBM
// traffic stays 1x because neighbouring tiles share halos through L2.
//
// fp32-parity mode (SPLIT): x = hi + lo (two bf16 planes), w = hi + lo; D = Ahi*Whi + Alo*Whi + Ahi*Wlo with
// fp32 accumulation in TMEM: ~2^-17 relative per product, i.e. an fp32 convolution to ~1e-5.
// Measured (tools/probe_mma_rate.cu, profiles/r01b_probe_mma_rate.txt): an M=128 MMA costs
// max(N/2, (4096 + 32 N)/128) clk - tcgen05 reads its shared-memory operands at 128 B/clk/SM, so N = 64 is
// operand-bandwidth bound (48 clk, 66 % of the tensor peak) and N = 128 runs at 99.8 %.  The two products that
// share A (Ahi*Whi and Ahi*Wlo) are therefore issued as ONE N = 128 MMA against the stacked filter [Whi; Wlo]
// (the per-tap filter block holds its hi rows then its lo rows), accumulating into 128 TMEM columns that the
// epilogue folds (col c + col 64+c); Alo*Whi is an N = 64 MMA into the first 64 columns.
// Per (tap, k-step): 64 + 48 = 112 clk instead of 3 x 48.  Plain bf16 mode issues Ahi*Whi only.
//
// Persistent CTAs (grid = #SMs, 1 CTA/SM, 192 threads):
//   warp 0   TMA producer: all 9 filter taps once (resident, 72 KB per plane), then the activation ring
//   warp 1   TMEM alloc + single-thread tcgen05.mma issue; tcgen05.commit releases ring slots / publishes tiles
//   warps 2-9 epilogue: tcgen05.ld (lane = pixel, column = cout), BN scale/shift, residual, ReLU, NHWC stores;
//            two warps per TMEM lane group, 32 output channels each (one warp per scheduler was latency-bound)
// TMEM accumulator is double-buffered (2 x 64 columns): the epilogue of tile i overlaps the MMAs of tile i+1.
// The activation ring holds single PLANES (23 KB slots): per tile the hi plane feeds Ahi*Whi + Ahi*Wlo (72 MMAs),
// the lo plane Alo*Whi (36 MMAs); 3 slots (split) / 6 slots (bf16) keep >= 2 loads in flight next to the
// 144 KB / 72 KB resident filter.
//
// Bound: tensor pipe / its operand feed: SPLIT 36 x (64 + 48) = 4032 clk per tile, bf16 36 x 48 = 1728 clk per tile
// (vs 26-52 MB of HBM traffic per page: 1700-3400 clk per tile-equivalent at the measured 6.5 TB/s).
// Algorithmic work: 2*9*64*64 = 73,728 FLOP per output pixel (7.55 GFLOP per 320x320 page).
#include "common.cuh"
#include "ptx.cuh"
#include "tma_host.cuh"

namespace cova {

constexpr int CT_C = 64;                       // Cin = Cout
constexpr int CT_TW = 8, CT_TH = 16;           // output tile
constexpr int CT_HW = CT_TW + 2, CT_HH = CT_TH + 2;        // halo patch 10 x 18 pixels
constexpr int CT_PATCH_BYTES = CT_HW * CT_HH * 128;        // 23,040 bytes landed per plane load
constexpr int CT_SLOT_BYTES = 23 * 1024;                   // ring slot (1024-B aligned for the swizzle atoms)
constexpr int CT_GROUP_STRIDE = CT_HW * 128;               // 8-row-group stride of a tap view: one patch row
constexpr int CT_W_TAP_BYTES = CT_C * 128;                 // 8,192: one tap of one plane (64 cout rows x 128 B)
constexpr int CT_THREADS = 352;                 // warp 0 TMA, warps 1 and 10 MMA issuers (even / odd tiles), warps 2-9 epilogue
constexpr int CT_ISSUER_B = 10;                 // (2 per TMEM lane group)
constexpr int CT_EPI_THREADS = 256;

template <bool SPLIT>
struct ConvTcCfg {
  static constexpr int NPLANE = SPLIT ? 2 : 1;
  static constexpr int NSTAGE = SPLIT ? 3 : 6;
  static constexpr int STAGE_BYTES = CT_SLOT_BYTES;
  static constexpr int W_BYTES = 9 * CT_W_TAP_BYTES * NPLANE;   // [tap][plane][64 rows][128 B]
  static constexpr int ACC_COLS = SPLIT ? 128 : 64;             // fp32 accumulator columns per buffer
  static constexpr int TMEM_COLS = 2 * ACC_COLS;
  static constexpr int SMEM_BYTES = W_BYTES + NSTAGE * STAGE_BYTES + 1024 /*tail*/ + 1024 /*align slack*/;
};

struct ConvTcTail {   // lives after the operand buffers
  float scale[CT_C], shift[CT_C];
  uint64_t full[8], empty[8], tmem_full[2], tmem_empty[2], wbar;
  uint32_t tmem_base;
};

struct ConvTcParams {
  int B, H, W, tiles_w, tiles_h, n_tiles, relu;
  const float* bn_scale;
  const float* bn_shift;
  const __nv_bfloat16* res_hi;
  const __nv_bfloat16* res_lo;
  void* y0;
  void* y1;
  int pf_tiles;       // TMA producer: L2-prefetch the halo boxes of the tile this many iterations ahead (0 = off)
  int res_load;       // residual loads: 0 = ld.global.nc, 1 = ld.global, 2 = ld.global.L1::no_allocate (default: the
                      // one-sector-per-line access pattern thrashes L1 line allocation; measured -13 % cycles)
  int res_pf_tiles;   // epilogue: L2-prefetch the residual rows this many iterations ahead (0 = off; 1 is the register prefetch)
  unsigned long long* dbg;   // optional per-CTA wait-cycle counters (cova_debug_buffer)
  double* stats;             // optional [2][64]: sum y, sum y^2 over all pixels (raw mode: BatchNorm batch statistics), += here
  const float* res_f32;      // optional fp32 NHWC residual added to an fp32 output (training dgrad + the skip branch's gradient)
};

// mbarrier wait that adds the cycles spent to `acc` when instrumentation is on
__device__ __forceinline__ void mbar_wait_t(uint64_t* bar, uint32_t parity, bool timed, unsigned long long& acc) {
  if (timed) {
    const long long t0 = clock64();
    ptx::mbar_wait(bar, parity);
    acc += (unsigned long long)(clock64() - t0);
  } else {
    ptx::mbar_wait(bar, parity);
  }
}

// HALF: the planes and the filter are fp16 instead of bf16.  !SPLIT: COVA_F16, one product per MMA (11-bit significand:
// ~8x tighter than bf16 - measured ~5e-4 on the logits, inside BASELINE.json's 1e-3 bar).  SPLIT: COVA_F16X2, the same
// three products on split-fp16 planes (22 significand bits: an fp32 convolution to ~1e-6); the filter comes scaled by
// SPLIT_F16_WSCALE, undone in the epilogue's scale.
template <bool SPLIT, int OUT_DTYPE, bool HALF, bool RES32 = false>
__global__ void __launch_bounds__(CT_THREADS, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tm_x_hi, const __grid_constant__ CUtensorMap tm_x_lo,
                  const __grid_constant__ CUtensorMap tm_w_hi, const __grid_constant__ CUtensorMap tm_w_lo,
                  const ConvTcParams p) {
  using Cfg = ConvTcCfg<SPLIT>;
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  unsigned char* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);   // 1024-B aligned (swizzle atoms)
  unsigned char* sm_w = smem;
  unsigned char* sm_a = smem + Cfg::W_BYTES;
  ConvTcTail& tail = *reinterpret_cast<ConvTcTail*>(smem + Cfg::W_BYTES + Cfg::NSTAGE * Cfg::STAGE_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool timed = p.dbg != nullptr;
  const long long t_start = timed ? clock64() : 0;
  unsigned long long wc0 = 0, wc1 = 0;   // wait-cycle counters of this warp's role

  if (threadIdx.x < CT_C) {
    tail.scale[threadIdx.x] = p.bn_scale[threadIdx.x] * ((SPLIT && HALF) ? 1.f / SPLIT_F16_WSCALE : 1.f);
    tail.shift[threadIdx.x] = p.bn_shift[threadIdx.x];
  }
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < Cfg::NSTAGE; ++i) {
      ptx::mbar_init(&tail.full[i], 1);
      ptx::mbar_init(&tail.empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tail.tmem_full[i], 1);
      ptx::mbar_init(&tail.tmem_empty[i], CT_EPI_THREADS);
    }
    ptx::mbar_init(&tail.wbar, 1);
    ptx::fence_barrier_init();
    ptx::prefetch_tensormap(&tm_x_hi);
    ptx::prefetch_tensormap(&tm_w_hi);
    if (SPLIT) {
      ptx::prefetch_tensormap(&tm_x_lo);
      ptx::prefetch_tensormap(&tm_w_lo);
    }
  }
  if (warp == 1) {
    ptx::tmem_alloc(&tail.tmem_base, Cfg::TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tail.tmem_base;

  if (warp == 0) {
    // ======================= TMA producer (warp converged; one elected lane issues) =======================
    if (ptx::elect_one()) {
      // resident filter, [tap][plane][64 cout rows][128 B]: a tap's hi rows are followed by its lo rows
      ptx::mbar_arrive_expect_tx(&tail.wbar, Cfg::W_BYTES);
      for (int tap = 0; tap < 9; ++tap) {
        ptx::tma_load_2d(sm_w + tap * Cfg::NPLANE * CT_W_TAP_BYTES, &tm_w_hi, &tail.wbar, 0, tap * CT_C);
        if (SPLIT)
          ptx::tma_load_2d(sm_w + (tap * 2 + 1) * CT_W_TAP_BYTES, &tm_w_lo, &tail.wbar, 0, tap * CT_C);
      }
    }
    __syncwarp();
    uint32_t stage = 0, phase = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      const int b = tile / (p.tiles_h * p.tiles_w);
      const int th = (tile / p.tiles_w) % p.tiles_h, tw = tile % p.tiles_w;
      const int h0 = th * CT_TH, w0 = tw * CT_TW;
      const int pft = tile + p.pf_tiles * (int)gridDim.x;
      if (p.pf_tiles > 0 && pft < p.n_tiles && ptx::elect_one()) {   // warm L2 for a later iteration's halo boxes
        const int pb = pft / (p.tiles_h * p.tiles_w);
        const int ph0 = ((pft / p.tiles_w) % p.tiles_h) * CT_TH, pw0 = (pft % p.tiles_w) * CT_TW;
        ptx::tma_prefetch_4d(&tm_x_hi, 0, pw0 - 1, ph0 - 1, pb);
        if (SPLIT) ptx::tma_prefetch_4d(&tm_x_lo, 0, pw0 - 1, ph0 - 1, pb);
      }
      __syncwarp();
      for (int pl = 0; pl < Cfg::NPLANE; ++pl) {
        mbar_wait_t(&tail.empty[stage], phase ^ 1, timed, wc0);
        if (ptx::elect_one()) {
          ptx::mbar_arrive_expect_tx(&tail.full[stage], CT_PATCH_BYTES);
          ptx::tma_load_4d(sm_a + stage * Cfg::STAGE_BYTES, pl == 0 ? &tm_x_hi : &tm_x_lo, &tail.full[stage], 0, w0 - 1,
                           h0 - 1, b);
        }
        __syncwarp();
        if (++stage == Cfg::NSTAGE) { stage = 0; phase ^= 1; }
      }
    }
    if (timed && lane == 0) atomicAdd(p.dbg + blockIdx.x * 8 + 2, wc0);
  } else if (warp == 1 || warp == CT_ISSUER_B) {
    // ======================= MMA issuers (warp converged; one elected lane issues) =======================
    // Two of them, one per accumulator buffer: the barrier checks, commits and loop overhead between two planes / tiles are
    // serial latency in the issuing thread during which the tensor pipe drains its queue (stem_tc.cu, round-2 probes); the
    // other issuer's MMAs fill it.  A tile's MMAs all come from one thread, so every tcgen05.commit covers what it must.
    const uint32_t ipar = warp == 1 ? 0u : 1u;
    // Descriptors differ only in their 14-bit start-address field, so each MMA costs two 32-bit adds on
    // warp-uniform values (the elect.sync guard lets the compiler keep them in uniform registers; an
    // `if (lane == 0)` region would wrap every UTCHMMA in a per-lane serialisation loop).
    constexpr uint32_t idesc64 = HALF ? ptx::umma_idesc_f16(128, CT_C) : ptx::umma_idesc_bf16(128, CT_C);
    constexpr uint32_t idesc128 = HALF ? ptx::umma_idesc_f16(128, 2 * CT_C) : ptx::umma_idesc_bf16(128, 2 * CT_C);
    const uint64_t da0 = ptx::umma_desc_sw128(ptx::smem_u32(sm_a), CT_GROUP_STRIDE);
    const uint64_t db0 = ptx::umma_desc_sw128(ptx::smem_u32(sm_w), 1024);
    ptx::mbar_wait(&tail.wbar, 0);
    uint32_t stage = 0, phase = 0, it = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
      if ((it & 1u) != ipar) {   // the other issuer's tile: step over its ring slots
        for (int pl = 0; pl < Cfg::NPLANE; ++pl)
          if (++stage == Cfg::NSTAGE) { stage = 0; phase ^= 1; }
        continue;
      }
      const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
      mbar_wait_t(&tail.tmem_empty[acc], acc_phase ^ 1, timed, wc1);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * Cfg::ACC_COLS;
      for (int pl = 0; pl < Cfg::NPLANE; ++pl) {   // pl 0: A = hi plane x [Whi; Wlo]; pl 1: A = lo plane x Whi
        mbar_wait_t(&tail.full[stage], phase, timed, wc0);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          uint64_t da_s = da0 + ((stage * Cfg::STAGE_BYTES) >> 4);
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            uint64_t db_t = db0;
            asm volatile("" : "+l"(da_s), "+l"(db_t));   // keep the 36 x 2 descriptors from being hoisted and spilled
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {   // 4 x K=16 bf16 (32 B) inside the 128-B swizzle row
              const uint64_t da = da_s + ((((tap / 3) * CT_HW + (tap % 3)) * 128 + kk * 32) >> 4);
              const uint64_t db = db_t + ((tap * Cfg::NPLANE * CT_W_TAP_BYTES + kk * 32) >> 4);
              ptx::umma_bf16(d_tmem, da, db, (SPLIT && pl == 0) ? idesc128 : idesc64, (pl | tap | kk) != 0);
            }
          }
          ptx::umma_commit(&tail.empty[stage]);                           // slot free once these MMAs have read it
          if (pl == Cfg::NPLANE - 1) ptx::umma_commit(&tail.tmem_full[acc]);   // accumulator complete
        }
        __syncwarp();
        if (++stage == Cfg::NSTAGE) { stage = 0; phase ^= 1; }
      }
    }
    if (timed && lane == 0 && warp == 1) {
      atomicAdd(p.dbg + blockIdx.x * 8 + 0, wc0);
      atomicAdd(p.dbg + blockIdx.x * 8 + 1, wc1);
      atomicAdd(p.dbg + blockIdx.x * 8 + 5, (unsigned long long)it);
    }
  } else {
    // ======================= epilogue warps (TMEM lane group = warp % 4; channel half = (warp-2)/4) ==========
    const int lg = warp & 3;
    const int ch0 = ((warp - 2) >> 2) * 32;       // this warp's 32 output channels
    const int m = lg * 32 + lane;                 // accumulator row = pixel within the tile
    const int py = m >> 3, px = m & 7;
    // The residual of tile i+1 is requested before tile i is processed (software pipeline): with the loads issued
    // only one barrier wait ahead, every tile exposed a full DRAM round trip on the epilogue's critical path.
    const bool has_res = p.res_hi != nullptr;
    uint32_t rh[2][8], rl[2][8], rh_n[2][8], rl_n[2][8];      // 32 bf16 per plane = 2 x 32-byte sectors
    auto tile_pix = [&](int tile, bool& inb) -> size_t {
      const int b = tile / (p.tiles_h * p.tiles_w);
      const int th = (tile / p.tiles_w) % p.tiles_h, tw = tile % p.tiles_w;
      const int oh = th * CT_TH + py, ow = tw * CT_TW + px;
      inb = oh < p.H && ow < p.W;
      return (((size_t)b * p.H + oh) * p.W + ow) * CT_C + ch0;
    };
    auto load_res = [&](int tile) {
      bool ok;
      const size_t px_off = tile_pix(tile, ok);
      if (has_res && tile < p.n_tiles && ok) {
        if (p.res_load == 1) {
#pragma unroll
          for (int j = 0; j < 2; ++j) ld_global_v8(p.res_hi + px_off + j * 16, rh_n[j]);
          if (SPLIT) {
#pragma unroll
            for (int j = 0; j < 2; ++j) ld_global_v8(p.res_lo + px_off + j * 16, rl_n[j]);
          }
        } else if (p.res_load == 2) {
#pragma unroll
          for (int j = 0; j < 2; ++j) ld_global_na_v8(p.res_hi + px_off + j * 16, rh_n[j]);
          if (SPLIT) {
#pragma unroll
            for (int j = 0; j < 2; ++j) ld_global_na_v8(p.res_lo + px_off + j * 16, rl_n[j]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 2; ++j) ld_global_nc_v8(p.res_hi + px_off + j * 16, rh_n[j]);
          if (SPLIT) {
#pragma unroll
            for (int j = 0; j < 2; ++j) ld_global_nc_v8(p.res_lo + px_off + j * 16, rl_n[j]);
          }
        }
      }
    };
    auto prefetch_res = [&](int tile) {   // L2 only: one 64-byte half row per thread and plane
      bool ok;
      const size_t px_off = tile_pix(tile, ok);
      if (tile < p.n_tiles && ok) {
        ptx::prefetch_l2(p.res_hi + px_off);
        if (SPLIT) ptx::prefetch_l2(p.res_lo + px_off);
      }
    };
    load_res(blockIdx.x);
    const bool res_pf = has_res && p.res_pf_tiles > 1;
    double st_s = 0.0, st_q = 0.0;                  // running statistics of channel ch0 + lane
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
      const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
      bool inb;
      const size_t pix = tile_pix(tile, inb);
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int e = 0; e < 8; ++e) { rh[j][e] = rh_n[j][e]; rl[j][e] = rl_n[j][e]; }
      load_res(tile + gridDim.x);
      if (res_pf) prefetch_res(tile + p.res_pf_tiles * (int)gridDim.x);

      mbar_wait_t(&tail.tmem_full[acc], acc_phase, timed && warp == 2, wc0);
      ptx::tc_fence_after();
      uint32_t v[2][16];
      const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + acc * Cfg::ACC_COLS + ch0;
#pragma unroll
      for (int q = 0; q < 2; ++q) ptx::tmem_ld16(taddr + q * 16, v[q]);
      ptx::tmem_ld_wait();
      float o[32];
#pragma unroll
      for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int j = 0; j < 16; ++j) o[q * 16 + j] = __uint_as_float(v[q][j]);
      if (SPLIT) {   // columns 64..127 hold Ahi*Wlo
#pragma unroll
        for (int q = 0; q < 2; ++q) ptx::tmem_ld16(taddr + CT_C + q * 16, v[q]);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
          for (int j = 0; j < 16; ++j) o[q * 16 + j] += __uint_as_float(v[q][j]);
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(&tail.tmem_empty[acc]);     // accumulator buffer is free for tile it+2

      if (p.stats != nullptr) {   // raw mode (no residual / ReLU): statistics of the values as they are STORED (warp-collective)
        float z[32], z2[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          float t = fmaf(o[c], tail.scale[ch0 + c], tail.shift[ch0 + c]);
          if (OUT_DTYPE == COVA_BF16 && !HALF) t = round_bf16(t);
          z[c] = inb ? t : 0.f;
          z2[c] = z[c] * z[c];
        }
        st_s += (double)warp_transpose_sum32(z, lane);
        st_q += (double)warp_transpose_sum32(z2, lane);
      }
      if (!inb) continue;
#pragma unroll
      for (int c = 0; c < 32; ++c) o[c] = fmaf(o[c], tail.scale[ch0 + c], tail.shift[ch0 + c]);
      if (has_res) {
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            if (HALF) {
              const float2 r2 = unpack2_f16(rh[j][e]);
              o[j * 16 + 2 * e] += r2.x;
              o[j * 16 + 2 * e + 1] += r2.y;
              if (SPLIT) {
                const float2 l2 = unpack2_f16(rl[j][e]);
                o[j * 16 + 2 * e] += l2.x;
                o[j * 16 + 2 * e + 1] += l2.y;
              }
              continue;
            }
            o[j * 16 + 2 * e] += bf16lo_to_f32(rh[j][e]);
            o[j * 16 + 2 * e + 1] += bf16hi_to_f32(rh[j][e]);
            if (SPLIT) {
              o[j * 16 + 2 * e] += bf16lo_to_f32(rl[j][e]);
              o[j * 16 + 2 * e + 1] += bf16hi_to_f32(rl[j][e]);
            }
          }
      }
      if (RES32) {   // fp32 residual row segment, 8 channels at a time (32 more live registers spilled; the 8 epilogue warps hide the loads)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t rf[8];
          ld_global_na_v8(p.res_f32 + pix + j * 8, rf);
#pragma unroll
          for (int e = 0; e < 8; ++e) o[j * 8 + e] += __uint_as_float(rf[e]);
        }
      }
      if (p.relu) {
#pragma unroll
        for (int c = 0; c < 32; ++c) o[c] = fmaxf(o[c], 0.f);
      }
      if (OUT_DTYPE == COVA_F32) {
        float* dst = reinterpret_cast<float*>(p.y0) + pix;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t w8[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) w8[e] = __float_as_uint(o[j * 8 + e]);
          st_global_v8(dst + j * 8, w8);
        }
      } else {
        __nv_bfloat16* dh = reinterpret_cast<__nv_bfloat16*>(p.y0) + pix;
        __nv_bfloat16* dl = reinterpret_cast<__nv_bfloat16*>(p.y1) + pix;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          uint32_t hw[8], lw[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            if (OUT_DTYPE == COVA_BF16X2 && HALF) split_f16x2(o[j * 16 + 2 * e], o[j * 16 + 2 * e + 1], hw[e], lw[e]);
            else if (OUT_DTYPE == COVA_BF16X2) split_bf16x2(o[j * 16 + 2 * e], o[j * 16 + 2 * e + 1], hw[e], lw[e]);
            else if (HALF) hw[e] = pack2_f16(o[j * 16 + 2 * e], o[j * 16 + 2 * e + 1]);
            else hw[e] = pack2_bf16(o[j * 16 + 2 * e], o[j * 16 + 2 * e + 1]);
          }
          st_global_v8(dh + j * 16, hw);
          if (OUT_DTYPE == COVA_BF16X2) st_global_v8(dl + j * 16, lw);
        }
      }
    }
    if (p.stats != nullptr) {
      atomicAdd(p.stats + ch0 + lane, st_s);
      atomicAdd(p.stats + CT_C + ch0 + lane, st_q);
    }
  }

  if (timed && warp == 2 && lane == 0) atomicAdd(p.dbg + blockIdx.x * 8 + 3, wc0);
  ptx::tc_fence_before();
  __syncthreads();
  if (timed && threadIdx.x == 0) atomicAdd(p.dbg + blockIdx.x * 8 + 4, (unsigned long long)(clock64() - t_start));
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <bool SPLIT, int OUT_DTYPE, bool HALF = false, bool RES32 = false>
static int launch_conv_tc(const CUtensorMap& xh, const CUtensorMap& xl, const CUtensorMap& wh, const CUtensorMap& wl,
                          const ConvTcParams& p, cudaStream_t st) {
  using Cfg = ConvTcCfg<SPLIT>;
  static_assert(sizeof(ConvTcTail) <= 1024, "tail too large");
  auto kern = conv3x3_tc_kernel<SPLIT, OUT_DTYPE, HALF, RES32>;
  COVA_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  const int grid = p.n_tiles < sm_count() ? p.n_tiles : sm_count();
  kern<<<grid, CT_THREADS, Cfg::SMEM_BYTES, st>>>(xh, xl, wh, wl, p);
  COVA_LAUNCH_OK();
  return COVA_OK;
}

int conv3x3_tc(const void* x_hi, const void* x_lo, int split, int half, int B, int H, int W, const void* w_hi, const void* w_lo,
               const float* bn_scale, const float* bn_shift, const void* res_hi, const void* res_lo, int relu,
               int out_dtype, void* y0, void* y1, cudaStream_t st, double* stats, const float* res_f32) {
  CUtensorMap tx_hi, tx_lo, tw_hi, tw_lo;
  const uint64_t xd[4] = {(uint64_t)CT_C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
  const uint64_t xs[3] = {(uint64_t)CT_C * 2, (uint64_t)W * CT_C * 2, (uint64_t)H * W * CT_C * 2};
  const uint32_t xb[4] = {CT_C, CT_HW, CT_HH, 1};
  const uint64_t wd[2] = {(uint64_t)CT_C, (uint64_t)9 * CT_C};
  const uint64_t ws[1] = {(uint64_t)CT_C * 2};
  const uint32_t wb[2] = {CT_C, CT_C};   // one tap of one plane per load
  int rc;
  if ((rc = make_tmap_bf16(&tx_hi, x_hi, 4, xd, xs, xb))) return rc;
  if ((rc = make_tmap_bf16(&tw_hi, w_hi, 2, wd, ws, wb))) return rc;
  tx_lo = tx_hi;
  tw_lo = tw_hi;
  if (split) {
    if ((rc = make_tmap_bf16(&tx_lo, x_lo, 4, xd, xs, xb))) return rc;
    if ((rc = make_tmap_bf16(&tw_lo, w_lo, 2, wd, ws, wb))) return rc;
  }
  ConvTcParams p;
  p.B = B; p.H = H; p.W = W;
  p.tiles_w = ceil_div(W, CT_TW);
  p.tiles_h = ceil_div(H, CT_TH);
  p.n_tiles = B * p.tiles_w * p.tiles_h;
  p.relu = relu;
  p.bn_scale = bn_scale; p.bn_shift = bn_shift;
  p.res_hi = (const __nv_bfloat16*)res_hi;
  p.res_lo = (const __nv_bfloat16*)res_lo;
  p.y0 = y0; p.y1 = y1;
  p.pf_tiles = knob(COVA_KNOB_CONV_L2_PREFETCH, 0);
  p.res_pf_tiles = knob(COVA_KNOB_CONV_RES_PREFETCH, 1);
  p.res_load = knob(COVA_KNOB_CONV_RES_LOAD, 2);
  p.dbg = debug_words(8LL * sm_count());
  p.stats = stats;
  p.res_f32 = res_f32;
  if (stats) COVA_CUDA_OK(cudaMemsetAsync(stats, 0, 2 * CT_C * sizeof(double), st));
#define DISPATCH(SP)                                                                               \
  switch (out_dtype) {                                                                             \
    case COVA_F32: return launch_conv_tc<SP, COVA_F32>(tx_hi, tx_lo, tw_hi, tw_lo, p, st);         \
    case COVA_BF16: return launch_conv_tc<SP, COVA_BF16>(tx_hi, tx_lo, tw_hi, tw_lo, p, st);       \
    default: return launch_conv_tc<SP, COVA_BF16X2>(tx_hi, tx_lo, tw_hi, tw_lo, p, st);            \
  }
  if (res_f32 != nullptr) {   // training dgrad: split planes in, fp32 out + fp32 residual (its own instantiation: the 32
    // residual registers must not burden the inference kernels)
    if (!(split && out_dtype == COVA_F32)) { set_error("conv3x3_tc: an fp32 residual goes with split planes in and fp32 out"); return COVA_ERR_ARG; }
    return half ? launch_conv_tc<true, COVA_F32, true, true>(tx_hi, tx_lo, tw_hi, tw_lo, p, st)
                : launch_conv_tc<true, COVA_F32, false, true>(tx_hi, tx_lo, tw_hi, tw_lo, p, st);
  }
  if (half && split) {   // split-fp16 planes in; split-fp16 planes or fp32 out (stored through the COVA_BF16X2 code path)
    if (out_dtype == COVA_F32) return launch_conv_tc<true, COVA_F32, true>(tx_hi, tx_lo, tw_hi, tw_lo, p, st);
    return launch_conv_tc<true, COVA_BF16X2, true>(tx_hi, tx_lo, tw_hi, tw_lo, p, st);
  }
  if (half) {   // fp16 plane in; fp16 plane or fp32 out (out_dtype COVA_F16 is stored through the COVA_BF16 code path)
    if (out_dtype == COVA_F32) return launch_conv_tc<false, COVA_F32, true>(tx_hi, tx_lo, tw_hi, tw_lo, p, st);
    return launch_conv_tc<false, COVA_BF16, true>(tx_hi, tx_lo, tw_hi, tw_lo, p, st);
  }
  if (split) { DISPATCH(true) } else { DISPATCH(false) }
#undef DISPATCH
}

}  // namespace cova
