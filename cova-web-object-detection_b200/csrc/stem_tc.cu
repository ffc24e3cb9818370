// A2 stem on the 5th-gen tensor cores: conv 7x7 s2 p3 (3->64) + folded BN + ReLU + maxpool 3x3 s2 p1, ONE kernel
// (replaces `convnet[0:4]`, `/root/reference/models.py:49-51`, applied `:125`).  NCHW fp32 images in, NHWC planes out.
//
// The convolution is an implicit GEMM whose A operand is never materialised ("im2col staged in shared memory"
// degenerates to staging the raw image rows once):
//   * converter warps read image rows (NCHW fp32, coalesced along x) and write them to a shared-memory ring as
//     4-channel bf16 pixels (8 B: c0,c1,c2,0), hi plane and lo plane (x = hi + lo, split-bf16);
//   * for filter row r the GEMM row of output pixel ox is the 8 input pixels 2ox-3 .. 2ox+4 of image row 2oy+r-3:
//     32 K-elements (s = 0..7, c = 0..3; the s = 7 and c = 3 weights are zero) that sit CONTIGUOUSLY in the ring row,
//     16 B further along for each next output pixel (stride 2 x 8 B).  A SWIZZLE_NONE K-major tcgen05 descriptor
//     with {rows 16 B apart (core matrix), K-chunk stride LBO = 16 B, 8-row-group stride SBO = 128 B} reads exactly
//     that sliding window - K-chunk 1 of row j aliases K-chunk 0 of row j+1 (tools/probe_umma_noswz.cu).
//     So one conv row of 128 output pixels = 7 filter rows x 2 MMAs (K = 16 each) on the ring rows, no copies.
//   * accumulators (128 px x 64 cout fp32) live in TMEM, double buffered; epilogue warps apply BN + ReLU and
//     max-pool on the fly: vertical 3-max as a per-lane running maximum in registers while the CTA marches down conv
//     rows, then one horizontal 3-max with warp shuffles (lane = conv pixel) per pooled row; only pooled pixels are
//     written.  ReLU is folded into the maximum (the running maximum starts from 0).
//
// Work decomposition: CTA = (page, band of pooled rows), processed strip by strip (128 conv columns = 64 pooled
// columns).  The right-most conv column of strip i is kept per conv row in shared memory and is the left
// neighbour of strip i+1, so no conv column is recomputed; each band recomputes one conv row (its top halo).
// HBM traffic: image read once (+2 % strip / +3.5 % band halos), pooled map written once.
//
// Roles (26 warps): two MMA issuer warps (even / odd conv rows, one accumulator buffer each), 16 epilogue warps
// (4 TMEM lane groups x 4 channel groups), 8 converter warps.  Round-2 measurements (tools/probe_mma_contention.cu,
// profiles/r02e_*): with ONE issuer the tensor pipe idled - per conv row the issuing thread spent ~1000 clk outside its 14-28
// MMAs (barrier checks, commits, descriptor arithmetic), more than the MMAs themselves need; descriptors are now
// "uniform base + immediate" (one specialisation per ring phase) and the second issuer hides the rest.
//
// Bound, fp32 images (3 products): tensor pipe / operand feed, 14 x (64 + 48) = 1568 clk per 128 conv pixels; measured
// 1623 clk with the pooled-row emission switched off, ~1870 with it.  uint8 images (integer pixels are exact in one plane:
// 2 products, 14 x 64 = 896 clk): the epilogue's issue slots - ~300 warp instructions per warp and conv row, 16 warps.
// Algorithmic work: 2*64*147 FLOP per conv pixel (7.707 GFLOP per 1280x1280 page).
#include "common.cuh"
#include "ptx.cuh"

namespace cova {

constexpr int SX_TM = 128;                 // conv pixels per MMA tile (one conv-row segment)
constexpr int SX_NPX = 264;                // staged input pixels per ring row (2*127 + 8 = 262, padded)
constexpr int SX_ROW_BYTES = SX_NPX * 8;   // 2112 (4 bf16 channels per pixel)
constexpr int SX_R = 16;                   // ring slots (7 live + 4 in flight + slack)
constexpr int SX_KCHUNKS = 28;             // 7 filter rows x 4 chunks of 8 K-elements
constexpr int SX_W_CHUNK = 2 * 64 * 16;    // 2,048 B: one K-chunk = 64 hi rows then 64 lo rows of 8 bf16
constexpr int SX_W_BYTES = SX_KCHUNKS * SX_W_CHUNK;   // 57,344 B: [chunk][plane][cout][8] bf16
constexpr int SX_MAXROWS = 81;             // conv rows per band (2*40 + 1)
constexpr int SX_ND = 16;                  // tile-row completion barriers
constexpr int SX_NACC = 2;                 // accumulator buffers in TMEM (4 measured no faster: the epilogue's issue slots bound the kernel,
                                           // not the hand-back latency)
constexpr int SX_EPI_WARPS = 16;            // 4 per TMEM lane group, 16 output channels each
constexpr int SX_EPI_THREADS = SX_EPI_WARPS * 32;
constexpr int SX_CH = 64 / (SX_EPI_WARPS / 4);   // channels per epilogue warp
constexpr int sx_threads(int ncv) { return 32 * (2 + SX_EPI_WARPS + ncv); }   // warp 0 and the last warp: MMA issuers (even / odd conv
                                                                              // rows); epilogue warps; ncv converter warps


template <bool SPLIT>
struct StemTcSmem {
  static constexpr int NPLANE = SPLIT ? 2 : 1;
  alignas(128) unsigned char w[SX_W_BYTES];   // both planes always (the lo rows are simply unused in bf16 mode)
  alignas(128) unsigned char ring[NPLANE][SX_R][SX_ROW_BYTES];
  alignas(16) float edge[3][SX_MAXROWS][64];           // right-most conv column of strips s-1, s (three buffers: with ONE emitted
                                                       // row per strip, strip s+2's write is not ordered after strip s+1's read of a two-buffer ring)
  float xch[2][4][64];                     // lane-31 rows exchanged between the 4 epilogue warps
  alignas(16) float scale[64], shift[64];
  uint32_t lut[256];                       // uint8 images: bf16 hi (low half) | bf16 lo (high half) of v/255
  uint64_t in_full[SX_R], pair_full[SX_ND], mma_done[SX_ND], tmem_empty[SX_NACC], wbar;
  uint32_t tmem_base;
};

struct StemTcParams {
  const void* img;
  int B, H, W, Hc, Wc, Hp, Wp;
  int bands_per_page, nb;                  // pooled rows per band
  const unsigned char* w_packed;           // [28][2][64][8] bf16
  const float* bn_scale;
  const float* bn_shift;
  void* out0;
  void* out1;
  float* raw_out;                          // training mode: un-normalised, un-pooled conv output [B,Hc,Wc,64] fp32 (or NULL)
  int raw_bf16;                            // ... stored as bf16 instead (the bf16 training mode)
  double* stats;                           // raw mode, optional [2][64]: sum y, sum y^2 over all conv pixels (BatchNorm statistics)
  int u8_int;                              // uint8 images, split modes: integer pixels in ONE plane (exact), 1/255 in the epilogue scale,
                                           // two products instead of three (the Alo x Whi MMA and the lo ring stores are skipped)
  unsigned long long* dbg;                 // optional wait-cycle counters (cova_debug_buffer), 8 words per CTA
  int pf_rows;                             // converter warps L2-prefetch the image row they will load this many turns ahead
};

__device__ __forceinline__ uint64_t desc_noswz(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;   // K-chunk (8 elements) stride
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;   // 8-row-group stride
  d |= (uint64_t)1 << 46;                       // version 1; layout_type 0 = SWIZZLE_NONE
  return d;
}

// Converter stores.  A lane owns ring pixels 4wi-2 .. 4wi+1 (8 bytes each) = two aligned 16-byte pairs, 32 bytes from its
// neighbour lane's.  Eight-byte stores at that pitch were 4-way bank-conflicted (8 wavefronts per instruction; the LSU
// wavefronts come out of the same 128 B/clk the tensor core fetches its operands with).  A 128-bit store is processed per
// quarter-warp; lanes 0-3 of each quarter store their first pair while lanes 4-7 store their second one, which covers all 32
// banks exactly once.  Pair A = pixels (4wi-2, 4wi-1) exists for wi >= 1; pair B = (4wi, 4wi+1) for 4wi+1 < SX_NPX.
template <bool SPLIT>
__device__ __forceinline__ void stem_store_pairs(unsigned char* dst_hi, int wi, int lane, const uint2 (&hi)[4], const uint2 (&lo)[4],
                                                 bool store_lo) {
  const bool sw = (lane >> 2) & 1;
  const bool ok_a = wi >= 1 && wi < 67, ok_b = wi < 67 && 4 * wi + 1 < SX_NPX;
#pragma unroll
  for (int step = 0; step < 2; ++step) {
    const bool second = (step == 1) != sw;          // this step stores pair B
    const bool ok = second ? ok_b : ok_a;
    const int off = (second ? 4 * wi : 4 * wi - 2) * 8;
    if (ok) {
      *reinterpret_cast<uint4*>(dst_hi + off) = second ? make_uint4(hi[2].x, hi[2].y, hi[3].x, hi[3].y) : make_uint4(hi[0].x, hi[0].y, hi[1].x, hi[1].y);
      if (SPLIT && store_lo)
        *reinterpret_cast<uint4*>(dst_hi + SX_R * SX_ROW_BYTES + off) =
            second ? make_uint4(lo[2].x, lo[2].y, lo[3].x, lo[3].y) : make_uint4(lo[0].x, lo[0].y, lo[1].x, lo[1].y);
    }
  }
}

// The 14 (filter row, half) MMA steps of one conv-row tile whose first input row sits in ring slot S0 (compile time).
// Ahi x [Whi; Wlo] is ONE N = 128 MMA (operand feed: 64 clk instead of 2 x 48, see conv_tc.cu), Alo x Whi N = 64.
template <int S0, bool SPLIT, bool HALF, bool LO_MMA>
__device__ __forceinline__ void stem_issue_tile(uint32_t d_tmem, uint64_t da0, uint64_t db0) {
  static_assert(SX_R == 16, "one specialisation per ring phase");
  constexpr uint32_t idesc64 = HALF ? ptx::umma_idesc_f16(128, 64) : ptx::umma_idesc_bf16(128, 64);
  constexpr uint32_t idesc128 = HALF ? ptx::umma_idesc_f16(128, 128) : ptx::umma_idesc_bf16(128, 128);
#pragma unroll
  for (int r = 0; r < 7; ++r) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {               // s = 0..3 / 4..7 -> 4 pixels = 32 B further
      const uint64_t da_hi = da0 + (uint64_t)(((((S0 + r) % SX_R) * SX_ROW_BYTES) + half * 32) >> 4);
      const uint64_t db = db0 + (uint64_t)(((r * 4 + half * 2) * SX_W_CHUNK) >> 4);
      ptx::umma_bf16(d_tmem, da_hi, db, SPLIT ? idesc128 : idesc64, (r | half) != 0);
      if (LO_MMA) ptx::umma_bf16(d_tmem, da_hi + ((SX_R * SX_ROW_BYTES) >> 4), db, idesc64, 1);
    }
  }
}

// HALF (single-plane mode only): ring pixels, filter and output plane are fp16 instead of bf16 (COVA_F16).
// RAW: training mode (raw conv output + optional batch statistics instead of BN + ReLU + pool) - its own instantiations, so the
// inference kernels carry none of that code.
template <bool SPLIT, int OUT_DTYPE, bool U8, int NCV, bool HALF, bool RAW = false>
__global__ void __launch_bounds__(sx_threads(NCV), 1)
stem_tc_kernel(const StemTcParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  StemTcSmem<SPLIT>& sm = *reinterpret_cast<StemTcSmem<SPLIT>*>(smem_raw);
  constexpr int ACC_COLS = SPLIT ? 128 : 64;      // fp32 accumulator columns per buffer
  constexpr int TMEM_COLS = SX_NACC * ACC_COLS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- band geometry (identical in every role)
  const int b = blockIdx.x / p.bands_per_page, band = blockIdx.x % p.bands_per_page;
  const int py0 = band * p.nb, py1 = min(py0 + p.nb, p.Hp);
  const int oy_begin = py0 > 0 ? 2 * py0 - 1 : 0;
  const int oy_end = min(2 * py1, p.Hc);             // exclusive
  const int n_conv = oy_end - oy_begin;              // conv rows per strip
  const int NQ = 2 * (n_conv - 1) + 7;               // input rows per strip
  const int y_first = 2 * oy_begin - 3;
  const int n_strips = (p.Wc + SX_TM - 1) / SX_TM;

  if (threadIdx.x < 64) {
    sm.scale[threadIdx.x] = (RAW ? 1.f : p.bn_scale[threadIdx.x]) * ((SPLIT && HALF) ? 1.f / SPLIT_F16_WSCALE : 1.f) *
                            ((U8 && SPLIT && p.u8_int) ? 1.f / 255.f : 1.f);
    sm.shift[threadIdx.x] = RAW ? 0.f : p.bn_shift[threadIdx.x];
  }
  if (U8 && threadIdx.x < 256) {   // v/255 with IEEE division == torchvision ToTensor; split once per pixel value
    const float pv = (SPLIT && p.u8_int) ? (float)threadIdx.x : __fdiv_rn((float)threadIdx.x, 255.f);   // u8_int: lo = 0 exactly
    if (HALF && SPLIT) {
      uint32_t h, l;
      split_f16x2(pv, 0.f, h, l);
      sm.lut[threadIdx.x] = (h & 0xffffu) | (l << 16);
    } else if (HALF) {
      sm.lut[threadIdx.x] = pack2_f16(pv, 0.f);
    } else {
      __nv_bfloat16 h, l;
      split_bf16(pv, h, l);
      sm.lut[threadIdx.x] = pack_bf16x2(h, l);
    }
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < SX_R; ++i) ptx::mbar_init(&sm.in_full[i], 1);
    for (int i = 0; i < SX_ND; ++i) {
      ptx::mbar_init(&sm.mma_done[i], 2);        // tile t done AND tile t-1 done (the two issuers commit independently)
      ptx::mbar_init(&sm.pair_full[i], 2);       // the two input rows that are new for one conv row
    }
    for (int i = 0; i < SX_NACC; ++i) {
      ptx::mbar_init(&sm.tmem_empty[i], SX_EPI_WARPS);   // one arrival per epilogue warp (512 per-thread arrivals on one
                                                          // shared-memory word per tile were serialised in the barrier unit)
    }
    ptx::mbar_init(&sm.wbar, 1);
    ptx::mbar_arrive(&sm.mma_done[0]);           // tile 0 has no predecessor
    ptx::fence_barrier_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc(&sm.tmem_base, TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = sm.tmem_base;
  if (n_conv <= 0) goto teardown;

  if (warp == 0 || warp == 1 + SX_EPI_WARPS + NCV) {
    // ======================= MMA issuers (warp converged; one elected lane issues) =======================
    // Two of them, one per accumulator buffer (even / odd tiles).  What an issuer does per tile besides the MMAs - two barrier
    // waits at ~170 clk each even when ready, the commits, the dispatch - is ~1000 clk of serial latency in ONE thread, more than
    // the tensor pipe needs for the tile's 14 N = 128 MMAs (896 clk); with two issuers that latency hides behind the other's MMAs.
    const uint32_t ipar = warp == 0 ? 0u : 1u;
    if (warp == 0 && ptx::elect_one()) {
      ptx::mbar_arrive_expect_tx(&sm.wbar, SX_W_BYTES);
      ptx::tma_bulk_g2s(sm.w, p.w_packed, SX_W_BYTES, &sm.wbar);
    }
    __syncwarp();
    ptx::mbar_wait(&sm.wbar, 0);
    const uint64_t da0 = desc_noswz(ptx::smem_u32(&sm.ring[0][0][0]), 16, 128);
    const uint64_t db0 = desc_noswz(ptx::smem_u32(&sm.w[0]), SX_W_CHUNK, 128);
    uint32_t t = 0;
    const bool timed = p.dbg != nullptr && warp == 0;
    const long long t_start = timed ? clock64() : 0;
    long long wt0 = 0, wt1 = 0, wt2 = 0, wt3 = 0;
    for (int strip = 0; strip < n_strips; ++strip) {
      for (int i = 0; i < n_conv; ++i, ++t) {
        if ((t & 1u) != ipar) continue;
        const uint32_t g0 = (uint32_t)strip * NQ + 2 * i;
        long long c0 = timed ? clock64() : 0;
        // Rows that are new for this conv row: 2i+5 and 2i+6 arrive on ONE pair barrier.  The rows below them belong to the pair
        // barriers of tiles t-1 (the other issuer's), t-2, t-3: waiting for t and t-1 on every tile covers all of them.  The
        // first two conv rows of a strip also need rows 0..4, which keep their per-slot barriers.
        if (i <= 1) {
          for (int r = 0; r < 5 - 2 * i; ++r) {
            const uint32_t g = g0 + r;
            ptx::mbar_wait(&sm.in_full[g % SX_R], (g / SX_R) & 1);
          }
        }
        const uint32_t acc = t % SX_NACC;
        const uint32_t tp = t > 0 ? t - 1 : 0;
        // one polling loop, the three checks in flight together (a ready mbarrier check costs the issuing thread ~170 clk)
        {
          uint32_t spins = 0;
          while (true) {
            const bool a = ptx::mbar_try_wait(&sm.pair_full[t % SX_ND], (t / SX_ND) & 1);
            const bool b = ptx::mbar_try_wait(&sm.pair_full[tp % SX_ND], (tp / SX_ND) & 1);
            const bool c = ptx::mbar_try_wait(&sm.tmem_empty[acc], ((t / SX_NACC) & 1) ^ 1);
            if (a & b & c) break;
            if (++spins > (1u << 26)) __trap();
          }
        }
        if (timed) { const long long c1 = clock64(); wt0 += c1 - c0; c0 = c1; }
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * ACC_COLS;
        if (ptx::elect_one()) {
          // The ring slot of filter row r is (g0 + r) % SX_R.  Computing 7 descriptors per tile from a run-time g0 cost the issuing
          // thread ~60 integer / R2UR instructions between MMAs (the tensor pipe idled ~40 clk after every MMA); one
          // specialisation per value of g0 % SX_R turns every descriptor into "uniform base + immediate".
          const long long m0 = timed ? clock64() : 0;
          // (uint8 integer mode: the lo plane is identically zero, so its MMA is not issued)
#define COVA_STEM_CASE(S, LO) case S: stem_issue_tile<S, SPLIT, HALF, LO>(d_tmem, da0, db0); break;
#define COVA_STEM_SWITCH(LO)                                                                                                  \
  switch (g0 % SX_R) {                                                                                                        \
    COVA_STEM_CASE(0, LO) COVA_STEM_CASE(1, LO) COVA_STEM_CASE(2, LO) COVA_STEM_CASE(3, LO) COVA_STEM_CASE(4, LO)             \
    COVA_STEM_CASE(5, LO) COVA_STEM_CASE(6, LO) COVA_STEM_CASE(7, LO) COVA_STEM_CASE(8, LO) COVA_STEM_CASE(9, LO)             \
    COVA_STEM_CASE(10, LO) COVA_STEM_CASE(11, LO) COVA_STEM_CASE(12, LO) COVA_STEM_CASE(13, LO) COVA_STEM_CASE(14, LO)        \
    COVA_STEM_CASE(15, LO)                                                                                                    \
  }
          if (SPLIT && !(U8 && p.u8_int)) { COVA_STEM_SWITCH(true) } else { COVA_STEM_SWITCH(false) }
#undef COVA_STEM_SWITCH
#undef COVA_STEM_CASE
          const long long m1 = timed ? clock64() : 0;
          // tcgen05.commit tracks the issuing thread's MMAs only: tile t's barrier also needs tile t-1 (the other issuer's),
          // so every tile arrives on its own barrier and on its successor's
          ptx::umma_commit(&sm.mma_done[t % SX_ND]);
          ptx::umma_commit(&sm.mma_done[(t + 1) % SX_ND]);
          if (timed) { wt2 += m1 - m0; wt3 += clock64() - m1; }
        }
        __syncwarp();
      }
    }
    if (timed && lane == 0) {
      unsigned long long* d = p.dbg + (size_t)blockIdx.x * 8;
      atomicAdd(d + 0, (unsigned long long)wt0);
      atomicAdd(d + 1, (unsigned long long)wt1);
      atomicAdd(d + 2, (unsigned long long)wt2);
      atomicAdd(d + 3, (unsigned long long)wt3);
      atomicAdd(d + 4, (unsigned long long)(clock64() - t_start));
      atomicAdd(d + 5, (unsigned long long)t);
    }
  } else if (warp <= SX_EPI_WARPS) {
    // ======================= epilogue: BN + ReLU + 3x3/s2 max-pool =======================
    const int lg = warp & 3;                       // TMEM lane group this warp may read
    const int ch0 = ((warp - 1) >> 2) * SX_CH;     // this warp's output channels (SX_EPI_WARPS/4 warps per lane group)
    const int m = lg * 32 + lane;                  // conv column within the strip
    float acc_v[SX_CH];                            // running vertical max of the current pooling window
    double st_s = 0.0, st_q = 0.0;                 // raw mode: running statistics of channel ch0 + (lane & 15)
    uint32_t t = 0, n_emit = 0;
    for (int strip = 0; strip < n_strips; ++strip) {
      const int ox = strip * SX_TM + m;
      // ReLU is folded into the pooling: max(0, max_window(bn(x))) == max_window(relu(bn(x))), and 0 is also what
      // pool padding contributes after ReLU - so the running maximum simply starts from 0.
#pragma unroll
      for (int c = 0; c < SX_CH; ++c) acc_v[c] = 0.f;
      for (int i = 0; i < n_conv; ++i, ++t) {
        const int oy = oy_begin + i;
        const uint32_t acc = t % SX_NACC;
        ptx::mbar_wait(&sm.mma_done[t % SX_ND], (t / SX_ND) & 1);
        ptx::tc_fence_after();
        uint32_t raw[SX_CH], raw2[SX_CH];
        float v[SX_CH];
        const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + acc * ACC_COLS + ch0;
        ptx::tmem_ld16(taddr, *reinterpret_cast<uint32_t(*)[16]>(raw));
        if (SPLIT) ptx::tmem_ld16(taddr + 64, *reinterpret_cast<uint32_t(*)[16]>(raw2));   // columns 64..127 hold Ahi*Wlo
        ptx::tmem_ld_wait();                                                                // both loads in flight, one wait
#pragma unroll
        for (int j = 0; j < SX_CH; ++j) v[j] = SPLIT ? __uint_as_float(raw[j]) + __uint_as_float(raw2[j]) : __uint_as_float(raw[j]);
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&sm.tmem_empty[acc]);

        if (RAW) {
          // training mode (BatchNorm needs batch statistics of THIS tensor): write the raw conv row and skip the
          // pooling.  A band recomputes the conv row above it as pooling halo; only the owner band writes a row.
          if (p.stats != nullptr) {   // statistics of the values as they are STORED; every conv pixel is counted by its owner band
            static_assert(SX_CH == 16, "warp_transpose_sum16");
            const bool valid = oy >= 2 * py0 && ox < p.Wc;
            float z[16], z2[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) {
              float tv = (SPLIT && HALF) ? v[c] * (1.f / SPLIT_F16_WSCALE) : v[c];
              if (p.raw_bf16) tv = round_bf16(tv);
              z[c] = valid ? tv : 0.f;
              z2[c] = z[c] * z[c];
            }
            st_s += (double)warp_transpose_sum16(z, lane);
            st_q += (double)warp_transpose_sum16(z2, lane);
          }
          if (oy >= 2 * py0 && ox < p.Wc && p.raw_bf16) {
            __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.raw_out) + (((size_t)b * p.Hc + oy) * p.Wc + ox) * 64 + ch0;
#pragma unroll
            for (int hlf = 0; hlf < SX_CH / 8; ++hlf)
              *reinterpret_cast<uint4*>(dst + hlf * 8) = make_uint4(pack2_bf16(v[hlf * 8], v[hlf * 8 + 1]), pack2_bf16(v[hlf * 8 + 2], v[hlf * 8 + 3]),
                                                                    pack2_bf16(v[hlf * 8 + 4], v[hlf * 8 + 5]), pack2_bf16(v[hlf * 8 + 6], v[hlf * 8 + 7]));
          } else if (oy >= 2 * py0 && ox < p.Wc) {
            float* dst = p.raw_out + (((size_t)b * p.Hc + oy) * p.Wc + ox) * 64 + ch0;
#pragma unroll
            for (int hlf = 0; hlf < SX_CH / 8; ++hlf) {
              uint32_t w8[8];
#pragma unroll
              for (int e = 0; e < 8; ++e)
                w8[e] = __float_as_uint((SPLIT && HALF) ? v[hlf * 8 + e] * (1.f / SPLIT_F16_WSCALE) : v[hlf * 8 + e]);
              st_global_v8(dst + hlf * 8, w8);
            }
          }
          continue;
        }
        // BN (ReLU comes with the max).  The tensor core's operand fetch needs the whole 128 B/clk of shared memory while an
        // N = 128 MMA runs, so every LSU wavefront is a stolen MMA cycle: 128-bit loads (8 wavefronts per warp and tile, not 32).
#pragma unroll
        for (int c = 0; c < SX_CH; c += 4) {
          const float4 s4 = *reinterpret_cast<const float4*>(sm.scale + ch0 + c), h4 = *reinterpret_cast<const float4*>(sm.shift + ch0 + c);
          v[c] = fmaf(v[c], s4.x, h4.x); v[c + 1] = fmaf(v[c + 1], s4.y, h4.y);
          v[c + 2] = fmaf(v[c + 2], s4.z, h4.z); v[c + 3] = fmaf(v[c + 3], s4.w, h4.w);
        }
        if (ox >= p.Wc) {   // conv columns past the image (last, partial strip only) act as pool padding
#pragma unroll
          for (int c = 0; c < SX_CH; ++c) v[c] = 0.f;
        }
        // Pooling order: vertical first (a per-lane running maximum over the window's conv rows - no cross-lane
        // traffic), horizontal once per POOLED row.  max is associative, so this equals the 3x3 window maximum, and it
        // halves the shuffles, the lane-31 exchange and the epilogue barrier (they dominated the epilogue).
        const bool odd = oy & 1;
        const bool emit = odd || (oy == p.Hc - 1);                 // the conv row that completes pooled row py
        if (!emit) {
#pragma unroll
          for (int c = 0; c < SX_CH; ++c) acc_v[c] = fmaxf(acc_v[c], v[c]);
          continue;
        }
        float w[SX_CH];
#pragma unroll
        for (int c = 0; c < SX_CH; ++c) w[c] = fmaxf(acc_v[c], v[c]);        // column maximum of the window (>= 0: ReLU)
        // lane 31 publishes its column for the next warp (and, from the last lane group, for the next strip)
        const int py = oy >> 1;
        // exchange buffers alternate per EMITTED row (not per py: the last pooled row of a strip and the first of the next can
        // have the same parity, and nothing but the barrier of the row in between orders a write after the previous reads)
        const uint32_t xb = n_emit & 1u;
        ++n_emit;
        float* xrow = sm.xch[xb][lg] + ch0;
        if (lane == 31) {
#pragma unroll
          for (int c = 0; c < SX_CH; c += 4) *reinterpret_cast<float4*>(xrow + c) = make_float4(w[c], w[c + 1], w[c + 2], w[c + 3]);
          if (lg == 3) {
            float* e = sm.edge[strip % 3][i] + ch0;
#pragma unroll
            for (int c = 0; c < SX_CH; c += 4) *reinterpret_cast<float4*>(e + c) = make_float4(w[c], w[c + 1], w[c + 2], w[c + 3]);
          }
        }
        // only the four lane-group warps of one channel group exchange columns: one named barrier per channel group (a single
        // 512-thread barrier kept all 16 epilogue warps in lockstep, so none of their latencies overlapped)
        asm volatile("bar.sync %0, 128;" ::"r"(1 + ((warp - 1) >> 2)) : "memory");
        const float* left = (lg > 0 ? sm.xch[xb][lg - 1] : sm.edge[(strip + 2) % 3][i]) + ch0;
        const bool left_zero = (lg == 0 && strip == 0);            // conv column -1 = pool padding
        // Horizontal 3-max + store.  The 16 pooled pixels of this warp's 32 columns take all 32 lanes: the even lane of a column
        // pair produces channels 0..7 of the pixel, the odd lane channels 8..15 (with lane = column, the odd lanes had nothing to
        // write and the warp issued the whole split / pack / store sequence for half its lanes).  Window of pixel px = columns
        // 2px-1, 2px, 2px+1 = (L-1, L, L+1) seen from the even lane L and (L-2, L-1, L) seen from the odd lane L.
        static_assert(SX_CH == 16, "even / odd lanes split 16 channels");
        const bool odd_lane = lane & 1;
        const int px = ox >> 1;
        const bool writer = py >= py0 && py < py1 && px < p.Wp;
        const size_t opix = (((size_t)b * p.Hp + py) * p.Wp + px) * 64 + ch0 + (odd_lane ? 8 : 0);
        float lf[8];                                               // lanes 0 / 1: the column left of this warp's first one
        if (lane < 2) {
          if (left_zero) {
#pragma unroll
            for (int k = 0; k < 8; ++k) lf[k] = 0.f;
          } else {
            const float4 a4 = *reinterpret_cast<const float4*>(left + 8 * lane), b4 = *reinterpret_cast<const float4*>(left + 8 * lane + 4);
            lf[0] = a4.x; lf[1] = a4.y; lf[2] = a4.z; lf[3] = a4.w; lf[4] = b4.x; lf[5] = b4.y; lf[6] = b4.z; lf[7] = b4.w;
          }
        }
        float o[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float x = w[k], y = w[k + 8];
          float up1 = __shfl_up_sync(0xffffffffu, odd_lane ? x : y, 1);   // even lane: x[L-1]; odd lane: y[L-1]
          const float dn1 = __shfl_down_sync(0xffffffffu, x, 1);           // even lane: x[L+1]
          float up2 = __shfl_up_sync(0xffffffffu, y, 2);                   // odd lane: y[L-2]
          if (lane == 0) up1 = lf[k];
          if (lane == 1) up2 = lf[k];
          o[k] = odd_lane ? fmaxf(fmaxf(up2, up1), y) : fmaxf(fmaxf(up1, x), dn1);
        }
#pragma unroll
        for (int c = 0; c < SX_CH; ++c) acc_v[c] = fmaxf(v[c], 0.f);      // an odd conv row also opens pooled row py + 1
        if (writer) {
          if (OUT_DTYPE == COVA_F32) {
            uint32_t w8[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) w8[e] = __float_as_uint(o[e]);
            st_global_v8(reinterpret_cast<float*>(p.out0) + opix, w8);
          } else {
            uint32_t hw[4], lw[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              if (OUT_DTYPE == COVA_BF16X2 && HALF) split_f16x2(o[2 * e], o[2 * e + 1], hw[e], lw[e]);
              else if (OUT_DTYPE == COVA_BF16X2) split_bf16x2(o[2 * e], o[2 * e + 1], hw[e], lw[e]);
              else if (HALF) hw[e] = pack2_f16(o[2 * e], o[2 * e + 1]);
              else hw[e] = pack2_bf16(o[2 * e], o[2 * e + 1]);
            }
            *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out0) + opix) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
            if (OUT_DTYPE == COVA_BF16X2)
              *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out1) + opix) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
          }
        }
      }
    }
    if (RAW && p.stats != nullptr && lane < 16) {
      atomicAdd(p.stats + ch0 + lane, st_s);
      atomicAdd(p.stats + 64 + ch0 + lane, st_q);
    }
  } else if (warp <= SX_EPI_WARPS + NCV) {
    // ======================= converters: NCHW fp32 rows -> ring of 4-channel bf16 pixels =======================
    const int cw = warp - 1 - SX_EPI_WARPS;         // this warp owns input rows g with g % NCV == cw
    const size_t plane = (size_t)p.H * p.W;
    const size_t img_b = (size_t)b * 3 * plane;
    const uint32_t n_rows_total = (uint32_t)n_strips * NQ;
    for (uint32_t g = cw; g < n_rows_total; g += NCV) {
      const int strip = g / NQ, q = g % NQ;
      const int y = y_first + q;
      const int x0 = 2 * strip * SX_TM - 3;
      if (g >= SX_R) {   // the previous occupant of this slot must have been consumed by its last conv row
        const uint32_t gp = g - SX_R;
        const int sp = gp / NQ, qp = gp % NQ;
        const uint32_t t_last = (uint32_t)sp * n_conv + min(qp >> 1, n_conv - 1);
        ptx::mbar_wait(&sm.mma_done[t_last % SX_ND], (t_last / SX_ND) & 1);
      }
      const bool row_ok = y >= 0 && y < p.H;
      unsigned char* dst_hi = sm.ring[0][g % SX_R];
      if (p.pf_rows > 0) {   // warm L2 for a later row of this warp: one 128-byte line per lane (3 channels x <= 10 lines)
        const uint32_t gn = g + (uint32_t)(NCV * p.pf_rows);
        if (gn < n_rows_total) {
          const int yn = y_first + (int)(gn % NQ);
          const int xn = 2 * (int)(gn / NQ) * SX_TM - 3;
          const int esz = U8 ? 1 : 4, per_line = 128 / esz;
          const int c = lane / 10, j = lane % 10;
          const int x = xn + j * per_line;
          if (c < 3 && yn >= 0 && yn < p.H && x < p.W && x + per_line > 0 && j * per_line < SX_NPX + per_line) {
            const size_t idx = img_b + c * plane + (size_t)yn * p.W + (x < 0 ? 0 : x);
            ptx::prefetch_l2(reinterpret_cast<const unsigned char*>(p.img) + idx * esz);
          }
        }
      }
      if (U8 && (p.W & 3) == 0) {
        // uint8 fast path: x0 - 1 is a multiple of 4, so the row segment is 67 aligned 4-pixel words per channel
        // (9 word loads per lane instead of 27 byte loads); the value -> (hi, lo) bf16 split comes from the LUT.
        const unsigned char* img8 = reinterpret_cast<const unsigned char*>(p.img) + img_b;
        uint32_t wv[3][3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const int wi = lane + 32 * j, xw = x0 - 1 + 4 * wi;
          const bool ok = row_ok && wi < 67 && xw >= 0 && xw < p.W;
#pragma unroll
          for (int c = 0; c < 3; ++c)
            wv[j][c] = ok ? __ldg(reinterpret_cast<const uint32_t*>(img8 + c * plane + (size_t)y * p.W + xw)) : 0u;
        }
        // Word wi holds ring pixels 4wi-1 .. 4wi+2; a lane STORES pixels 4wi-2 .. 4wi+1 (the first one is the neighbour word's
        // last pixel, fetched by shuffle) as two aligned 16-byte pairs - see stem_store_pairs.
        const bool u8i = SPLIT && p.u8_int;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const int wi = lane + 32 * j;
          const uint32_t last = (wv[j][0] >> 24) | ((wv[j][1] >> 24) << 8) | ((wv[j][2] >> 24) << 16);   // pixel 4wi+2, 3 channels
          uint32_t prev = __shfl_up_sync(0xffffffffu, last, 1);
          if (j > 0) {
            const uint32_t plast = (wv[j > 0 ? j - 1 : 0][0] >> 24) | ((wv[j > 0 ? j - 1 : 0][1] >> 24) << 8) | ((wv[j > 0 ? j - 1 : 0][2] >> 24) << 16);
            const uint32_t carry = __shfl_sync(0xffffffffu, plast, 31);
            if (lane == 0) prev = carry;
          }
          uint2 hi[4], lo[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) {                   // t = 0: pixel 4wi-2 (from prev), t = 1..3: own bytes 0..2
            const uint32_t b0 = t == 0 ? (prev & 255u) : ((wv[j][0] >> (8 * (t - 1))) & 255u);
            const uint32_t b1 = t == 0 ? ((prev >> 8) & 255u) : ((wv[j][1] >> (8 * (t - 1))) & 255u);
            const uint32_t b2 = t == 0 ? ((prev >> 16) & 255u) : ((wv[j][2] >> (8 * (t - 1))) & 255u);
            if (u8i) {
              // integer mode: the pixel value itself, exact in 16 bits, by arithmetic (0x4B000000 | v is the float 2^23 + v)
              const float f0 = __uint_as_float(0x4B000000u | b0) - 8388608.f, f1 = __uint_as_float(0x4B000000u | b1) - 8388608.f,
                          f2 = __uint_as_float(0x4B000000u | b2) - 8388608.f;
              hi[t] = HALF ? make_uint2(pack2_f16(f0, f1), pack2_f16(f2, 0.f)) : make_uint2(pack2_bf16(f0, f1), pack2_bf16(f2, 0.f));
              lo[t] = make_uint2(0u, 0u);
            } else {
              const uint32_t l0 = sm.lut[b0], l1 = sm.lut[b1], l2 = sm.lut[b2];
              hi[t] = make_uint2(__byte_perm(l0, l1, 0x5410), l2 & 0xffffu);
              lo[t] = make_uint2(__byte_perm(l0, l1, 0x7632), l2 >> 16);
            }
          }
          stem_store_pairs<SPLIT>(dst_hi, wi, lane, hi, lo, SPLIT && !u8i);
        }
      } else if (!U8 && (p.W & 3) == 0 && (reinterpret_cast<uintptr_t>(p.img) & 15) == 0) {
        // fp32 fast path: x0 - 1 is a multiple of 4, so the row segment is 67 aligned float4 per channel
        // (9 x LDG.128 per lane instead of 27 x LDG.32 - the converters were what the MMA issuer waited for)
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
          float pv[3];                                    // pixel 4wi-2 = the neighbour word's last pixel
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            pv[c] = __shfl_up_sync(0xffffffffu, fv[j][c].w, 1);
            if (j > 0) {
              const float carry = __shfl_sync(0xffffffffu, fv[j > 0 ? j - 1 : 0][c].w, 31);
              if (lane == 0) pv[c] = carry;
            }
          }
          uint2 hi[4], lo[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float e0 = t == 0 ? pv[0] : (&fv[j][0].x)[t - 1], e1 = t == 0 ? pv[1] : (&fv[j][1].x)[t - 1],
                        e2 = t == 0 ? pv[2] : (&fv[j][2].x)[t - 1];
            uint32_t h01, l01, h2, l2;
            if (HALF && SPLIT) {
              split_f16x2(e0, e1, h01, l01);
              split_f16x2(e2, 0.f, h2, l2);
            } else if (HALF) {
              h01 = pack2_f16(e0, e1); h2 = pack2_f16(e2, 0.f); l01 = l2 = 0u;
            } else {
              split_bf16x2(e0, e1, h01, l01);
              split_bf16x2(e2, 0.f, h2, l2);
            }
            hi[t] = make_uint2(h01, h2);
            lo[t] = make_uint2(l01, l2);
          }
          stem_store_pairs<SPLIT>(dst_hi, wi, lane, hi, lo, SPLIT);
        }
      } else {
        float f[9][3];
#pragma unroll
        for (int j = 0; j < 9; ++j) {
          const int i = lane + 32 * j, x = x0 + i;
          const bool ok = row_ok && i < SX_NPX && x >= 0 && x < p.W;
#pragma unroll
          for (int c = 0; c < 3; ++c) f[j][c] = ok ? load_pixel<U8>(p.img, img_b + c * plane + (size_t)y * p.W + x) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < 9; ++j) {
          const int i = lane + 32 * j;
          if (i < SX_NPX) {
            uint32_t h01, l01, h2, l2;
            if (HALF && SPLIT) {
              split_f16x2(f[j][0], f[j][1], h01, l01);
              split_f16x2(f[j][2], 0.f, h2, l2);
            } else if (HALF) {
              h01 = pack2_f16(f[j][0], f[j][1]); h2 = pack2_f16(f[j][2], 0.f); l01 = l2 = 0u;
            } else {
              split_bf16x2(f[j][0], f[j][1], h01, l01);
              split_bf16x2(f[j][2], 0.f, h2, l2);
            }
            *reinterpret_cast<uint2*>(dst_hi + i * 8) = make_uint2(h01, h2);
            if (SPLIT) *reinterpret_cast<uint2*>(dst_hi + SX_R * SX_ROW_BYTES + i * 8) = make_uint2(l01, l2);
          }
        }
      }
      ptx::fence_proxy_async();      // generic-proxy writes -> visible to tcgen05 (async proxy) reads
      __syncwarp();
      if (lane == 0) {
        ptx::mbar_arrive(&sm.in_full[g % SX_R]);
        if (q >= 5) {                                  // first needed by conv row (q - 5) / 2 of this strip
          const uint32_t tq = (uint32_t)strip * n_conv + ((q - 5) >> 1);
          ptx::mbar_arrive(&sm.pair_full[tq % SX_ND]);
        }
      }
    }
  }

teardown:
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// OIHW fp32 [64,3,7,7] -> [28 chunks][plane][64 cout][8] bf16, K index = r*32 + s*4 + c (s = 7 and c = 3 are zero)
__global__ void pack_stem_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= SX_KCHUNKS * 64 * 8) return;
  const int e = i & 7, co = (i >> 3) & 63, chunk = i >> 9;
  const int r = chunk >> 2, s = (chunk & 3) * 2 + (e >> 2), c = e & 3;
  float v = 0.f;
  if (s < 7 && c < 3) v = w[((co * 3 + c) * 7 + r) * 7 + s];
  __nv_bfloat16 h, l;
  split_bf16(v, h, l);
  out[((chunk * 2 + 0) * 64 + co) * 8 + e] = h;
  out[((chunk * 2 + 1) * 64 + co) * 8 + e] = l;
}

// same layout, plane 0 = fp16(w), plane 1 = 0 (single-plane fp16 mode)
__global__ void pack_stem_weight_f16_kernel(const float* __restrict__ w, __half* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= SX_KCHUNKS * 64 * 8) return;
  const int e = i & 7, co = (i >> 3) & 63, chunk = i >> 9;
  const int r = chunk >> 2, s = (chunk & 3) * 2 + (e >> 2), c = e & 3;
  float v = 0.f;
  if (s < 7 && c < 3) v = w[((co * 3 + c) * 7 + r) * 7 + s];
  out[((chunk * 2 + 0) * 64 + co) * 8 + e] = __float2half_rn(v);
  out[((chunk * 2 + 1) * 64 + co) * 8 + e] = __float2half_rn(0.f);
}

// split-fp16 filter: planes hi = f16(256 w), lo = f16(256 w - hi)
__global__ void pack_stem_weight_f16x2_kernel(const float* __restrict__ w, __half* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= SX_KCHUNKS * 64 * 8) return;
  const int e = i & 7, co = (i >> 3) & 63, chunk = i >> 9;
  const int r = chunk >> 2, s = (chunk & 3) * 2 + (e >> 2), c = e & 3;
  float v = 0.f;
  if (s < 7 && c < 3) v = w[((co * 3 + c) * 7 + r) * 7 + s] * SPLIT_F16_WSCALE;
  const __half h = __float2half_rn(v);
  out[((chunk * 2 + 0) * 64 + co) * 8 + e] = h;
  out[((chunk * 2 + 1) * 64 + co) * 8 + e] = __float2half_rn(v - __half2float(h));
}

template <bool SPLIT, int OUT_DTYPE, bool U8, int NCV, bool HALF, bool RAW = false>
static int launch_stem_tc_n(const StemTcParams& p, int grid, cudaStream_t st) {
  auto kern = stem_tc_kernel<SPLIT, OUT_DTYPE, U8, NCV, HALF, RAW>;
  const int smem = (int)sizeof(StemTcSmem<SPLIT>) + 128;
  COVA_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  kern<<<grid, sx_threads(NCV), smem, st>>>(p);
  COVA_LAUNCH_OK();
  return COVA_OK;
}
template <bool SPLIT, int OUT_DTYPE, bool U8, bool HALF = false>
static int launch_stem_tc(const StemTcParams& p, int grid, cudaStream_t st) {
  if (p.raw_out != nullptr) {
    if constexpr (OUT_DTYPE == COVA_F32 || (OUT_DTYPE == COVA_BF16 && !HALF))
      return launch_stem_tc_n<SPLIT, OUT_DTYPE, U8, 8, HALF, true>(p, grid, st);
    set_error("stem_tc: raw (training) output is fp32 or one bf16 plane");
    return COVA_ERR_ARG;
  }
  if (knob(COVA_KNOB_STEM_CONVERTERS, 8) >= 8) return launch_stem_tc_n<SPLIT, OUT_DTYPE, U8, 8, HALF>(p, grid, st);
  return launch_stem_tc_n<SPLIT, OUT_DTYPE, U8, 4, HALF>(p, grid, st);
}

static int stem_tc_impl(const void* images, int img_u8, int B, int H, int W, const void* w_packed, const float* bn_scale,
                        const float* bn_shift, int out_dtype, void* out0, void* out1, cudaStream_t st, float* raw_out,
                        bool split_f16 = false, bool raw_bf16 = false, double* stats = nullptr) {
  StemTcParams p;
  p.raw_out = raw_out;
  p.raw_bf16 = raw_bf16 ? 1 : 0;
  p.stats = raw_out ? stats : nullptr;
  p.u8_int = (img_u8 && !raw_out && (W & 3) == 0 && knob(COVA_KNOB_STEM_U8_EXACT, 0) == 0) ? 1 : 0;   // (the LUT fast path only)
  if (p.stats) COVA_CUDA_OK(cudaMemsetAsync(p.stats, 0, 2 * 64 * sizeof(double), st));
  p.img = images; p.B = B; p.H = H; p.W = W;
  p.Hc = (H + 6 - 7) / 2 + 1; p.Wc = (W + 6 - 7) / 2 + 1;
  p.Hp = (p.Hc + 2 - 3) / 2 + 1; p.Wp = (p.Wc + 2 - 3) / 2 + 1;
  int bands = sm_count() / B;
  if (bands < 1) bands = 1;
  if (bands > p.Hp) bands = p.Hp;
  int nb = ceil_div(p.Hp, bands);
  const int nb_max = (SX_MAXROWS - 1) / 2;
  if (nb > nb_max) nb = nb_max;
  p.nb = nb;
  p.bands_per_page = ceil_div(p.Hp, nb);
  p.w_packed = (const unsigned char*)w_packed;
  p.bn_scale = bn_scale; p.bn_shift = bn_shift;
  p.out0 = out0; p.out1 = out1;
  p.pf_rows = knob(COVA_KNOB_STEM_L2_PREFETCH, 1);
  p.dbg = debug_words(8LL * B * p.bands_per_page);
  const int grid = B * p.bands_per_page;
  if (out_dtype == COVA_F16)   // fp16 mode: the filter comes from cova_pack_stem_weight_f16
    return img_u8 ? launch_stem_tc<false, COVA_BF16, true, true>(p, grid, st)
                  : launch_stem_tc<false, COVA_BF16, false, true>(p, grid, st);
  if (out_dtype == COVA_F16X2 || split_f16)   // split-fp16 mode: the filter comes from cova_pack_stem_weight_f16x2
    return out_dtype == COVA_F32
               ? (img_u8 ? launch_stem_tc<true, COVA_F32, true, true>(p, grid, st)
                         : launch_stem_tc<true, COVA_F32, false, true>(p, grid, st))
               : (img_u8 ? launch_stem_tc<true, COVA_BF16X2, true, true>(p, grid, st)
                         : launch_stem_tc<true, COVA_BF16X2, false, true>(p, grid, st));
  const bool split = out_dtype != COVA_BF16;   // bf16 output <=> bf16 mode; fp32 / split outputs use the 3-product mode
#define GO(SP, DT) (img_u8 ? launch_stem_tc<SP, DT, true>(p, grid, st) : launch_stem_tc<SP, DT, false>(p, grid, st))
  if (split) {
    if (out_dtype == COVA_F32) return GO(true, COVA_F32);
    return GO(true, COVA_BF16X2);
  }
  return GO(false, COVA_BF16);
#undef GO
}

int stem_tc(const void* images, int img_u8, int B, int H, int W, const void* w_packed, const float* bn_scale,
            const float* bn_shift, int out_dtype, void* out0, void* out1, cudaStream_t st) {
  return stem_tc_impl(images, img_u8, B, H, W, w_packed, bn_scale, bn_shift, out_dtype, out0, out1, st, nullptr);
}

}  // namespace cova

extern "C" int cova_pack_stem_weight(const float* w_oihw, void* packed, void* stream) {
  COVA_REQUIRE(w_oihw && packed, "cova_pack_stem_weight: null pointer");
  cova::pack_stem_weight_kernel<<<cova::ceil_div(cova::SX_KCHUNKS * 64 * 8, 256), 256, 0, (cudaStream_t)stream>>>(
      w_oihw, (__nv_bfloat16*)packed);
  COVA_LAUNCH_OK();
  return COVA_OK;
}

extern "C" int cova_pack_stem_weight_f16(const float* w_oihw, void* packed, void* stream) {
  COVA_REQUIRE(w_oihw && packed, "cova_pack_stem_weight_f16: null pointer");
  cova::pack_stem_weight_f16_kernel<<<cova::ceil_div(cova::SX_KCHUNKS * 64 * 8, 256), 256, 0, (cudaStream_t)stream>>>(
      w_oihw, (__half*)packed);
  COVA_LAUNCH_OK();
  return COVA_OK;
}

// A2, training mode: conv1 alone (7x7 s2 p3, fp32-parity tensor-core mode) -> [B, H/2, W/2, 64] fp32 NHWC, no BN, no
// ReLU, no pooling: nn.BatchNorm2d in train mode needs the batch statistics of this tensor (`train.py:27`).
extern "C" int cova_stem_conv_raw_fwd(const void* images, int img_dtype, int B, int H, int W, const void* w_packed,
                                      int w_dtype, float* out, void* stream) {
  COVA_REQUIRE(w_dtype == COVA_BF16X2 || w_dtype == COVA_F16X2, "cova_stem_conv_raw_fwd: w_dtype is the packed filter's format");
  COVA_REQUIRE(images && w_packed && out && B > 0 && H >= 7 && W >= 7, "cova_stem_conv_raw_fwd: bad arguments");
  COVA_REQUIRE(img_dtype == COVA_F32 || img_dtype == COVA_U8, "cova_stem_conv_raw_fwd: images must be fp32 or uint8");
  COVA_REQUIRE(((uintptr_t)out & 31) == 0, "cova_stem_conv_raw_fwd: out must be 32-byte aligned");
  return cova::stem_tc_impl(images, img_dtype == COVA_U8, B, H, W, w_packed, nullptr, nullptr, COVA_F32, out, nullptr,
                            (cudaStream_t)stream, out, w_dtype == COVA_F16X2);
}

// The raw conv1 entry points with the BatchNorm batch statistics of the output accumulated by the epilogue
// (stats_ws: [2][64] doubles = sum y, sum y^2, zeroed here; feeds cova_bn_train_finalize directly - no separate statistics pass)
extern "C" int cova_stem_conv_raw_stats_fwd(const void* images, int img_dtype, int B, int H, int W, const void* w_packed,
                                            int w_dtype, void* out, double* stats_ws, void* stream) {
  COVA_REQUIRE(w_dtype == COVA_BF16X2 || w_dtype == COVA_F16X2 || w_dtype == COVA_BF16,
               "cova_stem_conv_raw_stats_fwd: w_dtype is the packed filter's format (COVA_BF16 = one bf16 product, bf16 output)");
  COVA_REQUIRE(images && w_packed && out && B > 0 && H >= 7 && W >= 7, "cova_stem_conv_raw_stats_fwd: bad arguments");
  COVA_REQUIRE(img_dtype == COVA_F32 || img_dtype == COVA_U8, "cova_stem_conv_raw_stats_fwd: images must be fp32 or uint8");
  COVA_REQUIRE(((uintptr_t)out & 31) == 0, "cova_stem_conv_raw_stats_fwd: out must be 32-byte aligned");
  if (w_dtype == COVA_BF16)
    return cova::stem_tc_impl(images, img_dtype == COVA_U8, B, H, W, w_packed, nullptr, nullptr, COVA_BF16, out, nullptr,
                              (cudaStream_t)stream, reinterpret_cast<float*>(out), false, true, stats_ws);
  return cova::stem_tc_impl(images, img_dtype == COVA_U8, B, H, W, w_packed, nullptr, nullptr, COVA_F32, out, nullptr,
                            (cudaStream_t)stream, reinterpret_cast<float*>(out), w_dtype == COVA_F16X2, false, stats_ws);
}

// bf16 training mode: conv1 in one bf16 product (filter from cova_pack_stem_weight: its hi plane), raw output stored as bf16
extern "C" int cova_stem_conv_raw_fwd_bf16(const void* images, int img_dtype, int B, int H, int W, const void* w_packed,
                                           void* out_bf16, void* stream) {
  COVA_REQUIRE(images && w_packed && out_bf16 && B > 0 && H >= 7 && W >= 7, "cova_stem_conv_raw_fwd_bf16: bad arguments");
  COVA_REQUIRE(img_dtype == COVA_F32 || img_dtype == COVA_U8, "cova_stem_conv_raw_fwd_bf16: images must be fp32 or uint8");
  COVA_REQUIRE(((uintptr_t)out_bf16 & 15) == 0, "cova_stem_conv_raw_fwd_bf16: out must be 16-byte aligned");
  return cova::stem_tc_impl(images, img_dtype == COVA_U8, B, H, W, w_packed, nullptr, nullptr, COVA_BF16, out_bf16, nullptr,
                            (cudaStream_t)stream, reinterpret_cast<float*>(out_bf16), false, true);
}

extern "C" int cova_pack_stem_weight_f16x2(const float* w_oihw, void* packed, void* stream) {
  COVA_REQUIRE(w_oihw && packed, "cova_pack_stem_weight_f16x2: null pointer");
  cova::pack_stem_weight_f16x2_kernel<<<cova::ceil_div(cova::SX_KCHUNKS * 64 * 8, 256), 256, 0, (cudaStream_t)stream>>>(
      w_oihw, (__half*)packed);
  COVA_LAUNCH_OK();
  return COVA_OK;
}
