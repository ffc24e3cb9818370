// ResNet-50 Bottleneck 1x1 convolutions on the tensor cores (torchvision `Bottleneck.forward`
// conv1 / conv3 / downsample + folded BN (+ residual) (+ ReLU)), fp32-parity split-bf16 arithmetic.
//
// A 1x1 convolution over NHWC activations is the GEMM  Y[M, Cout] = X[M, Cin] W[Cout, Cin]^T  with
// M = B*H*W pixel rows (1.6 M rows at B = 16): tiny K and N, enormous M - the opposite of linear_tc.cu's shape,
// so this kernel is persistent (one CTA per SM walks 128-row tiles) and keeps the whole split filter resident in
// shared memory (<= 64 KB).  Activations are split-bf16 planes read by TMA ({64 ch, 128 rows} boxes, 128-B
// swizzle); products as in conv_tc.cu: Cout = 64 -> Ahi x [Whi;Wlo] as one N = 128 MMA + Alo x Whi (N = 64);
// Cout = 256 -> three N = 256 MMAs into one 256-column accumulator.  TMEM accumulators are double buffered;
// eight epilogue warps apply BN (+ residual planes) (+ ReLU) and write split planes or fp32 rows with 32-byte
// stores, 32 channels at a time.
//
// Bound: HBM.  Algorithmic bytes per pixel (split planes, 4 B/elem): 4*(Cin + Cout) (+ 4*Cout residual);
// MMA time is 448 / 1536 / 1792 clk per 128 pixels for 64->64 / 64->256 / 256->64 against ~2800-12500 clk of HBM time.
#include "common.cuh"
#include "ptx.cuh"
#include "tma_host.cuh"

namespace cova {

constexpr int PW_BM = 128;
constexpr int PW_A_PLANE = PW_BM * 128;            // 16 KB: 128 rows x 64 channels bf16
constexpr int PW_STAGE = 2 * PW_A_PLANE;           // hi + lo
constexpr int PW_NSTAGE = 4;
constexpr int PW_THREADS = 320;                    // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue

struct PwTail {
  uint64_t full[PW_NSTAGE], empty[PW_NSTAGE], tmem_full[2], tmem_empty[2], wbar;
  uint32_t tmem_base;
};

struct PwParams {
  int M, n_tiles, relu;
  const float* bn_scale;
  const float* bn_shift;
  const __nv_bfloat16* res_hi;
  const __nv_bfloat16* res_lo;
  void* y0;
  void* y1;
  double* stats;     // optional [2][COUT]: sum y, sum y^2 over all pixel rows (BatchNorm batch statistics of the output), += here
  const float* res_f32;   // optional fp32 residual rows [M, COUT] added to an fp32 output (training dgrad + skip gradient)
};

template <int KB, int COUT, bool SINGLE = false>
struct PwCfg {
  static constexpr int W_KB_BYTES = (SINGLE ? 1 : 2) * COUT * 128;   // one k-block of the filter: hi rows then lo rows
  static constexpr int W_BYTES = KB * W_KB_BYTES;
  static constexpr int ACC_COLS = COUT == 64 ? (SINGLE ? 64 : 128) : 256;
  static constexpr int TMEM_COLS = 2 * ACC_COLS;
  static constexpr int SMEM = W_BYTES + PW_NSTAGE * PW_STAGE + 2 * COUT * 4 + 1024 + 1024;
};

// HALF: split-fp16 planes and filter (training path: forward and dgrad in the three-product fp16 mode; the filter comes
// scaled by SPLIT_F16_WSCALE, which the caller folds into bn_scale); fp32 output, no residual.
// SINGLE (bf16 training mode): one bf16 plane in, one bf16 product, one bf16 plane out (OUT_DTYPE = COVA_BF16): half the
// bytes of this HBM-bound kernel; the lo maps are not touched.
template <int KB, int COUT, int OUT_DTYPE, bool HALF = false, bool SINGLE = false, bool RES32 = false>
__global__ void __launch_bounds__(PW_THREADS, 1)
pw_tc_kernel(const __grid_constant__ CUtensorMap tm_x_hi, const __grid_constant__ CUtensorMap tm_x_lo,
             const __grid_constant__ CUtensorMap tm_w_hi, const __grid_constant__ CUtensorMap tm_w_lo, const PwParams p) {
  using Cfg = PwCfg<KB, COUT, SINGLE>;
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  unsigned char* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  unsigned char* sm_w = smem;
  unsigned char* sm_a = smem + Cfg::W_BYTES;
  float* sm_scale = reinterpret_cast<float*>(sm_a + PW_NSTAGE * PW_STAGE);
  float* sm_shift = sm_scale + COUT;
  PwTail& tail = *reinterpret_cast<PwTail*>(sm_shift + COUT);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < COUT; i += PW_THREADS) {
    sm_scale[i] = p.bn_scale[i];
    sm_shift[i] = p.bn_shift[i];
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < PW_NSTAGE; ++i) {
      ptx::mbar_init(&tail.full[i], 1);
      ptx::mbar_init(&tail.empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tail.tmem_full[i], 1);
      ptx::mbar_init(&tail.tmem_empty[i], 256);
    }
    ptx::mbar_init(&tail.wbar, 1);
    ptx::fence_barrier_init();
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
    // ---------------- TMA producer ----------------
    if (ptx::elect_one()) {
      ptx::mbar_arrive_expect_tx(&tail.wbar, Cfg::W_BYTES);
      for (int kb = 0; kb < KB; ++kb) {
        ptx::tma_load_2d(sm_w + kb * Cfg::W_KB_BYTES, &tm_w_hi, &tail.wbar, kb * 64, 0);
        if (!SINGLE) ptx::tma_load_2d(sm_w + kb * Cfg::W_KB_BYTES + COUT * 128, &tm_w_lo, &tail.wbar, kb * 64, 0);
      }
    }
    __syncwarp();
    uint32_t stage = 0, phase = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      for (int kb = 0; kb < KB; ++kb) {
        ptx::mbar_wait(&tail.empty[stage], phase ^ 1);
        if (ptx::elect_one()) {
          unsigned char* dst = sm_a + stage * PW_STAGE;
          ptx::mbar_arrive_expect_tx(&tail.full[stage], SINGLE ? PW_A_PLANE : PW_STAGE);
          ptx::tma_load_2d(dst, &tm_x_hi, &tail.full[stage], kb * 64, tile * PW_BM);
          if (!SINGLE) ptx::tma_load_2d(dst + PW_A_PLANE, &tm_x_lo, &tail.full[stage], kb * 64, tile * PW_BM);
        }
        __syncwarp();
        if (++stage == PW_NSTAGE) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    constexpr uint32_t idesc_n = HALF ? ptx::umma_idesc_f16(128, COUT) : ptx::umma_idesc_bf16(128, COUT);
    constexpr uint32_t idesc_2n = HALF ? ptx::umma_idesc_f16(128, COUT == 64 ? 128 : 256)
                                       : ptx::umma_idesc_bf16(128, COUT == 64 ? 128 : 256);
    const uint64_t da0 = ptx::umma_desc_sw128(ptx::smem_u32(sm_a), 1024);
    const uint64_t db0 = ptx::umma_desc_sw128(ptx::smem_u32(sm_w), 1024);
    ptx::mbar_wait(&tail.wbar, 0);
    uint32_t stage = 0, phase = 0, it = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
      const uint32_t acc = it & 1;
      ptx::mbar_wait(&tail.tmem_empty[acc], ((it >> 1) & 1) ^ 1);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * Cfg::ACC_COLS;
#pragma unroll
      for (int kb = 0; kb < KB; ++kb) {
        ptx::mbar_wait(&tail.full[stage], phase);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const uint64_t da_s = da0 + ((stage * PW_STAGE) >> 4);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t a_hi = da_s + ((kk * 32) >> 4), a_lo = a_hi + (PW_A_PLANE >> 4);
            const uint64_t b_hi = db0 + ((kb * Cfg::W_KB_BYTES + kk * 32) >> 4), b_lo = b_hi + ((COUT * 128) >> 4);
            if (SINGLE) {
              ptx::umma_bf16(d_tmem, a_hi, b_hi, idesc_n, (kb | kk) != 0);
            } else if (COUT == 64) {
              ptx::umma_bf16(d_tmem, a_hi, b_hi, idesc_2n, (kb | kk) != 0);    // Ahi x [Whi; Wlo] -> 128 columns
              ptx::umma_bf16(d_tmem, a_lo, b_hi, idesc_n, 1);                  // Alo x Whi -> first 64 columns
            } else {
              ptx::umma_bf16(d_tmem, a_hi, b_hi, idesc_n, (kb | kk) != 0);
              ptx::umma_bf16(d_tmem, a_lo, b_hi, idesc_n, 1);
              ptx::umma_bf16(d_tmem, a_hi, b_lo, idesc_n, 1);
            }
          }
          ptx::umma_commit(&tail.empty[stage]);
          if (kb == KB - 1) ptx::umma_commit(&tail.tmem_full[acc]);
        }
        __syncwarp();
        if (++stage == PW_NSTAGE) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ---------------- epilogue: lane group = warp % 4, channel half = (warp-2)/4, 32 channels per pass ----------------
    const int lg = warp & 3;
    const int cbase = ((warp - 2) >> 2) * (COUT / 2);
    constexpr int NCHUNK = COUT / 64;              // 32-channel passes per warp
    const bool has_res = p.res_hi != nullptr;
    double st_s[NCHUNK], st_q[NCHUNK];             // running statistics of this lane's channel (cbase + ch * 32 + lane)
#pragma unroll
    for (int ch = 0; ch < NCHUNK; ++ch) st_s[ch] = st_q[ch] = 0.0;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
      const uint32_t acc = it & 1;
      const int m = tile * PW_BM + lg * 32 + lane;
      const bool inb = m < p.M;
      const size_t row = (size_t)m * COUT;
      ptx::mbar_wait(&tail.tmem_full[acc], (it >> 1) & 1);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + acc * Cfg::ACC_COLS;
#pragma unroll
      for (int ch = 0; ch < NCHUNK; ++ch) {
        const int c0 = cbase + ch * 32;
        uint32_t rh[2][8], rl[2][8];
        if (has_res && inb) {
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            ld_global_nc_v8(p.res_hi + row + c0 + j * 16, rh[j]);
            if (!SINGLE) ld_global_nc_v8(p.res_lo + row + c0 + j * 16, rl[j]);
          }
        }
        uint32_t v[2][16];
        float o[32];
#pragma unroll
        for (int q = 0; q < 2; ++q) ptx::tmem_ld16(taddr + c0 + q * 16, v[q]);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
          for (int j = 0; j < 16; ++j) o[q * 16 + j] = __uint_as_float(v[q][j]);
        if (COUT == 64 && !SINGLE) {   // columns 64..127 hold Ahi*Wlo
#pragma unroll
          for (int q = 0; q < 2; ++q) ptx::tmem_ld16(taddr + 64 + c0 + q * 16, v[q]);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int j = 0; j < 16; ++j) o[q * 16 + j] += __uint_as_float(v[q][j]);
        }
        if (ch == NCHUNK - 1) {
          ptx::tc_fence_before();
          ptx::mbar_arrive(&tail.tmem_empty[acc]);
        }
        if (p.stats != nullptr) {   // raw mode (no residual / ReLU): statistics of the values as they are STORED (warp-collective)
          float z[32], z2[32];
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            float t = fmaf(o[c], sm_scale[c0 + c], sm_shift[c0 + c]);
            if (OUT_DTYPE == COVA_BF16) t = round_bf16(t);
            z[c] = inb ? t : 0.f;
            z2[c] = z[c] * z[c];
          }
          st_s[ch] += (double)warp_transpose_sum32(z, lane);
          st_q[ch] += (double)warp_transpose_sum32(z2, lane);
        }
        if (!inb) continue;
#pragma unroll
        for (int c = 0; c < 32; ++c) o[c] = fmaf(o[c], sm_scale[c0 + c], sm_shift[c0 + c]);
        if (has_res) {
#pragma unroll
          for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              if (SINGLE) {     // bf16 training mode: the residual (e.g. the skip branch's gradient) is one bf16 plane
                o[j * 16 + 2 * e] += bf16lo_to_f32(rh[j][e]);
                o[j * 16 + 2 * e + 1] += bf16hi_to_f32(rh[j][e]);
              } else {
                o[j * 16 + 2 * e] += bf16lo_to_f32(rh[j][e]) + bf16lo_to_f32(rl[j][e]);
                o[j * 16 + 2 * e + 1] += bf16hi_to_f32(rh[j][e]) + bf16hi_to_f32(rl[j][e]);
              }
            }
        }
        if (RES32) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint32_t rf[8];
            ld_global_na_v8(p.res_f32 + row + c0 + j * 8, rf);
#pragma unroll
            for (int e = 0; e < 8; ++e) o[j * 8 + e] += __uint_as_float(rf[e]);
          }
        }
        if (p.relu) {
#pragma unroll
          for (int c = 0; c < 32; ++c) o[c] = fmaxf(o[c], 0.f);
        }
        if (OUT_DTYPE == COVA_F32) {
          float* dst = reinterpret_cast<float*>(p.y0) + row + c0;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint32_t w8[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) w8[e] = __float_as_uint(o[j * 8 + e]);
            st_global_v8(dst + j * 8, w8);
          }
        } else if (OUT_DTYPE == COVA_BF16) {
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            uint32_t hw[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) hw[e] = pack2_bf16(o[j * 16 + 2 * e], o[j * 16 + 2 * e + 1]);
            st_global_v8(reinterpret_cast<__nv_bfloat16*>(p.y0) + row + c0 + j * 16, hw);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            uint32_t hw[8], lw[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) split_bf16x2(o[j * 16 + 2 * e], o[j * 16 + 2 * e + 1], hw[e], lw[e]);
            st_global_v8(reinterpret_cast<__nv_bfloat16*>(p.y0) + row + c0 + j * 16, hw);
            st_global_v8(reinterpret_cast<__nv_bfloat16*>(p.y1) + row + c0 + j * 16, lw);
          }
        }
      }
    }
    if (p.stats != nullptr) {
#pragma unroll
      for (int ch = 0; ch < NCHUNK; ++ch) {
        atomicAdd(p.stats + cbase + ch * 32 + lane, st_s[ch]);
        atomicAdd(p.stats + COUT + cbase + ch * 32 + lane, st_q[ch]);
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int KB, int COUT, int OUT_DTYPE, bool HALF = false, bool SINGLE = false, bool RES32 = false>
static int launch_pw(const CUtensorMap& xh, const CUtensorMap& xl, const CUtensorMap& wh, const CUtensorMap& wl,
                     const PwParams& p, cudaStream_t st) {
  using Cfg = PwCfg<KB, COUT, SINGLE>;
  auto kern = pw_tc_kernel<KB, COUT, OUT_DTYPE, HALF, SINGLE, RES32>;
  COVA_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
  const int grid = p.n_tiles < sm_count() ? p.n_tiles : sm_count();
  kern<<<grid, PW_THREADS, Cfg::SMEM, st>>>(xh, xl, wh, wl, p);
  COVA_LAUNCH_OK();
  return COVA_OK;
}

}  // namespace cova

static int pw_run(const void* x_hi, const void* x_lo, int64_t M, int Cin, int Cout, const void* w_packed, const float* bn_scale,
                  const float* bn_shift, const void* res_hi, const void* res_lo, int relu, int out_dtype, void* y0, void* y1,
                  bool half, void* stream, double* stats = nullptr, const float* res_f32 = nullptr);

extern "C" int cova_conv1x1_bn_act_fwd(const void* x_hi, const void* x_lo, int64_t M, int Cin, int Cout,
                                       const void* w_packed, const float* bn_scale, const float* bn_shift,
                                       const void* res_hi, const void* res_lo, int relu, int out_dtype, void* y0,
                                       void* y1, void* stream) {
  return pw_run(x_hi, x_lo, M, Cin, Cout, w_packed, bn_scale, bn_shift, res_hi, res_lo, relu, out_dtype, y0, y1, false, stream);
}

extern "C" int cova_conv1x1_raw_stats_fwd(const void* x_hi, const void* x_lo, int planes_dtype, int64_t M, int Cin, int Cout,
                                          const void* w_packed, const float* scale, const float* zero_shift, void* y,
                                          double* stats_ws, void* stream);
extern "C" int cova_conv1x1_raw_res_fwd(const void* x, int64_t M, int Cin, int Cout, const void* w_bf16, const float* scale,
                                        const float* zero_shift, const void* res_bf16, void* y_bf16, void* stream);

extern "C" int cova_conv1x1_raw_fwd(const void* x_hi, const void* x_lo, int planes_dtype, int64_t M, int Cin, int Cout,
                                    const void* w_packed, const float* scale, const float* zero_shift, void* y, void* stream) {
  return cova_conv1x1_raw_stats_fwd(x_hi, x_lo, planes_dtype, M, Cin, Cout, w_packed, scale, zero_shift, y, nullptr, stream);
}

extern "C" int cova_conv1x1_raw_stats_fwd(const void* x_hi, const void* x_lo, int planes_dtype, int64_t M, int Cin, int Cout,
                                          const void* w_packed, const float* scale, const float* zero_shift, void* y,
                                          double* stats_ws, void* stream) {
  COVA_REQUIRE(planes_dtype == COVA_F16X2 || planes_dtype == COVA_BF16X2 || planes_dtype == COVA_BF16,
               "cova_conv1x1_raw_fwd: planes are split-fp16, split-bf16 or one bf16 plane");
  if (planes_dtype == COVA_BF16)   // bf16 training mode: x_lo unused, w_packed = bf16 [Cout][Cin], y = bf16 rows
    return pw_run(x_hi, x_hi, M, Cin, Cout, w_packed, scale, zero_shift, nullptr, nullptr, 0, COVA_BF16, y, nullptr, false, stream,
                  stats_ws);
  return pw_run(x_hi, x_lo, M, Cin, Cout, w_packed, scale, zero_shift, nullptr, nullptr, 0, COVA_F32, y, nullptr,
                planes_dtype == COVA_F16X2, stream, stats_ws);
}

// fp32-parity training mode: y = scale * (x W^T) + zero_shift + res_f32, split planes in, fp32 rows out (dgrad + skip gradient)
extern "C" int cova_conv1x1_raw_res_f32_fwd(const void* x_hi, const void* x_lo, int planes_dtype, int64_t M, int Cin, int Cout,
                                            const void* w_packed, const float* scale, const float* zero_shift, const float* res_f32,
                                            float* y, void* stream) {
  COVA_REQUIRE(planes_dtype == COVA_F16X2 || planes_dtype == COVA_BF16X2, "cova_conv1x1_raw_res_f32_fwd: planes are split-fp16 or split-bf16");
  COVA_REQUIRE(x_hi && x_lo && w_packed && scale && zero_shift && res_f32 && y, "cova_conv1x1_raw_res_f32_fwd: null pointer");
  COVA_REQUIRE((((uintptr_t)res_f32 | (uintptr_t)y) & 31) == 0, "cova_conv1x1_raw_res_f32_fwd: 32-byte alignment");
  return pw_run(x_hi, x_lo, M, Cin, Cout, w_packed, scale, zero_shift, nullptr, nullptr, 0, COVA_F32, y, nullptr,
                planes_dtype == COVA_F16X2, stream, nullptr, res_f32);
}

// bf16 training mode: y = x W^T + res on single bf16 planes (a dgrad with the skip branch's gradient added in the epilogue: no
// separate elementwise add pass over the 256-channel map)
extern "C" int cova_conv1x1_raw_res_fwd(const void* x, int64_t M, int Cin, int Cout, const void* w_bf16, const float* scale,
                                        const float* zero_shift, const void* res_bf16, void* y_bf16, void* stream) {
  COVA_REQUIRE(x && w_bf16 && scale && zero_shift && res_bf16 && y_bf16, "cova_conv1x1_raw_res_fwd: null pointer");
  return pw_run(x, x, M, Cin, Cout, w_bf16, scale, zero_shift, res_bf16, res_bf16, 0, COVA_BF16, y_bf16, nullptr, false, stream, nullptr);
}

static int pw_run(const void* x_hi, const void* x_lo, int64_t M, int Cin, int Cout, const void* w_packed, const float* bn_scale,
                  const float* bn_shift, const void* res_hi, const void* res_lo, int relu, int out_dtype, void* y0, void* y1,
                  bool half, void* stream, double* stats, const float* res_f32) {
  using namespace cova;
  COVA_REQUIRE(x_hi && x_lo && w_packed && bn_scale && bn_shift && y0, "cova_conv1x1_bn_act_fwd: null pointer");
  COVA_REQUIRE((Cin == 64 && (Cout == 64 || Cout == 256)) || (Cin == 256 && Cout == 64),
               "cova_conv1x1_bn_act_fwd: built for 64->64, 64->256 and 256->64 (got %d->%d)", Cin, Cout);
  COVA_REQUIRE(out_dtype == COVA_F32 || out_dtype == COVA_BF16 || (out_dtype == COVA_BF16X2 && y1), "cova_conv1x1_bn_act_fwd: bad output dtype");
  COVA_REQUIRE((res_hi == nullptr) == (res_lo == nullptr), "cova_conv1x1_bn_act_fwd: residual needs both planes");
  COVA_REQUIRE(M >= 0 && M < (int64_t)1 << 31, "cova_conv1x1_bn_act_fwd: M out of range");
  if (M == 0) return COVA_OK;
  CUtensorMap tx_hi, tx_lo, tw_hi, tw_lo;
  const uint64_t xd[2] = {(uint64_t)Cin, (uint64_t)M}, xs[1] = {(uint64_t)Cin * 2};
  const uint32_t xb[2] = {64, PW_BM};
  const uint64_t wd[2] = {(uint64_t)Cin, (uint64_t)Cout}, ws[1] = {(uint64_t)Cin * 2};
  const uint32_t wb[2] = {64, (uint32_t)Cout};
  const __nv_bfloat16* wp = (const __nv_bfloat16*)w_packed;   // [2][Cout][Cin] from cova_pack_linear_weight
  int rc;
  if ((rc = make_tmap_bf16(&tx_hi, x_hi, 2, xd, xs, xb))) return rc;
  if ((rc = make_tmap_bf16(&tx_lo, x_lo, 2, xd, xs, xb))) return rc;
  if ((rc = make_tmap_bf16(&tw_hi, wp, 2, wd, ws, wb))) return rc;
  if ((rc = make_tmap_bf16(&tw_lo, out_dtype == COVA_BF16 ? wp : wp + (size_t)Cout * Cin, 2, wd, ws, wb))) return rc;
  PwParams p;
  p.M = (int)M;
  p.n_tiles = ceil_div((int)M, PW_BM);
  p.relu = relu;
  p.bn_scale = bn_scale; p.bn_shift = bn_shift;
  p.res_hi = (const __nv_bfloat16*)res_hi; p.res_lo = (const __nv_bfloat16*)res_lo;
  p.y0 = y0; p.y1 = y1;
  p.stats = stats;
  p.res_f32 = res_f32;
  cudaStream_t st = (cudaStream_t)stream;
  if (stats) COVA_CUDA_OK(cudaMemsetAsync(stats, 0, 2 * Cout * sizeof(double), st));
#define GO(KB, CO) (res_f32 != nullptr ? (half ? launch_pw<KB, CO, COVA_F32, true, false, true>(tx_hi, tx_lo, tw_hi, tw_lo, p, st)     \
                                              : launch_pw<KB, CO, COVA_F32, false, false, true>(tx_hi, tx_lo, tw_hi, tw_lo, p, st))   \
                    : out_dtype == COVA_BF16 ? launch_pw<KB, CO, COVA_BF16, false, true>(tx_hi, tx_lo, tw_hi, tw_lo, p, st) \
                    : half ? launch_pw<KB, CO, COVA_F32, true>(tx_hi, tx_lo, tw_hi, tw_lo, p, st)          \
                    : out_dtype == COVA_F32 ? launch_pw<KB, CO, COVA_F32>(tx_hi, tx_lo, tw_hi, tw_lo, p, st) \
                                            : launch_pw<KB, CO, COVA_BF16X2>(tx_hi, tx_lo, tw_hi, tw_lo, p, st))
  if (Cin == 64 && Cout == 64) return GO(1, 64);
  if (Cin == 64) return GO(1, 256);
  return GO(4, 64);
#undef GO
}
