// A9 / SURVEY D2: weight gradient of the ResNet-50 Bottleneck 1x1 convolutions (conv1 / conv3 / downsample,
// `loss.backward()` of /root/reference/train.py:59 through torchvision `Bottleneck.forward`) on the tensor cores:
//     dW[co][ci] = inv_scale * sum over the M = B*H*W pixel rows of  dY[m][co] * X[m][ci]
// The contraction runs over pixels, so - exactly as in wgrad_tc.cu - both NHWC operands are MN-major tcgen05 operands:
// a TMA box {64 channels, 64 pixel rows} with the 128-byte swizzle IS the operand image (row = K index).
// UMMA shape: A = X (M side), one M = 128 MMA per 64-channel block of Cin with the block's hi and lo planes as its two
// 64-row atoms; B = dY (N side), [hi ; lo] of one (N = 128) or two (N = 256) 64-channel blocks of Cout, the atoms laid
// out 8 KB apart.  Every MMA therefore yields all four plane products (hi*hi, hi*lo, lo*hi and the negligible lo*lo) at
// the full N >= 128 tensor rate; accumulators: (Cin/64) x (Cout/64) x 128 columns = 128 (64->64) or 512 (64->256, 256->64).
// Persistent CTAs walk 64-pixel tiles (4 K-steps), ring of 2-4 stages, accumulators drained every `drain_every` tiles with
// 128-bit reductions into the global fp32 sum [Cin][Cout] (rows ci of the hi and of the lo plane land on the same address).
//
// Bound: HBM - 4*(Cin+Cout) bytes per pixel (split planes) against 512-2048 tensor clocks per 128 pixels.
#include "common.cuh"
#include "ptx.cuh"
#include "tma_host.cuh"

namespace cova {

constexpr int PWG_BM = 64;                          // pixel rows per tile
constexpr int PWG_ATOM = PWG_BM * 128;              // 8 KB: 64 rows x 64 channels of one plane
constexpr int PWG_THREADS = 192;                    // warp 0 TMA, warp 1 MMA, warps 2-5 drain

struct PwgTail {
  uint64_t full[4], empty[4], acc_full, acc_empty;
  uint32_t tmem_base;
};

struct PwgParams {
  int M, n_tiles, drain_every;
  int single;                                       // bf16 training mode: one plane per operand - lo atoms zero-filled once, never loaded
  float* ws;                                        // [Cin][Cout] fp32, zeroed
};

template <int CIB, int COB>
struct PwgCfg {
  static constexpr int NSTAGE = (CIB + COB == 2) ? 4 : 2;
  static constexpr int STAGE_BYTES = (CIB + COB) * 2 * PWG_ATOM;
  static constexpr int X_BYTES = CIB * 2 * PWG_ATOM;
  static constexpr int TMEM_COLS = CIB * COB * 128;               // 128 or 512
  static constexpr int SMEM = NSTAGE * STAGE_BYTES + 1024 + 1024;
};

__device__ __forceinline__ uint64_t pwg_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t pwg_idesc(int M, int N, bool half) {
  return (1u << 4) | (half ? 0u : ((1u << 7) | (1u << 10))) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

template <int CIB, int COB, bool HALF>
__global__ void __launch_bounds__(PWG_THREADS, 1)
pw_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tm_x_hi, const __grid_constant__ CUtensorMap tm_x_lo,
                   const __grid_constant__ CUtensorMap tm_dy_hi, const __grid_constant__ CUtensorMap tm_dy_lo,
                   const PwgParams p) {
  using Cfg = PwgCfg<CIB, COB>;
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  unsigned char* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  PwgTail& tail = *reinterpret_cast<PwgTail*>(smem + Cfg::NSTAGE * Cfg::STAGE_BYTES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int my_tiles = p.n_tiles > (int)blockIdx.x ? (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < Cfg::NSTAGE; ++i) {
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
    ptx::tmem_alloc(&tail.tmem_base, Cfg::TMEM_COLS);
    ptx::tmem_relinquish();
  }
  if (p.single) {
    for (int st = 0; st < Cfg::NSTAGE; ++st)
      for (int a = 0; a < CIB + COB; ++a) {
        uint4* lo = reinterpret_cast<uint4*>(smem + st * Cfg::STAGE_BYTES + (a * 2 + 1) * PWG_ATOM);
        for (int i = threadIdx.x; i < PWG_ATOM / 16; i += PWG_THREADS) lo[i] = make_uint4(0u, 0u, 0u, 0u);
      }
    ptx::fence_proxy_async();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tail.tmem_base;

  if (warp == 0) {
    // ---------------- TMA producer: atoms [block][plane], 8 KB apart ----------------
    uint32_t stage = 0, phase = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      ptx::mbar_wait(&tail.empty[stage], phase ^ 1);
      if (ptx::elect_one()) {
        unsigned char* s = smem + stage * Cfg::STAGE_BYTES;
        ptx::mbar_arrive_expect_tx(&tail.full[stage], p.single ? Cfg::STAGE_BYTES / 2 : Cfg::STAGE_BYTES);
#pragma unroll
        for (int cb = 0; cb < CIB; ++cb) {
          ptx::tma_load_2d(s + (cb * 2) * PWG_ATOM, &tm_x_hi, &tail.full[stage], cb * 64, tile * PWG_BM);
          if (!p.single) ptx::tma_load_2d(s + (cb * 2 + 1) * PWG_ATOM, &tm_x_lo, &tail.full[stage], cb * 64, tile * PWG_BM);
        }
#pragma unroll
        for (int nb = 0; nb < COB; ++nb) {
          ptx::tma_load_2d(s + Cfg::X_BYTES + (nb * 2) * PWG_ATOM, &tm_dy_hi, &tail.full[stage], nb * 64, tile * PWG_BM);
          if (!p.single) ptx::tma_load_2d(s + Cfg::X_BYTES + (nb * 2 + 1) * PWG_ATOM, &tm_dy_lo, &tail.full[stage], nb * 64, tile * PWG_BM);
        }
      }
      __syncwarp();
      if (++stage == Cfg::NSTAGE) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    constexpr int NB_PER_MMA = COB >= 2 ? 2 : 1;                       // 64-channel blocks of Cout per MMA (N = 256 / 128)
    constexpr uint32_t idesc = pwg_idesc(128, NB_PER_MMA * 128, HALF);
    uint32_t stage = 0, phase = 0;
    int since_drain = 0, n_drain = 0;
    for (int i = 0; i < my_tiles; ++i) {
      if (since_drain == 0 && n_drain > 0) {
        ptx::mbar_wait(&tail.acc_empty, (n_drain - 1) & 1);
        ptx::tc_fence_after();
      }
      ptx::mbar_wait(&tail.full[stage], phase);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        const uint32_t s_addr = ptx::smem_u32(smem + stage * Cfg::STAGE_BYTES);
#pragma unroll
        for (int kk = 0; kk < PWG_BM / 16; ++kk) {
          const uint32_t acc = (since_drain > 0 || kk > 0) ? 1u : 0u;
#pragma unroll
          for (int cb = 0; cb < CIB; ++cb) {
            const uint64_t da = pwg_desc_mn(s_addr + cb * 2 * PWG_ATOM + kk * 2048, PWG_ATOM, 1024);
#pragma unroll
            for (int j = 0; j < COB / NB_PER_MMA; ++j) {
              const uint64_t db = pwg_desc_mn(s_addr + Cfg::X_BYTES + j * NB_PER_MMA * 2 * PWG_ATOM + kk * 2048, PWG_ATOM, 1024);
              ptx::umma_bf16(tmem_base + (cb * COB + j * NB_PER_MMA) * 128, da, db, idesc, acc);
            }
          }
        }
        ptx::umma_commit(&tail.empty[stage]);
        if (since_drain + 1 == p.drain_every || i + 1 == my_tiles) ptx::umma_commit(&tail.acc_full);
      }
      __syncwarp();
      if (++since_drain == p.drain_every || i + 1 == my_tiles) { since_drain = 0; ++n_drain; }
      if (++stage == Cfg::NSTAGE) { stage = 0; phase ^= 1; }
    }
  } else {
    // ---------------- drain: row = (plane of X, ci); columns = [block][plane of dY][co] ----------------
    const int lg = warp & 3;
    const int ci = (lg * 32 + lane) & 63;
    const int n_drains = (my_tiles + p.drain_every - 1) / p.drain_every;
    constexpr int COUT = COB * 64;
    for (int dr = 0; dr < n_drains; ++dr) {
      ptx::mbar_wait(&tail.acc_full, dr & 1);
      ptx::tc_fence_after();
#pragma unroll 1
      for (int blk = 0; blk < CIB * COB; ++blk) {
        const int cb = blk / COB, nb = blk % COB;
        const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + blk * 128;
        float* dst = p.ws + (size_t)(cb * 64 + ci) * COUT + nb * 64;
#pragma unroll 1
        for (int q = 0; q < 4; ++q) {
          uint32_t v[16], u[16];
          ptx::tmem_ld16(taddr + q * 16, v);
          ptx::tmem_ld16(taddr + 64 + q * 16, u);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float4 a;
            a.x = __uint_as_float(v[4 * j]) + __uint_as_float(u[4 * j]);
            a.y = __uint_as_float(v[4 * j + 1]) + __uint_as_float(u[4 * j + 1]);
            a.z = __uint_as_float(v[4 * j + 2]) + __uint_as_float(u[4 * j + 2]);
            a.w = __uint_as_float(v[4 * j + 3]) + __uint_as_float(u[4 * j + 3]);
            atomicAdd(reinterpret_cast<float4*>(dst + q * 16 + 4 * j), a);
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
    ptx::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

int launch_wgrad_finalize(const float* ws, const float* inv_scale, float mult, int taps, int Cin, int Cout, float* dw, cudaStream_t st);

template <int CIB, int COB>
static int launch_pwg(const CUtensorMap& xh, const CUtensorMap& xl, const CUtensorMap& dh, const CUtensorMap& dl,
                      const PwgParams& p, bool half, cudaStream_t st) {
  using Cfg = PwgCfg<CIB, COB>;
  static_assert(sizeof(PwgTail) <= 1024, "tail too large");
  const int grid = p.n_tiles < sm_count() ? p.n_tiles : sm_count();
  if (half) {
    auto kern = pw_wgrad_tc_kernel<CIB, COB, true>;
    COVA_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    kern<<<grid, PWG_THREADS, Cfg::SMEM, st>>>(xh, xl, dh, dl, p);
  } else {
    auto kern = pw_wgrad_tc_kernel<CIB, COB, false>;
    COVA_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    kern<<<grid, PWG_THREADS, Cfg::SMEM, st>>>(xh, xl, dh, dl, p);
  }
  COVA_LAUNCH_OK();
  return COVA_OK;
}

}  // namespace cova

extern "C" int cova_conv1x1_wgrad(const void* x_hi, const void* x_lo, const void* dy_hi, const void* dy_lo, int64_t M, int Cin,
                                  int Cout, int planes_dtype, const float* inv_scale, float* ws, float* dw, void* stream) {
  using namespace cova;
  COVA_REQUIRE(x_hi && dy_hi && ws && dw, "cova_conv1x1_wgrad: null pointer");
  COVA_REQUIRE(planes_dtype == COVA_BF16 || (x_lo && dy_lo), "cova_conv1x1_wgrad: split planes need their lo plane");
  COVA_REQUIRE((Cin == 64 && (Cout == 64 || Cout == 256)) || (Cin == 256 && Cout == 64),
               "cova_conv1x1_wgrad: built for 64->64, 64->256 and 256->64 (got %d->%d)", Cin, Cout);
  COVA_REQUIRE(planes_dtype == COVA_F16X2 || planes_dtype == COVA_BF16X2 || planes_dtype == COVA_BF16,
               "cova_conv1x1_wgrad: planes are split-fp16, split-bf16 or single bf16 planes (x_lo = dy_lo = NULL)");
  COVA_REQUIRE(M > 0 && M < (int64_t)1 << 31, "cova_conv1x1_wgrad: M out of range");
  cudaStream_t st = (cudaStream_t)stream;
  CUtensorMap tx_hi, tx_lo, td_hi, td_lo;
  const uint64_t xd[2] = {(uint64_t)Cin, (uint64_t)M}, xs[1] = {(uint64_t)Cin * 2};
  const uint64_t dd[2] = {(uint64_t)Cout, (uint64_t)M}, ds[1] = {(uint64_t)Cout * 2};
  const uint32_t bx[2] = {64, PWG_BM};
  int rc;
  if ((rc = make_tmap_bf16(&tx_hi, x_hi, 2, xd, xs, bx))) return rc;
  if ((rc = make_tmap_bf16(&tx_lo, x_lo ? x_lo : x_hi, 2, xd, xs, bx))) return rc;
  if ((rc = make_tmap_bf16(&td_hi, dy_hi, 2, dd, ds, bx))) return rc;
  if ((rc = make_tmap_bf16(&td_lo, dy_lo ? dy_lo : dy_hi, 2, dd, ds, bx))) return rc;
  PwgParams p;
  p.M = (int)M;
  p.n_tiles = ceil_div((int)M, PWG_BM);
  p.drain_every = 2 * knob(COVA_KNOB_WGRAD_DRAIN, 16);
  if (p.drain_every < 1) p.drain_every = 1;
  p.ws = ws;
  p.single = planes_dtype == COVA_BF16;
  COVA_CUDA_OK(cudaMemsetAsync(ws, 0, (size_t)Cin * Cout * sizeof(float), st));
  const bool half = planes_dtype == COVA_F16X2;
  if (Cin == 64 && Cout == 64) rc = launch_pwg<1, 1>(tx_hi, tx_lo, td_hi, td_lo, p, half, st);
  else if (Cin == 64) rc = launch_pwg<1, 4>(tx_hi, tx_lo, td_hi, td_lo, p, half, st);
  else rc = launch_pwg<4, 1>(tx_hi, tx_lo, td_hi, td_lo, p, half, st);
  if (rc) return rc;
  return launch_wgrad_finalize(ws, inv_scale, 1.f, 1, Cin, Cout, dw, st);
}
