// Row-major linear layer on the 5th-gen tensor cores, fp32-parity (split-bf16, 3 products, fp32 accumulate in TMEM):
//     Y[M,N] = act((X[M,K] @ W[N,K]^T + bias) * scale + shift + res)
// Replaces `nn.Linear` (+ folded BatchNorm1d + ReLU) of `/root/reference/models.py:160-164` (the GAT projections,
// fused into one extended weight matrix), `:85-87` (decoder.1 + decoder.2 + ReLU).
//
// X arrives as fp32 (it is produced by the RoI / GAT kernels); the split into bf16 hi/lo happens inside the kernel:
// TMA streams raw 128 x 64 fp32 tiles into a 3-deep shared-memory ring (so three tiles of loads are in flight without a
// register being held: with register-staged loads one tile ahead the kernel was a chain of exposed L2 round trips,
// ~1.7 us per k-block), converter warps read them and store the two bf16 planes in the K-major 128-byte-swizzled
// layout tcgen05 expects (chunk16 index XOR (row & 7) on the absolute address - the same rule TMA applies), then
// fence.proxy.async + mbarrier.  W is pre-split once per weight update (cova_pack_linear_weight) and streamed by TMA.
// One CTA = one 128 x 96 output tile (N = 992 -> 11 column tiles x 12 row tiles = 132 CTAs for the decoder: one wave
// of 148 SMs), 2-stage operand ring + 3-stage raw ring.
//
// Bound: tensor pipe (tiny GEMMs: < 1 % of the step's FLOPs); the point is to take them off the CUDA cores.
#include "common.cuh"
#include "ptx.cuh"
#include "tma_host.cuh"

namespace cova {

constexpr int LT_BM = 128, LT_BN = 96, LT_BK = 64;
constexpr int LT_STAGES = 2;                     // converted A planes (hi, lo)
constexpr int LT_WSTAGES = 4;                    // W tiles (hi, lo): TMA runs up to four k-blocks ahead
constexpr int LT_XSTAGES = 2;                    // raw fp32 X tiles
constexpr int LT_X_BYTES = LT_BM * LT_BK * 4;    // 32 KB: one raw fp32 tile
constexpr int LT_A_PLANE = LT_BM * 128;          // 16 KB
constexpr int LT_B_PLANE = LT_BN * 128;          // 12 KB
constexpr int LT_STAGE_BYTES = 2 * LT_A_PLANE;    // 32 KB: A planes of one k-block
constexpr int LT_W_BYTES = 2 * LT_B_PLANE;        // 24 KB: W planes of one k-block
constexpr int LT_CVT_WARPS = 8;                  // converter / epilogue warps: two per TMEM lane group (one scheduler each had a
                                                 // single dependent-instruction stream with four: the kernel was converter-bound)
constexpr int LT_THREADS = 64 + 32 * LT_CVT_WARPS;   // warp 0 TMA, warp 1 MMA, then the converter / epilogue warps
constexpr int LT_SMEM = LT_STAGES * LT_STAGE_BYTES + LT_WSTAGES * LT_W_BYTES + LT_XSTAGES * LT_X_BYTES + 1024 + 1024;

struct LinearTcTail {
  uint64_t a_full[LT_STAGES], empty[LT_STAGES], b_full[LT_WSTAGES], b_empty[LT_WSTAGES], acc_full, x_full[LT_XSTAGES],
      x_empty[LT_XSTAGES];
  uint32_t tmem_base;
};

struct LinearTcParams {
  const float* x;
  int64_t ldx;
  int M, K, N;
  const float* bias;
  const float* scale;
  const float* shift;
  const float* res;
  int64_t ldr;
  int relu;
  int out_dtype;          // COVA_F32: y0 fp32;  COVA_BF16X2: y0 = hi plane, y1 = lo plane (bf16), ldy in elements
  void* y0;
  void* y1;
  int64_t ldy;
};

// HALF: split-fp16 operands (hi = f16(x), lo = f16(x - hi): 22 significand bits; W packed x256 by
// cova_pack_conv_weight_f16x2, undone in the epilogue) instead of split-bf16 (16 bits) - the training forward, whose
// gradients are discontinuous in the forward precision at the 1e-5 level.
template <bool HALF>
__global__ void __launch_bounds__(LT_THREADS, 1)
linear_tc_kernel(const __grid_constant__ CUtensorMap tm_w_hi, const __grid_constant__ CUtensorMap tm_w_lo,
                 const __grid_constant__ CUtensorMap tm_x, const LinearTcParams p) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  unsigned char* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  unsigned char* sm_w = smem + LT_STAGES * LT_STAGE_BYTES;
  unsigned char* sm_x = sm_w + LT_WSTAGES * LT_W_BYTES;
  LinearTcTail& tail = *reinterpret_cast<LinearTcTail*>(sm_x + LT_XSTAGES * LT_X_BYTES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * LT_BM, n0 = blockIdx.x * LT_BN;
  const int nkb = (p.K + LT_BK - 1) / LT_BK;

  if (threadIdx.x == 0) {
    for (int i = 0; i < LT_STAGES; ++i) {
      ptx::mbar_init(&tail.a_full[i], LT_CVT_WARPS);     // one arrive per converter warp
      ptx::mbar_init(&tail.empty[i], 1);
    }
    for (int i = 0; i < LT_WSTAGES; ++i) {
      ptx::mbar_init(&tail.b_full[i], 1);
      ptx::mbar_init(&tail.b_empty[i], 1);
    }
    for (int i = 0; i < LT_XSTAGES; ++i) {
      ptx::mbar_init(&tail.x_full[i], 1);
      ptx::mbar_init(&tail.x_empty[i], LT_CVT_WARPS);    // one arrive per converter warp
    }
    ptx::mbar_init(&tail.acc_full, 1);
    ptx::fence_barrier_init();
    ptx::prefetch_tensormap(&tm_w_hi);
    ptx::prefetch_tensormap(&tm_w_lo);
    ptx::prefetch_tensormap(&tm_x);
  }
  if (warp == 1) {
    ptx::tmem_alloc(&tail.tmem_base, 128);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tail.tmem_base;

  if (warp == 0) {
    // ---------------- TMA producer: W tiles (hi, lo; up to 4 k-blocks ahead) and raw X tiles (2 ahead) ----------------
    uint32_t ws_ = 0, wph = 0, xs = 0, xph = 0;
    int kw = 0, kx = 0;
    while (kw < nkb || kx < nkb) {
      // issue whichever ring has a free slot; W first (it never blocks on the converters)
      if (kw < nkb && (kw - kx < LT_WSTAGES || kx >= nkb)) {
        ptx::mbar_wait(&tail.b_empty[ws_], wph ^ 1);
        if (ptx::elect_one()) {
          unsigned char* sb = sm_w + ws_ * LT_W_BYTES;
          ptx::mbar_arrive_expect_tx(&tail.b_full[ws_], LT_W_BYTES);
          ptx::tma_load_2d(sb, &tm_w_hi, &tail.b_full[ws_], kw * LT_BK, n0);
          ptx::tma_load_2d(sb + LT_B_PLANE, &tm_w_lo, &tail.b_full[ws_], kw * LT_BK, n0);
        }
        __syncwarp();
        ++kw;
        if (++ws_ == LT_WSTAGES) { ws_ = 0; wph ^= 1; }
      }
      if (kx < nkb && kx <= kw) {
        ptx::mbar_wait(&tail.x_empty[xs], xph ^ 1);
        if (ptx::elect_one()) {
          ptx::mbar_arrive_expect_tx(&tail.x_full[xs], LT_X_BYTES);
          ptx::tma_load_2d(sm_x + xs * LT_X_BYTES, &tm_x, &tail.x_full[xs], kx * LT_BK, m0);
        }
        __syncwarp();
        ++kx;
        if (++xs == LT_XSTAGES) { xs = 0; xph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    constexpr uint32_t idesc = HALF ? ptx::umma_idesc_f16(LT_BM, LT_BN) : ptx::umma_idesc_bf16(LT_BM, LT_BN);
    const uint64_t d0 = ptx::umma_desc_sw128(ptx::smem_u32(smem), 1024);
    const uint32_t hi32 = (uint32_t)(d0 >> 32), lo0 = (uint32_t)d0;
    const uint64_t dw0 = ptx::umma_desc_sw128(ptx::smem_u32(sm_w), 1024);
    const uint32_t whi32 = (uint32_t)(dw0 >> 32), wlo0 = (uint32_t)dw0;
    uint32_t stage = 0, phase = 0, ws_ = 0, wph = 0;
    for (int kb = 0; kb < nkb; ++kb) {
      ptx::mbar_wait(&tail.a_full[stage], phase);
      ptx::mbar_wait(&tail.b_full[ws_], wph);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        const uint32_t s_lo = lo0 + ((stage * LT_STAGE_BYTES) >> 4);
        const uint32_t w_lo = wlo0 + ((ws_ * LT_W_BYTES) >> 4);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint64_t a_hi = ((uint64_t)hi32 << 32) | (uint32_t)(s_lo + ((kk * 32) >> 4));
          const uint64_t a_lo = a_hi + (LT_A_PLANE >> 4);
          const uint64_t b_hi = ((uint64_t)whi32 << 32) | (uint32_t)(w_lo + ((kk * 32) >> 4));
          const uint64_t b_lo = b_hi + (LT_B_PLANE >> 4);
          ptx::umma_bf16(tmem_base, a_hi, b_hi, idesc, (kb | kk) != 0);
          ptx::umma_bf16(tmem_base, a_lo, b_hi, idesc, 1);
          ptx::umma_bf16(tmem_base, a_hi, b_lo, idesc, 1);
        }
        ptx::umma_commit(&tail.empty[stage]);
        ptx::umma_commit(&tail.b_empty[ws_]);
        if (kb == nkb - 1) ptx::umma_commit(&tail.acc_full);
      }
      __syncwarp();
      if (++stage == LT_STAGES) { stage = 0; phase ^= 1; }
      if (++ws_ == LT_WSTAGES) { ws_ = 0; wph ^= 1; }
    }
  } else {
    // ---------------- converters: raw fp32 X tile (shared memory) -> swizzled 16-bit hi/lo planes ----------------
    constexpr int NJ = 128 / (2 * LT_CVT_WARPS);     // rows per thread
    const int ct = threadIdx.x - 64;                 // 0 .. 32 * LT_CVT_WARPS - 1
    const int c = ct & 15, r0 = ct >> 4;             // 16-byte fp32 chunk c of rows r0 + (2 * LT_CVT_WARPS) j: a warp reads 512 contiguous bytes
    uint32_t stage = 0, phase = 0, xs = 0, xph = 0;
    for (int kb = 0; kb < nkb; ++kb) {
      ptx::mbar_wait(&tail.x_full[xs], xph);
      const float4* src = reinterpret_cast<const float4*>(sm_x + xs * LT_X_BYTES);
      float4 v[NJ];
#pragma unroll
      for (int j = 0; j < NJ; ++j) v[j] = src[(r0 + 2 * LT_CVT_WARPS * j) * 16 + c];
      ptx::mbar_wait(&tail.empty[stage], phase ^ 1);
      unsigned char* sa = smem + stage * LT_STAGE_BYTES;
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int r = r0 + 2 * LT_CVT_WARPS * j;
        uint32_t h01, l01, h23, l23;
        if (HALF) {
          split_f16x2(v[j].x, v[j].y, h01, l01);
          split_f16x2(v[j].z, v[j].w, h23, l23);
        } else {
          split_bf16x2(v[j].x, v[j].y, h01, l01);
          split_bf16x2(v[j].z, v[j].w, h23, l23);
        }
        const uint32_t off = r * 128 + (((c >> 1) ^ (r & 7)) << 4) + (c & 1) * 8;   // 128-B swizzle
        *reinterpret_cast<uint2*>(sa + off) = make_uint2(h01, h23);
        *reinterpret_cast<uint2*>(sa + LT_A_PLANE + off) = make_uint2(l01, l23);
      }
      ptx::fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        ptx::mbar_arrive(&tail.a_full[stage]);
        // the raw slot is released only now: every lane's shared-memory loads of it have been CONSUMED by the conversion above
        // (an arrive issued right after the loads let the next TMA tile overwrite the slot while loads were still in flight)
        ptx::mbar_arrive(&tail.x_empty[xs]);
      }
      if (++xs == LT_XSTAGES) { xs = 0; xph ^= 1; }
      if (++stage == LT_STAGES) { stage = 0; phase ^= 1; }
    }
    // ---------------- epilogue (same warps): TMEM lane group = warp % 4 ----------------
    const int lg = warp & 3;
    const int m = m0 + lg * 32 + lane;
    constexpr int QPW = (LT_BN / 16) / (LT_CVT_WARPS / 4);     // 16-column chunks per warp
    const int q0 = ((warp - 2) >> 2) * QPW;
    ptx::mbar_wait(&tail.acc_full, 0);
    ptx::tc_fence_after();
    const bool vec_ok = (p.ldy % 16 == 0) && ((reinterpret_cast<uintptr_t>(p.y0) & 31) == 0) &&
                        (p.y1 == nullptr || (reinterpret_cast<uintptr_t>(p.y1) & 31) == 0);
#pragma unroll 1
    for (int q = q0; q < q0 + QPW; ++q) {
      uint32_t raw[16];
      ptx::tmem_ld16(tmem_base + ((uint32_t)(lg * 32) << 16) + q * 16, raw);
      ptx::tmem_ld_wait();
      const int nb = n0 + q * 16;
      if (m < p.M && nb < p.N) {
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int n = nb + j;
          float x = __uint_as_float(raw[j]);
          if (HALF) x *= 1.f / SPLIT_F16_WSCALE;
          if (n < p.N) {
            if (p.bias) x += p.bias[n];
            if (p.scale) x = fmaf(x, p.scale[n], p.shift[n]);
            if (p.res) x += p.res[(size_t)m * p.ldr + n];
            if (p.relu) x = fmaxf(x, 0.f);
          }
          v[j] = x;
        }
        const size_t o = (size_t)m * p.ldy + nb;
        const bool full = nb + 16 <= p.N && vec_ok;
        if (p.out_dtype == COVA_F32) {
          float* y = reinterpret_cast<float*>(p.y0);
          if (full) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              uint32_t w8[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) w8[e] = __float_as_uint(v[h * 8 + e]);
              st_global_v8(y + o + h * 8, w8);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (nb + j < p.N) y[o + j] = v[j];
          }
        } else {
          __nv_bfloat16* yh = reinterpret_cast<__nv_bfloat16*>(p.y0);
          __nv_bfloat16* yl = reinterpret_cast<__nv_bfloat16*>(p.y1);
          if (full) {
            uint32_t hw[8], lw[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) split_bf16x2(v[2 * e], v[2 * e + 1], hw[e], lw[e]);
            st_global_v8(yh + o, hw);
            st_global_v8(yl + o, lw);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (nb + j < p.N) split_bf16(v[j], yh[o + j], yl[o + j]);
          }
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 128);
  }
}

// fp32 [N,K] -> bf16 [2][N][K] (hi plane, lo plane)
__global__ void pack_linear_weight_kernel(const float* __restrict__ w, int64_t n, __nv_bfloat16* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    __nv_bfloat16 h, l;
    split_bf16(w[i], h, l);
    out[i] = h;
    out[n + i] = l;
  }
}

bool linear_tc_supported(const float* x, int64_t ld_x, int K) {
  return (K % 8 == 0) && (ld_x % 4 == 0) && (((uintptr_t)x & 15) == 0);
}

int linear_tc(const float* x, int64_t ld_x, int M, int K, const void* w_packed, int N, const float* bias,
              const float* scale, const float* shift, const float* res, int64_t ld_res, int relu, int out_dtype,
              void* y0, void* y1, int64_t ld_y, cudaStream_t st, bool half) {
  CUtensorMap tw_hi, tw_lo;
  const uint64_t wd[2] = {(uint64_t)K, (uint64_t)N};
  const uint64_t ws[1] = {(uint64_t)K * 2};
  const uint32_t wb[2] = {LT_BK, LT_BN};
  const __nv_bfloat16* wp = (const __nv_bfloat16*)w_packed;
  int rc;
  if ((rc = make_tmap_bf16(&tw_hi, wp, 2, wd, ws, wb))) return rc;
  if ((rc = make_tmap_bf16(&tw_lo, wp + (size_t)N * K, 2, wd, ws, wb))) return rc;
  CUtensorMap tx;
  if ((rc = make_tmap_f32_2d(&tx, x, (uint64_t)K, (uint64_t)M, (uint64_t)ld_x * 4, LT_BK, LT_BM))) return rc;
  LinearTcParams p{x, ld_x, M, K, N, bias, scale, shift, res, ld_res, relu, out_dtype, y0, y1, ld_y};
  dim3 grid(ceil_div(N, LT_BN), ceil_div(M, LT_BM));
  if (half) {
    COVA_CUDA_OK(cudaFuncSetAttribute(linear_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, LT_SMEM));
    linear_tc_kernel<true><<<grid, LT_THREADS, LT_SMEM, st>>>(tw_hi, tw_lo, tx, p);
  } else {
    COVA_CUDA_OK(cudaFuncSetAttribute(linear_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, LT_SMEM));
    linear_tc_kernel<false><<<grid, LT_THREADS, LT_SMEM, st>>>(tw_hi, tw_lo, tx, p);
  }
  COVA_LAUNCH_OK();
  return COVA_OK;
}

}  // namespace cova

extern "C" int cova_pack_linear_weight(const float* w, int N, int K, void* packed, void* stream) {
  COVA_REQUIRE(w && packed && N > 0 && K > 0, "cova_pack_linear_weight: bad arguments");
  const int64_t n = (int64_t)N * K;
  cova::pack_linear_weight_kernel<<<(int)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184), 256, 0, (cudaStream_t)stream>>>(
      w, n, (__nv_bfloat16*)packed);
  COVA_LAUNCH_OK();
  return COVA_OK;
}
