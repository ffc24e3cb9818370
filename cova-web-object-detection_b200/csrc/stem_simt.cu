// A2 stem, CUDA-core fp32 engine: conv 7x7 s2 p3 (3->64) + folded BN + ReLU + maxpool 3x3 s2 p1 in ONE
// kernel (replaces `convnet[0:4]`, `/root/reference/models.py:49-51`, applied `:125`).
// The [B,64,H/2,W/2] conv output lives only in shared memory.
//
// One CTA = one 8x8 tile of POOLED pixels = a 17x17 tile of conv outputs = a 39x39x3 input patch.
// HBM traffic per CTA: 39*39*3*4 B in (1.49x the compulsory 32x32x3), 8*8*64*elt out.
#include <float.h>

#include "common.cuh"

namespace cova {

constexpr int ST_TP = 8;                 // pooled tile edge
constexpr int ST_TC = 2 * ST_TP + 1;     // conv tile edge (17)
constexpr int ST_TI = 2 * (ST_TC - 1) + 7;  // input patch edge (39)
constexpr int ST_TIP = ST_TI + 1;        // padded row pitch (40)
constexpr int ST_K = 3 * 7 * 7;          // 147
constexpr int ST_CO = 64;
constexpr int ST_THREADS = 256;
constexpr int ST_CT_PITCH = ST_CO + 4;   // conv-tile channel pitch (floats), keeps float4 alignment

struct StemSmem {
  float w[ST_K][ST_CO];                  // [k = c*49 + r*7 + s][cout]
  float in[3][ST_TI][ST_TIP];
  float ct[ST_TC * ST_TC][ST_CT_PITCH];  // conv tile after BN+ReLU
  float scale[ST_CO], shift[ST_CO];
};

template <int OUT_DTYPE, bool U8>
__global__ void __launch_bounds__(ST_THREADS, 1)
stem_simt_kernel(const void* __restrict__ img, int B, int H, int W, const float* __restrict__ wgt,
                 const float* __restrict__ bn_scale, const float* __restrict__ bn_shift, void* __restrict__ out0,
                 void* __restrict__ out1, int Hc, int Wc, int Hp, int Wp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  StemSmem& sm = *reinterpret_cast<StemSmem*>(smem_raw);
  const int tid = threadIdx.x;
  const int b = blockIdx.z;
  const int ph0 = blockIdx.y * ST_TP, pw0 = blockIdx.x * ST_TP;
  const int cr0 = 2 * ph0 - 1, cc0 = 2 * pw0 - 1;   // first conv row/col of the tile
  const int ir0 = 2 * cr0 - 3, ic0 = 2 * cc0 - 3;   // first input row/col of the patch

  // weights OIHW [64][147] -> smem [147][64]
  for (int i = tid; i < ST_K * ST_CO; i += ST_THREADS) {   // co fastest: conflict-free smem stores, L1-cached gathers
    int k = i / ST_CO, co = i % ST_CO;
    sm.w[k][co] = __ldg(wgt + co * ST_K + k);
  }
  if (tid < ST_CO) {
    sm.scale[tid] = bn_scale[tid];
    sm.shift[tid] = bn_shift[tid];
  }
  // input patch, zero outside the image (conv zero padding)
  const size_t imb = (size_t)b * 3 * H * W;
  for (int i = tid; i < 3 * ST_TI * ST_TI; i += ST_THREADS) {
    int c = i / (ST_TI * ST_TI), rem = i % (ST_TI * ST_TI);
    int r = rem / ST_TI, q = rem % ST_TI;
    int gr = ir0 + r, gc = ic0 + q;
    float v = 0.f;
    if (gr >= 0 && gr < H && gc >= 0 && gc < W) v = load_pixel<U8>(img, imb + ((size_t)c * H + gr) * W + gc);
    sm.in[c][r][q] = v;
  }
  __syncthreads();

  // conv + BN + ReLU into the smem conv tile. item = (channel quarter, conv position)
  constexpr int NPOS = ST_TC * ST_TC;
  for (int item = tid; item < 4 * NPOS; item += ST_THREADS) {
    const int q = item / NPOS, pos = item % NPOS;
    const int pr = pos / ST_TC, pc = pos % ST_TC;
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = 0.f;
    const int cbase = q * 16;
#pragma unroll 1
    for (int c = 0; c < 3; ++c) {
#pragma unroll 1
      for (int r = 0; r < 7; ++r) {
        const float* irow = &sm.in[c][2 * pr + r][2 * pc];
        const float* wrow = &sm.w[c * 49 + r * 7][cbase];
#pragma unroll
        for (int s = 0; s < 7; ++s) {
          const float x = irow[s];
          const float4* w4 = reinterpret_cast<const float4*>(wrow + s * ST_CO);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float4 wv = w4[j];
            acc[4 * j + 0] = fmaf(x, wv.x, acc[4 * j + 0]);
            acc[4 * j + 1] = fmaf(x, wv.y, acc[4 * j + 1]);
            acc[4 * j + 2] = fmaf(x, wv.z, acc[4 * j + 2]);
            acc[4 * j + 3] = fmaf(x, wv.w, acc[4 * j + 3]);
          }
        }
      }
    }
    const int gr = cr0 + pr, gc = cc0 + pc;
    const bool valid = gr >= 0 && gr < Hc && gc >= 0 && gc < Wc;   // outside = maxpool padding
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      float v = fmaxf(fmaf(acc[j], sm.scale[cbase + j], sm.shift[cbase + j]), 0.f);
      sm.ct[pos][cbase + j] = valid ? v : -FLT_MAX;
    }
  }
  __syncthreads();

  // 3x3 s2 max-pool out of smem, NHWC store (16 threads x float4 = one pixel's 64 channels)
  const int c4 = (tid & 15) * 4;
  for (int pp = tid >> 4; pp < ST_TP * ST_TP; pp += ST_THREADS / 16) {
    const int py = pp / ST_TP, px = pp % ST_TP;
    const int oh = ph0 + py, ow = pw0 + px;
    if (oh >= Hp || ow >= Wp) continue;
    float4 m = make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX);
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        float4 v = *reinterpret_cast<const float4*>(&sm.ct[(2 * py + r) * ST_TC + 2 * px + s][c4]);
        m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
      }
    const size_t o = (((size_t)b * Hp + oh) * Wp + ow) * ST_CO + c4;
    if (OUT_DTYPE == COVA_F32) {
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(out0) + o) = m;
    } else {
      __nv_bfloat16 h[4], l[4];
      split_bf16(m.x, h[0], l[0]); split_bf16(m.y, h[1], l[1]);
      split_bf16(m.z, h[2], l[2]); split_bf16(m.w, h[3], l[3]);
      *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out0) + o) =
          make_uint2(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]));
      if (OUT_DTYPE == COVA_BF16X2)
        *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out1) + o) =
            make_uint2(pack_bf16x2(l[0], l[1]), pack_bf16x2(l[2], l[3]));
    }
  }
}

int stem_simt(const void* images, int img_u8, int B, int H, int W, const float* w, const float* bn_scale,
              const float* bn_shift, int out_dtype, void* out0, void* out1, cudaStream_t st) {
  const int Hc = (H + 6 - 7) / 2 + 1, Wc = (W + 6 - 7) / 2 + 1;
  const int Hp = (Hc + 2 - 3) / 2 + 1, Wp = (Wc + 2 - 3) / 2 + 1;
  dim3 grid(ceil_div(Wp, ST_TP), ceil_div(Hp, ST_TP), B);
  const int smem = (int)sizeof(StemSmem);
#define LAUNCH2(DT, U)                                                                                        \
  do {                                                                                                      \
    COVA_CUDA_OK(cudaFuncSetAttribute(stem_simt_kernel<DT, U>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
    stem_simt_kernel<DT, U><<<grid, ST_THREADS, smem, st>>>(images, B, H, W, w, bn_scale, bn_shift, out0, out1, Hc, \
                                                            Wc, Hp, Wp);                                    \
  } while (0)
#define LAUNCH(DT) do { if (img_u8) LAUNCH2(DT, true); else LAUNCH2(DT, false); } while (0)
  if (out_dtype == COVA_F32) LAUNCH(COVA_F32);
  else if (out_dtype == COVA_BF16) LAUNCH(COVA_BF16);
  else LAUNCH(COVA_BF16X2);
#undef LAUNCH
#undef LAUNCH2
  COVA_LAUNCH_OK();
  return COVA_OK;
}

}  // namespace cova
