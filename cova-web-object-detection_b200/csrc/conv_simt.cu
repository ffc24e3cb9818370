// A2 BasicBlock half, CUDA-core fp32 engine: 3x3 s1 p1 conv (64->64) + folded BN (+ residual) (+ ReLU),
// NHWC fp32 in / out.  Replaces `convnet.4.{b}.conv{1,2}` + `bn{1,2}` (torchvision BasicBlock.forward via
// `/root/reference/models.py:49-51`).  This is the exact-fp32 engine and the on-device check for the
// tcgen05 engine (conv_tc.cu); it is not the fast path.
//
// Persistent CTAs (grid = #SMs): the 9x64x64 fp32 filter (144 KB) is staged in shared memory once per CTA,
// then the CTA walks 8x16-pixel output tiles; each tile stages its 10x18x64 halo patch once.
#include "common.cuh"

namespace cova {

constexpr int CS_C = 64;
constexpr int CS_TH = 8, CS_TW = 16;
constexpr int CS_HH = CS_TH + 2, CS_HW = CS_TW + 2;
constexpr int CS_PITCH = 66;   // floats per staged pixel: 4*66 mod 32 = 8 -> the 4 pixel groups of a warp hit 4 banks
constexpr int CS_THREADS = 256;

struct ConvSimtSmem {
  float w[9][CS_C][CS_C];              // [tap][cin][cout]
  float in[CS_HH * CS_HW][CS_PITCH];
  float scale[CS_C], shift[CS_C];
};

__global__ void __launch_bounds__(CS_THREADS, 1)
conv3x3_simt_kernel(const float* __restrict__ x, int B, int H, int W, const float* __restrict__ wgt,
                    const float* __restrict__ bn_scale, const float* __restrict__ bn_shift,
                    const float* __restrict__ res, int relu, float* __restrict__ y) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ConvSimtSmem& sm = *reinterpret_cast<ConvSimtSmem*>(smem_raw);
  const int tid = threadIdx.x;
  {
    const float4* src = reinterpret_cast<const float4*>(wgt);
    float4* dst = reinterpret_cast<float4*>(&sm.w[0][0][0]);
    for (int i = tid; i < 9 * CS_C * CS_C / 4; i += CS_THREADS) dst[i] = __ldg(src + i);
    if (tid < CS_C) {
      sm.scale[tid] = bn_scale[tid];
      sm.shift[tid] = bn_shift[tid];
    }
  }
  const int tiles_w = ceil_div(W, CS_TW), tiles_h = ceil_div(H, CS_TH);
  const int n_tiles = B * tiles_h * tiles_w;
  const int pg = tid >> 3, cg = tid & 7;
  const int prow = pg >> 2, pcol = (pg & 3) * 4;

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int b = tile / (tiles_h * tiles_w);
    const int th = (tile / tiles_w) % tiles_h, tw = tile % tiles_w;
    const int h0 = th * CS_TH, w0 = tw * CS_TW;
    __syncthreads();   // previous tile's readers are done with sm.in (also orders the weight staging)
    for (int i = tid; i < CS_HH * CS_HW * (CS_C / 2); i += CS_THREADS) {
      const int pix = i >> 5, c2 = (i & 31) * 2;
      const int gh = h0 - 1 + pix / CS_HW, gw = w0 - 1 + pix % CS_HW;
      float2 v = make_float2(0.f, 0.f);
      if (gh >= 0 && gh < H && gw >= 0 && gw < W)
        v = __ldg(reinterpret_cast<const float2*>(x + (((size_t)b * H + gh) * W + gw) * CS_C + c2));
      *reinterpret_cast<float2*>(&sm.in[pix][c2]) = v;
    }
    __syncthreads();

    float acc[4][8];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[p][j] = 0.f;

#pragma unroll 1
    for (int r = 0; r < 3; ++r) {
      const float* in_row = &sm.in[(prow + r) * CS_HW + pcol][0];
#pragma unroll 2
      for (int ci = 0; ci < CS_C; ++ci) {
        float xv[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) xv[j] = in_row[j * CS_PITCH + ci];
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          const float4* w4 = reinterpret_cast<const float4*>(&sm.w[r * 3 + s][ci][cg * 8]);
          const float4 wa = w4[0], wb = w4[1];
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            const float xx = xv[p + s];
            acc[p][0] = fmaf(xx, wa.x, acc[p][0]); acc[p][1] = fmaf(xx, wa.y, acc[p][1]);
            acc[p][2] = fmaf(xx, wa.z, acc[p][2]); acc[p][3] = fmaf(xx, wa.w, acc[p][3]);
            acc[p][4] = fmaf(xx, wb.x, acc[p][4]); acc[p][5] = fmaf(xx, wb.y, acc[p][5]);
            acc[p][6] = fmaf(xx, wb.z, acc[p][6]); acc[p][7] = fmaf(xx, wb.w, acc[p][7]);
          }
        }
      }
    }

    const int oh = h0 + prow;
    if (oh < H) {
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const int ow = w0 + pcol + p;
        if (ow >= W) continue;
        const size_t o = (((size_t)b * H + oh) * W + ow) * CS_C + cg * 8;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = fmaf(acc[p][j], sm.scale[cg * 8 + j], sm.shift[cg * 8 + j]);
        if (res != nullptr) {
          const float4 ra = __ldg(reinterpret_cast<const float4*>(res + o));
          const float4 rb = __ldg(reinterpret_cast<const float4*>(res + o + 4));
          v[0] += ra.x; v[1] += ra.y; v[2] += ra.z; v[3] += ra.w;
          v[4] += rb.x; v[5] += rb.y; v[6] += rb.z; v[7] += rb.w;
        }
        if (relu) {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        *reinterpret_cast<float4*>(y + o) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(y + o + 4) = make_float4(v[4], v[5], v[6], v[7]);
      }
    }
  }
}

int conv3x3_simt(const float* x, int B, int H, int W, const float* w, const float* bn_scale, const float* bn_shift,
                 const float* res, int relu, float* y, cudaStream_t st) {
  const int smem = (int)sizeof(ConvSimtSmem);
  COVA_CUDA_OK(cudaFuncSetAttribute(conv3x3_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int n_tiles = B * ceil_div(H, CS_TH) * ceil_div(W, CS_TW);
  const int grid = n_tiles < sm_count() ? n_tiles : sm_count();
  conv3x3_simt_kernel<<<grid, CS_THREADS, smem, st>>>(x, B, H, W, w, bn_scale, bn_shift, res, relu, y);
  COVA_LAUNCH_OK();
  return COVA_OK;
}

// OIHW fp32 -> fp32 [kh][kw][Cin][Cout] (SIMT engine) and/or split-bf16 [kh*kw][Cout][Cin] (tcgen05 engine)
__global__ void pack_conv_weight_kernel(const float* __restrict__ w, int Cout, int Cin, int KH, int KW,
                                        float* __restrict__ simt_out, __nv_bfloat16* __restrict__ tc_hi,
                                        __nv_bfloat16* __restrict__ tc_lo) {
  const int n = Cout * Cin * KH * KW;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int s = i % KW, r = (i / KW) % KH, ci = (i / (KW * KH)) % Cin, co = i / (KW * KH * Cin);
    const float v = w[i];
    const int tap = r * KW + s;
    if (simt_out) simt_out[((size_t)tap * Cin + ci) * Cout + co] = v;
    if (tc_hi) {
      __nv_bfloat16 h, l;
      split_bf16(v, h, l);
      const size_t o = ((size_t)tap * Cout + co) * Cin + ci;
      tc_hi[o] = h;
      if (tc_lo) tc_lo[o] = l;
    }
  }
}

// OIHW fp32 -> fp16 [kh*kw][Cout][Cin] (single-plane fp16 mode, COVA_F16)
__global__ void pack_conv_weight_f16_kernel(const float* __restrict__ w, int Cout, int Cin, int KH, int KW,
                                            __half* __restrict__ out) {
  const int n = Cout * Cin * KH * KW;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int s = i % KW, r = (i / KW) % KH, ci = (i / (KW * KH)) % Cin, co = i / (KW * KH * Cin);
    out[((size_t)(r * KW + s) * Cout + co) * Cin + ci] = __float2half_rn(w[i]);
  }
}

// OIHW fp32 -> split-fp16 [kh*kw][Cout][Cin] planes of SPLIT_F16_WSCALE * w (COVA_F16X2 mode)
__global__ void pack_conv_weight_f16x2_kernel(const float* __restrict__ w, int Cout, int Cin, int KH, int KW,
                                              __half* __restrict__ hi, __half* __restrict__ lo) {
  const int n = Cout * Cin * KH * KW;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int s = i % KW, r = (i / KW) % KH, ci = (i / (KW * KH)) % Cin, co = i / (KW * KH * Cin);
    const float v = w[i] * SPLIT_F16_WSCALE;
    const __half h = __float2half_rn(v);
    const size_t o = ((size_t)(r * KW + s) * Cout + co) * Cin + ci;
    hi[o] = h;
    lo[o] = __float2half_rn(v - __half2float(h));
  }
}

}  // namespace cova

extern "C" int cova_pack_conv_weight_f16x2(const float* w_oihw, int Cout, int Cin, int kh, int kw, void* tc_hi, void* tc_lo,
                                           void* stream) {
  COVA_REQUIRE(w_oihw && tc_hi && tc_lo && Cout > 0 && Cin > 0 && kh > 0 && kw > 0, "cova_pack_conv_weight_f16x2: bad arguments");
  const int n = Cout * Cin * kh * kw;
  cova::pack_conv_weight_f16x2_kernel<<<cova::ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(
      w_oihw, Cout, Cin, kh, kw, (__half*)tc_hi, (__half*)tc_lo);
  COVA_LAUNCH_OK();
  return COVA_OK;
}

extern "C" int cova_pack_conv_weight_f16(const float* w_oihw, int Cout, int Cin, int kh, int kw, void* tc_f16,
                                         void* stream) {
  COVA_REQUIRE(w_oihw && tc_f16 && Cout > 0 && Cin > 0 && kh > 0 && kw > 0, "cova_pack_conv_weight_f16: bad arguments");
  const int n = Cout * Cin * kh * kw;
  cova::pack_conv_weight_f16_kernel<<<cova::ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(w_oihw, Cout, Cin, kh, kw,
                                                                                             (__half*)tc_f16);
  COVA_LAUNCH_OK();
  return COVA_OK;
}

extern "C" int cova_pack_conv_weight(const float* w_oihw, int Cout, int Cin, int kh, int kw, float* simt_out,
                                     void* tc_hi, void* tc_lo, void* stream) {
  COVA_REQUIRE(w_oihw && Cout > 0 && Cin > 0 && kh > 0 && kw > 0, "cova_pack_conv_weight: bad arguments");
  const int n = Cout * Cin * kh * kw;
  cova::pack_conv_weight_kernel<<<cova::ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(
      w_oihw, Cout, Cin, kh, kw, simt_out, (__nv_bfloat16*)tc_hi, (__nv_bfloat16*)tc_lo);
  COVA_LAUNCH_OK();
  return COVA_OK;
}
