// extern "C" entry points that dispatch between the engines (include/cova_b200.h).
#include "common.cuh"

namespace cova {
int stem_simt(const void* images, int img_u8, int B, int H, int W, const float* w, const float* bn_scale,
              const float* bn_shift, int out_dtype, void* out0, void* out1, cudaStream_t st);
int stem_tc(const void* images, int img_u8, int B, int H, int W, const void* w_packed, const float* bn_scale,
            const float* bn_shift, int out_dtype, void* out0, void* out1, cudaStream_t st);
int conv3x3_simt(const float* x, int B, int H, int W, const float* w, const float* bn_scale, const float* bn_shift,
                 const float* res, int relu, float* y, cudaStream_t st);
int conv3x3_tc(const void* x_hi, const void* x_lo, int split, int half, int B, int H, int W, const void* w_hi, const void* w_lo,
               const float* bn_scale, const float* bn_shift, const void* res_hi, const void* res_lo, int relu,
               int out_dtype, void* y0, void* y1, cudaStream_t st, double* stats = nullptr, const float* res_f32 = nullptr);
int linear_simt(const float* x, int64_t ld_x, int M, int K, const float* w, int N, const float* bias,
                const float* scale, const float* shift, const float* res, int64_t ld_res, int relu, float* y,
                int64_t ld_y, cudaStream_t st);
bool linear_tc_supported(const float* x, int64_t ld_x, int K);
bool linear_skinny_supported(const float* x, int64_t ld_x, int K, int N, const void* w);
int linear_skinny(const float* x, int64_t ld_x, int M, int K, const void* w, bool packed, int N, const float* bias,
                  const float* scale, const float* shift, const float* res, int64_t ld_res, int relu, float* y,
                  int64_t ld_y, cudaStream_t st);
int linear_tc(const float* x, int64_t ld_x, int M, int K, const void* w_packed, int N, const float* bias,
              const float* scale, const float* shift, const float* res, int64_t ld_res, int relu, int out_dtype,
              void* y0, void* y1, int64_t ld_y, cudaStream_t st, bool half);
}  // namespace cova

using namespace cova;

static bool dtype_ok(int d) {
  return d == COVA_F32 || d == COVA_BF16 || d == COVA_BF16X2 || d == COVA_F16 || d == COVA_F16X2;
}

extern "C" int cova_stem_fwd(const void* images, int img_dtype, int B, int H, int W, const void* w, const float* bn_scale,
                             const float* bn_shift, int out_dtype, void* out0, void* out1, int engine, void* stream) {
  COVA_REQUIRE(images && w && bn_scale && bn_shift && out0, "cova_stem_fwd: null pointer");
  COVA_REQUIRE(B > 0 && H >= 7 && W >= 7, "cova_stem_fwd: bad dims B=%d H=%d W=%d", B, H, W);
  COVA_REQUIRE(dtype_ok(out_dtype), "cova_stem_fwd: bad out_dtype %d", out_dtype);
  COVA_REQUIRE(img_dtype == COVA_F32 || img_dtype == COVA_U8, "cova_stem_fwd: images must be fp32 or uint8");
  const int u8 = img_dtype == COVA_U8;
  COVA_REQUIRE((out_dtype != COVA_BF16X2 && out_dtype != COVA_F16X2) || out1, "cova_stem_fwd: split output needs the lo plane");
  COVA_REQUIRE(engine == COVA_ENGINE_SIMT || engine == COVA_ENGINE_TCGEN05, "cova_stem_fwd: bad engine");
  COVA_REQUIRE((out_dtype != COVA_F16 && out_dtype != COVA_F16X2) || engine == COVA_ENGINE_TCGEN05,
               "cova_stem_fwd: fp16 planes are a tcgen05-engine mode");
  if (engine == COVA_ENGINE_TCGEN05)
    return stem_tc(images, u8, B, H, W, w, bn_scale, bn_shift, out_dtype, out0, out1, (cudaStream_t)stream);
  return stem_simt(images, u8, B, H, W, (const float*)w, bn_scale, bn_shift, out_dtype, out0, out1,
                   (cudaStream_t)stream);
}

extern "C" int cova_conv3x3_bn_act_stats_fwd(const void* x0, const void* x1, int dtype, int B, int H, int W, int Cin, int Cout,
                                             const void* w_a, const void* w_b, const float* bn_scale, const float* bn_shift,
                                             const void* res0, const void* res1, int relu, int out_dtype, void* y0, void* y1,
                                             int engine, double* stats_ws, void* stream);

// fp32 output + fp32 NHWC residual (training dgrad on split planes with the skip branch's fp32 gradient added in the epilogue)
extern "C" int cova_conv3x3_scale_res_f32_fwd(const void* x0, const void* x1, int dtype, int B, int H, int W, const void* w_a,
                                              const void* w_b, const float* scale, const float* shift, const float* res_f32,
                                              float* y, void* stream) {
  COVA_REQUIRE(x0 && x1 && w_a && w_b && scale && shift && res_f32 && y && B > 0 && H > 0 && W > 0, "cova_conv3x3_scale_res_f32_fwd: bad arguments");
  COVA_REQUIRE(dtype == COVA_BF16X2 || dtype == COVA_F16X2, "cova_conv3x3_scale_res_f32_fwd: split planes in");
  COVA_REQUIRE((((uintptr_t)res_f32 | (uintptr_t)y) & 31) == 0, "cova_conv3x3_scale_res_f32_fwd: 32-byte alignment");
  return conv3x3_tc(x0, x1, 1, dtype == COVA_F16X2, B, H, W, w_a, w_b, scale, shift, nullptr, nullptr, 0, COVA_F32, y, nullptr,
                    (cudaStream_t)stream, nullptr, res_f32);
}

extern "C" int cova_conv3x3_bn_act_fwd(const void* x0, const void* x1, int dtype, int B, int H, int W, int Cin, int Cout,
                                       const void* w_a, const void* w_b, const float* bn_scale, const float* bn_shift,
                                       const void* res0, const void* res1, int relu, int out_dtype, void* y0, void* y1,
                                       int engine, void* stream) {
  return cova_conv3x3_bn_act_stats_fwd(x0, x1, dtype, B, H, W, Cin, Cout, w_a, w_b, bn_scale, bn_shift, res0, res1, relu, out_dtype,
                                       y0, y1, engine, nullptr, stream);
}

extern "C" int cova_conv3x3_bn_act_stats_fwd(const void* x0, const void* x1, int dtype, int B, int H, int W, int Cin, int Cout,
                                             const void* w_a, const void* w_b, const float* bn_scale, const float* bn_shift,
                                             const void* res0, const void* res1, int relu, int out_dtype, void* y0, void* y1,
                                             int engine, double* stats_ws, void* stream) {
  COVA_REQUIRE(!stats_ws || (engine == COVA_ENGINE_TCGEN05 && !res0 && !relu),
               "cova_conv3x3_bn_act_stats_fwd: output statistics come with the tensor-core engine, no residual, no ReLU");
  COVA_REQUIRE(x0 && w_a && bn_scale && bn_shift && y0, "cova_conv3x3_bn_act_fwd: null pointer");
  COVA_REQUIRE(B > 0 && H > 0 && W > 0, "cova_conv3x3_bn_act_fwd: bad dims");
  COVA_REQUIRE(Cin == 64 && Cout == 64, "cova_conv3x3_bn_act_fwd: only Cin=Cout=64 is built (got %d->%d)", Cin, Cout);
  COVA_REQUIRE(dtype_ok(dtype) && dtype_ok(out_dtype), "cova_conv3x3_bn_act_fwd: bad dtype");
  cudaStream_t st = (cudaStream_t)stream;
  if (engine == COVA_ENGINE_SIMT) {
    COVA_REQUIRE(dtype == COVA_F32 && out_dtype == COVA_F32, "cova_conv3x3_bn_act_fwd: the SIMT engine is fp32 in/out");
    return conv3x3_simt((const float*)x0, B, H, W, (const float*)w_a, bn_scale, bn_shift, (const float*)res0, relu,
                        (float*)y0, st);
  }
  COVA_REQUIRE(engine == COVA_ENGINE_TCGEN05, "cova_conv3x3_bn_act_fwd: bad engine %d", engine);
  COVA_REQUIRE(dtype == COVA_BF16 || dtype == COVA_BF16X2 || dtype == COVA_F16 || dtype == COVA_F16X2,
               "cova_conv3x3_bn_act_fwd: tcgen05 engine takes bf16 / split-bf16 / fp16 / split-fp16 planes");
  const int split = dtype == COVA_BF16X2 || dtype == COVA_F16X2, half = dtype == COVA_F16 || dtype == COVA_F16X2;
  COVA_REQUIRE(out_dtype == COVA_F32 || out_dtype == dtype,
               "cova_conv3x3_bn_act_fwd: tcgen05 output is fp32 or the input's plane format");
  COVA_REQUIRE(!split || (x1 && w_b), "cova_conv3x3_bn_act_fwd: BF16X2 needs lo planes for x and w");
  COVA_REQUIRE(!split || !res0 || res1, "cova_conv3x3_bn_act_fwd: BF16X2 residual needs its lo plane");
  COVA_REQUIRE((out_dtype != COVA_BF16X2 && out_dtype != COVA_F16X2) || y1, "cova_conv3x3_bn_act_fwd: split output needs the lo plane");
  return conv3x3_tc(x0, x1, split, half, B, H, W, w_a, w_b, bn_scale, bn_shift, res0, res1, relu, out_dtype, y0, y1, st, stats_ws);
}

extern "C" int cova_linear_fwd(const float* x, int64_t ld_x, int M, int K, const void* w, int N, const float* bias,
                               const float* scale, const float* shift, const float* res, int64_t ld_res, int relu,
                               int out_dtype, void* y0, void* y1, int64_t ld_y, int engine, void* stream) {
  COVA_REQUIRE(M >= 0 && K > 0 && N > 0, "cova_linear_fwd: bad dims");
  if (M == 0) return COVA_OK;
  COVA_REQUIRE(x && w && y0 && ld_x >= K && ld_y >= N, "cova_linear_fwd: bad arguments");
  COVA_REQUIRE((scale == nullptr) == (shift == nullptr), "cova_linear_fwd: scale/shift must come together");
  COVA_REQUIRE(engine == COVA_ENGINE_SIMT || engine == COVA_ENGINE_TCGEN05 || engine == COVA_ENGINE_TCGEN05_F16X2,
               "cova_linear_fwd: bad engine");
  COVA_REQUIRE(!res || ld_res >= N, "cova_linear_fwd: ld_res too small");
  COVA_REQUIRE(out_dtype == COVA_F32 || (out_dtype == COVA_BF16X2 && y1 && engine == COVA_ENGINE_TCGEN05),
               "cova_linear_fwd: output is fp32, or split-bf16 planes (y0 = hi, y1 = lo) on the tcgen05 engine");
  // skinny outputs (decoder.5: N = 4): exact-fp32 warp-per-row kernel on either engine (the split-fp16 training engine keeps its path)
  if (engine != COVA_ENGINE_TCGEN05_F16X2 && out_dtype == COVA_F32 && linear_skinny_supported(x, ld_x, K, N, w) &&
      (engine == COVA_ENGINE_SIMT || K % 8 == 0))
    return linear_skinny(x, ld_x, M, K, w, engine == COVA_ENGINE_TCGEN05, N, bias, scale, shift, res, ld_res, relu, (float*)y0, ld_y,
                         (cudaStream_t)stream);
  if (engine != COVA_ENGINE_SIMT) {
    COVA_REQUIRE(linear_tc_supported(x, ld_x, K),
                 "cova_linear_fwd: the tcgen05 engine needs K %% 8 == 0, ld_x %% 4 == 0 and a 16-byte aligned x "
                 "(K=%d, ld_x=%lld); use the SIMT engine for this shape", K, (long long)ld_x);
    return linear_tc(x, ld_x, M, K, w, N, bias, scale, shift, res, ld_res, relu, out_dtype, y0, y1, ld_y,
                     (cudaStream_t)stream, engine == COVA_ENGINE_TCGEN05_F16X2);
  }
  return linear_simt(x, ld_x, M, K, (const float*)w, N, bias, scale, shift, res, ld_res, relu, (float*)y0, ld_y,
                     (cudaStream_t)stream);
}
