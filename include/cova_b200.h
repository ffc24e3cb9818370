/*
 * cova_b200.h - C ABI of libcova_b200.so: the B200 (sm_100a) kernels behind CoVA's per-webpage
 * forward hot path.  The reference (`/root/reference`, pure Python) has no FFI of its own; each entry
 * point below replaces the torch/torchvision call the reference makes at the cited `models.py` line
 * (SURVEY.md section 8(a)/(b)).  The reference-side binding is a ctypes stub - see INTEGRATION.md.
 *
 * Conventions
 *  - Every pointer is a DEVICE pointer unless named `h_*`.  The caller (PyTorch) owns all memory:
 *    inputs, outputs and workspaces.  The library never allocates, frees or retains a pointer.
 *  - `stream` is a `cudaStream_t` passed as `void*` (pass `torch.cuda.current_stream().cuda_stream`).
 *    Launches are asynchronous; no entry point synchronises.
 *  - Return value: 0 = success, otherwise a COVA_ERR_* code; `cova_last_error()` returns a
 *    thread-local message.  Nothing throws or exits across this boundary.
 *  - Activation tensors are NHWC.  `dtype`:
 *      COVA_F32   one fp32 plane (`p0`; `p1` ignored)
 *      COVA_BF16  one bf16 plane (`p0`)
 *      COVA_BF16X2 split-bf16: `p0` = hi plane = bf16(x), `p1` = lo plane = bf16(x - hi).  hi+lo carries
 *                 ~16 mantissa bits; products hi*Whi + lo*Whi + hi*Wlo on tcgen05 tensor cores with fp32
 *                 accumulate reproduce an fp32 convolution to ~1e-5 (the "fp32-parity" tensor-core mode).
 *      COVA_F16   one fp16 plane (`p0`): the single-product throughput mode of the tcgen05 engine.  fp16 keeps
 *                 11 significand bits (bf16: 8), which is what brings a one-product pipeline inside the 1e-3 bar
 *                 (measured ~5e-4 on the logits); activations must stay below 65504 (post-BN/ReLU maps do).
 *      COVA_F16X2 split-fp16: `p0` = f16(x), `p1` = f16(x - p0): 22 significand bits, so the same three tensor-core
 *                 products reproduce an fp32 convolution to ~1e-6 (split-bf16: ~1e-5) at the same cost ("fp16x3").
 *                 Filters of this mode are packed pre-scaled by 256 (undone exactly in the kernels' epilogue) so that
 *                 their lo plane stays a normal fp16 number.
 *  - Eval-mode BatchNorm is passed folded: y = x*scale[c] + shift[c]
 *    (scale = weight/sqrt(running_var+eps), shift = bias - running_mean*scale).
 */
#ifndef COVA_B200_H
#define COVA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define COVA_ABI_VERSION 1

enum { COVA_OK = 0, COVA_ERR_ARG = 1, COVA_ERR_CUDA = 2, COVA_ERR_UNSUPPORTED = 3 };
enum { COVA_F32 = 0, COVA_BF16 = 1, COVA_BF16X2 = 2, COVA_U8 = 3 /* images only */, COVA_F16 = 4, COVA_F16X2 = 5 };
enum { COVA_ENGINE_SIMT = 0, COVA_ENGINE_TCGEN05 = 1,
       COVA_ENGINE_TCGEN05_F16X2 = 2 /* cova_linear_fwd only: split-fp16 operands, w from cova_pack_conv_weight_f16x2(w, N, K, 1, 1) */ };

int cova_abi_version(void);
const char* cova_last_error(void);
/* SM count / max dynamic smem of the current device (host query; used to size persistent grids). */
int cova_device_info(int* sm_count, int* max_smem_optin);

/* ---- Tuning / diagnostics (process-global, host side; not needed for correct results).
 * cova_set_knob: value < 0 restores the built-in default.
 *   COVA_KNOB_CONV_L2_PREFETCH   3x3 conv: tiles ahead for which the TMA producer issues an L2 prefetch of the halo
 *                                boxes (0 = off)
 *   COVA_KNOB_CONV_RES_PREFETCH  3x3 conv: tiles ahead for which the epilogue L2-prefetches the residual rows
 *   COVA_KNOB_STEM_L2_PREFETCH   stem: image rows ahead to L2-prefetch (0 = off)
 *   COVA_KNOB_STEM_CONVERTERS    stem: converter warps per CTA (4 or 8)
 *   COVA_KNOB_CONV_RES_LOAD      3x3 conv residual loads: 0 = ld.global.nc, 1 = ld.global, 2 = L1::no_allocate
 *   COVA_KNOB_ROI_ROWSPLIT       RoIPool: 1 = one CTA per (box, row bin) (default), 0 = one CTA per box
 *   COVA_KNOB_WGRAD_DRAIN        wgrad kernels: pixel tiles accumulated in TMEM between two drains to the global fp32 sum
 *   COVA_KNOB_STEM_U8_EXACT      stem, uint8 images: 1 = v/255 split through the look-up table (bit-identical to ToTensor, three
 *                                products); 0 (default) = integer pixels, exact in one 16-bit plane, 1/255 folded into the
 *                                epilogue scale: TWO products (stem output within one ulp of its split-bf16 format, 3e-5 absolute, of the ToTensor path)
 * cova_debug_buffer: a caller-owned device array of uint64 words; kernels that support it (the 3x3 tensor-core conv:
 *   8 words per CTA = cycles the MMA issuer waited for operands / for a free accumulator, the TMA producer for a free
 *   ring slot, epilogue warp 2 for a finished accumulator, CTA total, tiles) add their counters.  NULL disables. */
enum { COVA_KNOB_CONV_L2_PREFETCH = 0, COVA_KNOB_CONV_RES_PREFETCH = 1, COVA_KNOB_STEM_L2_PREFETCH = 2,
       COVA_KNOB_STEM_CONVERTERS = 5, COVA_KNOB_CONV_RES_LOAD = 6, COVA_KNOB_ROI_ROWSPLIT = 7, COVA_KNOB_WGRAD_DRAIN = 8,
       COVA_KNOB_STEM_U8_EXACT = 9, COVA_KNOB_COUNT = 10 };
int cova_set_knob(int id, int value);
int cova_debug_buffer(void* dev_words, int64_t n_words);

/* ---- A2: backbone stem.  Replaces `convnet[0:4]` = conv1 7x7 s2 p3 (no bias) -> bn1 -> relu ->
 * maxpool 3x3 s2 p1 (`models.py:49-51`, applied `models.py:125`).  ONE fused kernel: the
 * [B,64,H/2,W/2] conv output never reaches HBM.
 *   images  [B,3,H,W] NCHW; img_dtype COVA_F32 = fp32 in [0,1] (the reference's input contract, `models.py:96`)
 *           or COVA_U8 = raw 8-bit pixels at a quarter of the host->device bytes.  SIMT engine and
 *           COVA_KNOB_STEM_U8_EXACT = 1: converted as v/255 (IEEE division) on the way into shared memory, bit-identical
 *           to `torchvision.transforms.ToTensor` (`datasets.py:41-45`); TCGEN05 default: integer pixels, 1/255 in the
 *           epilogue scale (two tensor products instead of three; within one output ulp of the ToTensor path)
 *   w       engine SIMT   : [64,3,7,7] fp32 OIHW (`convnet.0.weight`)
 *           engine TCGEN05: the split-bf16 K-chunked filter written by cova_pack_stem_weight
 *   bn_scale/bn_shift [64] folded `convnet.1`
 *   out     NHWC [B,H/4,W/4,64] in `out_dtype` (planes out0/out1).  TCGEN05: COVA_BF16 / COVA_F16 output selects
 *           the single-product bf16 / fp16 mode, COVA_BF16X2 / COVA_F32 the 3-product split-bf16 mode, COVA_F16X2
 *           the 3-product split-fp16 mode (filter from cova_pack_stem_weight_f16x2).                      */
int cova_stem_fwd(const void* images, int img_dtype, int B, int H, int W, const void* w, const float* bn_scale,
                  const float* bn_shift, int out_dtype, void* out0, void* out1, int engine, void* stream);

/* Training mode (A9): conv1 ALONE in the fp32-parity tensor-core mode -> out [B, Hc, Wc, 64] fp32 NHWC (Hc = (H-1)/2+1),
 * no BN / ReLU / pooling: BatchNorm with batch statistics needs this tensor itself (`train.py:27`).  w = the
 * cova_pack_stem_weight filter.                                                                          */
int cova_stem_conv_raw_fwd(const void* images, int img_dtype, int B, int H, int W, const void* w_packed, int w_dtype,
                           float* out, void* stream);   /* w_dtype: COVA_BF16X2 (cova_pack_stem_weight) or COVA_F16X2 */

/* Same layout for the fp16 mode (out_dtype COVA_F16): plane 0 = fp16(w), plane 1 = 0. */
int cova_pack_stem_weight_f16(const float* w_oihw, void* packed, void* stream);
/* Same layout for the split-fp16 mode (out_dtype COVA_F16X2): planes of 256*w. */
int cova_pack_stem_weight_f16x2(const float* w_oihw, void* packed, void* stream);

/* OIHW fp32 [64,3,7,7] -> tcgen05 stem filter: bf16 [28 K-chunks][2 planes (hi, lo)][64 cout][8],
 * K index = r*32 + s*4 + c with zero weights at s = 7 and c = 3 (57,344 bytes).                       */
int cova_pack_stem_weight(const float* w_oihw, void* packed, void* stream);

/* ---- A2: 3x3 s1 p1 convolution + folded BN (+ residual) (+ ReLU): one BasicBlock half
 * (torchvision BasicBlock.forward; `convnet.4.{b}.conv{1,2}` + `bn{1,2}`), Cin = Cout = 64.
 *   x / res / y : NHWC [B,H,W,64]; `dtype` applies to x and res, `out_dtype` to y.
 *   w: engine SIMT   -> fp32 [3][3][Cin][Cout]           (repacked from OIHW by cova_pack_conv_weight)
 *      engine TCGEN05-> bf16 hi/lo [9][Cout][Cin] K-major (w_hi, w_lo; w_lo may be NULL for COVA_BF16)  */
int cova_conv3x3_bn_act_fwd(const void* x0, const void* x1, int dtype, int B, int H, int W, int Cin, int Cout,
                            const void* w_a, const void* w_b, const float* bn_scale, const float* bn_shift,
                            const void* res0, const void* res1, int relu, int out_dtype, void* y0, void* y1,
                            int engine, void* stream);

/* ---- A2 (ResNet-50): 1x1 convolution + folded BN (+ residual) (+ ReLU) over NHWC split-bf16 planes =
 * torchvision Bottleneck.forward conv1 / conv3 / downsample.  x planes [M, Cin], M = B*H*W pixel rows;
 * w = cova_pack_linear_weight of the [Cout, Cin] filter; residual planes [M, Cout]; output split planes
 * (COVA_BF16X2: y0 = hi, y1 = lo) or fp32 (COVA_F32: y0).  Built for 64->64, 64->256, 256->64.          */
int cova_conv1x1_bn_act_fwd(const void* x_hi, const void* x_lo, int64_t M, int Cin, int Cout, const void* w_packed,
                            const float* bn_scale, const float* bn_shift, const void* res_hi, const void* res_lo,
                            int relu, int out_dtype, void* y0, void* y1, void* stream);

/* Repack an OIHW fp32 conv weight [Cout,Cin,kh,kw] for the engines above.
 *   simt_out  : fp32 [kh][kw][Cin][Cout]                       (may be NULL)
 *   tc_hi/lo  : bf16 [kh*kw][Cout][Cin] hi/lo split            (may be NULL)                         */
int cova_pack_conv_weight(const float* w_oihw, int Cout, int Cin, int kh, int kw, float* simt_out,
                          void* tc_hi, void* tc_lo, void* stream);

/* OIHW fp32 -> fp16 [kh*kw][Cout][Cin] for the COVA_F16 mode of cova_conv3x3_bn_act_fwd (w_a = this, w_b = NULL). */
int cova_pack_conv_weight_f16(const float* w_oihw, int Cout, int Cin, int kh, int kw, void* tc_f16, void* stream);
/* OIHW fp32 -> split-fp16 planes [kh*kw][Cout][Cin] of 256*w for the COVA_F16X2 mode (w_a = hi, w_b = lo). */
int cova_pack_conv_weight_f16x2(const float* w_oihw, int Cout, int Cin, int kh, int kw, void* tc_hi, void* tc_lo,
                                void* stream);

/* ---- A4: RoIPool.  Replaces `torchvision.ops.RoIPool(P, scale)` (`models.py:58`, `:125-127`).
 * Bit-exact in fp32.  fm NHWC fp32 [B,Hf,Wf,C]; rois [T,5] fp32 = [batch_idx,x1,y1,x2,y2] image pixels.
 * Writes out[t*ld_out + c*PH*PW + ph*PW + pw] (the `.view(T, C*P*P)` flatten order of `models.py:125-127`).
 * argmax (optional, int32 [T,C,PH,PW], flat h*Wf+w index or -1) is what the backward needs.
 * mode 0 = RoIPool; mode 1 = RoIAlign(P, scale, sampling_ratio, aligned=False) (A4').      */
int cova_roi_fwd(const float* fm, int B, int Hf, int Wf, int C, const float* rois, int T, int PH, int PW,
                 float spatial_scale, int mode, int sampling_ratio, float* out, int64_t ld_out, int32_t* argmax,
                 void* stream);

/* ---- A4 backward (A9, `train.py:59`): scatter-add of the pooled gradients to their arg-max pixels
 * (torchvision roi_pool backward).  grad_fm: NHWC fp32 [B,Hf,Wf,C], ZEROED by the caller; argmax from cova_roi_fwd. */
int cova_roi_pool_bwd(const float* grad_out, int64_t ld_go, const int32_t* argmax, const float* rois, int T, int C,
                      int PH, int PW, int B, int Hf, int Wf, float* grad_fm, void* stream);

/* ---- A4' backward (D1 variant on the train path): torchvision roi_align backward - every output bin adds
 * grad / (gh*gw) x bilinear weight to the 4 taps of each of its samples (fp32 vector atomics).  grad_fm: NHWC fp32
 * [B,Hf,Wf,C], ZEROED by the caller, 16-byte aligned; C a multiple of 64; aligned=False semantics as the forward. */
int cova_roi_align_bwd(const float* grad_out, int64_t ld_go, const float* rois, int T, int C, int PH, int PW,
                       float spatial_scale, int sampling_ratio, int B, int Hf, int Wf, float* grad_fm, void* stream);

/* ---- A5: positional encoder.  Replaces `_get_bbox_features` + `bbox_feat_encoder`
 * (`models.py:129-148`, `:65-70`): [x1,y1,w,h,w/h] -> Linear(5,D) -> folded BN1d -> ReLU.
 * Writes out[t*ld_out + d], d < D (so it can land at column C*P*P of the `own` row: the concat of
 * `models.py:110` costs no extra pass).                                                             */
int cova_bbox_enc_fwd(const float* rois, int T, const float* w /*[D,5]*/, const float* b /*[D]*/,
                      const float* bn_scale, const float* bn_shift, int D, float* out, int64_t ld_out, void* stream);

/* ---- A5b: `bn_additional_feat` (folded BN1d, or identity when scale == NULL) + concat column copy
 * (`models.py:72-75`, `:109-110`).                                                                  */
int cova_affine_cols_fwd(const float* x, int T, int D, int64_t ld_x, const float* scale, const float* shift,
                         float* out, int64_t ld_out, void* stream);

/* ---- generic row-major linear layer  Y[M,N] = act((X[M,K] @ W[N,K]^T + bias) * scale + shift + res)
 * (`nn.Linear` = `models.py:160-164` W_i/W_j, `:85`, `:89`; also the 1x1 convolutions of the ResNet-50
 * Bottleneck on NHWC activations, where `res` is the identity branch); bias/scale/shift/res may be NULL.
 *   w: engine SIMT -> fp32 [N,K];  engine TCGEN05 -> split-bf16 [2][N][K] from cova_pack_linear_weight
 *      (needs K % 8 == 0, ld_x % 4 == 0, 16-byte aligned x; otherwise COVA_ERR_ARG - call the SIMT engine);
 *      engine TCGEN05_F16X2 -> split-fp16 [2][N][K] of 256*w (cova_pack_conv_weight_f16x2 with kh = kw = 1): 22 instead
 *      of 16 significand bits, same cost - the training forward (x must stay below 65504).                          */
int cova_linear_fwd(const float* x, int64_t ld_x, int M, int K, const void* w, int N, const float* bias,
                    const float* scale, const float* shift, const float* res, int64_t ld_res, int relu,
                    int out_dtype, void* y0, void* y1, int64_t ld_y, int engine, void* stream);
/*   out_dtype COVA_F32: y0 = fp32 [M, ld_y];  COVA_BF16X2 (tcgen05 engine only): y0 / y1 = bf16 hi / lo planes
 *   (ld_y in elements) - feeds cova_conv3x3_bn_act_fwd directly (ResNet-50 conv1 -> conv2).                 */

/* fp32 [N,K] -> bf16 [2][N][K] (hi plane = bf16(w), lo plane = bf16(w - hi)) for the tcgen05 linear engine. */
int cova_pack_linear_weight(const float* w, int N, int K, void* packed, void* stream);

/* ---- A6: graph-attention gather.  Replaces `GraphAttentionLayer.forward` lines `models.py:180-208`
 * after the once-per-node projections (algebraically identical restructuring):
 *   whj [T,Hd] = W_j h (ld_whj), s_i = s[i*ld_st] = a_i . W_i h, t_i = t[i*ld_st] = a_j . W_j h, att_b = bias
 *   e_ik = LeakyReLU_alpha(s_i + t_c(i,k) + att_b); masked (c<0) -> -9e15; softmax over K;
 *   out_i = sum_k alpha_ik whj[c(i,k)]   (c<0 rows contribute exactly 0)
 * ctx_idx: int64 [T,K] batch-global row ids or -1 (`datasets.py:117-128`, `:175`).
 * attn (optional) [T,K] = the `return_attn_wts=True` output (`models.py:210-211`).                  */
int cova_gat_fwd(const float* whj, int64_t ld_whj, const float* s, const float* t, int64_t ld_st, float att_b,
                 float alpha, const int64_t* ctx_idx, int T, int K, int Hd, float* out, int64_t ld_out, float* attn,
                 void* stream);

/* Multi-head form (one launch, grid.y = head): ext [T, ld_ext] = [whj_0 .. whj_{H-1} (Hd columns each) |
 * s_0 t_0 s_1 t_1 .. | pad] as produced by ONE projection GEMM; h_att_b = HOST array of the H attention biases;
 * out [T, H*Hd] (heads concatenated on dim 1); attn (optional) [H, T, K].                                      */
int cova_gat_multihead_fwd(const float* ext, int64_t ld_ext, int Hd, int n_heads, const float* h_att_b, float alpha,
                           const int64_t* ctx_idx, int T, int K, float* out, int64_t ld_out, float* attn, void* stream);

/* ---- A6 backward (A9): gradients of cova_gat_fwd w.r.t. whj, s, t and the bias, given grad_out [T,Hd] and the
 * attention weights saved by the forward.  d_whj / d_t / d_b are accumulated with atomics and must be ZEROED by the
 * caller; d_s is overwritten.  d_s / d_t: element i at [i*ld_dst].                                           */
int cova_gat_bwd(const float* grad_out, int64_t ld_go, const float* whj, int64_t ld_whj, const float* s, const float* t,
                 int64_t ld_st, float att_b, float alpha, const int64_t* ctx_idx, const float* attn, int T, int K, int Hd,
                 float* d_whj, int64_t ld_dw, float* d_s, float* d_t, int64_t ld_dst, float* d_b, void* stream);

/* ================= callers either side of the forward ================= */

/* ---- A9 / N2: `nn.CrossEntropyLoss(reduction="sum")` (`main.py:139`, `train.py:56`) forward AND gradient w.r.t. the
 * logits in one pass, plus the `(output.argmax(1) == labels).sum()` of `train.py:53-54`.
 *   logits [T, n_cls] fp32 (row stride ld), labels int64 [T]; rows whose label equals ignore_index (torch default -100)
 *   or lies outside [0, n_cls) contribute neither loss nor gradient.
 *   ws_acc : caller-owned scratch, one double (device); loss: [1] fp32, overwritten
 *   dlogits: optional [T, n_cls] (row stride ld_d) = softmax(logits) - onehot(labels)  (d loss / d logits)
 *   n_correct: optional [1] int32, overwritten.                                                          */
int cova_ce_sum_fwd_bwd(const float* logits, int64_t ld, const int64_t* labels, int T, int n_cls, int64_t ignore_index,
                        double* ws_acc, float* loss, float* dlogits, int64_t ld_d, int* n_correct, void* stream);

/* ---- A9 / N2: one `torch.optim.Adam` step (amsgrad off; L2 weight decay folded into the gradient, `main.py:133-135`,
 * `train.py:60`) over ONE flat fp32 bucket of n elements - the bucket the NCCL gradient all-reduce uses, so a whole
 * model steps in one launch.  grad is multiplied by grad_scale first (1.0; 1/world for mean-reduced data parallelism).
 * Hyper-parameters are doubles (torch keeps them as Python floats and derives 1-beta, the bias corrections and
 * lr/bc1 in double).  `step` is the 1-based step count AFTER the increment, as in torch.                                     */
int cova_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, double lr, double beta1,
                   double beta2, double eps, double weight_decay, int step, double grad_scale, void* stream);

/* ---- N3: the top-k test of `evaluate_model` (`train.py:131-154`) for every page and class in one launch.
 *   page_offsets int32 [B+1]: rows of page b are [page_offsets[b], page_offsets[b+1]) (pages are contiguous row ranges,
 *   `datasets.py:170-181`).  hits int32 [B, n_cls]: column c >= 1 = 1 if the first row labelled c is among the k rows
 *   with the largest logit[:, c] of its page (ties ranked as a stable ascending argsort), 0 if not, -1 if the page has
 *   no row labelled c (the reference raises there); column 0 is not written.                             */
int cova_topk_hits(const float* logits, int64_t ld, const int64_t* labels, const int* page_offsets, int B, int n_cls,
                   int k, int* hits, void* stream);

/* ---- N1: batch assembly on the device from per-page box counts: the +-context_size pre-order window of
 * `WebDataset.__getitem__` (`datasets.py:117-128`: left neighbours, then right neighbours, right-padded with -1) made
 * batch-global as `custom_collate_fn` does (`datasets.py:175`), and - when raw [x,y,w,h] boxes are given - the collated
 * box rows [page, x1, y1, x1+w, y1+h] (`datasets.py:114-115`, `:172-174`).
 *   context_indices int64 [T, 2*context_size]; boxes_xywh [T,4] / bboxes [T,5] both NULL or both set.      */
int cova_build_batch(const int* page_offsets, int B, int T, int context_size, const float* boxes_xywh, float* bboxes,
                     int64_t* context_indices, void* stream);

/* ---- A2 / A9, training mode: BatchNorm2d with BATCH statistics (+ residual) (+ ReLU) on NHWC fp32 maps x [M, C]
 * (M = B*H*W, C a power of two in [4, 1024]) - what torchvision's BasicBlock / Bottleneck run through nn.BatchNorm2d
 * under `model.train()` (`train.py:27`): biased variance for the normalisation, running statistics updated with
 * momentum on the unbiased variance.  ws: caller-owned scratch of 2*C doubles.
 *   cova_bn_train_stats    ws = [sum x | sum x^2]
 *   cova_bn_train_finalize mean, invstd = 1/sqrt(var+eps) (fp32 [C]); running_mean / running_var updated in place (or NULL)
 *   cova_bn_act_fwd        y = [relu]((x - mean) * invstd * gamma + beta [+ res]); relu_mask (optional, with relu): [M * C / 4]
 *                          bytes, bit k of byte i = [y > 0] of element 4 i + k, for cova_bn_act_bwd_planes (which then does not read res)
 *   cova_bn_act_bwd        g = dy * [y > 0] (y recomputed);  dx = gamma*invstd*(g - mean(g) - xhat*mean(g*xhat));
 *                          dres = g (optional);  dgamma = sum g*xhat;  dbeta = sum g                              */
int cova_bn_train_stats(const float* x, int64_t M, int C, double* ws, void* stream);
int cova_bn_train_finalize(const double* ws, int64_t M, int C, float eps, float momentum, float* mean, float* invstd,
                           float* running_mean, float* running_var, void* stream);
int cova_bn_act_fwd(const float* x, int64_t M, int C, const float* mean, const float* invstd, const float* gamma,
                    const float* beta, const float* res, int relu, float* y, void* y_hi, void* y_lo, int planes_dtype,
                    unsigned char* relu_mask, void* stream);
/*   y_hi / y_lo (optional, both or neither): the same result as split planes in `planes_dtype` (COVA_BF16X2 or
 *   COVA_F16X2) - the operand format of the tensor-core convolution that consumes it; y may be NULL when only the
 *   planes are wanted (a map whose single consumer is a tensor-core convolution).                                   */
int cova_bn_act_bwd(const float* dy, const float* x, const float* res, int64_t M, int C, const float* mean,
                    const float* invstd, const float* gamma, const float* beta, int relu, double* ws, float* dx,
                    float* dres, float* dgamma, float* dbeta, void* stream);
/* The same backward with dx emitted directly as SCALED split planes (hi, lo) of dx * s for the tensor-core dgrad / wgrad
 * kernels that consume it (no fp32 gradient map, no separate max / split passes): s = the power of two that brings an upper
 * bound of max|dx| (from per-channel max|x - mean| and max|g| gathered in the reduction pass) into
 * [2^target_log2, 2^(target_log2+1)); inv_scale_vec receives 256 copies of 1/s.  ws: 2*C doubles, ws_max: C+1 words.  */
int cova_bn_act_bwd_planes(const float* dy, const float* x, const float* res, int64_t M, int C, const float* mean,
                           const float* invstd, const float* gamma, const float* beta, int relu, double* ws,
                           unsigned int* ws_max, void* dx_hi, void* dx_lo, int planes_dtype, int target_log2,
                           float* inv_scale_vec, float* dres, float* dgamma, float* dbeta, const unsigned char* relu_mask, void* stream);

/* ---- A2 / A9: the stem's `nn.MaxPool2d(3, 2, 1)` on NHWC fp32 maps, forward and backward (the gradient of an output
 * goes to the FIRST maximum of its window in row-major scan order, as torch's max_pool2d_with_indices).
 * x [B,H,W,C] -> y [B,(H-1)/2+1,(W-1)/2+1,C]; code (optional, uint8, same shape as y) = position r*3+s of the winner
 * inside the window, which is all the backward needs; y_hi / y_lo optional split-bf16 planes of y.
 * dx [B,H,W,C] is overwritten (gather over the <= 2x2 windows containing a pixel: no atomics, no rescans).          */
int cova_maxpool3x3s2_fwd(const float* x, int B, int H, int W, int C, float* y, unsigned char* code, void* y_hi,
                          void* y_lo, int planes_dtype, void* stream);
int cova_maxpool3x3s2_bwd(const unsigned char* code, const float* dy, int B, int H, int W, int C, float* dx, void* stream);

/* ---- A2 / A9: the stem's bn1 + ReLU + maxpool FUSED for the training path: the normalised [B,H,W,C] map between them
 * (1.7 GB at B=16) is never written.  Forward: mean / invstd from cova_bn_train_stats + _finalize of the raw conv1
 * output x; y [B,Ho,Wo,C] = maxpool3x3s2p1(relu(bn(x))), winner codes, optional split planes.  Backward: from the pooled
 * gradient and the codes, dx of the raw conv1 output and dgamma / dbeta (ws: 2*C doubles of scratch).              */
int cova_bn_relu_pool_fwd(const float* x, int B, int H, int W, int C, const float* mean, const float* invstd,
                          const float* gamma, const float* beta, float* y, unsigned char* code, void* y_hi, void* y_lo,
                          int planes_dtype, void* stream);
int cova_bn_relu_pool_bwd(const float* x, const unsigned char* code, const float* dy_pooled, int B, int H, int W, int C,
                          const float* mean, const float* invstd, const float* gamma, const float* beta, double* ws,
                          float* dx, float* dgamma, float* dbeta, void* stream);

/* fp32 [n] -> split planes hi = r(x), lo = r(x - hi), r = bf16 (COVA_BF16X2) or fp16 (COVA_F16X2) rounding (n % 4 == 0). */
int cova_split_planes(const float* x, int64_t n, void* hi, void* lo, int planes_dtype, void* stream);

/* The same with a per-tensor power-of-two scale chosen on the device, for GRADIENT maps (their magnitude shrinks as the
 * sum-reduced loss converges; unscaled split-fp16 planes carry an absolute floor of 2^-25): s = 2^(target_log2 -
 * floor(log2(max|x|))), hi/lo = split(x * s), so max|x*s| lies in [2^target_log2, 2^(target_log2+1)).  The consumer
 * undoes it exactly with inv_scale_vec (256 floats, all = 1/s), e.g. as the epilogue `bn_scale` of the dgrad
 * convolution or the `inv_scale` of the wgrad kernels.  ws: 4 bytes of device scratch (the max|x| bits).
 * All-zero / non-finite maxima give s = 1.                                                                            */
int cova_split_planes_scaled(const float* x, int64_t n, void* hi, void* lo, int planes_dtype, int target_log2,
                             unsigned int* ws, float* inv_scale_vec, void* stream);

/* ---- A9 (`loss.backward()`, `train.py:59`): weight gradient of a 3x3 s1 p1 64->64 convolution (torchvision BasicBlock
 * conv1/conv2, Bottleneck conv2) on the tensor cores, contraction over the B*H*W pixels of NHWC split planes:
 *   dw[co][ci][r][s] = inv_scale * sum_{b,h,w} x[b,h+r-1,w+s-1,ci] * dy[b,h,w,co]     (zero padding)
 * x_hi/x_lo: split planes of the convolution's input (what the forward consumed); dy_hi/dy_lo: split planes of the output
 * gradient (cova_split_planes_scaled); inv_scale: device float (1/s of the dy planes) or NULL; ws: 9*64*64 floats of
 * device scratch (zeroed here); dw_oihw: [64,64,3,3] fp32, overwritten.                                              */
int cova_conv3x3_wgrad(const void* x_hi, const void* x_lo, const void* dy_hi, const void* dy_lo, int B, int H, int W,
                       int planes_dtype, const float* inv_scale, float* ws, float* dw_oihw, void* stream);

/* ---- A9 / SURVEY D2 (ResNet-50 Bottleneck 1x1 convolutions on the train path, torchvision `Bottleneck.forward`):
 * cova_conv1x1_raw_fwd: y[m][co] = scale[co] * sum_ci x[m][ci] * w[co][ci] + zero_shift[co]  (fp32 rows, no activation)
 *   from split planes (COVA_F16X2: w_packed = [2][Cout][Cin] fp16 hi/lo of 256*w, fold 1/256 into `scale`; COVA_BF16X2:
 *   cova_pack_linear_weight).  Forward (w) and dgrad (the transposed filter on the planes of dy) of a 1x1 convolution.
 * cova_conv1x1_wgrad: dw[co][ci] = inv_scale * sum_m dy[m][co] * x[m][ci] from the split planes of x and dy
 *   (ws: Cin*Cout floats of device scratch, zeroed here).  Both built for 64->64, 64->256, 256->64.                   */
int cova_conv1x1_raw_fwd(const void* x_hi, const void* x_lo, int planes_dtype, int64_t M, int Cin, int Cout,
                         const void* w_packed, const float* scale, const float* zero_shift, void* y, void* stream);
int cova_conv1x1_wgrad(const void* x_hi, const void* x_lo, const void* dy_hi, const void* dy_lo, int64_t M, int Cin,
                       int Cout, int planes_dtype, const float* inv_scale, float* ws, float* dw, void* stream);

/* ---- A9: weight gradient of conv1 (7x7 s2 p3, 3->64, `convnet[0]`, /root/reference/models.py:49-51; it needs no dgrad)
 * on the tensor cores, contraction over the B*Hc*Wc output pixels (stem_wgrad_tc.cu):
 *   dw[co][c][r][s] = inv_scale * sum_{b,oy,ox} dy[b,oy,ox,co] * x[b,c,2oy+r-3,2ox+s-3]      (zero padding)
 * images: the forward's input, NCHW fp32 (or uint8, read as v/255); dy_hi/dy_lo: split planes [B,Hc,Wc,64] of the output
 * gradient (cova_split_planes_scaled); inv_scale: device float (1/s of the dy planes) or NULL; ws: 64*224 floats of device
 * scratch (zeroed here); dw_oihw: [64,3,7,7] fp32, overwritten.                                                       */
int cova_stem_wgrad(const void* images, int img_dtype, int B, int H, int W, const void* dy_hi, const void* dy_lo,
                    int planes_dtype, const float* inv_scale, float* ws, float* dw_oihw, void* stream);

/* ---- bf16 training mode (BASELINE config 3: "ResNet-50 backbone bf16 ... train step"): the same passes with the maps
 * stored as bf16 (COVA_BF16) instead of fp32 - raw convolution outputs, activations (one bf16 plane, which is also the
 * operand of the next tensor-core convolution: one product per MMA), gradient maps; statistics, parameters and weight
 * gradients stay fp32, arithmetic inside every kernel is fp32.  dtype arguments are COVA_F32 or COVA_BF16.
 *   cova_bn_train_stats_t : ws = [sum x | sum x^2] of x [M, C] (x_dtype)
 *   cova_bn_act_fwd_t     : y (y_dtype) = [relu](BN(x) [+ res]); x and res are s_dtype.  relu_mask (optional, bf16 storage with
 *                           relu): [M * C / 8] bytes, bit k of byte i = [y > 0] of element 8 i + k - the forward's ReLU decisions
 *   cova_bn_act_bwd_t     : dx, dres (s_dtype) from dy (dy_dtype), x / res (s_dtype); dgamma / dbeta fp32.  With relu_mask (the
 *                           forward's) neither backward pass reads res: 1 bit instead of 16 per element and pass
 *   cova_maxpool3x3s2_*_t : nn.MaxPool2d(3,2,1) forward (+ winner codes) / backward on maps of `dtype`
 *   cova_stem_conv_raw_fwd_bf16 : conv1 (7x7 s2) in one bf16 product (w = cova_pack_stem_weight), raw output as bf16
 * The convolutions' bf16 entry points are the existing ones: cova_conv3x3_bn_act_fwd (COVA_BF16 in / out, identity
 * scale / shift), cova_conv1x1_raw_fwd (planes_dtype COVA_BF16: x_lo unused, w = bf16 [Cout][Cin], y = bf16 rows) and
 * the three wgrad calls with planes_dtype COVA_BF16 and NULL lo planes.                                               */
int cova_bn_train_stats_t(const void* x, int x_dtype, int64_t M, int C, double* ws, void* stream);
int cova_bn_act_fwd_t(const void* x, int s_dtype, int64_t M, int C, const float* mean, const float* invstd,
                      const float* gamma, const float* beta, const void* res, int relu, void* y, int y_dtype,
                      unsigned char* relu_mask, void* stream);
int cova_bn_act_bwd_t(const void* dy, int dy_dtype, const void* x, const void* res, int s_dtype, int64_t M, int C,
                      const float* mean, const float* invstd, const float* gamma, const float* beta, int relu, double* ws,
                      void* dx, void* dres, float* dgamma, float* dbeta, const unsigned char* relu_mask, void* stream);
int cova_maxpool3x3s2_fwd_t(const void* x, int dtype, int B, int H, int W, int C, void* y, unsigned char* code, void* stream);
int cova_maxpool3x3s2_bwd_t(const unsigned char* code, const void* dy, int dtype, int B, int H, int W, int C, void* dx,
                            void* stream);
int cova_stem_conv_raw_fwd_bf16(const void* images, int img_dtype, int B, int H, int W, const void* w_packed,
                                void* out_bf16, void* stream);

/* ---- BatchNorm batch statistics accumulated by the producing convolution's epilogue (A9): the raw-output convolutions of the
 * train path with one more argument, stats_ws = [2][Cout] doubles (sum y, sum y^2 over all output pixels of the values as they
 * are stored; zeroed by the call), which feeds cova_bn_train_finalize directly - the separate cova_bn_train_stats pass over the
 * raw map is not needed.  cova_conv3x3_bn_act_stats_fwd needs the tensor-core engine, no residual, no ReLU;
 * cova_stem_conv_raw_stats_fwd: w_dtype COVA_BF16X2 / COVA_F16X2 (fp32 output) or COVA_BF16 (one bf16 product, bf16 output). */
int cova_conv3x3_bn_act_stats_fwd(const void* x0, const void* x1, int dtype, int B, int H, int W, int Cin, int Cout,
                                  const void* w_a, const void* w_b, const float* bn_scale, const float* bn_shift,
                                  const void* res0, const void* res1, int relu, int out_dtype, void* y0, void* y1,
                                  int engine, double* stats_ws, void* stream);
int cova_conv1x1_raw_stats_fwd(const void* x_hi, const void* x_lo, int planes_dtype, int64_t M, int Cin, int Cout,
                               const void* w_packed, const float* scale, const float* zero_shift, void* y, double* stats_ws,
                               void* stream);
int cova_stem_conv_raw_stats_fwd(const void* images, int img_dtype, int B, int H, int W, const void* w_packed, int w_dtype,
                                 void* out, double* stats_ws, void* stream);

/* ---- The stem's tail fused, typed (default of the training path): BatchNorm(batch statistics) + ReLU + MaxPool2d(3,2,1) of the
 * raw conv1 output without the normalised 640^2 map, forward and backward, on fp32 or bf16 maps (x / y / dy share the type).
 *   fwd: y (pooled, may be NULL when the planes are wanted alone), winner codes, optional split planes of y (fp32 maps).
 *   bwd: from the pooled gradient and the codes, dx of the raw map either in the map's type (dx) or - fp32 maps - as scaled
 *        split planes for conv1's wgrad (dx_hi / dx_lo, ws_max = C+1 words, inv_scale_vec = 256 x 1/s); dgamma / dbeta.  */
int cova_bn_relu_pool_fwd_t(const void* x, int s_dtype, int B, int H, int W, int C, const float* mean, const float* invstd,
                            const float* gamma, const float* beta, void* y, int y_dtype, unsigned char* code, void* y_hi,
                            void* y_lo, int planes_dtype, void* stream);
int cova_bn_relu_pool_bwd_t(const void* x, int s_dtype, const unsigned char* code, const void* dy_pooled, int dy_dtype, int B,
                            int H, int W, int C, const float* mean, const float* invstd, const float* gamma, const float* beta,
                            double* ws, unsigned int* ws_max, void* dx, void* dx_hi, void* dx_lo, int planes_dtype,
                            int target_log2, float* inv_scale_vec, float* dgamma, float* dbeta, void* stream);

/* bf16 training mode: y = scale * (x W^T) + zero_shift + res on single bf16 planes [M, C] - the dgrad of a 1x1 convolution with the
 * skip branch's gradient added in the epilogue (no separate elementwise add pass over the 256-channel map).              */
int cova_conv1x1_raw_res_fwd(const void* x, int64_t M, int Cin, int Cout, const void* w_bf16, const float* scale,
                             const float* zero_shift, const void* res_bf16, void* y_bf16, void* stream);

/* fp32-parity training mode, the same idea on split planes: fp32 output = scale * conv(x) + shift + res_f32 (fp32 NHWC / rows),
 * i.e. a dgrad with the skip branch's fp32 gradient added by the epilogue.                                                */
int cova_conv3x3_scale_res_f32_fwd(const void* x0, const void* x1, int dtype, int B, int H, int W, const void* w_a,
                                   const void* w_b, const float* scale, const float* shift, const float* res_f32, float* y,
                                   void* stream);
int cova_conv1x1_raw_res_f32_fwd(const void* x_hi, const void* x_lo, int planes_dtype, int64_t M, int Cin, int Cout,
                                 const void* w_packed, const float* scale, const float* zero_shift, const float* res_f32,
                                 float* y, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* COVA_B200_H */
