"""Python wrappers over the C ABI (``include/cova_b200.h``).  PyTorch is used for device memory and the
current stream only; every op below is one call into ``libcova_b200.so``.  CUDA tensors only - there is no
CPU path here (the CPU restatement lives in ``oracle/`` and is test infrastructure)."""
import torch

from . import _lib
from ._lib import BF16, BF16X2, ENGINE_SIMT, ENGINE_TCGEN05, ENGINE_TCGEN05_F16X2, F16, F16X2, F32, U8  # noqa: F401

ENGINES = {"simt": ENGINE_SIMT, "tcgen05": ENGINE_TCGEN05}

# kernels launched through this module since the last reset (bench.py's `gpu_launches` claim)
launch_count = 0


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _cuda(t, dtype=None, name="tensor"):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"cova_b200: `{name}` must be a CUDA tensor (the hot path has no CPU implementation)")
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f"cova_b200: `{name}` must be {dtype}, got {t.dtype}")
    return t


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _call(name, *args):
    global launch_count
    launch_count += 1
    _lib.check(getattr(_lib.lib(), name)(*args), name)


class Planes:
    """An NHWC activation [B,H,W,C] in one of the ABI dtypes: fp32 (p0), bf16 (p0) or split-bf16 (p0=hi, p1=lo)."""

    __slots__ = ("dtype", "p0", "p1", "shape")

    def __init__(self, dtype, shape, device):
        self.dtype, self.shape = dtype, tuple(shape)
        td = torch.float32 if dtype == F32 else (torch.float16 if dtype in (F16, F16X2) else torch.bfloat16)
        self.p0 = torch.empty(self.shape, dtype=td, device=device)
        self.p1 = torch.empty(self.shape, dtype=td, device=device) if dtype in (BF16X2, F16X2) else None

    def float(self):
        """fp32 NHWC view of the value (hi + lo for split-bf16)."""
        if self.dtype == F32:
            return self.p0
        return self.p0.float() + self.p1.float() if self.p1 is not None else self.p0.float()


def device_info():
    sm, smem = _lib._I(), _lib._I()
    _lib.check(_lib.lib().cova_device_info(sm, smem), "cova_device_info")
    return sm.value, smem.value


KNOBS = {"wgrad_drain": 8, "conv_l2_prefetch": 0, "conv_res_prefetch": 1, "stem_l2_prefetch": 2, "stem_converters": 5, "conv_res_load": 6, "roi_rowsplit": 7, "stem_u8_exact": 9}


def set_knob(name, value):
    """Tuning knob of the library (include/cova_b200.h COVA_KNOB_*); value < 0 restores the default."""
    _lib.check(_lib.lib().cova_set_knob(KNOBS[name], int(value)), "cova_set_knob")


def debug_buffer(t):
    """Attach (or detach with None) a CUDA int64 tensor the instrumented kernels add their wait-cycle counters to."""
    _lib.check(_lib.lib().cova_debug_buffer(_ptr(t), 0 if t is None else t.numel()), "cova_debug_buffer")


def stem_fwd(images, w, bn_scale, bn_shift, out_dtype=F32, engine=ENGINE_SIMT):
    """conv7x7 s2 p3 + BN + ReLU + maxpool3x3 s2 p1; images [B,3,H,W] NCHW, fp32 in [0,1] or uint8 raw pixels
    (converted as v/255 in the kernel) -> Planes [B,H/4,W/4,64]."""
    _cuda(images, None, "images")
    if images.dtype not in (torch.float32, torch.uint8):
        raise RuntimeError(f"cova_b200: images must be float32 or uint8, got {images.dtype}")
    images = images.contiguous()
    B, C, H, W = images.shape
    if C != 3:
        raise RuntimeError("cova_b200: images must be [B,3,H,W]")
    Hc, Wc = (H + 6 - 7) // 2 + 1, (W + 6 - 7) // 2 + 1
    Hp, Wp = (Hc + 2 - 3) // 2 + 1, (Wc + 2 - 3) // 2 + 1
    out = Planes(out_dtype, (B, Hp, Wp, 64), images.device)
    _call("cova_stem_fwd", images.data_ptr(), U8 if images.dtype == torch.uint8 else F32, B, H, W, w.data_ptr(), bn_scale.data_ptr(), bn_shift.data_ptr(),
          out_dtype, out.p0.data_ptr(), _ptr(out.p1), engine, _stream())
    return out


def pack_stem_weight(w_oihw):
    """OIHW fp32 [64,3,7,7] -> packed split-bf16 filter of the tcgen05 stem ([28,2,64,8] bf16)."""
    _cuda(w_oihw, torch.float32, "w")
    out = torch.empty((28, 2, 64, 8), dtype=torch.bfloat16, device=w_oihw.device)
    _call("cova_pack_stem_weight", w_oihw.contiguous().data_ptr(), out.data_ptr(), _stream())
    return out


def pack_stem_weight_f16(w_oihw):
    """fp16-mode stem filter ([28,2,64,8] fp16, plane 1 zero)."""
    _cuda(w_oihw, torch.float32, "w")
    out = torch.empty((28, 2, 64, 8), dtype=torch.float16, device=w_oihw.device)
    _call("cova_pack_stem_weight_f16", w_oihw.contiguous().data_ptr(), out.data_ptr(), _stream())
    return out


def pack_stem_weight_f16x2(w_oihw):
    """split-fp16 stem filter ([28,2,64,8] fp16: hi / lo planes of 256*w)."""
    _cuda(w_oihw, torch.float32, "w")
    out = torch.empty((28, 2, 64, 8), dtype=torch.float16, device=w_oihw.device)
    _call("cova_pack_stem_weight_f16x2", w_oihw.contiguous().data_ptr(), out.data_ptr(), _stream())
    return out


def pack_conv_weight_f16x2(w_oihw):
    """OIHW fp32 -> (hi, lo) fp16 [kh*kw, Cout, Cin] planes of 256*w for the split-fp16 mode."""
    _cuda(w_oihw, torch.float32, "w")
    w = w_oihw.contiguous()
    Co, Ci, kh, kw = w.shape
    hi = torch.empty((kh * kw, Co, Ci), dtype=torch.float16, device=w.device)
    lo = torch.empty_like(hi)
    _call("cova_pack_conv_weight_f16x2", w.data_ptr(), Co, Ci, kh, kw, hi.data_ptr(), lo.data_ptr(), _stream())
    return hi, lo


def pack_conv_weight_f16(w_oihw):
    """OIHW fp32 -> fp16 [kh*kw, Cout, Cin] for the fp16 mode of conv3x3_bn_act_fwd."""
    _cuda(w_oihw, torch.float32, "w")
    w = w_oihw.contiguous()
    Co, Ci, kh, kw = w.shape
    out = torch.empty((kh * kw, Co, Ci), dtype=torch.float16, device=w.device)
    _call("cova_pack_conv_weight_f16", w.data_ptr(), Co, Ci, kh, kw, out.data_ptr(), _stream())
    return out


def pack_conv_weight(w_oihw, simt=True, tc=True, split=True):
    """OIHW fp32 -> (simt fp32 [kh,kw,Cin,Cout] | None, tc_hi bf16 [kh*kw,Cout,Cin] | None, tc_lo | None)."""
    _cuda(w_oihw, torch.float32, "w")
    w = w_oihw.contiguous()
    Co, Ci, kh, kw = w.shape
    s = torch.empty((kh, kw, Ci, Co), dtype=torch.float32, device=w.device) if simt else None
    hi = torch.empty((kh * kw, Co, Ci), dtype=torch.bfloat16, device=w.device) if tc else None
    lo = torch.empty_like(hi) if (tc and split) else None
    _call("cova_pack_conv_weight", w.data_ptr(), Co, Ci, kh, kw, _ptr(s), _ptr(hi), _ptr(lo), _stream())
    return s, hi, lo


def new_stats_ws(C, device):
    """[2][C] doubles a raw-output convolution fills with sum y / sum y^2 of its output (-> bn_train_fwd(stats_ws=...))."""
    return torch.empty(2 * C, dtype=torch.float64, device=device)


def conv3x3_bn_act_fwd(x, w_a, w_b, bn_scale, bn_shift, res=None, relu=True, out_dtype=None, engine=ENGINE_SIMT, stats_ws=None):
    """3x3 s1 p1 conv (64->64) + folded BN (+ residual) (+ ReLU) on NHWC Planes.  stats_ws (tensor-core engine, no residual /
    ReLU): the epilogue also accumulates the BatchNorm batch statistics of the output into it."""
    B, H, W, C = x.shape
    out_dtype = x.dtype if out_dtype is None else out_dtype
    y = Planes(out_dtype, (B, H, W, C), x.p0.device)
    if res is not None and res.dtype != x.dtype:
        raise RuntimeError("cova_b200: residual must have the input's dtype")
    if stats_ws is not None:
        _call("cova_conv3x3_bn_act_stats_fwd", x.p0.data_ptr(), _ptr(x.p1), x.dtype, B, H, W, C, C, w_a.data_ptr(), _ptr(w_b),
              bn_scale.data_ptr(), bn_shift.data_ptr(), 0, 0, int(relu), out_dtype, y.p0.data_ptr(), _ptr(y.p1), engine,
              stats_ws.data_ptr(), _stream())
        return y
    _call("cova_conv3x3_bn_act_fwd", x.p0.data_ptr(), _ptr(x.p1), x.dtype, B, H, W, C, C, w_a.data_ptr(), _ptr(w_b),
          bn_scale.data_ptr(), bn_shift.data_ptr(), _ptr(res.p0) if res is not None else 0,
          _ptr(res.p1) if res is not None else 0, int(relu), out_dtype, y.p0.data_ptr(), _ptr(y.p1), engine, _stream())
    return y


def conv1x1_bn_act_fwd(x, w_packed, bn_scale, bn_shift, res=None, relu=True, out_dtype=BF16X2):
    """ResNet-50 Bottleneck 1x1 conv + folded BN (+ residual) (+ ReLU) on split-bf16 NHWC Planes (tcgen05)."""
    B, H, W, Cin = x.shape
    Cout = w_packed.shape[1]
    y = Planes(out_dtype, (B, H, W, Cout), x.p0.device)
    _call("cova_conv1x1_bn_act_fwd", x.p0.data_ptr(), x.p1.data_ptr(), B * H * W, Cin, Cout, w_packed.data_ptr(),
          bn_scale.data_ptr(), bn_shift.data_ptr(), _ptr(res.p0) if res is not None else 0,
          _ptr(res.p1) if res is not None else 0, int(relu), out_dtype, y.p0.data_ptr(), _ptr(y.p1), _stream())
    return y


def roi_fwd(fm, rois, P, spatial_scale, out, mode="pool", sampling_ratio=2, want_argmax=False):
    """fm NHWC fp32 [B,Hf,Wf,C]; rois [T,5]; writes out[:, :C*PH*PW] (out may be a wider row-major buffer)."""
    _cuda(fm, torch.float32, "fm")
    _cuda(rois, torch.float32, "bboxes")
    rois = rois.contiguous()
    B, Hf, Wf, C = fm.shape
    T = rois.shape[0]
    PH, PW = P
    assert out.stride(1) == 1 and out.dtype == torch.float32
    argmax = torch.empty((T, C, PH, PW), dtype=torch.int32, device=fm.device) if want_argmax else None
    _call("cova_roi_fwd", fm.data_ptr(), B, Hf, Wf, C, rois.data_ptr(), T, PH, PW, float(spatial_scale),
          0 if mode == "pool" else 1, int(sampling_ratio), out.data_ptr(), out.stride(0), _ptr(argmax), _stream())
    return argmax


def roi_pool_bwd(grad_out, argmax, rois, fm_shape):
    """Scatter-add of the pooled gradients [T, C*PH*PW] to their arg-max pixels -> NHWC fp32 [B,Hf,Wf,C]."""
    _cuda(grad_out, torch.float32, "grad_out")
    B, Hf, Wf, C = fm_shape
    T, _, PH, PW = argmax.shape
    if grad_out.stride(1) != 1:
        grad_out = grad_out.contiguous()
    grad_fm = torch.zeros(fm_shape, dtype=torch.float32, device=grad_out.device)
    _call("cova_roi_pool_bwd", grad_out.data_ptr(), grad_out.stride(0) if T else C * PH * PW, _ptr(argmax),
          _ptr(rois.contiguous()), T, C, PH, PW, B, Hf, Wf, grad_fm.data_ptr(), _stream())
    return grad_fm


def roi_align_bwd(grad_out, rois, P, spatial_scale, fm_shape, sampling_ratio=2):
    """Backward of RoIAlign(P, scale, sampling_ratio, aligned=False): [T, C*PH*PW] -> NHWC fp32 [B,Hf,Wf,C]."""
    _cuda(grad_out, torch.float32, "grad_out")
    B, Hf, Wf, C = fm_shape
    PH, PW = P
    T = rois.shape[0]
    if grad_out.stride(1) != 1:
        grad_out = grad_out.contiguous()
    grad_fm = torch.zeros(fm_shape, dtype=torch.float32, device=grad_out.device)
    _call("cova_roi_align_bwd", grad_out.data_ptr(), grad_out.stride(0) if T else C * PH * PW, _ptr(rois.contiguous()), T, C,
          PH, PW, float(spatial_scale), int(sampling_ratio), B, Hf, Wf, grad_fm.data_ptr(), _stream())
    return grad_fm


def bbox_enc_fwd(rois, w, b, bn_scale, bn_shift, out):
    _cuda(rois, torch.float32, "bboxes")
    rois = rois.contiguous()
    assert out.stride(1) == 1
    _call("cova_bbox_enc_fwd", rois.data_ptr(), rois.shape[0], w.data_ptr(), b.data_ptr(), _ptr(bn_scale),
          _ptr(bn_shift), w.shape[0], out.data_ptr(), out.stride(0), _stream())


def affine_cols_fwd(x, scale, shift, out):
    _cuda(x, torch.float32, "additional_feats")
    T, D = x.shape
    if T == 0 or D == 0:
        return
    if x.stride(1) != 1:
        x = x.contiguous()
    _call("cova_affine_cols_fwd", x.data_ptr(), T, D, x.stride(0), _ptr(scale), _ptr(shift), out.data_ptr(),
          out.stride(0), _stream())


def pack_linear_weight(w):
    """fp32 [N,K] -> split-bf16 [2,N,K] for the tcgen05 linear engine."""
    _cuda(w, torch.float32, "w")
    w = w.contiguous()
    out = torch.empty((2,) + tuple(w.shape), dtype=torch.bfloat16, device=w.device)
    _call("cova_pack_linear_weight", w.data_ptr(), w.shape[0], w.shape[1], out.data_ptr(), _stream())
    return out


def linear_tc_ok(x):
    """Shape/alignment contract of the tcgen05 linear engine (see include/cova_b200.h)."""
    return x.shape[1] % 8 == 0 and x.stride(0) % 4 == 0 and x.data_ptr() % 16 == 0


def linear_fwd(x, w, bias=None, scale=None, shift=None, res=None, relu=False, out=None, engine=ENGINE_SIMT,
               out_planes=False):
    """Y = act((X @ W^T + bias) * scale + shift + res); x [M,K] (row stride free);
    w = fp32 [N,K] (SIMT engine) or the packed split-bf16 [2,N,K] (tcgen05 engine).
    out_planes=True (tcgen05 only): returns split-bf16 `Planes` [1,1,M,N] instead of an fp32 tensor."""
    _cuda(x, torch.float32, "x")
    M, K = x.shape
    N = w.shape[-2]
    assert x.stride(1) == 1 and w.is_contiguous() and w.shape[-1] == K
    assert (w.dtype == (torch.float16 if engine == ENGINE_TCGEN05_F16X2 else torch.bfloat16) and w.dim() == 3) == (engine != ENGINE_SIMT)
    if out_planes:
        pl = Planes(BF16X2, (1, 1, M, N), x.device)
        y0, y1, ldy, odt = pl.p0.data_ptr(), pl.p1.data_ptr(), N, BF16X2
    else:
        if out is None:
            out = torch.empty((M, N), dtype=torch.float32, device=x.device)
        assert out.stride(1) == 1
        y0, y1, ldy, odt = out.data_ptr(), 0, out.stride(0), F32
    _call("cova_linear_fwd", x.data_ptr(), x.stride(0), M, K, w.data_ptr(), N, _ptr(bias), _ptr(scale), _ptr(shift),
          _ptr(res), res.stride(0) if res is not None else 0, int(relu), odt, y0, y1, ldy, engine, _stream())
    return pl if out_planes else out


def gat_fwd(whj, s, t, att_b, alpha, ctx_idx, out, want_attn=False):
    """Fused neighbour gather + masked softmax + weighted sum.  whj [T,Hd] (row stride free), s/t [T] (stride free)."""
    _cuda(ctx_idx, torch.int64, "context_indices")
    ctx_idx = ctx_idx.contiguous()
    T, K = ctx_idx.shape
    Hd = whj.shape[1]
    assert s.stride(0) == t.stride(0) and whj.stride(1) == 1 and out.stride(1) == 1
    attn = torch.empty((T, K), dtype=torch.float32, device=whj.device) if want_attn else None
    _call("cova_gat_fwd", whj.data_ptr(), whj.stride(0), s.data_ptr(), t.data_ptr(), s.stride(0), float(att_b),
          float(alpha), ctx_idx.data_ptr(), T, K, Hd, out.data_ptr(), out.stride(0), _ptr(attn), _stream())
    return attn


def gat_multihead_fwd(ext, Hd, att_b, alpha, ctx_idx, out, want_attn=False):
    """All heads in one launch.  ext [T, >= H*(Hd+2)] = [whj_0..whj_{H-1} | s_0 t_0 s_1 t_1 ..]; att_b = list of H floats;
    out [T, H*Hd] (row stride free).  Returns attention weights [H,T,K] or None."""
    _cuda(ctx_idx, torch.int64, "context_indices")
    ctx_idx = ctx_idx.contiguous()
    T, K = ctx_idx.shape
    H = len(att_b)
    assert ext.stride(1) == 1 and out.stride(1) == 1
    attn = torch.empty((H, T, K), dtype=torch.float32, device=ext.device) if want_attn else None
    b = (_lib._F * H)(*[float(x) for x in att_b])
    _call("cova_gat_multihead_fwd", ext.data_ptr(), ext.stride(0), Hd, H, b, float(alpha), ctx_idx.data_ptr(), T, K,
          out.data_ptr(), out.stride(0), _ptr(attn), _stream())
    return attn


def gat_bwd(grad_out, ext, Hd, att_b, alpha, ctx_idx, attn):
    """Backward of `gat_fwd` for ext = [whj | s | t | pad] ([T, Hd+4]): returns (d_ext [T,Hd+4], d_bias [1])."""
    _cuda(grad_out, torch.float32, "grad_out")
    grad_out = grad_out.contiguous()
    ctx_idx = ctx_idx.contiguous()
    T, K = ctx_idx.shape
    d_ext = torch.zeros_like(ext)
    d_b = torch.zeros(1, dtype=torch.float32, device=ext.device)
    ld = ext.stride(0)
    _call("cova_gat_bwd", grad_out.data_ptr(), grad_out.stride(0), ext.data_ptr(), ld, ext[:, Hd].data_ptr(),
          ext[:, Hd + 1].data_ptr(), ld, float(att_b), float(alpha), ctx_idx.data_ptr(), attn.data_ptr(), T, K, Hd,
          d_ext.data_ptr(), ld, d_ext[:, Hd].data_ptr(), d_ext[:, Hd + 1].data_ptr(), ld, d_b.data_ptr(), _stream())
    return d_ext, d_b


# ------------------------------------------------------------------ callers either side of the forward (A9, N1-N3)
# bumped by every native in-place parameter update (the Adam kernel writes through raw pointers, which does not touch
# Tensor._version): engine.NativeForward keys its derived weight caches on it
param_generation = 0


def ce_sum_fwd_bwd(logits, labels, want_grad=True, want_correct=False, ignore_index=-100):
    """CrossEntropyLoss(reduction="sum") forward + d loss/d logits (+ number of argmax hits) in one launch.
    Returns (loss [1] fp32, dlogits [T,C] or None, n_correct [1] int32 or None)."""
    _cuda(logits, torch.float32, "logits")
    _cuda(labels, torch.int64, "labels")
    if logits.stride(1) != 1:
        logits = logits.contiguous()
    labels = labels.contiguous()
    T, C = logits.shape
    dev = logits.device
    ws = torch.empty(1, dtype=torch.float64, device=dev)
    loss = torch.empty(1, dtype=torch.float32, device=dev)
    dl = torch.empty((T, C), dtype=torch.float32, device=dev) if want_grad else None
    nc = torch.empty(1, dtype=torch.int32, device=dev) if want_correct else None
    _call("cova_ce_sum_fwd_bwd", logits.data_ptr(), logits.stride(0) if T else C, labels.data_ptr(), T, C,
          int(ignore_index), ws.data_ptr(), loss.data_ptr(), _ptr(dl), C, _ptr(nc), _stream())
    return loss, dl, nc


def adam_step(param, grad, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0):
    """One torch.optim.Adam step over flat fp32 buffers (in place)."""
    global param_generation
    for t, n in ((param, "param"), (grad, "grad"), (exp_avg, "exp_avg"), (exp_avg_sq, "exp_avg_sq")):
        _cuda(t, torch.float32, n)
        assert t.is_contiguous() and t.numel() == param.numel()
    _call("cova_adam_step", param.data_ptr(), grad.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(), param.numel(),
          float(lr), float(beta1), float(beta2), float(eps), float(weight_decay), int(step), float(grad_scale), _stream())
    param_generation += 1


def topk_hits(logits, labels, page_offsets, k=1):
    """evaluate_model's top-k test for every (page, class): int32 [B, C] (column 0 unused; -1 = class absent)."""
    _cuda(logits, torch.float32, "logits")
    _cuda(labels, torch.int64, "labels")
    _cuda(page_offsets, torch.int32, "page_offsets")
    if logits.stride(1) != 1:
        logits = logits.contiguous()
    B, C = page_offsets.numel() - 1, logits.shape[1]
    hits = torch.zeros((B, C), dtype=torch.int32, device=logits.device)
    _call("cova_topk_hits", logits.data_ptr(), logits.stride(0) if logits.shape[0] else C, labels.contiguous().data_ptr(),
          page_offsets.contiguous().data_ptr(), B, C, int(k), hits.data_ptr(), _stream())
    return hits


def build_batch(page_offsets, context_size, boxes_xywh=None):
    """Device-side batch assembly from per-page row offsets (int32 [B+1]): returns (bboxes [T,5] or None,
    context_indices int64 [T, 2*context_size])."""
    _cuda(page_offsets, torch.int32, "page_offsets")
    page_offsets = page_offsets.contiguous()
    B = page_offsets.numel() - 1
    if boxes_xywh is not None:
        _cuda(boxes_xywh, torch.float32, "boxes_xywh")
        boxes_xywh = boxes_xywh.contiguous()
        T = boxes_xywh.shape[0]
    else:
        T = int(page_offsets[-1].item())
    dev = page_offsets.device
    ctx = torch.empty((T, 2 * context_size), dtype=torch.int64, device=dev)
    bb = torch.empty((T, 5), dtype=torch.float32, device=dev) if boxes_xywh is not None else None
    _call("cova_build_batch", page_offsets.data_ptr(), B, T, int(context_size), _ptr(boxes_xywh), _ptr(bb),
          ctx.data_ptr() if context_size > 0 else 0, _stream())
    return bb, ctx


# ------------------------------------------------------------------ training-mode backbone pieces (bn_train.cu)
def _nhwc(t, name):
    _cuda(t, torch.float32, name)
    if not t.is_contiguous():
        raise RuntimeError(f"cova_b200: `{name}` must be a contiguous NHWC fp32 tensor")
    return t


def _planes_like(x, dtype=BF16X2):
    return Planes(dtype, tuple(x.shape), x.device)


def split_planes(x, dtype=BF16X2):
    """fp32 NHWC tensor -> split `Planes` (hi = r(x), lo = r(x - hi); r = bf16 for BF16X2, fp16 for F16X2)."""
    _nhwc(x, "x")
    pl = _planes_like(x, dtype)
    _call("cova_split_planes", x.data_ptr(), x.numel(), pl.p0.data_ptr(), pl.p1.data_ptr(), dtype, _stream())
    return pl


def split_planes_scaled(x, dtype=F16X2, target_log2=10):
    """Gradient map -> split `Planes` of x * s with a per-tensor power-of-two s chosen on the device (no host sync), and
    the [256] vector of 1/s the consuming kernel multiplies back in.  Returns (Planes, inv_scale_vec)."""
    _nhwc(x, "x")
    pl = _planes_like(x, dtype)
    ws = torch.empty(1, dtype=torch.int32, device=x.device)
    inv = torch.empty(256, dtype=torch.float32, device=x.device)
    _call("cova_split_planes_scaled", x.data_ptr(), x.numel(), pl.p0.data_ptr(), pl.p1.data_ptr(), dtype, int(target_log2),
          ws.data_ptr(), inv.data_ptr(), _stream())
    return pl, inv


def conv3x3_wgrad(x_planes, dy_planes, inv_scale=None):
    """Weight gradient [64,64,3,3] (OIHW fp32) of a 3x3 s1 p1 64->64 convolution from the split planes of its input and of
    its (scaled) output gradient; inv_scale = device tensor holding 1/s of the dy planes (or None)."""
    B, H, W, C = x_planes.shape
    if C != 64 or tuple(dy_planes.shape) != (B, H, W, 64) or x_planes.dtype != dy_planes.dtype:
        raise RuntimeError("cova_b200: conv3x3_wgrad needs [B,H,W,64] split planes of one format for x and dy")
    dev = x_planes.p0.device
    ws = torch.empty(9 * 64 * 64, dtype=torch.float32, device=dev)
    dw = torch.empty((64, 64, 3, 3), dtype=torch.float32, device=dev)
    _call("cova_conv3x3_wgrad", x_planes.p0.data_ptr(), _ptr(x_planes.p1), dy_planes.p0.data_ptr(), _ptr(dy_planes.p1),
          B, H, W, x_planes.dtype, _ptr(inv_scale), ws.data_ptr(), dw.data_ptr(), _stream())
    return dw


def pack_linear_weight_f16x2(w):
    """fp32 [N,K] -> split-fp16 [2,N,K] planes of 256*w (the 1x1 convolutions of the training path)."""
    _cuda(w, torch.float32, "w")
    w = w.contiguous()
    N, K = w.shape
    out = torch.empty((2, N, K), dtype=torch.float16, device=w.device)
    _call("cova_pack_conv_weight_f16x2", w.data_ptr(), N, K, 1, 1, out[0].data_ptr(), out[1].data_ptr(), _stream())
    return out


_W256 = {}


def _inv256(device):
    k = str(device)
    if k not in _W256:
        _W256[k] = (torch.full((256,), 1.0 / 256.0, device=device), torch.zeros(256, device=device))
    return _W256[k]


_ONES256 = {}


def _ones256(device):
    k = str(device)
    if k not in _ONES256:
        _ONES256[k] = (torch.ones(256, device=device), torch.zeros(256, device=device))
    return _ONES256[k]


def conv3x3_scale_res_f32_fwd(x_planes, w_hi, w_lo, scale, shift, res_f32):
    """fp32-parity training dgrad: split planes in -> fp32 [B,H,W,64] = scale * conv3x3(x) + shift + res_f32."""
    B, H, W, C = x_planes.shape
    y = torch.empty((B, H, W, C), dtype=torch.float32, device=x_planes.p0.device)
    _call("cova_conv3x3_scale_res_f32_fwd", x_planes.p0.data_ptr(), x_planes.p1.data_ptr(), x_planes.dtype, B, H, W, w_hi.data_ptr(),
          w_lo.data_ptr(), scale.data_ptr(), shift.data_ptr(), _nhwc(res_f32, "res").data_ptr(), y.data_ptr(), _stream())
    return y


def conv1x1_raw_res_f32_fwd(x_planes, w_packed, scale, res_f32):
    """fp32-parity training dgrad of a 1x1 convolution: fp32 rows = (scale / 256) * (x W^T) + res_f32."""
    Cin, Cout = x_planes.shape[-1], w_packed.shape[-2]
    M = 1
    for d in x_planes.shape[:-1]:
        M *= d
    dev = x_planes.p0.device
    inv, zero = _inv256(dev)
    sc = inv if scale is None else scale[:256] * (1.0 / 256.0)
    y = torch.empty(tuple(x_planes.shape[:-1]) + (Cout,), dtype=torch.float32, device=dev)
    _call("cova_conv1x1_raw_res_f32_fwd", x_planes.p0.data_ptr(), x_planes.p1.data_ptr(), x_planes.dtype, M, Cin, Cout,
          w_packed.data_ptr(), sc.data_ptr(), zero.data_ptr(), _nhwc(res_f32, "res").data_ptr(), y.data_ptr(), _stream())
    return y


def conv1x1_raw_res_fwd_bf16(x, w_bf16, res):
    """bf16 training mode: x [.., Cin] bf16 @ w_bf16 [Cout, Cin]^T + res [.., Cout] bf16 -> bf16 (dgrad + skip-branch gradient)."""
    Cin, Cout = x.shape[-1], w_bf16.shape[0]
    M = x.numel() // Cin
    one, zero = _ones256(x.device)
    y = torch.empty(tuple(x.shape[:-1]) + (Cout,), dtype=torch.bfloat16, device=x.device)
    _call("cova_conv1x1_raw_res_fwd", x.data_ptr(), M, Cin, Cout, w_bf16.data_ptr(), one.data_ptr(), zero.data_ptr(),
          _map(res, "res").data_ptr(), y.data_ptr(), _stream())
    return y


def conv1x1_raw_fwd(x_planes, w_packed, scale=None, stats_ws=None):
    """1x1 convolution of split-fp16 NHWC planes [..., Cin] with `pack_linear_weight_f16x2(w [Cout,Cin])` -> raw fp32
    [..., Cout].  `scale` ([>= Cout] device floats, e.g. the 1/s of scaled gradient planes) multiplies the result."""
    Cin = x_planes.shape[-1]
    Cout = w_packed.shape[-2]
    M = 1
    for d in x_planes.shape[:-1]:
        M *= d
    dev = x_planes.p0.device
    inv, zero = _inv256(dev)
    if x_planes.dtype == BF16:           # bf16 training mode: w_packed = bf16 [Cout, Cin] (unscaled), bf16 rows out
        one, _ = _ones256(dev)
        y = torch.empty(tuple(x_planes.shape[:-1]) + (Cout,), dtype=torch.bfloat16, device=dev)
        _call("cova_conv1x1_raw_stats_fwd", x_planes.p0.data_ptr(), 0, BF16, M, Cin, Cout, w_packed.data_ptr(), one.data_ptr(),
              zero.data_ptr(), y.data_ptr(), _ptr(stats_ws), _stream())
        return y
    sc = inv if scale is None else scale[:256] * (1.0 / 256.0)
    y = torch.empty(tuple(x_planes.shape[:-1]) + (Cout,), dtype=torch.float32, device=dev)
    _call("cova_conv1x1_raw_stats_fwd", x_planes.p0.data_ptr(), x_planes.p1.data_ptr(), x_planes.dtype, M, Cin, Cout,
          w_packed.data_ptr(), sc.data_ptr(), zero.data_ptr(), y.data_ptr(), _ptr(stats_ws), _stream())
    return y


def conv1x1_wgrad(x_planes, dy_planes, inv_scale=None):
    """Weight gradient [Cout, Cin, 1, 1] of a 1x1 convolution from the split planes of its input and output gradient."""
    Cin, Cout = x_planes.shape[-1], dy_planes.shape[-1]
    M = 1
    for d in x_planes.shape[:-1]:
        M *= d
    dev = x_planes.p0.device
    ws = torch.empty(Cin * Cout, dtype=torch.float32, device=dev)
    dw = torch.empty((Cout, Cin, 1, 1), dtype=torch.float32, device=dev)
    _call("cova_conv1x1_wgrad", x_planes.p0.data_ptr(), _ptr(x_planes.p1), dy_planes.p0.data_ptr(), _ptr(dy_planes.p1),
          M, Cin, Cout, x_planes.dtype, _ptr(inv_scale), ws.data_ptr(), dw.data_ptr(), _stream())
    return dw


def stem_conv_raw_fwd(images, w_packed, stats_ws=None):
    """conv1 alone (training mode): images [B,3,H,W] fp32 / uint8 NCHW -> raw conv output [B,H/2,W/2,64] fp32 NHWC.
    w_packed from pack_stem_weight (bf16: split-bf16 products) or pack_stem_weight_f16x2 (fp16: split-fp16)."""
    _cuda(images, None, "images")
    images = images.contiguous()
    B, C, H, W = images.shape
    out = torch.empty((B, (H - 1) // 2 + 1, (W - 1) // 2 + 1, 64), dtype=torch.float32, device=images.device)
    _call("cova_stem_conv_raw_stats_fwd", images.data_ptr(), U8 if images.dtype == torch.uint8 else F32, B, H, W,
          w_packed.data_ptr(), F16X2 if w_packed.dtype == torch.float16 else BF16X2, out.data_ptr(), _ptr(stats_ws), _stream())
    return out


def stem_wgrad(images, dy_planes, inv_scale=None):
    """Weight gradient [64,3,7,7] (OIHW fp32) of conv1 (7x7 s2 p3) from the forward's images ([B,3,H,W] fp32 / uint8 NCHW)
    and the split planes [B,Hc,Wc,64] of the (scaled) output gradient; inv_scale = device tensor holding 1/s (or None)."""
    _cuda(images, None, "images")
    images = images.contiguous()
    B, C, H, W = images.shape
    Hc, Wc = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    if C != 3 or tuple(dy_planes.shape) != (B, Hc, Wc, 64) or dy_planes.dtype not in (F16X2, BF16X2, BF16):
        raise RuntimeError("cova_b200: stem_wgrad needs [B,3,H,W] images and [B,H/2,W/2,64] split planes of dy")
    dev = images.device
    ws = torch.empty(64 * 224, dtype=torch.float32, device=dev)
    dw = torch.empty((64, 3, 7, 7), dtype=torch.float32, device=dev)
    _call("cova_stem_wgrad", images.data_ptr(), U8 if images.dtype == torch.uint8 else F32, B, H, W, dy_planes.p0.data_ptr(),
          _ptr(dy_planes.p1), dy_planes.dtype, _ptr(inv_scale), ws.data_ptr(), dw.data_ptr(), _stream())
    return dw


def bn_train_fwd(x, gamma, beta, running_mean, running_var, momentum, eps, res=None, relu=True, want_planes=False,
                 planes_dtype=BF16X2, want_y=True, stats_ws=None, relu_mask=None):
    """BatchNorm2d with batch statistics (+ residual) (+ ReLU) on an NHWC fp32 map x [..., C]; updates the running
    statistics in place (pass None to skip).  Returns (y, mean [C], invstd [C], split-bf16 Planes of y or None)."""
    _nhwc(x, "x")
    C = x.shape[-1]
    M = x.numel() // C
    dev = x.device
    ws = stats_ws if stats_ws is not None else torch.empty(2 * C, dtype=torch.float64, device=dev)
    mean, inv = torch.empty(C, dtype=torch.float32, device=dev), torch.empty(C, dtype=torch.float32, device=dev)
    y = torch.empty_like(x) if (want_y or not want_planes) else None
    pl = _planes_like(x, planes_dtype) if want_planes else None
    if stats_ws is None:                 # (else: the producing convolution's epilogue already accumulated them)
        _call("cova_bn_train_stats", x.data_ptr(), M, C, ws.data_ptr(), _stream())
    _call("cova_bn_train_finalize", ws.data_ptr(), M, C, float(eps), float(momentum), mean.data_ptr(), inv.data_ptr(),
          _ptr(running_mean), _ptr(running_var), _stream())
    if running_mean is not None:          # raw-pointer buffer update: tell the derived-weight caches
        global param_generation
        param_generation += 1
    _call("cova_bn_act_fwd", x.data_ptr(), M, C, mean.data_ptr(), inv.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
          _ptr(None if res is None else _nhwc(res, "res")), int(relu), _ptr(y),
          pl.p0.data_ptr() if pl else 0, pl.p1.data_ptr() if pl else 0, planes_dtype, _ptr(relu_mask), _stream())
    return y, mean, inv, pl


def bn_train_bwd(dy, x, mean, invstd, gamma, beta, res=None, relu=True, want_dres=False):
    """Backward of `bn_train_fwd`: returns (dx, dres or None, dgamma [C], dbeta [C])."""
    _nhwc(dy, "dy"); _nhwc(x, "x")
    C = x.shape[-1]
    M = x.numel() // C
    dev = x.device
    ws = torch.empty(2 * C, dtype=torch.float64, device=dev)
    dx = torch.empty_like(x)
    dres = torch.empty_like(x) if want_dres else None
    dg, db = torch.empty(C, dtype=torch.float32, device=dev), torch.empty(C, dtype=torch.float32, device=dev)
    _call("cova_bn_act_bwd", dy.data_ptr(), x.data_ptr(), _ptr(res), M, C, mean.data_ptr(), invstd.data_ptr(),
          gamma.data_ptr(), beta.data_ptr(), int(relu), ws.data_ptr(), dx.data_ptr(), _ptr(dres), dg.data_ptr(),
          db.data_ptr(), _stream())
    return dx, dres, dg, db


def bn_train_bwd_planes(dy, x, mean, invstd, gamma, beta, res=None, relu=True, want_dres=False, planes_dtype=F16X2,
                        target_log2=10, relu_mask=None):
    """Backward of `bn_train_fwd` with dx emitted as scaled split planes for the tensor-core dgrad / wgrad kernels:
    returns (Planes of dx * s, inv_scale_vec [256] = 1/s, dres or None, dgamma [C], dbeta [C])."""
    _nhwc(dy, "dy"); _nhwc(x, "x")
    C = x.shape[-1]
    M = x.numel() // C
    dev = x.device
    ws = torch.empty(2 * C, dtype=torch.float64, device=dev)
    wmax = torch.empty(C + 1, dtype=torch.int32, device=dev)
    pl = _planes_like(x, planes_dtype)
    inv = torch.empty(256, dtype=torch.float32, device=dev)
    dres = torch.empty_like(x) if want_dres else None
    dg, db = torch.empty(C, dtype=torch.float32, device=dev), torch.empty(C, dtype=torch.float32, device=dev)
    _call("cova_bn_act_bwd_planes", dy.data_ptr(), x.data_ptr(), _ptr(res), M, C, mean.data_ptr(), invstd.data_ptr(),
          gamma.data_ptr(), beta.data_ptr(), int(relu), ws.data_ptr(), wmax.data_ptr(), pl.p0.data_ptr(), pl.p1.data_ptr(),
          planes_dtype, int(target_log2), inv.data_ptr(), _ptr(dres), dg.data_ptr(), db.data_ptr(), _ptr(relu_mask), _stream())
    return pl, inv, dres, dg, db


def maxpool3x3s2_fwd(x, want_planes=False, planes_dtype=BF16X2):
    """nn.MaxPool2d(3, 2, 1) on an NHWC fp32 map [B,H,W,C]: returns (y, winner codes uint8 like y, Planes of y or None)."""
    _nhwc(x, "x")
    B, H, W, C = x.shape
    y = torch.empty((B, (H - 1) // 2 + 1, (W - 1) // 2 + 1, C), dtype=torch.float32, device=x.device)
    code = torch.empty(y.shape, dtype=torch.uint8, device=x.device)
    pl = _planes_like(y, planes_dtype) if want_planes else None
    _call("cova_maxpool3x3s2_fwd", x.data_ptr(), B, H, W, C, y.data_ptr(), code.data_ptr(),
          pl.p0.data_ptr() if pl else 0, pl.p1.data_ptr() if pl else 0, planes_dtype, _stream())
    return y, code, pl


def maxpool3x3s2_bwd(code, dy, in_shape):
    _nhwc(dy, "dy")
    B, H, W, C = in_shape
    dx = torch.empty(in_shape, dtype=torch.float32, device=dy.device)
    _call("cova_maxpool3x3s2_bwd", code.data_ptr(), dy.data_ptr(), B, H, W, C, dx.data_ptr(), _stream())
    return dx


def bn_relu_pool_fwd(x, gamma, beta, running_mean, running_var, momentum, eps, want_planes=False, planes_dtype=BF16X2):
    """Stem of the training path, fused: BatchNorm2d(batch statistics) + ReLU + MaxPool2d(3,2,1) of the raw conv1 output
    x [B,H,W,C] NHWC fp32 without writing the normalised map.  Returns (y pooled, codes uint8, mean, invstd, Planes|None)."""
    _nhwc(x, "x")
    B, H, W, C = x.shape
    M = B * H * W
    dev = x.device
    ws = torch.empty(2 * C, dtype=torch.float64, device=dev)
    mean, inv = torch.empty(C, dtype=torch.float32, device=dev), torch.empty(C, dtype=torch.float32, device=dev)
    _call("cova_bn_train_stats", x.data_ptr(), M, C, ws.data_ptr(), _stream())
    _call("cova_bn_train_finalize", ws.data_ptr(), M, C, float(eps), float(momentum), mean.data_ptr(), inv.data_ptr(),
          _ptr(running_mean), _ptr(running_var), _stream())
    y = torch.empty((B, (H - 1) // 2 + 1, (W - 1) // 2 + 1, C), dtype=torch.float32, device=dev)
    code = torch.empty(y.shape, dtype=torch.uint8, device=dev)
    pl = _planes_like(y, planes_dtype) if want_planes else None
    _call("cova_bn_relu_pool_fwd", x.data_ptr(), B, H, W, C, mean.data_ptr(), inv.data_ptr(), gamma.data_ptr(),
          beta.data_ptr(), y.data_ptr(), code.data_ptr(), pl.p0.data_ptr() if pl else 0, pl.p1.data_ptr() if pl else 0,
          planes_dtype, _stream())
    return y, code, mean, inv, pl


def bn_relu_pool_bwd(x, code, dy_pooled, mean, invstd, gamma, beta):
    """Backward of `bn_relu_pool_fwd`: returns (dx of the raw conv1 output, dgamma, dbeta)."""
    _nhwc(x, "x"); _nhwc(dy_pooled, "dy")
    B, H, W, C = x.shape
    dev = x.device
    ws = torch.empty(2 * C, dtype=torch.float64, device=dev)
    dx = torch.empty_like(x)
    dg, db = torch.empty(C, dtype=torch.float32, device=dev), torch.empty(C, dtype=torch.float32, device=dev)
    _call("cova_bn_relu_pool_bwd", x.data_ptr(), code.data_ptr(), dy_pooled.data_ptr(), B, H, W, C, mean.data_ptr(),
          invstd.data_ptr(), gamma.data_ptr(), beta.data_ptr(), ws.data_ptr(), dx.data_ptr(), dg.data_ptr(), db.data_ptr(),
          _stream())
    return dx, dg, db


# ----------------------------------------------------------------------------- bf16 training mode (typed entry points)
def _dt(t):
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise RuntimeError("cova_b200: maps of the training path are fp32 or bf16")


def _map(t, name):
    if not t.is_cuda or not t.is_contiguous():
        raise RuntimeError(f"cova_b200: `{name}` must be a contiguous CUDA NHWC tensor")
    return t


def stem_conv_raw_fwd_bf16(images, w_packed, stats_ws=None):
    """conv1 in one bf16 product: images [B,3,H,W] fp32 / uint8 -> raw conv output [B,H/2,W/2,64] bf16 NHWC."""
    _cuda(images, None, "images")
    images = images.contiguous()
    B, C, H, W = images.shape
    out = torch.empty((B, (H - 1) // 2 + 1, (W - 1) // 2 + 1, 64), dtype=torch.bfloat16, device=images.device)
    _call("cova_stem_conv_raw_stats_fwd", images.data_ptr(), U8 if images.dtype == torch.uint8 else F32, B, H, W,
          w_packed.data_ptr(), BF16, out.data_ptr(), _ptr(stats_ws), _stream())
    return out


def bn_train_fwd_t(x, gamma, beta, running_mean, running_var, momentum, eps, res=None, relu=True, out_dtype=None, stats_ws=None,
                   relu_mask=None):
    """`bn_train_fwd` on a map of either storage type (fp32 / bf16); y in `out_dtype` (default: x's).  Returns (y, mean, invstd).
    relu_mask (bf16 storage, relu): uint8 [numel / 8], filled with the ReLU decisions for `bn_train_bwd_t`."""
    _map(x, "x")
    C = x.shape[-1]
    M = x.numel() // C
    dev = x.device
    ws = stats_ws if stats_ws is not None else torch.empty(2 * C, dtype=torch.float64, device=dev)
    mean, inv = torch.empty(C, dtype=torch.float32, device=dev), torch.empty(C, dtype=torch.float32, device=dev)
    y = torch.empty(x.shape, dtype=out_dtype or x.dtype, device=dev)
    if res is not None and _map(res, "res").dtype != x.dtype:
        raise RuntimeError("cova_b200: the residual has the storage type of x")
    if stats_ws is None:
        _call("cova_bn_train_stats_t", x.data_ptr(), _dt(x), M, C, ws.data_ptr(), _stream())
    _call("cova_bn_train_finalize", ws.data_ptr(), M, C, float(eps), float(momentum), mean.data_ptr(), inv.data_ptr(),
          _ptr(running_mean), _ptr(running_var), _stream())
    if running_mean is not None:
        global param_generation
        param_generation += 1
    _call("cova_bn_act_fwd_t", x.data_ptr(), _dt(x), M, C, mean.data_ptr(), inv.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
          _ptr(res), int(relu), y.data_ptr(), _dt(y), _ptr(relu_mask), _stream())
    return y, mean, inv


def bn_train_bwd_t(dy, x, mean, invstd, gamma, beta, res=None, relu=True, want_dres=False, relu_mask=None):
    """Backward of `bn_train_fwd_t`: dx / dres in x's storage type; dy fp32 or bf16.  Returns (dx, dres | None, dgamma, dbeta).
    relu_mask = the forward's bit mask: the residual map is then not needed (nor read)."""
    _map(dy, "dy"); _map(x, "x")
    C = x.shape[-1]
    M = x.numel() // C
    dev = x.device
    ws = torch.empty(2 * C, dtype=torch.float64, device=dev)
    dx = torch.empty_like(x)
    dres = torch.empty_like(x) if want_dres else None
    dg, db = torch.empty(C, dtype=torch.float32, device=dev), torch.empty(C, dtype=torch.float32, device=dev)
    _call("cova_bn_act_bwd_t", dy.data_ptr(), _dt(dy), x.data_ptr(), _ptr(res), _dt(x), M, C, mean.data_ptr(), invstd.data_ptr(),
          gamma.data_ptr(), beta.data_ptr(), int(relu), ws.data_ptr(), dx.data_ptr(), _ptr(dres), dg.data_ptr(), db.data_ptr(),
          _ptr(relu_mask), _stream())
    return dx, dres, dg, db


def maxpool3x3s2_fwd_t(x):
    _map(x, "x")
    B, H, W, C = x.shape
    y = torch.empty((B, (H - 1) // 2 + 1, (W - 1) // 2 + 1, C), dtype=x.dtype, device=x.device)
    code = torch.empty(y.shape, dtype=torch.uint8, device=x.device)
    _call("cova_maxpool3x3s2_fwd_t", x.data_ptr(), _dt(x), B, H, W, C, y.data_ptr(), code.data_ptr(), _stream())
    return y, code


def maxpool3x3s2_bwd_t(code, dy, in_shape):
    _map(dy, "dy")
    B, H, W, C = in_shape
    dx = torch.empty(in_shape, dtype=dy.dtype, device=dy.device)
    _call("cova_maxpool3x3s2_bwd_t", code.data_ptr(), dy.data_ptr(), _dt(dy), B, H, W, C, dx.data_ptr(), _stream())
    return dx


def bf16_plane(t):
    """A contiguous bf16 NHWC tensor as single-plane `Planes` (the operand format of the bf16 training mode)."""
    pl = Planes.__new__(Planes)
    pl.dtype, pl.shape, pl.p0, pl.p1 = BF16, tuple(t.shape), t, None
    return pl


def bn_relu_pool_fwd_t(x, gamma, beta, running_mean, running_var, momentum, eps, want_planes=False, planes_dtype=F16X2, stats_ws=None):
    """The stem's tail fused on an fp32 or bf16 map: BatchNorm2d(batch statistics) + ReLU + MaxPool2d(3,2,1) of the raw conv1
    output x [B,H,W,C] without the normalised map.  Returns (y pooled in x's type, codes uint8, mean, invstd, Planes of y | None)."""
    _map(x, "x")
    B, H, W, C = x.shape
    M = B * H * W
    dev = x.device
    ws = stats_ws if stats_ws is not None else torch.empty(2 * C, dtype=torch.float64, device=dev)
    mean, inv = torch.empty(C, dtype=torch.float32, device=dev), torch.empty(C, dtype=torch.float32, device=dev)
    if stats_ws is None:
        _call("cova_bn_train_stats_t", x.data_ptr(), _dt(x), M, C, ws.data_ptr(), _stream())
    _call("cova_bn_train_finalize", ws.data_ptr(), M, C, float(eps), float(momentum), mean.data_ptr(), inv.data_ptr(),
          _ptr(running_mean), _ptr(running_var), _stream())
    if running_mean is not None:
        global param_generation
        param_generation += 1
    y = torch.empty((B, (H - 1) // 2 + 1, (W - 1) // 2 + 1, C), dtype=x.dtype, device=dev)
    code = torch.empty(y.shape, dtype=torch.uint8, device=dev)
    pl = _planes_like(y, planes_dtype) if want_planes else None
    _call("cova_bn_relu_pool_fwd_t", x.data_ptr(), _dt(x), B, H, W, C, mean.data_ptr(), inv.data_ptr(), gamma.data_ptr(),
          beta.data_ptr(), y.data_ptr(), _dt(y), code.data_ptr(), pl.p0.data_ptr() if pl else 0, pl.p1.data_ptr() if pl else 0,
          planes_dtype, _stream())
    return y, code, mean, inv, pl


def bn_relu_pool_bwd_t(x, code, dy_pooled, mean, invstd, gamma, beta, planes=False, planes_dtype=F16X2, target_log2=10):
    """Backward of `bn_relu_pool_fwd_t`.  planes=False: (dx in x's type, None, dgamma, dbeta); planes=True (fp32 maps): (scaled
    split Planes of dx, inv_scale_vec, dgamma, dbeta) - the operand of conv1's tensor-core wgrad."""
    _map(x, "x"); _map(dy_pooled, "dy")
    B, H, W, C = x.shape
    dev = x.device
    ws = torch.empty(2 * C, dtype=torch.float64, device=dev)
    dg, db = torch.empty(C, dtype=torch.float32, device=dev), torch.empty(C, dtype=torch.float32, device=dev)
    if planes:
        wmax = torch.empty(C + 1, dtype=torch.int32, device=dev)
        pl = _planes_like(x, planes_dtype)
        inv = torch.empty(256, dtype=torch.float32, device=dev)
        _call("cova_bn_relu_pool_bwd_t", x.data_ptr(), _dt(x), code.data_ptr(), dy_pooled.data_ptr(), _dt(dy_pooled), B, H, W, C,
              mean.data_ptr(), invstd.data_ptr(), gamma.data_ptr(), beta.data_ptr(), ws.data_ptr(), wmax.data_ptr(), 0,
              pl.p0.data_ptr(), pl.p1.data_ptr(), planes_dtype, int(target_log2), inv.data_ptr(), dg.data_ptr(), db.data_ptr(), _stream())
        return pl, inv, dg, db
    dx = torch.empty_like(x)
    _call("cova_bn_relu_pool_bwd_t", x.data_ptr(), _dt(x), code.data_ptr(), dy_pooled.data_ptr(), _dt(dy_pooled), B, H, W, C,
          mean.data_ptr(), invstd.data_ptr(), gamma.data_ptr(), beta.data_ptr(), ws.data_ptr(), 0, dx.data_ptr(), 0, 0, planes_dtype,
          int(target_log2), 0, dg.data_ptr(), db.data_ptr(), _stream())
    return dx, None, dg, db
