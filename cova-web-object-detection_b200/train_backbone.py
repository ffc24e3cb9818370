"""Training-mode backbone: the truncated ResNet of `/root/reference/models.py:49-51` under
`model.train()` (`train.py:27`), i.e. BatchNorm with BATCH statistics and autograd, NHWC end to end.

Native (libcova_b200.so): every FORWARD convolution that has a tensor-core kernel here - conv1 (`cova_stem_conv_raw_fwd`)
and the 3x3 64->64 convolutions (`cova_conv3x3_bn_act_fwd` with an identity epilogue), both in the split-fp16
three-product mode (activations to ~1e-6); BatchNorm(batch statistics) + residual + ReLU forward / backward (`cova_bn_train_*`, `cova_bn_act_*`),
which also emit the split planes the next convolution consumes; the stem's maxpool forward / backward.
The dgrad of those 3x3 convolutions is native as well: the same tcgen05 kernel run on the split-fp16 planes of the
output gradient with the rotated, channel-swapped filter.
Library, interim: the convolutions' wgrad (and conv1's, which needs no dgrad) through
`aten.convolution_backward` (cuDNN), and the ResNet-50 1x1 convolutions.

Profile that motivated this (tools/prof_train.py, B=16, before): cuDNN BatchNorm 29 ms, fp32 forward convolutions
13-21 ms, maxpool backward 4.4 ms, NCHW<->NHWC transposes 5 ms of a 61 ms step."""
import os

import torch
import torch.nn.functional as F

from . import ops
from .ops import ENGINE_TCGEN05, F16X2, F32

_CONST = {}


def _ones_zeros(device):
    k = str(device)
    if k not in _CONST:
        _CONST[k] = (torch.ones(64, device=device), torch.zeros(64, device=device))
    return _CONST[k]


def _momentum(bn, track):
    """nn.BatchNorm's exponential_average_factor: `momentum`, or the cumulative average 1 / num_batches_tracked (counted
    after this batch) when momentum is None."""
    if bn.momentum is not None:
        return bn.momentum
    if track and bn.num_batches_tracked is not None:
        return 1.0 / float(int(bn.num_batches_tracked) + 1)
    return 0.0


class _BnActFn(torch.autograd.Function):
    """(y, y_hi, y_lo) = [relu](BN_batchstats(x) [+ res]) on NHWC fp32; running statistics updated like nn.BatchNorm2d.
    The planes are the split-bf16 copy of y for the tensor-core convolution that follows (or None)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, res, bn, relu, want_planes):
        ctx.set_materialize_grads(False)      # no zero-filled 400 MB "gradients" for the non-differentiable planes
        x = x.contiguous()
        res_c = None if res is None else res.contiguous()
        track = bn.track_running_stats and bn.running_mean is not None
        mom = _momentum(bn, track)
        y, mean, inv, pl = ops.bn_train_fwd(x, gamma.detach(), beta.detach(), bn.running_mean if track else None,
                                            bn.running_var if track else None, mom, bn.eps, res=res_c, relu=relu,
                                            want_planes=want_planes, planes_dtype=F16X2)
        if track and bn.num_batches_tracked is not None:
            bn.num_batches_tracked += 1
        ctx.save_for_backward(x, mean, inv, gamma.detach(), beta.detach(), res_c if (relu and res is not None) else None)
        ctx.relu, ctx.has_res = relu, res is not None
        if pl is None:
            return y, None, None
        ctx.mark_non_differentiable(pl.p0, pl.p1)
        return y, pl.p0, pl.p1

    @staticmethod
    def backward(ctx, dy, _dhi, _dlo):
        x, mean, inv, gamma, beta, res = ctx.saved_tensors
        if dy is None:
            dy = torch.zeros_like(x)
        want_dres = ctx.has_res and ctx.needs_input_grad[3]
        dx, dres, dg, db = ops.bn_train_bwd(dy.contiguous(), x, mean, inv, gamma, beta, res=res, relu=ctx.relu,
                                            want_dres=want_dres)
        return dx, dg, db, dres, None, None, None


class _BnReluPoolFn(torch.autograd.Function):
    """The stem after conv1, fused: (y, y_hi, y_lo) = maxpool3x3s2p1(relu(BN_batchstats(x))) on NHWC fp32.  The
    normalised map between bn1 and the maxpool (1.7 GB at B=16) is never written, forward or backward."""

    @staticmethod
    def forward(ctx, x, gamma, beta, bn, want_planes):
        ctx.set_materialize_grads(False)
        x = x.contiguous()
        track = bn.track_running_stats and bn.running_mean is not None
        mom = _momentum(bn, track)
        y, code, mean, inv, pl = ops.bn_relu_pool_fwd(x, gamma.detach(), beta.detach(), bn.running_mean if track else None,
                                                      bn.running_var if track else None, mom, bn.eps,
                                                      want_planes=want_planes, planes_dtype=F16X2)
        if track and bn.num_batches_tracked is not None:
            bn.num_batches_tracked += 1
        ctx.save_for_backward(x, code, mean, inv, gamma.detach(), beta.detach())
        if pl is None:
            return y, None, None
        ctx.mark_non_differentiable(pl.p0, pl.p1)
        return y, pl.p0, pl.p1

    @staticmethod
    def backward(ctx, dy, _dhi, _dlo):
        x, code, mean, inv, gamma, beta = ctx.saved_tensors
        if dy is None:
            return None, None, None, None, None
        dx, dg, db = ops.bn_relu_pool_bwd(x, code, dy.contiguous(), mean, inv, gamma, beta)
        return dx, dg, db, None, None


class _MaxPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, want_planes):
        ctx.set_materialize_grads(False)
        x = x.contiguous()
        y, code, pl = ops.maxpool3x3s2_fwd(x, want_planes=want_planes, planes_dtype=F16X2)
        ctx.save_for_backward(code)
        ctx.in_shape = tuple(x.shape)
        if pl is None:
            return y, None, None
        ctx.mark_non_differentiable(pl.p0, pl.p1)
        return y, pl.p0, pl.p1

    @staticmethod
    def backward(ctx, dy, _dhi, _dlo):
        (code,) = ctx.saved_tensors
        if dy is None:
            return None, None
        return ops.maxpool3x3s2_bwd(code, dy.contiguous(), ctx.in_shape), None


def _conv_bwd(dy_nhwc, x_nchw, weight, stride, padding, need_input, want_w=True):
    """Library backward of a convolution (cuDNN dgrad / wgrad through ATen) in plain fp32: TF32 is switched OFF here
    whatever the process default is (torch's default lets cuDNN use TF32 - narrower than the reference's CPU fp32);
    COVA_B200_TRAIN_TF32=1 turns it on for speed comparisons."""
    args = (dy_nhwc.permute(0, 3, 1, 2), x_nchw, weight, None, stride, padding, [1, 1], False, [0, 0], 1,
            [need_input, want_w, False])
    tf32 = os.environ.get("COVA_B200_TRAIN_TF32", "0") == "1"
    with torch.backends.cudnn.flags(enabled=True, allow_tf32=tf32):
        gi, gw, _ = torch.ops.aten.convolution_backward(*args)
    return gi, gw


class _StemConvFn(torch.autograd.Function):
    """conv1 (7x7 s2 p3, no bias) forward and wgrad on the tensor cores (raw fp32 NHWC output; the image needs no gradient).
    COVA_B200_TRAIN_WGRAD=library selects cuDNN (fp32, TF32 off) for comparison runs."""

    @staticmethod
    def forward(ctx, images, weight):
        out = ops.stem_conv_raw_fwd(images, ops.pack_stem_weight_f16x2(weight.detach().float().contiguous()))
        ctx.save_for_backward(images, weight)
        return out

    @staticmethod
    def backward(ctx, dy):
        images, weight = ctx.saved_tensors
        if os.environ.get("COVA_B200_TRAIN_WGRAD", "native") == "native":
            # pixel-contraction tcgen05 kernel on the raw image rows and the scaled split-fp16 planes of dy (stem_wgrad_tc.cu)
            dyp, inv = ops.split_planes_scaled(dy.contiguous(), F16X2)
            return None, ops.stem_wgrad(images, dyp, inv)
        x = images.float().div(255) if images.dtype == torch.uint8 else images
        # the gradient arrives NHWC: give cuDNN the (3-channel, cheap) image in channels_last as well, otherwise it
        # transposes the 1.7 GB gradient map to NCHW first (3.9 ms of a 21 ms step)
        x = x.contiguous(memory_format=torch.channels_last)
        _, gw = _conv_bwd(dy.contiguous(), x, weight, [2, 2], [3, 3], False)
        return None, gw


class _Conv3x3Fn(torch.autograd.Function):
    """3x3 s1 p1 64->64 convolution (no bias), forward, dgrad and wgrad on the tensor cores from split-fp16 planes.
    The output gradient is split ONCE into planes scaled by a per-tensor power of two (chosen on the device from max|dy|:
    gradient magnitudes shrink as the sum-reduced loss converges, and unscaled fp16 planes would floor them at 2^-25);
    dgrad = the forward kernel on those planes with the rotated, channel-swapped filter and 1/s as its epilogue scale;
    wgrad = the pixel-contraction kernel (wgrad_tc.cu) on the planes of x and dy, 1/s applied when the sum is finalised.
    COVA_B200_TRAIN_DGRAD / COVA_B200_TRAIN_WGRAD = library select cuDNN (fp32, TF32 off) for comparison runs."""

    @staticmethod
    def forward(ctx, x, x_hi, x_lo, weight):
        pl = ops.Planes.__new__(ops.Planes)
        pl.dtype, pl.shape, pl.p0, pl.p1 = F16X2, tuple(x.shape), x_hi, x_lo
        w_hi, w_lo = ops.pack_conv_weight_f16x2(weight.detach().float())
        one, zero = _ones_zeros(x.device)
        y = ops.conv3x3_bn_act_fwd(pl, w_hi, w_lo, one, zero, res=None, relu=False, out_dtype=F32, engine=ENGINE_TCGEN05)
        native_w = os.environ.get("COVA_B200_TRAIN_WGRAD", "native") == "native"
        ctx.native_w = native_w
        if native_w:
            ctx.save_for_backward(x_hi, x_lo, weight)        # the planes the forward consumed are the wgrad operand
        else:
            ctx.save_for_backward(x, weight)
        ctx.x_shape = tuple(x.shape)
        return y.p0

    @staticmethod
    def backward(ctx, dy):
        dy = dy.contiguous()
        native_d = os.environ.get("COVA_B200_TRAIN_DGRAD", "native") == "native"
        if ctx.native_w:
            x_hi, x_lo, weight = ctx.saved_tensors
            x = None
        else:
            x, weight = ctx.saved_tensors
        need_dx = ctx.needs_input_grad[0]
        gi = gw = None
        dyp = inv = None
        if ctx.native_w or (native_d and need_dx):
            dyp, inv = ops.split_planes_scaled(dy, F16X2)
        if need_dx:
            if native_d:
                # dgrad on the tensor cores: the gradient w.r.t. the input of a 3x3 s1 p1 convolution IS a 3x3 s1 p1
                # convolution of the output gradient with the filter rotated by 180 degrees and its channel axes swapped
                w_rot = weight.detach().float().flip(2, 3).transpose(0, 1).contiguous()
                w_hi, w_lo = ops.pack_conv_weight_f16x2(w_rot)
                _, zero = _ones_zeros(dy.device)
                gi = ops.conv3x3_bn_act_fwd(dyp, w_hi, w_lo, inv[:64], zero, res=None, relu=False, out_dtype=F32,
                                            engine=ENGINE_TCGEN05).p0
            else:
                xs = x if x is not None else (x_hi.float() + x_lo.float())
                gi, _ = _conv_bwd(dy, xs.permute(0, 3, 1, 2), weight, [1, 1], [1, 1], True, want_w=False)
                gi = gi.permute(0, 2, 3, 1)
        if ctx.native_w:
            xp = ops.Planes.__new__(ops.Planes)
            xp.dtype, xp.shape, xp.p0, xp.p1 = F16X2, ctx.x_shape, x_hi, x_lo
            gw = ops.conv3x3_wgrad(xp, dyp, inv)
        else:
            _, gw = _conv_bwd(dy, x.permute(0, 3, 1, 2), weight, [1, 1], [1, 1], False)
        return gi, None, None, gw


def _planes(dtype, shape, hi, lo):
    pl = ops.Planes.__new__(ops.Planes)
    pl.dtype, pl.shape, pl.p0, pl.p1 = dtype, tuple(shape), hi, lo
    return pl


class _Conv1x1Fn(torch.autograd.Function):
    """ResNet-50 Bottleneck 1x1 convolution (conv1 / conv3 / downsample; 64->64, 64->256, 256->64), forward, dgrad and wgrad
    on the tensor cores in the split-fp16 three-product mode: forward = persistent GEMM over the pixel rows of the input
    planes (pw_tc.cu), dgrad = the same kernel with the transposed filter on the scaled planes of the output gradient,
    wgrad = the pixel-contraction kernel (pw_wgrad_tc.cu)."""

    @staticmethod
    def forward(ctx, x, x_hi, x_lo, weight):
        w2 = weight.detach().float().flatten(1)
        y = ops.conv1x1_raw_fwd(_planes(F16X2, x.shape, x_hi, x_lo), ops.pack_linear_weight_f16x2(w2))
        ctx.save_for_backward(x_hi, x_lo, weight)
        ctx.x_shape = tuple(x.shape)
        return y

    @staticmethod
    def backward(ctx, dy):
        x_hi, x_lo, weight = ctx.saved_tensors
        dyp, inv = ops.split_planes_scaled(dy.contiguous(), F16X2)
        gi = None
        if ctx.needs_input_grad[0]:
            wt = weight.detach().float().flatten(1).t().contiguous()                  # [Cin, Cout]: dX = dY W
            gi = ops.conv1x1_raw_fwd(dyp, ops.pack_linear_weight_f16x2(wt), scale=inv)
        gw = ops.conv1x1_wgrad(_planes(F16X2, ctx.x_shape, x_hi, x_lo), dyp, inv)
        return gi, None, None, gw


class _LibConvFn(torch.autograd.Function):
    """A convolution that has no kernel of this library yet (interim): cuDNN through ATen on NHWC views, forward AND
    backward pinned to plain fp32 (the backward of an `F.conv2d` would otherwise run under whatever TF32 flag is current
    when `loss.backward()` executes)."""

    @staticmethod
    def forward(ctx, x, weight, stride, padding):
        tf32 = os.environ.get("COVA_B200_TRAIN_TF32", "0") == "1"
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=tf32):
            y = F.conv2d(x.permute(0, 3, 1, 2), weight, None, stride, padding)             # channels_last in and out
        ctx.save_for_backward(x, weight)
        ctx.meta = (list(stride), list(padding))
        return y.permute(0, 2, 3, 1)

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        stride, padding = ctx.meta
        gi, gw = _conv_bwd(dy.contiguous(), x.permute(0, 3, 1, 2), weight, stride, padding, ctx.needs_input_grad[0])
        return (gi.permute(0, 2, 3, 1) if gi is not None else None), gw, None, None


def tc_forward_convs():
    """The training FORWARD convolutions run on the tensor cores in the split-fp16 three-product mode (default;
    COVA_B200_TRAIN_CONV=cudnn selects plain fp32 library convolutions instead).  Precision matters here more than on
    the inference path: the decoder's train-mode BatchNorm1d makes the GRADIENTS jump by 1-2 % when the feature map is
    perturbed at the 1e-5 level (measured both with the split-bf16 forward and with 1e-5 noise injected into the
    all-library path, profiles/r01l_train_grad_conditioning.txt), while the split-fp16 forward (logits to 5e-6) leaves
    every gradient within 1.2e-5 of the live-reference fixture - the same as the all-library path."""
    return os.environ.get("COVA_B200_TRAIN_CONV", "tcgen05") == "tcgen05"


def _tc_conv_ok(conv):
    """3 = the 3x3 64->64 kernel, 1 = the 1x1 kernels (64->64, 64->256, 256->64), 0 = no tensor-core kernel here."""
    if not tc_forward_convs() or conv.bias is not None or conv.groups != 1 or conv.stride != (1, 1):
        return 0
    if conv.kernel_size == (3, 3) and conv.padding == (1, 1) and conv.in_channels == 64 and conv.out_channels == 64:
        return 3
    if (conv.kernel_size == (1, 1) and conv.padding == (0, 0) and os.environ.get("COVA_B200_TRAIN_CONV1X1", "native") == "native"
            and (conv.in_channels, conv.out_channels) in ((64, 64), (64, 256), (256, 64))):
        return 1
    return 0


def _conv(x, planes, conv):
    """x NHWC fp32 (+ its split planes, or None) -> raw conv output NHWC fp32."""
    kind = _tc_conv_ok(conv) if planes is not None else 0
    if kind == 3:
        return _Conv3x3Fn.apply(x, planes[0], planes[1], conv.weight)
    if kind == 1:
        return _Conv1x1Fn.apply(x, planes[0], planes[1], conv.weight)
    return _LibConvFn.apply(x, conv.weight, tuple(conv.stride), tuple(conv.padding))


def _bn_act(x, bn, res=None, relu=True, planes_for=None):
    """Returns (y, planes or None); planes are produced when the consumer `planes_for` is a tensor-core convolution."""
    want = planes_for is not None and bool(_tc_conv_ok(planes_for))
    y, hi, lo = _BnActFn.apply(x, bn.weight, bn.bias, res, bn, relu, want)
    return y, ((hi, lo) if want else None)


def feature_map_train_modular(convnet, images):
    """`convnet(images)` (`models.py:125`) in training mode -> NHWC fp32 feature map [B, H/4, W/4, C]."""
    blocks = list(convnet[4])
    first = blocks[0].conv1
    if tc_forward_convs():
        x = _StemConvFn.apply(images if images.dtype == torch.uint8 else images.float(), convnet[0].weight)   # conv1
    else:
        img = images.float().div(255) if images.dtype == torch.uint8 else images.float()
        x = _conv(img.permute(0, 2, 3, 1).contiguous(), None, convnet[0])
    want = bool(_tc_conv_ok(first))
    # COVA_B200_TRAIN_STEM_FUSED=1: bn1 + ReLU + maxpool in one forward / two backward kernels that never write the
    # normalised 640x640 map nor its gradient (3.4 GB less memory at B=16).  Measured no faster than the separate
    # passes (19.2 vs 18.1 ms/step: the backward gather is index-math bound), so it is a memory option, off by default.
    if os.environ.get("COVA_B200_TRAIN_STEM_FUSED", "0") == "1":
        x, hi, lo = _BnReluPoolFn.apply(x, convnet[1].weight, convnet[1].bias, convnet[1], want)          # bn1+relu+maxpool
    else:
        x, _ = _bn_act(x, convnet[1], relu=True)                                                          # bn1 + relu
        x, hi, lo = _MaxPoolFn.apply(x, want)                                                             # maxpool
    xp = (hi, lo) if want else None
    for bi, blk in enumerate(blocks):                                   # layer1
        nxt = blocks[bi + 1].conv1 if bi + 1 < len(blocks) else None
        if hasattr(blk, "conv3"):                                       # Bottleneck (torchvision resnet.py:143-163)
            o, op = _bn_act(_conv(x, xp, blk.conv1), blk.bn1, planes_for=blk.conv2)
            o, op = _bn_act(_conv(o, op, blk.conv2), blk.bn2, planes_for=blk.conv3)
            if blk.downsample is None:
                idt = x
            else:
                idt, _ = _bn_act(_conv(x, xp, blk.downsample[0]), blk.downsample[1], relu=False)
            x, xp = _bn_act(_conv(o, op, blk.conv3), blk.bn3, res=idt, planes_for=nxt)
        else:                                                           # BasicBlock (resnet.py:89-105)
            o, op = _bn_act(_conv(x, xp, blk.conv1), blk.bn1, planes_for=blk.conv2)
            x, xp = _bn_act(_conv(o, op, blk.conv2), blk.bn2, res=x, planes_for=nxt)
    return x


# ----------------------------------------------------------------------------- whole-backbone autograd function
class _Unit:
    """Saved state of one conv -> BatchNorm(batch statistics) (+ residual) (+ ReLU) unit of the backbone."""
    __slots__ = ("kind", "xin", "weight", "bn", "raw", "mean", "inv", "res", "relu", "has_res", "y", "planes", "x_shape", "mask")


def _unit_fwd(kind, xin, conv, bn, res=None, relu=True, want_y=True, want_planes=False):
    """kind "stem": xin = images; 3 / 1: xin = split-fp16 Planes of the input map.  Returns the unit (y fp32 and / or the
    split planes of y for the tensor-core convolution that follows)."""
    u = _Unit()
    u.kind, u.xin, u.weight, u.bn, u.relu, u.has_res = kind, xin, conv.weight, bn, relu, res is not None
    w = conv.weight.detach().float()
    # the convolution's epilogue accumulates the BatchNorm batch statistics of its output (COVA_B200_TRAIN_EPI_STATS=0: separate pass)
    sw = ops.new_stats_ws(conv.out_channels, w.device) if _epi_stats(kind) else None
    if kind == "stem":
        u.raw = ops.stem_conv_raw_fwd(xin, ops.pack_stem_weight_f16x2(w.contiguous()), stats_ws=sw)
    elif kind == 3:
        w_hi, w_lo = ops.pack_conv_weight_f16x2(w)
        one, zero = _ones_zeros(w.device)
        u.raw = ops.conv3x3_bn_act_fwd(xin, w_hi, w_lo, one, zero, res=None, relu=False, out_dtype=F32, engine=ENGINE_TCGEN05,
                                       stats_ws=sw).p0
    else:
        u.raw = ops.conv1x1_raw_fwd(xin, ops.pack_linear_weight_f16x2(w.flatten(1)), stats_ws=sw)
    track = bn.track_running_stats and bn.running_mean is not None
    mom = _momentum(bn, track)
    u.res = None
    u.mask = (torch.empty(u.raw.numel() // 4, dtype=torch.uint8, device=u.raw.device)
              if (relu and res is not None and _relu_mask()) else None)    # 4 decisions per byte (one per thread of the passes)
    if u.mask is None and relu and res is not None:
        u.res = res
    u.y, u.mean, u.inv, u.planes = ops.bn_train_fwd(u.raw, bn.weight.detach(), bn.bias.detach(), bn.running_mean if track else None,
                                                    bn.running_var if track else None, mom, bn.eps, res=res, relu=relu,
                                                    want_planes=want_planes, planes_dtype=F16X2, want_y=want_y, stats_ws=sw,
                                                    relu_mask=u.mask)
    if track and bn.num_batches_tracked is not None:
        bn.num_batches_tracked += 1
    return u


def _epi_stats(kind=3):
    """Which convolutions accumulate their output's BatchNorm statistics in the epilogue: COVA_B200_TRAIN_EPI_STATS = "3" (the
    tensor-bound 3x3 convolutions, default), "all" (also the HBM-bound 1x1 / conv1 kernels, whose epilogues have no issue slots
    to spare: measured slower than the separate pass) or "0" (none)."""
    v = os.environ.get("COVA_B200_TRAIN_EPI_STATS", "3")
    return v == "all" or (v == "3" and kind == 3)


def _unit_bwd(u, dy, need_dx=True, add=None):
    """dy = gradient of the unit's output (fp32 NHWC).  BatchNorm backward emits the gradient of the raw convolution output
    directly as scaled split-fp16 planes (no fp32 map, no max / split passes); dgrad and wgrad consume them on the tensor
    cores.  Returns (dx fp32 or None, dw, dgamma, dbeta, dres or None)."""
    bn = u.bn
    dyp, inv, dres, dg, db = ops.bn_train_bwd_planes(dy, u.raw, u.mean, u.inv, bn.weight.detach(), bn.bias.detach(), res=u.res,
                                                     relu=u.relu, want_dres=u.has_res, planes_dtype=F16X2,
                                                     relu_mask=getattr(u, "mask", None))
    u.mask = None
    w = u.weight.detach().float()
    dx = None
    if u.kind == "stem":
        dw = ops.stem_wgrad(u.xin, dyp, inv)
    elif u.kind == 3:
        if need_dx:        # `add` (the skip branch's fp32 gradient) rides in the dgrad convolution's epilogue: no add pass
            w_hi, w_lo = ops.pack_conv_weight_f16x2(w.flip(2, 3).transpose(0, 1).contiguous())
            _, zero = _ones_zeros(dy.device)
            if add is None:
                dx = ops.conv3x3_bn_act_fwd(dyp, w_hi, w_lo, inv[:64], zero, res=None, relu=False, out_dtype=F32,
                                            engine=ENGINE_TCGEN05).p0
            else:
                dx = ops.conv3x3_scale_res_f32_fwd(dyp, w_hi, w_lo, inv[:64], zero, add)
        dw = ops.conv3x3_wgrad(u.xin, dyp, inv)
    else:
        if need_dx:
            wt = ops.pack_linear_weight_f16x2(w.flatten(1).t().contiguous())
            dx = ops.conv1x1_raw_fwd(dyp, wt, scale=inv) if add is None else ops.conv1x1_raw_res_f32_fwd(dyp, wt, inv, add)
        dw = ops.conv1x1_wgrad(u.xin, dyp, inv)
    u.raw = u.res = u.xin = u.mean = u.inv = None        # release the unit's saved maps as the backward walks up the network
    return dx, dw, dg, db, dres


def _relu_mask():
    """Residual units hand a ReLU bit mask to their backward (COVA_B200_TRAIN_RELU_MASK=0: re-read the residual)."""
    return os.environ.get("COVA_B200_TRAIN_RELU_MASK", "1") != "0"


def _unit_fwd16(kind, xin, conv, bn, res=None, relu=True, out_fp32=False):
    """bf16 training mode (BASELINE config 3): the unit with every map stored as bf16 - raw convolution output, y (one bf16
    plane = the operand of the next convolution, one product per MMA) - statistics / parameters fp32.  xin = images (stem)
    or the bf16 NHWC input map.  out_fp32: y as fp32 (the final feature map, read by RoIPool)."""
    u = _Unit()
    u.kind, u.xin, u.weight, u.bn, u.relu, u.has_res = kind, xin, conv.weight, bn, relu, res is not None
    w = conv.weight.detach().float()
    sw = ops.new_stats_ws(conv.out_channels, w.device) if _epi_stats(kind) else None
    if kind == "stem":
        u.raw = ops.stem_conv_raw_fwd_bf16(xin, ops.pack_stem_weight(w.contiguous()), stats_ws=sw)
    elif kind == 3:
        _, w_hi, _ = ops.pack_conv_weight(w, simt=False, tc=True, split=False)
        one, zero = _ones_zeros(w.device)
        u.raw = ops.conv3x3_bn_act_fwd(ops.bf16_plane(xin), w_hi, None, one, zero, res=None, relu=False, out_dtype=ops.BF16,
                                       engine=ENGINE_TCGEN05, stats_ws=sw).p0
    else:
        u.raw = ops.conv1x1_raw_fwd(ops.bf16_plane(xin), w.flatten(1).to(torch.bfloat16).contiguous(), stats_ws=sw)
    track = bn.track_running_stats and bn.running_mean is not None
    mom = _momentum(bn, track)
    # residual units keep the forward's ReLU decisions as a bit mask (1/16 of the residual map's bytes): the two backward passes
    # then do not read the residual at all
    u.res = None
    u.mask = (torch.empty(u.raw.numel() // 8, dtype=torch.uint8, device=u.raw.device)
              if (relu and res is not None and _relu_mask()) else None)
    if u.mask is None and relu and res is not None:
        u.res = res
    u.y, u.mean, u.inv = ops.bn_train_fwd_t(u.raw, bn.weight.detach(), bn.bias.detach(), bn.running_mean if track else None,
                                            bn.running_var if track else None, mom, bn.eps, res=res, relu=relu,
                                            out_dtype=torch.float32 if out_fp32 else torch.bfloat16, stats_ws=sw, relu_mask=u.mask)
    u.planes = None
    if track and bn.num_batches_tracked is not None:
        bn.num_batches_tracked += 1
    return u


def _unit_bwd16(u, dy, need_dx=True, add=None):
    """Backward of `_unit_fwd16`: dy fp32 (the feature-map gradient) or bf16; bf16 gradient maps throughout (bf16 has fp32's
    exponent range: no scaling), fp32 weight gradients from the single-plane tensor-core wgrad kernels."""
    bn = u.bn
    dxr, dres, dg, db = ops.bn_train_bwd_t(dy, u.raw, u.mean, u.inv, bn.weight.detach(), bn.bias.detach(), res=u.res, relu=u.relu,
                                           want_dres=u.has_res, relu_mask=getattr(u, "mask", None))
    u.mask = None
    dyp = ops.bf16_plane(dxr)
    w = u.weight.detach().float()
    dx = None
    if u.kind == "stem":
        dw = ops.stem_wgrad(u.xin, dyp, None)
    elif u.kind == 3:
        if need_dx:        # `add` (the skip branch's gradient) rides in the dgrad convolution's residual epilogue: no add pass
            _, w_hi, _ = ops.pack_conv_weight(w.flip(2, 3).transpose(0, 1).contiguous(), simt=False, tc=True, split=False)
            one, zero = _ones_zeros(dxr.device)
            dx = ops.conv3x3_bn_act_fwd(dyp, w_hi, None, one, zero, res=None if add is None else ops.bf16_plane(add), relu=False,
                                        out_dtype=ops.BF16, engine=ENGINE_TCGEN05).p0
        dw = ops.conv3x3_wgrad(ops.bf16_plane(u.xin), dyp, None)
    else:
        if need_dx:
            wt = w.flatten(1).t().to(torch.bfloat16).contiguous()
            dx = ops.conv1x1_raw_fwd(dyp, wt) if add is None else ops.conv1x1_raw_res_fwd_bf16(dxr, wt, add)
        dw = ops.conv1x1_wgrad(ops.bf16_plane(u.xin), dyp, None)
    u.raw = u.res = u.xin = u.mean = u.inv = None
    return dx, dw, dg, db, dres


class _BackboneFn(torch.autograd.Function):
    """`convnet(images)` in training mode as ONE autograd node: every convolution (forward, dgrad, wgrad) on the tensor
    cores in the split-fp16 three-product mode, BatchNorm(batch statistics) + residual + ReLU and the maxpool as the native
    NHWC passes, maps handed from kernel to kernel in the format the consumer reads (split planes for a convolution, fp32
    for a residual / the pooled map / RoIPool).  Parameters arrive as (conv.weight, bn.weight, bn.bias) per unit, in unit order."""

    @staticmethod
    def forward(ctx, images, convnet, bf16, *params):
        blocks = list(convnet[4])
        units = []
        ctx.bf16 = bf16
        if bf16:
            def unit(kind, src, conv, bn, res=None, relu=True, want_y=True, want_planes=False, last=False):
                return _unit_fwd16(kind, src, conv, bn, res=res, relu=relu, out_fp32=last)
            operand = lambda u: u.y
        else:
            def unit(kind, src, conv, bn, res=None, relu=True, want_y=True, want_planes=False, last=False):
                return _unit_fwd(kind, src, conv, bn, res=res, relu=relu, want_y=want_y, want_planes=want_planes)
            operand = lambda u: u.planes
        ctx.fused_stem = os.environ.get("COVA_B200_TRAIN_STEM_FUSED", "1") == "1"
        if ctx.fused_stem:
            # conv1 -> [bn1 + ReLU + maxpool fused: the normalised 640^2 map is never written, forward or backward]
            stem = _Unit()
            bn, w = convnet[1], convnet[0].weight.detach().float().contiguous()
            stem.kind, stem.xin, stem.weight, stem.bn = "stem", images, convnet[0].weight, bn
            sw = ops.new_stats_ws(64, w.device) if _epi_stats("stem") else None
            stem.raw = (ops.stem_conv_raw_fwd_bf16(images, ops.pack_stem_weight(w), stats_ws=sw) if bf16
                        else ops.stem_conv_raw_fwd(images, ops.pack_stem_weight_f16x2(w), stats_ws=sw))
            track = bn.track_running_stats and bn.running_mean is not None
            x, code, stem.mean, stem.inv, pl = ops.bn_relu_pool_fwd_t(
                stem.raw, bn.weight.detach(), bn.bias.detach(), bn.running_mean if track else None,
                bn.running_var if track else None, _momentum(bn, track), bn.eps, want_planes=not bf16, planes_dtype=F16X2, stats_ws=sw)
            if track and bn.num_batches_tracked is not None:
                bn.num_batches_tracked += 1
            stem.y = stem.planes = stem.res = None
            xp = x if bf16 else pl
            units.append(stem)
            ctx.pool = (code, None)
        else:
            stem = unit("stem", images, convnet[0], convnet[1], relu=True, want_y=True)
            units.append(stem)
            if bf16:
                x, code = ops.maxpool3x3s2_fwd_t(stem.y)
                xp = x
            else:
                x, code, xp = ops.maxpool3x3s2_fwd(stem.y, want_planes=True, planes_dtype=F16X2)
            ctx.pool = (code, tuple(stem.y.shape))
            stem.y = None                                            # the 640^2 map is not needed again (the ReLU mask comes from raw)
        plan = []
        for bi, blk in enumerate(blocks):
            last = bi + 1 == len(blocks)
            if hasattr(blk, "conv3"):                            # Bottleneck (torchvision resnet.py:143-163)
                u1 = unit(1, xp, blk.conv1, blk.bn1, want_y=False, want_planes=True)
                u2 = unit(3, operand(u1), blk.conv2, blk.bn2, want_y=False, want_planes=True)
                ud = None
                if blk.downsample is not None:
                    ud = unit(1, xp, blk.downsample[0], blk.downsample[1], relu=False, want_y=True)
                u3 = unit(1, operand(u2), blk.conv3, blk.bn3, res=(x if ud is None else ud.y), want_y=True, want_planes=not last,
                          last=last)
                group = [u1, u2, u3] + ([ud] if ud is not None else [])
                x, xp = u3.y, operand(u3)
            else:                                                # BasicBlock (resnet.py:89-105)
                u1 = unit(3, xp, blk.conv1, blk.bn1, want_y=False, want_planes=True)
                u2 = unit(3, operand(u1), blk.conv2, blk.bn2, res=x, want_y=True, want_planes=not last, last=last)
                group = [u1, u2]
                x, xp = u2.y, operand(u2)
            plan.append((len(units), len(group), hasattr(blk, "conv3"), blk.downsample is not None))
            units.extend(group)
        for u in units:            # outputs live on only where a consumer holds them: `xin` (operand), `res` (residual), the caller (x)
            u.y = u.planes = None
        ctx.units, ctx.plan = units, plan
        return x

    @staticmethod
    def backward(ctx, g):
        units, plan = ctx.units, ctx.plan
        ubwd = _unit_bwd16 if ctx.bf16 else _unit_bwd
        grads = [None] * len(units)
        g = g.contiguous()
        for start, n, bottleneck, has_ds in reversed(plan):
            us = units[start:start + n]
            if bottleneck:
                d3, *p3, dres = ubwd(us[2], g)
                d2, *p2, _ = ubwd(us[1], d3)
                if has_ds:
                    dd, *pd, _ = ubwd(us[3], dres)
                    grads[start + 3] = pd
                    dres = dd
                g, *p1, _ = ubwd(us[0], d2, add=dres)          # the skip gradient is added by conv1's dgrad epilogue
                grads[start + 2], grads[start + 1], grads[start] = p3, p2, p1
            else:
                d2, *p2, dres = ubwd(us[1], g)
                g, *p1, _ = ubwd(us[0], d2, add=dres)
                grads[start + 1], grads[start] = p2, p1
        code, shape = ctx.pool
        if ctx.fused_stem:
            u = units[0]
            bn = u.bn
            gam, bet = bn.weight.detach(), bn.bias.detach()
            in_shape = tuple(u.raw.shape)
            if os.environ.get("COVA_B200_TRAIN_STEM_FUSED_BWD", "0") == "1":
                # gather-in-the-passes backward (no dense pooled-gradient map at all): fewer bytes, but ~25 instructions per
                # element against an issue budget of ~20 - measured slower (2.6 vs 1.9 ms at B=16), kept as a memory option
                if ctx.bf16:
                    dxr, _, dg0, db0 = ops.bn_relu_pool_bwd_t(u.raw, code, g, u.mean, u.inv, gam, bet)
                    dyp, inv = ops.bf16_plane(dxr), None
                else:
                    dyp, inv, dg0, db0 = ops.bn_relu_pool_bwd_t(u.raw, code, g, u.mean, u.inv, gam, bet, planes=True, planes_dtype=F16X2)
            elif ctx.bf16:
                gd = ops.maxpool3x3s2_bwd_t(code, g, in_shape)
                dxr, _, dg0, db0 = ops.bn_train_bwd_t(gd, u.raw, u.mean, u.inv, gam, bet, res=None, relu=True)
                dyp, inv = ops.bf16_plane(dxr), None
            else:
                gd = ops.maxpool3x3s2_bwd(code, g, in_shape)
                dyp, inv, _, dg0, db0 = ops.bn_train_bwd_planes(gd, u.raw, u.mean, u.inv, gam, bet, res=None, relu=True, planes_dtype=F16X2)
            dw0 = ops.stem_wgrad(u.xin, dyp, inv)
            u.raw = u.xin = None
            grads[0] = [dw0, dg0, db0]
        else:
            g = ops.maxpool3x3s2_bwd_t(code, g, shape) if ctx.bf16 else ops.maxpool3x3s2_bwd(code, g, shape)
            _, *p0, _ = ubwd(units[0], g, need_dx=False)
            grads[0] = p0
        flat = []
        for dw, dg, db in grads:
            flat += [dw, dg, db]
        ctx.units = ctx.plan = ctx.pool = None
        return (None, None, None) + tuple(gr if need else None for gr, need in zip(flat, ctx.needs_input_grad[3:]))


def _fused_ok(convnet):
    if not tc_forward_convs() or os.environ.get("COVA_B200_TRAIN_BACKBONE", "native") == "modular":
        return False
    if os.environ.get("COVA_B200_TRAIN_WGRAD", "native") != "native" or os.environ.get("COVA_B200_TRAIN_DGRAD", "native") != "native":
        return False
    c0 = convnet[0]
    if not (c0.kernel_size == (7, 7) and c0.stride == (2, 2) and c0.padding == (3, 3) and c0.bias is None
            and c0.in_channels == 3 and c0.out_channels == 64):
        return False
    for blk in convnet[4]:
        convs = [(blk.conv1, 1 if hasattr(blk, "conv3") else 3), (blk.conv2, 3)]
        if hasattr(blk, "conv3"):
            convs.append((blk.conv3, 1))
        if blk.downsample is not None:
            convs.append((blk.downsample[0], 1))
        if any(_tc_conv_ok(c) != k for c, k in convs):
            return False
    return True


def _unit_params(convnet):
    mods = [(convnet[0], convnet[1])]
    for blk in convnet[4]:
        mods += [(blk.conv1, blk.bn1), (blk.conv2, blk.bn2)]
        if hasattr(blk, "conv3"):
            mods.append((blk.conv3, blk.bn3))
        if blk.downsample is not None:
            mods.append((blk.downsample[0], blk.downsample[1]))
    out = []
    for conv, bn in mods:
        out += [conv.weight, bn.weight, bn.bias]
    return out


def feature_map_train(convnet, images, precision="fp32"):
    """`convnet(images)` (`models.py:125`) in training mode -> NHWC fp32 feature map [B, H/4, W/4, C].  Default: the
    whole-backbone autograd node (`_BackboneFn`) in the fp32-parity mode (split-fp16 three-product convolutions, fp32 maps);
    precision="bf16" = the bf16 training mode (bf16 maps and gradient maps, one bf16 product per MMA; fp32 statistics,
    parameters and weight gradients).  COVA_B200_TRAIN_BACKBONE=modular (or any library-convolution switch) selects the
    per-operator autograd functions above."""
    if _fused_ok(convnet) and all(p is not None for p in _unit_params(convnet)):
        img = images if images.dtype == torch.uint8 else images.float()
        return _BackboneFn.apply(img, convnet, precision == "bf16", *_unit_params(convnet))
    if precision == "bf16":
        raise RuntimeError("cova_b200: the bf16 training mode needs the tensor-core backbone (a ResNet-18 / ResNet-50 layer1 stack)")
    return feature_map_train_modular(convnet, images)
