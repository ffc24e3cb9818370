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


def feature_map_train(convnet, images):
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
