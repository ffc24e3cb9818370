"""Drop-in replacement for the reference's ``models.py``: same ``CoVA`` / ``GraphAttentionLayer`` classes,
constructor arguments, ``forward`` signature, attributes and ``state_dict`` keys
(`/root/reference/models.py:9-212`), so the reference's ``main.py`` / ``train.py`` /
``evaluate.py`` / ``extract_attn_wts_and_visualize.py`` run unchanged and its checkpoints load.

Execution:
  * inference (``torch.no_grad()``, ``model.eval()``, CUDA tensors) runs the hand-written sm_100a kernels of
    ``libcova_b200.so`` through ``engine.NativeForward`` - engine ``"tcgen05"`` (tensor-core convolutions,
    precision ``"fp32"`` = split-bf16 3-product fp32-parity mode (1e-5), ``"fp32x"`` = the same three products on
    split-fp16 planes (22 significand bits: ~1e-6, same cost; ResNet-18 stack), ``"fp16"`` = one fp16 product (ResNet-18 backbone;
    ~5e-4 on the logits, inside the 1e-3 bar), or ``"bf16"`` (3-4e-3: outside the bar)) or ``"simt"`` (exact fp32
    CUDA cores).
  * anything that needs autograd (``train.py:45-60``) runs the autograd path.  In ``model.train()`` the backbone goes
    through ``train_backbone.feature_map_train`` (NHWC end to end: tensor-core forward convolutions in the split-fp16
    three-product mode, native BatchNorm(batch statistics) + residual + ReLU and maxpool, forward and backward); RoIPool
    and the GAT gather have native forward / backward kernels, the 3x3 convolutions a native dgrad.  The convolutions'
    wgrad, the small linear layers and
    BatchNorm1d are PyTorch operators (cuDNN / cuBLAS = library code, interim - DESIGN.md section 9).  With autograd on
    in ``eval()`` mode (running statistics) the backbone is the plain PyTorch module.
  * non-CUDA inputs raise: this package has no CPU path.

Extra keyword-only constructor arguments (all with reference-preserving defaults):
``backbone`` ("resnet18" | "resnet50"), ``n_heads``, ``roi_mode`` ("pool" | "align"), ``engine``, ``precision``.
"""
import os
import warnings

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .engine import NativeForward


def count_parameters(model):
    """`/root/reference/utils.py:37-41`."""
    return sum(p.numel() for p in model.parameters() if p.requires_grad)


# ----------------------------------------------------------------------------- backbone blocks
# Same attribute names as torchvision's BasicBlock / Bottleneck so `convnet.4.{b}.*` checkpoint keys match.
class BasicBlock(nn.Module):
    def __init__(self, planes=64):
        super().__init__()
        self.conv1 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = None

    def forward(self, x):
        out = self.relu(self.bn1(self.conv1(x)))
        out = self.bn2(self.conv2(out))
        return self.relu(out + x)


class Bottleneck(nn.Module):
    def __init__(self, inplanes, planes=64, downsample=False):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = (nn.Sequential(nn.Conv2d(inplanes, planes * 4, 1, bias=False), nn.BatchNorm2d(planes * 4))
                           if downsample else None)

    def forward(self, x):
        out = self.relu(self.bn1(self.conv1(x)))
        out = self.relu(self.bn2(self.conv2(out)))
        out = self.bn3(self.conv3(out))
        idt = x if self.downsample is None else self.downsample(x)
        return self.relu(out + idt)


def _truncated_resnet(backbone):
    """`list(resnet.children())[:-5]` (`models.py:49-51`): conv1, bn1, relu, maxpool, layer1."""
    if backbone == "resnet18":
        layer1, C = nn.Sequential(BasicBlock(), BasicBlock()), 64
    elif backbone == "resnet50":
        layer1, C = nn.Sequential(Bottleneck(64, 64, True), Bottleneck(256), Bottleneck(256)), 256
    else:
        raise ValueError("backbone must be 'resnet18' or 'resnet50'")
    net = nn.Sequential(nn.Conv2d(3, 64, 7, 2, 3, bias=False), nn.BatchNorm2d(64), nn.ReLU(inplace=True),
                        nn.MaxPool2d(3, 2, 1), layer1)
    for mod in net.modules():   # torchvision ResNet init
        if isinstance(mod, nn.Conv2d):
            nn.init.kaiming_normal_(mod.weight, mode="fan_out", nonlinearity="relu")
    return net, C


_PRETRAINED_FILES = {"resnet18": "resnet18-f37072fd.pth", "resnet50": "resnet50-0676ba61.pth"}


def _try_load_pretrained(convnet, backbone):
    """The reference starts from ImageNet weights (`models.py:49`, `pretrained=True`).  Use them when the
    torchvision checkpoint is already in the local torch-hub cache; never touch the network.  Otherwise
    keep the random init and say so."""
    path = os.path.join(torch.hub.get_dir(), "checkpoints", _PRETRAINED_FILES[backbone])
    if not os.path.exists(path):
        warnings.warn("cova_b200: pretrained %s weights not in %s - backbone is randomly initialised; "
                      "load a checkpoint with load_state_dict()" % (backbone, os.path.dirname(path)))
        return False
    src = torch.load(path, map_location="cpu")
    remap = {"conv1.": "0.", "bn1.": "1.", "layer1.": "4."}
    sd = {new + k[len(old):]: v for k, v in src.items() for old, new in remap.items() if k.startswith(old)}
    convnet.load_state_dict(sd, strict=True)
    return True


# ----------------------------------------------------------------------------- RoI autograd (training path)
class _RoIPoolFn(torch.autograd.Function):
    """Native RoIPool forward (NHWC fp32 feature map, arg-max kept) and native backward (atomic scatter-add of
    every output's gradient to its arg-max pixel, as torchvision's roi_pool backward)."""

    @staticmethod
    def forward(ctx, fm_nhwc, rois, P, scale):
        B, Hf, Wf, C = fm_nhwc.shape
        out = torch.empty((rois.shape[0], C * P[0] * P[1]), dtype=torch.float32, device=fm_nhwc.device)
        argmax = ops.roi_fwd(fm_nhwc.contiguous(), rois, P, scale, out, mode="pool", want_argmax=True)
        ctx.save_for_backward(argmax, rois)
        ctx.shape = (B, Hf, Wf, C)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        argmax, rois = ctx.saved_tensors
        return ops.roi_pool_bwd(grad_out.float(), argmax, rois, ctx.shape), None, None, None


class _RoIAlignFn(torch.autograd.Function):
    """Native RoIAlign(P, scale, sampling_ratio=2, aligned=False) forward and backward (the RoI op
    BASELINE.json's north_star names; atomic scatter of the 4 bilinear taps of every sample, as torchvision's)."""

    @staticmethod
    def forward(ctx, fm_nhwc, rois, P, scale):
        B, Hf, Wf, C = fm_nhwc.shape
        out = torch.empty((rois.shape[0], C * P[0] * P[1]), dtype=torch.float32, device=fm_nhwc.device)
        ops.roi_fwd(fm_nhwc.contiguous(), rois, P, scale, out, mode="align", sampling_ratio=2)
        ctx.save_for_backward(rois)
        ctx.meta = ((B, Hf, Wf, C), P, scale)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (rois,) = ctx.saved_tensors
        shape, P, scale = ctx.meta
        return ops.roi_align_bwd(grad_out.float(), rois, P, scale, shape, 2), None, None, None


class _GatGatherFn(torch.autograd.Function):
    """Native fused neighbour gather / masked softmax / weighted sum (cova_gat_fwd) with its native backward
    (cova_gat_bwd).  Input `ext` = [W_j h | a_i.W_i h + b | a_j.W_j h | pad] comes from a differentiable GEMM whose bias
    row carries the attention bias b, so the gradients of W_i, W_j, the attention vector and b flow through autograd
    and no scalar has to be read back to the host (the kernels get att_b = 0)."""

    @staticmethod
    def forward(ctx, ext, ctx_idx, Hd, alpha):
        ext = ext.contiguous()
        out = torch.empty((ext.shape[0], Hd), dtype=torch.float32, device=ext.device)
        attn = ops.gat_fwd(ext[:, :Hd], ext[:, Hd], ext[:, Hd + 1], 0.0, alpha, ctx_idx, out, want_attn=True)
        ctx.save_for_backward(ext, ctx_idx, attn)
        ctx.meta = (Hd, alpha)
        ctx.mark_non_differentiable(attn)
        return out, attn

    @staticmethod
    def backward(ctx, grad_out, _grad_attn):
        ext, ctx_idx, attn = ctx.saved_tensors
        Hd, alpha = ctx.meta
        d_ext, _ = ops.gat_bwd(grad_out.float(), ext, Hd, 0.0, alpha, ctx_idx, attn)
        return d_ext, None, None, None


class _LinearFn(torch.autograd.Function):
    """`nn.Linear` on the autograd path (`models.py:65-70,82-90,160-164` under `train.py:45-60`) without cuBLAS: forward,
    dX = dY W and dW = dY^T X all run the library's own GEMM kernel (`cova_linear_fwd`: tcgen05 split-bf16 three-product,
    or the exact CUDA-core engine when K % 8 != 0 - the 5-feature bbox encoder, the 4-class output layer)."""

    @staticmethod
    def _mm(a, b_nk, bias=None, precise=False):
        """a [M,K] @ b_nk[N,K]^T (+ bias) through cova_linear_fwd.  precise = split-fp16 operands (22 bits; the forward:
        the gradients of the train-mode BatchNorm1d that follows jump by 1-2 % when its input moves at the 1e-5 level),
        else split-bf16 (16 bits, full fp32 range: the gradient GEMMs)."""
        a = a.contiguous()
        b_nk = b_nk.contiguous()
        if ops.linear_tc_ok(a):
            if precise:
                return ops.linear_fwd(a, ops.pack_linear_weight_f16x2(b_nk), bias=bias, engine=ops.ENGINE_TCGEN05_F16X2)
            return ops.linear_fwd(a, ops.pack_linear_weight(b_nk), bias=bias, engine=ops.ENGINE_TCGEN05)
        return ops.linear_fwd(a, b_nk, bias=bias, engine=ops.ENGINE_SIMT)

    @staticmethod
    def forward(ctx, x, weight, bias):
        x = x.float()
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return _LinearFn._mm(x, weight.detach().float(), None if bias is None else bias.detach().float().contiguous(),
                             precise=True)

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dy = dy.float().contiguous()
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = _LinearFn._mm(dy, weight.detach().float().t())                 # [M,N] @ [K,N]^T
        if ctx.needs_input_grad[1]:
            dw = _LinearFn._mm(dy.t(), x.t())                                   # [N,M] @ [K,M]^T
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = dy.sum(0)
        return dx, dw, db


def _native_linear(x, weight, bias):
    if os.environ.get("COVA_B200_TRAIN_LINEAR", "native") == "native" and x.is_cuda and x.shape[0] > 0:
        return _LinearFn.apply(x, weight, bias)
    return F.linear(x, weight, bias)


def _run_sequential(seq, x):
    """`seq(x)` with every nn.Linear routed through `_LinearFn` (same modules, same parameters, same order)."""
    for mod in seq:
        x = _native_linear(x, mod.weight, mod.bias) if isinstance(mod, nn.Linear) else mod(x)
    return x


# ----------------------------------------------------------------------------- GAT
class GraphAttentionLayer(nn.Module):
    """Simple GAT layer, similar to https://arxiv.org/abs/1710.10903 (`/root/reference/models.py:151-212`)."""

    def __init__(self, in_features, hidden_dim, alpha=0.2):
        super().__init__()
        self.in_features = in_features
        self.hidden_dim = hidden_dim
        self.W_i = nn.Linear(in_features, hidden_dim, bias=False)
        self.W_j = nn.Linear(in_features, hidden_dim, bias=False)
        self.attention_layer = nn.Linear(2 * hidden_dim, 1)
        self.leakyrelu = nn.LeakyReLU(alpha)
        self._native = None

    def forward(self, h_i, context_indices, return_attn_wts=False):
        """h_i [N,in_features]; context_indices int64 [N,n_context] (ids 0..N-1, -1 = padding)."""
        if not h_i.is_cuda:
            raise RuntimeError("cova_b200: GraphAttentionLayer needs CUDA tensors (no CPU path)")
        if torch.is_grad_enabled() and (h_i.requires_grad or any(p.requires_grad for p in self.parameters())):
            return self._forward_composite(h_i, context_indices, return_attn_wts)
        key = (ops.param_generation,) + tuple((p.data_ptr(), p._version) for p in self.parameters())
        if self._native is None or self._native[0] != key:
            self._native = (key, NativeForward._gat_cache(self))
        g = self._native[1]
        h = h_i.float()
        if h.stride(-1) != 1:
            h = h.contiguous()
        out = torch.empty((h.shape[0], self.hidden_dim), dtype=torch.float32, device=h.device)
        ext = ops.linear_fwd(h, g["ext"])
        Hd = self.hidden_dim
        attn = ops.gat_fwd(ext[:, :Hd], ext[:, Hd], ext[:, Hd + 1], g["b"], g["alpha"], context_indices, out,
                           want_attn=return_attn_wts)
        return (out, attn) if return_attn_wts else out

    def _forward_composite(self, h_i, context_indices, return_attn_wts=False):
        """Autograd path: the projections are one differentiable GEMM against [W_j ; a_i W_i ; a_j W_j] (the same
        restructuring as the inference path), the gather / softmax / weighted sum and its backward are the native
        kernels.  COVA_B200_GAT_TRAIN=torch selects the pure PyTorch-operator formulation below instead."""
        if os.environ.get("COVA_B200_GAT_TRAIN", "native") != "torch":
            Hd = self.hidden_dim
            a = self.attention_layer.weight[0]
            pad = torch.zeros((2, self.in_features), dtype=h_i.dtype, device=h_i.device)
            ext_w = torch.cat((self.W_j.weight, (a[:Hd] @ self.W_i.weight)[None], (a[Hd:] @ self.W_j.weight)[None], pad), 0)
            zb = torch.zeros(Hd + 4, dtype=h_i.dtype, device=h_i.device)
            ext_b = torch.cat((zb[:Hd], self.attention_layer.bias, zb[:3]))      # b rides in the s column
            out, attn = _GatGatherFn.apply(_native_linear(h_i, ext_w, ext_b), context_indices, Hd,
                                           float(self.leakyrelu.negative_slope))
            return (out, attn) if return_attn_wts else out
        N, K = context_indices.shape
        Hd = self.hidden_dim
        a = self.attention_layer.weight[0]
        Wh_i, Wh_j = self.W_i(h_i), self.W_j(h_i)
        s, t = Wh_i @ a[:Hd], Wh_j @ a[Hd:]
        valid = context_indices >= 0
        idx = context_indices.clamp_min(0)
        e = self.leakyrelu(s[:, None] + t[idx] + self.attention_layer.bias)
        e = torch.where(valid, e, torch.full_like(e, -9e15))
        attn = torch.softmax(e, dim=1)
        nb = Wh_j[idx.reshape(-1)].view(N, K, Hd) * valid.unsqueeze(-1)
        out = (attn.unsqueeze(-1) * nb).sum(1)
        return (out, attn) if return_attn_wts else out


class MultiHeadGAT(nn.Module):
    """`n_heads` independent reference layers of width hidden_dim // n_heads, concatenated."""

    def __init__(self, in_features, hidden_dim, n_heads):
        super().__init__()
        assert hidden_dim % n_heads == 0
        self.heads = nn.ModuleList(GraphAttentionLayer(in_features, hidden_dim // n_heads) for _ in range(n_heads))

    def forward(self, h_i, context_indices, return_attn_wts=False):
        if h_i.is_cuda and not (torch.is_grad_enabled() and (h_i.requires_grad or any(p.requires_grad for p in self.parameters()))):
            # inference: one projection GEMM + one gather launch for all heads
            key = (ops.param_generation,) + tuple((p.data_ptr(), p._version) for p in self.parameters())
            if getattr(self, "_native", None) is None or self._native[0] != key:
                caches = [NativeForward._gat_cache(h) for h in self.heads]
                Hd = caches[0]["Hd"]
                ext = torch.cat([g["ext"][:Hd] for g in caches] + [g["ext"][Hd:Hd + 2] for g in caches], 0)
                pad = (-ext.shape[0]) % 4
                if pad:
                    ext = torch.cat((ext, torch.zeros((pad, ext.shape[1]), device=ext.device)), 0)
                self._native = (key, dict(ext=ext.contiguous(), Hd=Hd, b=[g["b"] for g in caches], alpha=caches[0]["alpha"]))
            g = self._native[1]
            h = h_i.float() if h_i.stride(-1) == 1 else h_i.float().contiguous()
            out = torch.empty((h.shape[0], g["Hd"] * len(self.heads)), dtype=torch.float32, device=h.device)
            attn = ops.gat_multihead_fwd(ops.linear_fwd(h, g["ext"]), g["Hd"], g["b"], g["alpha"], context_indices, out,
                                         want_attn=return_attn_wts)
            return (out, attn.permute(1, 0, 2)) if return_attn_wts else out
        outs = [h(h_i, context_indices, return_attn_wts) for h in self.heads]
        if return_attn_wts:
            return torch.cat([o[0] for o in outs], 1), torch.stack([o[1] for o in outs], 1)
        return torch.cat(outs, 1)


# ----------------------------------------------------------------------------- CoVA
class CoVA(nn.Module):
    def __init__(self, roi_output_size, img_H, n_classes, use_context=True, hidden_dim=384, bbox_hidden_dim=32,
                 n_additional_feat=0, drop_prob=0.2, class_names=None, *, backbone="resnet18", n_heads=1,
                 roi_mode="pool", engine=None, precision=None, pretrained=True):
        """Same positional arguments as `/root/reference/models.py:10-21` (callers pass all nine positionally:
        `main.py:122-132`).  The feature-map shape is computed analytically (no probe forward on
        uninitialised memory as in `models.py:53-54`)."""
        super().__init__()
        self.n_classes = n_classes
        self.use_context = use_context
        self.hidden_dim = hidden_dim
        self.bbox_hidden_dim = bbox_hidden_dim
        self.n_additional_feat = n_additional_feat
        self.class_names = np.arange(self.n_classes).astype(str) if class_names is None else class_names
        self.img_H = img_H
        self.roi_output_size = tuple(roi_output_size)
        self.backbone = backbone
        self.roi_mode = roi_mode
        self.engine = engine or os.environ.get("COVA_B200_ENGINE", "tcgen05")
        self.precision = precision or os.environ.get("COVA_B200_PRECISION", "fp32")
        if self.engine not in ("tcgen05", "simt") or self.precision not in ("fp32", "fp32x", "bf16", "fp16") or roi_mode not in ("pool", "align"):
            raise ValueError("engine in {tcgen05, simt}, precision in {fp32, fp32x, bf16, fp16}, roi_mode in {pool, align}")

        ##### REPRESENTATION NETWORK (RN) #####
        self.convnet, C = _truncated_resnet(backbone)
        if pretrained:
            _try_load_pretrained(self.convnet, backbone)
        Hc = (img_H + 6 - 7) // 2 + 1
        self.fm_H = (Hc + 2 - 3) // 2 + 1
        self.spatial_scale = self.fm_H / img_H                       # models.py:56
        self.n_visual_feat = C * self.roi_output_size[0] * self.roi_output_size[1]
        self.n_feat = self.n_visual_feat + self.bbox_hidden_dim + self.n_additional_feat

        if self.bbox_hidden_dim > 0:
            self.bbox_feat_encoder = nn.Sequential(nn.Linear(5, self.bbox_hidden_dim),
                                                   nn.BatchNorm1d(self.bbox_hidden_dim), nn.ReLU())
        if self.n_additional_feat > 0:
            self.bn_additional_feat = nn.BatchNorm1d(self.n_additional_feat)
        else:
            self.bn_additional_feat = lambda x: x

        ##### GRAPH ATTENTION LAYER (GAT) #####
        if self.use_context:
            self.gat = (GraphAttentionLayer(self.n_feat, self.hidden_dim) if n_heads == 1
                        else MultiHeadGAT(self.n_feat, self.hidden_dim, n_heads))

        ##### FC LAYERS #####
        self.n_total_feat = self.n_feat + self.hidden_dim
        self.decoder = nn.Sequential(nn.Dropout(drop_prob), nn.Linear(self.n_total_feat, self.n_total_feat),
                                     nn.BatchNorm1d(self.n_total_feat), nn.ReLU(), nn.Dropout(drop_prob),
                                     nn.Linear(self.n_total_feat, self.n_classes))
        self._native = NativeForward(self)
        print("Model Parameters:", count_parameters(self))

    def invalidate_native_cache(self):
        """Call after modifying parameters / buffers through `.data` or raw pointers (writes `Tensor._version` cannot
        see): the native inference path rebuilds its packed / folded weight copies on the next forward."""
        self._native.invalidate()
        for h in (self.gat_heads() if self.use_context else []):
            h._native = None
        if self.use_context and hasattr(self.gat, "heads"):
            self.gat._native = None

    def gat_heads(self):
        return list(self.gat.heads) if isinstance(self.gat, MultiHeadGAT) else [self.gat]

    # ------------------------------------------------------------------ dispatch
    def _use_native(self, *tensors):
        for t in tensors:
            if not t.is_cuda:
                raise RuntimeError("cova_b200: CoVA needs CUDA tensors - the hot path has no CPU implementation")
        return not self.training and not torch.is_grad_enabled()

    def forward(self, images, bboxes, additional_feats, context_indices):
        """images [B,3,img_H,img_H] f32; bboxes [N,5] = [batch_idx,x1,y1,x2,y2]; additional_feats [N,n_add];
        context_indices int64 [N,n_context] (-1 padded) -> logits [N,n_classes] (`models.py:94-122`)."""
        if self._use_native(images, bboxes):
            return self._native.forward(self._img(images), bboxes.float(), additional_feats, context_indices)
        return self._forward_composite(images, bboxes, additional_feats, context_indices)

    @staticmethod
    def _img(images):
        """fp32 in [0,1] (the reference contract) or raw uint8 pixels: the native stem applies the 1/255 itself
        (a quarter of the host->device bytes - SURVEY.md 8(f) N1; `ops.set_knob("stem_u8_exact", 1)` = the per-pixel
        v/255 path that is bit-identical to `ToTensor`)."""
        return images if images.dtype == torch.uint8 else images.float()

    def _forward_composite(self, images, bboxes, additional_feats, context_indices):
        # fp32 like the reference by default; COVA_B200_TRAIN_TF32=1 lets cuDNN / cuBLAS use TF32 tensor cores on this
        # (library) autograd path: ~1e-3 relative on the gradients instead of 1e-6, several times faster
        tf32 = os.environ.get("COVA_B200_TRAIN_TF32", "0") == "1"
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        try:
            with torch.backends.cudnn.flags(enabled=True, allow_tf32=tf32):
                return self._forward_composite_impl(images, bboxes, additional_feats, context_indices)
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev

    def _forward_composite_impl(self, images, bboxes, additional_feats, context_indices):
        visual_feats = self._get_visual_features(images, bboxes)
        bbox_feats = self._get_bbox_features(bboxes)
        additional_feats = self.bn_additional_feat(additional_feats)
        own_features = torch.cat((visual_feats, bbox_feats, additional_feats), dim=1)
        if self.use_context:
            context_representation = self.gat(own_features, context_indices)
        else:
            context_representation = own_features[:, :0]
        return _run_sequential(self.decoder, torch.cat((own_features, context_representation), dim=1))

    def _get_visual_features(self, images, bboxes):
        """`models.py:124-127` -> [N, C*P*P]."""
        if self._use_native(images, bboxes):
            return self._native.visual_features(self._img(images), bboxes.float())
        if images.dtype == torch.uint8:
            images = images.float().div(255)
        if self.training and os.environ.get("COVA_B200_TRAIN_BACKBONE", "native") != "torch":
            from .train_backbone import feature_map_train          # NHWC end to end, native BatchNorm / maxpool
            fm = feature_map_train(self.convnet, images, self.precision)
        else:
            fm = self.convnet(images).permute(0, 2, 3, 1)            # NCHW -> NHWC view for the native RoI kernel
        if self.roi_mode == "pool":
            return _RoIPoolFn.apply(fm, bboxes.float(), self.roi_output_size, self.spatial_scale)
        return _RoIAlignFn.apply(fm, bboxes.float(), self.roi_output_size, self.spatial_scale)

    def _get_bbox_features(self, bboxes):
        """`models.py:129-148`: [x,y,w,h,asp_ratio] -> bbox_hidden_dim features (or [N,0])."""
        if self.bbox_hidden_dim <= 0:
            return bboxes[:, :0]
        if self._use_native(bboxes):
            return self._native.bbox_features(bboxes.float())
        f = bboxes[:, 1:].clone()
        f[:, 2:] -= f[:, :2]
        f = torch.cat((f, (f[:, 2] / f[:, 3]).view(-1, 1)), dim=1)
        return _run_sequential(self.bbox_feat_encoder, f)
