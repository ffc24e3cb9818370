"""Native (libcova_b200.so) inference forward of the CoVA hot path.

Owns the derived weight caches the kernels need - folded eval-mode BatchNorm (scale/shift), repacked /
split-bf16 convolution filters, the extended GAT projection matrix - and rebuilds them whenever a parameter
or buffer of the model changed (``optimizer.step()``, ``load_state_dict`` bump ``Tensor._version``).
The fp32 ``nn.Parameter``s in PyTorch layout stay the single source of truth.

Stage order = ``CoVA.forward`` (`/root/reference/models.py:94-122`):
  stem -> BasicBlock x2 (or Bottleneck x3) -> RoIPool/RoIAlign + positional encoder + additional feats written
  straight into one [T, n_feat + hidden] row buffer (the two `torch.cat`s of `:110` / `:119` never run)
  -> GAT projection GEMM -> fused gather/softmax/weighted-sum -> decoder GEMM + BN + ReLU -> logits GEMM.
"""
import torch

from . import ops
from .ops import BF16, BF16X2, ENGINE_SIMT, ENGINE_TCGEN05, F16, F16X2, F32


def fold_bn(bn):
    """Eval-mode BatchNorm as y = x*scale + shift (eps from the module)."""
    scale = bn.weight.detach() / torch.sqrt(bn.running_var.detach() + bn.eps)
    shift = bn.bias.detach() - bn.running_mean.detach() * scale
    return scale.float().contiguous(), shift.float().contiguous()


class NativeForward:
    def __init__(self, model):
        self.m = model
        self._key = None
        self.c = {}

    # ------------------------------------------------------------------ caches
    def _state_key(self):
        m = self.m
        return (m.engine, m.precision, ops.param_generation) + tuple((t.data_ptr(), t._version) for t in
                                               list(m.parameters()) + list(m.buffers()))

    def invalidate(self):
        """Drop the derived weight caches (folded BN, packed filters, GAT matrix).  Needed after writes that bypass
        `Tensor._version` (`p.data.copy_`, raw-pointer updates by foreign code); the native optimizer and the native
        BatchNorm already signal theirs through `ops.param_generation`."""
        self._key = None

    def prepare(self):
        key = self._state_key()
        if key == self._key:
            return
        m, c = self.m, {}
        r50 = m.backbone == "resnet50"
        tc = m.engine == "tcgen05"
        split = m.precision in ("fp32", "fp32x") or r50   # the ResNet-50 tensor-core path is split-bf16 only
        half = tc and m.precision == "fp16" and not r50
        split16 = tc and m.precision == "fp32x" and not r50          # split-fp16 ("fp16x3"): ResNet-18 stack
        c["tc"] = tc
        c["act_dtype"] = ((F16X2 if split16 else BF16X2) if split else (F16 if half else BF16)) if tc else F32
        cn = m.convnet
        c["stem_w"] = cn[0].weight.detach().float().contiguous()
        if tc:
            c["stem_w"] = (ops.pack_stem_weight_f16(c["stem_w"]) if half else
                           ops.pack_stem_weight_f16x2(c["stem_w"]) if split16 else ops.pack_stem_weight(c["stem_w"]))
        c["stem_bn"] = fold_bn(cn[1])
        blocks = []
        for blk in cn[4]:
            d = {}
            if m.backbone == "resnet18":
                for i in (1, 2):
                    conv, bn = getattr(blk, f"conv{i}"), getattr(blk, f"bn{i}")
                    if half:
                        d[f"w{i}"] = (ops.pack_conv_weight_f16(conv.weight.detach().float()), None)
                    elif split16:
                        d[f"w{i}"] = ops.pack_conv_weight_f16x2(conv.weight.detach().float())
                    else:
                        s, hi, lo = ops.pack_conv_weight(conv.weight.detach().float(), simt=not tc, tc=tc, split=split)
                        d[f"w{i}"] = (hi, lo) if tc else (s, None)
                    d[f"bn{i}"] = fold_bn(bn)
            else:  # Bottleneck: the 1x1 convs are row-major GEMMs on the NHWC pixel rows
                d["w1"] = blk.conv1.weight.detach().float().flatten(1).contiguous()
                s2, hi2, lo2 = ops.pack_conv_weight(blk.conv2.weight.detach().float(), simt=not tc, tc=tc, split=True)
                d["w2"] = (hi2, lo2) if tc else (s2, None)
                d["w3"] = blk.conv3.weight.detach().float().flatten(1).contiguous()
                if tc:
                    d["w1_p"], d["w3_p"] = ops.pack_linear_weight(d["w1"]), ops.pack_linear_weight(d["w3"])
                    if blk.downsample is not None:
                        d["wd_p"] = ops.pack_linear_weight(blk.downsample[0].weight.detach().float().flatten(1).contiguous())
                for i in (1, 2, 3):
                    d[f"bn{i}"] = fold_bn(getattr(blk, f"bn{i}"))
                if blk.downsample is not None:
                    d["wd"] = blk.downsample[0].weight.detach().float().flatten(1).contiguous()
                    d["bnd"] = fold_bn(blk.downsample[1])
            blocks.append(d)
        c["blocks"] = blocks
        if m.bbox_hidden_dim > 0:
            enc = m.bbox_feat_encoder
            c["bbox"] = (enc[0].weight.detach().float().contiguous(), enc[0].bias.detach().float().contiguous(),
                         *fold_bn(enc[1]))
        if m.n_additional_feat > 0:
            c["add_bn"] = fold_bn(m.bn_additional_feat)
        if m.use_context:
            c["gat"] = [self._gat_cache(h) for h in m.gat_heads()]
            if len(c["gat"]) > 1:   # all heads: ONE projection GEMM [W_j(0);..;W_j(H-1); s_0;t_0;..] and ONE gather launch
                Hd = c["gat"][0]["Hd"]
                rows = [g["ext"][:Hd] for g in c["gat"]] + [g["ext"][Hd:Hd + 2] for g in c["gat"]]
                ext = torch.cat(rows, 0)
                pad = (-ext.shape[0]) % 4
                if pad:
                    ext = torch.cat((ext, torch.zeros((pad, ext.shape[1]), device=ext.device)), 0)
                c["gat_mh"] = dict(ext=ext.contiguous(), Hd=Hd, b=[g["b"] for g in c["gat"]], alpha=c["gat"][0]["alpha"])
        dec = m.decoder
        c["dec1"] = (dec[1].weight.detach().float().contiguous(), dec[1].bias.detach().float().contiguous(),
                     *fold_bn(dec[2]))
        c["dec2"] = (dec[5].weight.detach().float().contiguous(), dec[5].bias.detach().float().contiguous())
        if m.engine == "tcgen05":   # split-bf16 copies for the tensor-core GEMMs
            c["dec1_p"] = ops.pack_linear_weight(c["dec1"][0])
            c["dec2_p"] = ops.pack_linear_weight(c["dec2"][0])
            for g in c.get("gat", []) + ([c["gat_mh"]] if "gat_mh" in c else []):
                g["ext_p"] = ops.pack_linear_weight(g["ext"])
        self.c, self._key = c, key

    def _linear(self, x, w, w_packed, *args, **kw):
        """Tensor-core GEMM when the engine is tcgen05 and the shape meets its contract, CUDA cores otherwise."""
        if w_packed is not None and ops.linear_tc_ok(x):
            return ops.linear_fwd(x, w_packed, *args, engine=ENGINE_TCGEN05, **kw)
        return ops.linear_fwd(x, w, *args, **kw)

    @staticmethod
    def _gat_cache(layer):
        """[W_j ; a_i^T W_i ; a_j^T W_j ; 0 ; 0]: one GEMM yields whj, s and t."""
        Hd = layer.hidden_dim
        Wi, Wj = layer.W_i.weight.detach().float(), layer.W_j.weight.detach().float()
        a = layer.attention_layer.weight.detach().float()[0]
        ext = torch.zeros((Hd + 4, Wj.shape[1]), dtype=torch.float32, device=Wj.device)
        ext[:Hd] = Wj
        ext[Hd] = a[:Hd] @ Wi
        ext[Hd + 1] = a[Hd:] @ Wj
        return dict(ext=ext.contiguous(), Hd=Hd, b=float(layer.attention_layer.bias.detach().float().item()),
                    alpha=float(layer.leakyrelu.negative_slope))

    # ------------------------------------------------------------------ stages
    def _eng(self):
        return ENGINE_TCGEN05 if self.c["tc"] else ENGINE_SIMT

    def feature_map(self, images):
        """A2: [B,3,H,W] fp32 NCHW -> fp32 NHWC feature map [B,H/4,W/4,C]."""
        self.prepare()
        c, m = self.c, self.m
        eng = self._eng()
        x = ops.stem_fwd(images, c["stem_w"], *c["stem_bn"], out_dtype=c["act_dtype"], engine=eng)
        if m.backbone == "resnet18":
            nb = len(c["blocks"])
            for bi, d in enumerate(c["blocks"]):
                y = ops.conv3x3_bn_act_fwd(x, d["w1"][0], d["w1"][1], *d["bn1"], relu=True, engine=eng)
                last = bi == nb - 1
                x = ops.conv3x3_bn_act_fwd(y, d["w2"][0], d["w2"][1], *d["bn2"], res=x, relu=True,
                                           out_dtype=F32 if last else None, engine=eng)
            return x.p0
        # resnet50 Bottlenecks
        nb = len(c["blocks"])
        if c["tc"]:   # split-bf16 planes end to end: 1x1 convs = persistent tcgen05 GEMMs over the pixel rows
            for bi, d in enumerate(c["blocks"]):
                o = ops.conv1x1_bn_act_fwd(x, d["w1_p"], *d["bn1"], relu=True)
                o = ops.conv3x3_bn_act_fwd(o, d["w2"][0], d["w2"][1], *d["bn2"], relu=True, engine=ENGINE_TCGEN05)
                idt = ops.conv1x1_bn_act_fwd(x, d["wd_p"], *d["bnd"], relu=False) if "wd_p" in d else x
                x = ops.conv1x1_bn_act_fwd(o, d["w3_p"], *d["bn3"], res=idt, relu=True,
                                           out_dtype=F32 if bi == nb - 1 else BF16X2)
            return x.p0
        # CUDA-core fp32 engine: 1x1 conv = GEMM over the [B*H*W, C] pixel rows
        B, H, W, _ = x.shape
        cur = x.p0.view(B * H * W, 64)
        for d in c["blocks"]:
            o = ops.linear_fwd(cur, d["w1"], None, *d["bn1"], relu=True)
            p = ops.Planes.__new__(ops.Planes)
            p.dtype, p.shape, p.p0, p.p1 = F32, (B, H, W, 64), o.view(B, H, W, 64), None
            o = ops.conv3x3_bn_act_fwd(p, d["w2"][0], None, *d["bn2"], relu=True, engine=ENGINE_SIMT).p0.view(B * H * W, 64)
            idt = ops.linear_fwd(cur, d["wd"], None, *d["bnd"]) if "wd" in d else cur
            cur = ops.linear_fwd(o, d["w3"], None, *d["bn3"], res=idt, relu=True)
        return cur.view(B, H, W, 256)

    def own_into(self, fm, bboxes, additional_feats, comb):
        """A4 + A5 + A5b: fill comb[:, :n_feat] (visual | bbox | additional)."""
        c, m = self.c, self.m
        scale = m.spatial_scale                                      # fixed at construction like models.py:56
        ops.roi_fwd(fm, bboxes, m.roi_output_size, scale, comb, mode=m.roi_mode, sampling_ratio=2)
        col = m.n_visual_feat
        if m.bbox_hidden_dim > 0:
            ops.bbox_enc_fwd(bboxes, *c["bbox"], comb[:, col:])
            col += m.bbox_hidden_dim
        if m.n_additional_feat > 0:
            ops.affine_cols_fwd(additional_feats.float(), *c["add_bn"], comb[:, col:])

    def gat_into(self, own, context_indices, out, want_attn=False, heads=None):
        """A6: own [T,n_feat] (strided view ok) -> out [T, hidden] ; returns attn of the (single) head or None."""
        self.prepare()
        if heads is None and "gat_mh" in self.c:
            g = self.c["gat_mh"]
            ext = self._linear(own, g["ext"], g.get("ext_p"))
            return ops.gat_multihead_fwd(ext, g["Hd"], g["b"], g["alpha"], context_indices, out, want_attn=want_attn)
        attn, col = None, 0
        for g in (self.c["gat"] if heads is None else heads):
            ext = self._linear(own, g["ext"], g.get("ext_p"))
            Hd = g["Hd"]
            attn = ops.gat_fwd(ext[:, :Hd], ext[:, Hd], ext[:, Hd + 1], g["b"], g["alpha"], context_indices,
                               out[:, col:col + Hd], want_attn=want_attn)
            col += Hd
        return attn

    def visual_features(self, images, bboxes):
        """`CoVA._get_visual_features` (`models.py:124-127`)."""
        fm = self.feature_map(images)
        out = torch.empty((bboxes.shape[0], self.m.n_visual_feat), dtype=torch.float32, device=fm.device)
        ops.roi_fwd(fm, bboxes, self.m.roi_output_size, self.m.spatial_scale, out, mode=self.m.roi_mode)
        return out

    def bbox_features(self, bboxes):
        """`CoVA._get_bbox_features` (`models.py:129-148`)."""
        self.prepare()
        out = torch.empty((bboxes.shape[0], self.m.bbox_hidden_dim), dtype=torch.float32, device=bboxes.device)
        if self.m.bbox_hidden_dim > 0:
            ops.bbox_enc_fwd(bboxes, *self.c["bbox"], out)
        return out

    def forward(self, images, bboxes, additional_feats, context_indices, return_intermediates=False):
        return self.forward_from_fm(self.feature_map(images), bboxes, additional_feats, context_indices, return_intermediates)

    def forward_from_fm(self, fm, bboxes, additional_feats, context_indices, return_intermediates=False):
        """A3-A8 on a given NHWC fp32 feature map (RoI pooling, positional encoder, GAT, decoder)."""
        m = self.m
        self.prepare()
        T = bboxes.shape[0]
        comb = torch.empty((T, m.n_total_feat), dtype=torch.float32, device=fm.device)
        self.own_into(fm, bboxes, additional_feats, comb)
        if m.use_context:
            self.gat_into(comb[:, :m.n_feat], context_indices, comb[:, m.n_feat:])
        c = self.c
        h1 = self._linear(comb, c["dec1"][0], c.get("dec1_p"), c["dec1"][1], c["dec1"][2], c["dec1"][3], relu=True)
        logits = self._linear(h1, c["dec2"][0], c.get("dec2_p"), c["dec2"][1])
        if return_intermediates:
            return dict(fm=fm, own=comb[:, :m.n_feat], ctx=comb[:, m.n_feat:], logits=logits)
        return logits
