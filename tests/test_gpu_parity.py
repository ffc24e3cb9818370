"""GPU parity tests: every kernel, called through the C ABI (``cova_b200.ops`` -> ctypes -> libcova_b200.so),
against the oracle (``oracle/cova_oracle.py``) on the same seeded inputs and against the fixtures frozen from
the live reference (``tests/golden``).  Tolerances (max|a-b| / max|b| unless stated):
  * RoIPool ....................................... bit-exact
  * exact-fp32 engine (simt), every stage ......... 1e-5
  * tcgen05 fp32-parity mode (split-bf16 x3) ...... 1e-4 on the feature map, 1e-4 on logits
  * tcgen05 fp32x mode (split-fp16 x3) ............ 1e-5 on the feature map, 2e-5 on logits (the linear layers stay
    split-bf16; the backbone alone is ~1e-6)
  * tcgen05 fp16 mode (one fp16 product) .......... 2e-3 on the feature map, 1e-3 on logits (= BASELINE.json's bar)
  * tcgen05 bf16 mode ............................. 2e-2 on the feature map, 1e-2 on logits.  Measured 3-4e-3 on
    logits: single-pass bf16 does NOT meet BASELINE.json's 1e-3 bar, which is why the fp32-parity mode is the
    default and the one bench.py measures.
"""
import warnings

import numpy as np
import pytest
import torch

import cova_b200.synth as synth
from conftest import load_golden, rel_err
from oracle import cova_oracle as O

pytestmark = pytest.mark.gpu
warnings.filterwarnings("ignore")
DEV = "cuda:0"


def ops():
    from cova_b200 import ops as _ops
    return _ops


def t2n(t):
    return t.detach().float().cpu().numpy()


def torch_conv(x, w, stride, pad):
    """Oracle conv for the larger cases: same arithmetic as cova_oracle.conv2d_nchw, CPU fp32, via ATen."""
    return torch.nn.functional.conv2d(torch.from_numpy(np.ascontiguousarray(x)), torch.from_numpy(np.ascontiguousarray(w)),
                                      None, stride, pad).numpy()


def make_model(engine, precision="fp32", img=1280, **kw):
    from cova_b200.models import CoVA
    cfg = dict(roi=(3, 3), use_context=True, hidden=384, bbhd=32, n_add=0)
    cfg.update({k: kw.pop(k) for k in list(kw) if k in cfg})
    m = CoVA(cfg["roi"], img, 4, cfg["use_context"], cfg["hidden"], cfg["bbhd"], cfg["n_add"], 0.2, None,
             pretrained=False, engine=engine, precision=precision, **kw)
    sd = synth.make_state_dict(123, backbone=kw.get("backbone", "resnet18"), roi_output_size=cfg["roi"],
                               hidden_dim=cfg["hidden"], bbox_hidden_dim=cfg["bbhd"], n_additional_feat=cfg["n_add"],
                               use_context=cfg["use_context"], n_heads=kw.get("n_heads", 1))
    m.load_state_dict(sd, strict=True)
    return m.to(DEV).eval(), {k: v.numpy() for k, v in sd.items()}


def to_dev(inp):
    return [t.to(DEV) for t in inp]


# ----------------------------------------------------------------------------- single kernels
@pytest.mark.parametrize("out_dtype", ["f32", "bf16x2"])
def test_stem_kernel(out_dtype):
    o = ops()
    g = torch.Generator().manual_seed(1)
    B, H = 2, 104                       # 104 -> conv 52 -> pooled 26: partial 8x8 tiles on both edges
    img = torch.rand(B, 3, H, H, generator=g)
    w = torch.randn(64, 3, 7, 7, generator=g) * 0.1
    scale, shift = 0.5 + torch.rand(64, generator=g), 0.1 * torch.randn(64, generator=g)
    ref = O.conv2d_nchw(img.numpy(), w.numpy(), 2, 3) * scale.numpy().reshape(1, -1, 1, 1) + shift.numpy().reshape(1, -1, 1, 1)
    ref = O.maxpool3x3s2p1(np.maximum(ref, 0)).transpose(0, 2, 3, 1)
    out = o.stem_fwd(img.to(DEV), w.to(DEV), scale.to(DEV), shift.to(DEV),
                     out_dtype=o.F32 if out_dtype == "f32" else o.BF16X2)
    assert out.shape == (B, 26, 26, 64)
    assert rel_err(t2n(out.float()), ref) < (1e-5 if out_dtype == "f32" else 3e-5)


@pytest.mark.parametrize("out_dtype", ["f32", "bf16x2", "bf16", "f16", "f16x2"])
@pytest.mark.parametrize("B,H", [(2, 104), (1, 260), (3, 64), (1, 102)])
def test_stem_tcgen05_kernel(out_dtype, B, H):
    """Tensor-core stem (ring of raw image rows + sliding-window descriptors + fused pooling) vs the oracle.
    Sizes: partial strips/bands (104), two strips (260 -> 130 conv columns), more pages than bands (3 x 64), a width
    that is not a multiple of 4 (102: the scalar converter path instead of the float4 one)."""
    o = ops()
    g = torch.Generator().manual_seed(1)
    img = torch.rand(B, 3, H, H, generator=g)
    w = torch.randn(64, 3, 7, 7, generator=g) * 0.1
    scale, shift = 0.5 + torch.rand(64, generator=g), 0.1 * torch.randn(64, generator=g)
    if out_dtype == "bf16":   # single-product mode: the oracle sees the bf16-rounded image and filter
        ref = O.conv2d_nchw(img.bfloat16().float().numpy(), w.bfloat16().float().numpy(), 2, 3)
    elif out_dtype == "f16":  # fp16 mode: fp16-rounded image and filter
        ref = O.conv2d_nchw(img.half().float().numpy(), w.half().float().numpy(), 2, 3)
    else:
        ref = O.conv2d_nchw(img.numpy(), w.numpy(), 2, 3)
    ref = ref * scale.numpy().reshape(1, -1, 1, 1) + shift.numpy().reshape(1, -1, 1, 1)
    ref = O.maxpool3x3s2p1(np.maximum(ref, 0)).transpose(0, 2, 3, 1)
    wp = {"f16": o.pack_stem_weight_f16, "f16x2": o.pack_stem_weight_f16x2}.get(out_dtype, o.pack_stem_weight)(w.to(DEV))
    out = o.stem_fwd(img.to(DEV), wp, scale.to(DEV), shift.to(DEV),
                     out_dtype={"f32": o.F32, "bf16x2": o.BF16X2, "bf16": o.BF16, "f16": o.F16, "f16x2": o.F16X2}[out_dtype],
                     engine=o.ENGINE_TCGEN05)
    torch.cuda.synchronize()
    assert out.shape == ref.shape
    # split-fp16 ("fp16x3"): 22 significand bits -> an order of magnitude tighter than split-bf16
    assert rel_err(t2n(out.float()), ref) < {"f32": 3e-5, "bf16x2": 5e-5, "bf16": 6e-3, "f16": 6e-4, "f16x2": 3e-6}[out_dtype]


@pytest.mark.parametrize("W", [104, 102])
def test_stem_tcgen05_uint8_images(W):
    """uint8 pixels: the aligned 4-pixel-word + LUT converter (W % 4 == 0) and the byte fallback (W % 4 != 0)."""
    o = ops()
    g = torch.Generator().manual_seed(3)
    u8 = torch.randint(0, 256, (2, 3, W, W), dtype=torch.uint8, generator=g)
    w = torch.randn(64, 3, 7, 7, generator=g) * 0.1
    scale, shift = 0.5 + torch.rand(64, generator=g), 0.1 * torch.randn(64, generator=g)
    ref = O.conv2d_nchw(u8.float().div(255).numpy(), w.numpy(), 2, 3)
    ref = ref * scale.numpy().reshape(1, -1, 1, 1) + shift.numpy().reshape(1, -1, 1, 1)
    ref = O.maxpool3x3s2p1(np.maximum(ref, 0)).transpose(0, 2, 3, 1)
    out = o.stem_fwd(u8.to(DEV), o.pack_stem_weight(w.to(DEV)), scale.to(DEV), shift.to(DEV), out_dtype=o.F32,
                     engine=o.ENGINE_TCGEN05)
    assert rel_err(t2n(out.p0), ref) < 3e-5


@pytest.mark.parametrize("u8", [False, True])
@pytest.mark.parametrize("B,H,W", [(1, 8, 1600), (1, 18, 1028), (2, 10, 772), (1, 1280, 36), (5, 38, 516), (1, 14, 2052)])
def test_stem_tcgen05_band_and_strip_geometries(B, H, W, u8):
    """Rectangular pages against the exact-fp32 CUDA-core stem: one or two emitted pooled rows per band with up to 9 strips
    (the strip-edge ring and the exchange buffers are reused with almost no barrier in between), an odd number of conv
    rows (H = 10, 14, 18, 38: the last conv row is an even one that still completes a pooled row), a 36-pixel-wide page
    with 148 one-row bands, and consecutive tiles whose two MMA issuers start on either accumulator."""
    o = ops()
    g = torch.Generator().manual_seed(B * 1000 + H + W)
    if u8:
        img = torch.randint(0, 256, (B, 3, H, W), dtype=torch.uint8, generator=g).to(DEV)
    else:
        img = torch.rand(B, 3, H, W, generator=g).to(DEV)
    w = (torch.randn(64, 3, 7, 7, generator=g) * 0.1).to(DEV)
    scale, shift = (0.5 + torch.rand(64, generator=g)).to(DEV), (0.1 * torch.randn(64, generator=g)).to(DEV)
    a = o.stem_fwd(img, w, scale, shift, out_dtype=o.F32, engine=o.ENGINE_SIMT)
    wp = o.pack_stem_weight(w)
    for _ in range(3):      # a race would not show on every launch
        b = o.stem_fwd(img, wp, scale, shift, out_dtype=o.F32, engine=o.ENGINE_TCGEN05)
        torch.cuda.synchronize()
        assert b.p0.shape == a.p0.shape
        assert rel_err(t2n(b.p0), t2n(a.p0)) < 3e-5


def test_stem_tcgen05_full_size_vs_simt():
    """1280x1280 (5 strips x 9 bands per page): the tensor-core stem against the exact-fp32 CUDA-core stem."""
    o = ops()
    g = torch.Generator().manual_seed(2)
    img = torch.rand(2, 3, 1280, 1280, generator=g).to(DEV)
    w = (torch.randn(64, 3, 7, 7, generator=g) * 0.1).to(DEV)
    scale, shift = (0.5 + torch.rand(64, generator=g)).to(DEV), (0.1 * torch.randn(64, generator=g)).to(DEV)
    a = o.stem_fwd(img, w, scale, shift, out_dtype=o.F32, engine=o.ENGINE_SIMT)
    b = o.stem_fwd(img, o.pack_stem_weight(w), scale, shift, out_dtype=o.F32, engine=o.ENGINE_TCGEN05)
    torch.cuda.synchronize()
    assert rel_err(t2n(b.p0), t2n(a.p0)) < 3e-5


def _conv_case(seed, B, H, W, with_res):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, H, W, 64, generator=g)
    w = torch.randn(64, 64, 3, 3, generator=g) * (2.0 / 576) ** 0.5
    scale, shift = 0.5 + torch.rand(64, generator=g), 0.1 * torch.randn(64, generator=g)
    res = torch.randn(B, H, W, 64, generator=g) if with_res else None
    ref = torch_conv(x.permute(0, 3, 1, 2).numpy(), w.numpy(), 1, 1).transpose(0, 2, 3, 1)
    ref = ref * scale.numpy() + shift.numpy()
    if with_res:
        ref = ref + res.numpy()
    return x, w, scale, shift, res, np.maximum(ref, 0)


@pytest.mark.parametrize("with_res", [False, True])
def test_conv3x3_simt_kernel(with_res):
    o = ops()
    x, w, scale, shift, res, ref = _conv_case(2, 2, 21, 35, with_res)     # ragged: partial 8x16 tiles
    ws, _, _ = o.pack_conv_weight(w.to(DEV), simt=True, tc=False)
    xp = o.Planes(o.F32, x.shape, DEV); xp.p0.copy_(x)
    rp = None
    if with_res:
        rp = o.Planes(o.F32, x.shape, DEV); rp.p0.copy_(res)
    y = o.conv3x3_bn_act_fwd(xp, ws, None, scale.to(DEV), shift.to(DEV), res=rp, relu=True, engine=o.ENGINE_SIMT)
    assert rel_err(t2n(y.p0), ref) < 1e-5


def _split(t):
    hi = t.bfloat16()
    lo = (t - hi.float()).bfloat16()
    return hi, lo


@pytest.mark.parametrize("mode,out", [("split", "f32"), ("split", "bf16x2"), ("bf16", "bf16"), ("bf16", "f32"),
                                      ("f16", "f16"), ("f16", "f32"), ("f16x2", "f16x2"), ("f16x2", "f32")])
@pytest.mark.parametrize("shape", [(1, 16, 8), (2, 37, 45), (1, 80, 64)])
def test_conv3x3_tcgen05_kernel(mode, out, shape):
    """tcgen05 implicit-GEMM conv vs the oracle conv; shapes cover 1 tile, ragged tiles, multi-tile."""
    o = ops()
    B, H, W = shape
    x, w, scale, shift, res, ref = _conv_case(3, B, H, W, True)
    split, half, split16 = mode == "split", mode == "f16", mode == "f16x2"
    if half:
        whi, wlo = o.pack_conv_weight_f16(w.to(DEV)), None
    elif split16:
        whi, wlo = o.pack_conv_weight_f16x2(w.to(DEV))
    else:
        _, whi, wlo = o.pack_conv_weight(w.to(DEV), simt=False, tc=True, split=split)
    dt = o.BF16X2 if split else (o.F16 if half else (o.F16X2 if split16 else o.BF16))
    xp, rp = o.Planes(dt, x.shape, DEV), o.Planes(dt, x.shape, DEV)
    for p, t in ((xp, x), (rp, res)):
        hi, lo = _split(t)
        if split16:
            hi = t.half()
            lo = (t - hi.float()).half()
        p.p0.copy_(t.half() if half else hi)
        if split or split16:
            p.p1.copy_(lo)
    split = split or split16
    if not split:   # oracle sees what the kernel sees: bf16- (fp16-) rounded input, residual and weights
        q = (lambda t: t.half().float()) if half else (lambda t: t.bfloat16().float())
        xq, rq, wq = q(x), q(res), q(w)
        ref = torch_conv(xq.permute(0, 3, 1, 2).numpy(), wq.numpy(), 1, 1).transpose(0, 2, 3, 1)
        ref = np.maximum(ref * scale.numpy() + shift.numpy() + rq.numpy(), 0)
    od = {"f32": o.F32, "bf16": o.BF16, "bf16x2": o.BF16X2, "f16": o.F16, "f16x2": o.F16X2}[out]
    y = o.conv3x3_bn_act_fwd(xp, whi, wlo, scale.to(DEV), shift.to(DEV), res=rp, relu=True, out_dtype=od,
                             engine=o.ENGINE_TCGEN05)
    torch.cuda.synchronize()
    tol = {"f32": 3e-5, "bf16x2": 5e-5, "bf16": 6e-3, "f16": 6e-4, "f16x2": 3e-6}[out]   # bf16 / fp16 out: one rounding of the result
    if split16:
        tol = 3e-6
    assert rel_err(t2n(y.float()), ref) < tol


@pytest.mark.parametrize("P", [(1, 1), (3, 3), (7, 7), (2, 5)])
def test_roi_pool_adversarial_bit_exact(P):
    """Fixture from `torchvision.ops.roi_pool` itself: negative coords, sub-pixel boxes, .5 ties, empty bins.
    C=8 in the fixture -> tiled x8 to the kernel's 64-channel blocks."""
    o = ops()
    g = load_golden("g_roi")
    fm = np.tile(g["fm"], (1, 8, 1, 1))
    want = np.tile(g["pool_%dx%d" % P], (1, 8, 1, 1)).reshape(len(g["boxes"]), -1)
    fm_d = torch.from_numpy(fm).permute(0, 2, 3, 1).contiguous().to(DEV)
    out = torch.empty((len(g["boxes"]), 64 * P[0] * P[1]), device=DEV)
    am = o.roi_fwd(fm_d, torch.from_numpy(g["boxes"]).to(DEV), P, 0.25, out, want_argmax=True)
    assert np.array_equal(t2n(out), want)
    out2 = torch.empty_like(out)
    o.roi_fwd(fm_d, torch.from_numpy(g["boxes"]).to(DEV), P, 0.25, out2)
    assert torch.equal(out, out2)                                       # argmax and plain variants agree
    # argmax points at a pixel holding the max (or -1 for empty bins)
    am = am.cpu().numpy().reshape(len(g["boxes"]), 64, -1)
    v = want.reshape(len(g["boxes"]), 64, -1)
    flat = fm.reshape(2, 64, -1)
    bidx = g["boxes"][:, 0].astype(int)
    for n in range(0, len(bidx), 7):
        for c in (0, 13, 63):
            for k in range(v.shape[2]):
                if am[n, c, k] >= 0:
                    assert flat[bidx[n], c, am[n, c, k]] == v[n, c, k]
                else:
                    assert v[n, c, k] == 0


@pytest.mark.parametrize("P", [(3, 3), (7, 7), (2, 5)])
def test_roi_align_adversarial(P):
    o = ops()
    g = load_golden("g_roi")
    fm = np.tile(g["fm"], (1, 8, 1, 1))
    want = np.tile(g["align_%dx%d" % P], (1, 8, 1, 1)).reshape(len(g["boxes"]), -1)
    fm_d = torch.from_numpy(fm).permute(0, 2, 3, 1).contiguous().to(DEV)
    out = torch.empty((len(g["boxes"]), 64 * P[0] * P[1]), device=DEV)
    o.roi_fwd(fm_d, torch.from_numpy(g["boxes"]).to(DEV), P, 0.25, out, mode="align")
    assert np.abs(t2n(out) - want).max() < 3e-5   # fm ~ N(0,1): a few fp32 ulps (FMA contraction differs)


def test_roi_pool_wide_row_buffer_and_c256():
    """Writes into a wider row (the fused concat) and walks 4 channel blocks (ResNet-50 C=256)."""
    o = ops()
    g = torch.Generator().manual_seed(4)
    fm = torch.randn(2, 30, 40, 256, generator=g)
    _, bboxes, _, _ = synth.gen(2, 25, 8, seed=4, img=160)
    bboxes[:, 1:] *= 0.75
    want = O.roi_pool(fm.permute(0, 3, 1, 2).numpy(), bboxes.numpy(), (3, 3), 0.25).reshape(50, -1)
    buf = torch.full((50, 2304 + 40), -7.0, device=DEV)
    o.roi_fwd(fm.to(DEV), bboxes.to(DEV), (3, 3), 0.25, buf)
    assert np.array_equal(t2n(buf[:, :2304]), want) and bool((buf[:, 2304:] == -7).all())


def test_bbox_encoder_and_affine_cols():
    o = ops()
    sd = {k: v for k, v in synth.make_state_dict(123).items()}
    _, bboxes, _, _ = synth.gen(3, 0, 8, seed=3, img=256, counts=[11, 1, 30])
    want = O.bbox_encoder(bboxes.numpy(), {k: v.numpy() for k, v in sd.items()})
    rv, rm = sd["bbox_feat_encoder.1.running_var"], sd["bbox_feat_encoder.1.running_mean"]
    scale = sd["bbox_feat_encoder.1.weight"] / torch.sqrt(rv + 1e-5)
    shift = sd["bbox_feat_encoder.1.bias"] - rm * scale
    out = torch.zeros((42, 40), device=DEV)
    o.bbox_enc_fwd(bboxes.to(DEV), sd["bbox_feat_encoder.0.weight"].to(DEV), sd["bbox_feat_encoder.0.bias"].to(DEV),
                   scale.to(DEV), shift.to(DEV), out[:, 4:36])
    assert rel_err(t2n(out[:, 4:36]), want) < 1e-5 and bool((out[:, :4] == 0).all()) and bool((out[:, 36:] == 0).all())
    x = torch.randn(42, 7)
    o.affine_cols_fwd(x.to(DEV), scale[:7].to(DEV), shift[:7].to(DEV), out[:, 10:])
    assert np.allclose(t2n(out[:, 10:17]), x.numpy() * scale[:7].numpy() + shift[:7].numpy(), rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("M,K,N", [(1, 5, 32), (77, 608, 388), (1440, 992, 992), (130, 992, 4)])
def test_linear_kernel(M, K, N):
    o = ops()
    g = torch.Generator().manual_seed(5)
    x, w = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / K ** 0.5
    b, sc, sh = torch.randn(N, generator=g), 0.5 + torch.rand(N, generator=g), torch.randn(N, generator=g)
    res = torch.randn(M, N, generator=g)
    want = np.maximum((O.linear(x.numpy(), w.numpy(), b.numpy()) * sc.numpy() + sh.numpy()) + res.numpy(), 0)
    y = o.linear_fwd(x.to(DEV), w.to(DEV), b.to(DEV), sc.to(DEV), sh.to(DEV), res=res.to(DEV), relu=True)
    assert rel_err(t2n(y), want) < 1e-5
    y = o.linear_fwd(x.to(DEV), w.to(DEV))
    assert rel_err(t2n(y), O.linear(x.numpy(), w.numpy())) < 1e-5


@pytest.mark.parametrize("M,K,N", [(1, 8, 32), (77, 608, 388), (1440, 992, 992), (130, 992, 4), (300, 2720, 200)])
def test_linear_tcgen05_kernel(M, K, N):
    """Tensor-core GEMM (in-kernel fp32 -> swizzled split-bf16 conversion of X, TMA-streamed split W) vs oracle:
    ragged M/N/K tiles, strided X rows (a column slice of a wider buffer), full epilogue."""
    o = ops()
    g = torch.Generator().manual_seed(5)
    xw = torch.randn(M, K + 8, generator=g)
    x, w = xw[:, :K], torch.randn(N, K, generator=g) / K ** 0.5
    b, sc, sh = torch.randn(N, generator=g), 0.5 + torch.rand(N, generator=g), torch.randn(N, generator=g)
    res = torch.randn(M, N, generator=g)
    want = np.maximum((O.linear(x.numpy(), w.numpy(), b.numpy()) * sc.numpy() + sh.numpy()) + res.numpy(), 0)
    xd = xw.to(DEV)[:, :K]
    wp = o.pack_linear_weight(w.to(DEV))
    y = o.linear_fwd(xd, wp, b.to(DEV), sc.to(DEV), sh.to(DEV), res=res.to(DEV), relu=True, engine=o.ENGINE_TCGEN05)
    torch.cuda.synchronize()
    assert rel_err(t2n(y), want) < 3e-5
    y = o.linear_fwd(xd, wp, engine=o.ENGINE_TCGEN05)
    assert rel_err(t2n(y), O.linear(x.numpy(), w.numpy())) < 3e-5


def test_gat_layer_golden_edge_cases_and_heads():
    """The layer fixture of the live reference: arbitrary (non-window) ids, an all -1 row, partial padding,
    attention weights, and the 2-head composition."""
    from cova_b200.models import GraphAttentionLayer
    g = load_golden("g_gat")
    layer = GraphAttentionLayer(96, 64).to(DEV).eval()
    layer.load_state_dict({"W_i.weight": torch.from_numpy(g["W_i"]), "W_j.weight": torch.from_numpy(g["W_j"]),
                           "attention_layer.weight": torch.from_numpy(g["att_w"]),
                           "attention_layer.bias": torch.from_numpy(g["att_b"])})
    h, ci = torch.from_numpy(g["h"]).to(DEV), torch.from_numpy(g["ci"]).to(DEV)
    with torch.no_grad():
        out, attn = layer(h, ci, return_attn_wts=True)
        out_only = layer(h, ci)
    assert np.abs(t2n(out) - g["out"]).max() < 5e-6 and np.abs(t2n(attn) - g["attn"]).max() < 2e-6
    assert torch.equal(out, out_only)
    assert bool((out[3] == 0).all()) and np.allclose(t2n(attn[3]), 0.1)          # all -1 row: uniform, exact 0 out
    outs = []
    for i in range(2):
        hd = GraphAttentionLayer(96, 32).to(DEV).eval()
        hd.load_state_dict({"W_i.weight": torch.from_numpy(g[f"h{i}_W_i"]), "W_j.weight": torch.from_numpy(g[f"h{i}_W_j"]),
                            "attention_layer.weight": torch.from_numpy(g[f"h{i}_att_w"]),
                            "attention_layer.bias": torch.from_numpy(g[f"h{i}_att_b"])})
        with torch.no_grad():
            outs.append(hd(h, ci))
    assert np.abs(t2n(torch.cat(outs, 1)) - g["out_2head"]).max() < 5e-6


@pytest.mark.parametrize("T,K,Hd", [(5, 1, 4), (300, 48, 192), (1000, 128, 384)])
def test_gat_kernel_shapes_vs_oracle(T, K, Hd):
    """K=1, the stress K=48 and the kernel's K limit; ids far apart force the un-staged (direct L2) gather."""
    o = ops()
    g = torch.Generator().manual_seed(6)
    h = torch.randn(T, 64, generator=g)
    Wi, Wj = torch.randn(Hd, 64, generator=g) / 8, torch.randn(Hd, 64, generator=g) / 8
    aw, ab = torch.randn(1, 2 * Hd, generator=g) / Hd ** 0.5, torch.randn(1, generator=g)
    ci = torch.randint(-1, T, (T, K), generator=g)
    want, want_attn = O.gat(h.numpy(), ci.numpy(), Wi.numpy(), Wj.numpy(), aw.numpy(), ab.numpy(), return_attn_wts=True)
    whj = (h @ Wj.T).contiguous()
    s, t = (h @ Wi.T) @ aw[0, :Hd], whj @ aw[0, Hd:]
    out = torch.empty(T, Hd, device=DEV)
    attn = o.gat_fwd(whj.to(DEV), s.to(DEV), t.to(DEV), float(ab), 0.2, ci.to(DEV), out, want_attn=True)
    assert np.abs(t2n(out) - want).max() < 2e-5 and np.abs(t2n(attn) - want_attn).max() < 5e-6


def test_gat_stress_config5_shape():
    """BASELINE config 5 shape per page group: N=300 boxes/page, K=48, 2 heads of 192 (4 pages here so the
    as-written oracle stays small): window ids, staged smem path with 56-row windows."""
    m, sd = make_model("tcgen05", img=128, n_heads=2)
    own = torch.randn(1200, 608, generator=torch.Generator().manual_seed(3))
    ci = np.concatenate([synth.context_window(300, 24) + np.where(synth.context_window(300, 24) >= 0, p * 300, 0)
                         for p in range(4)], 0)
    heads = [(sd[f"gat.heads.{i}.W_i.weight"], sd[f"gat.heads.{i}.W_j.weight"],
              sd[f"gat.heads.{i}.attention_layer.weight"], sd[f"gat.heads.{i}.attention_layer.bias"]) for i in range(2)]
    want = O.gat_multihead(own.numpy(), ci, heads)
    with torch.no_grad():
        got = m.gat(own.to(DEV), torch.from_numpy(ci).to(DEV))
    assert rel_err(t2n(got), want) < 5e-5


# ----------------------------------------------------------------------------- whole forward vs the live reference
# fp16 mode: the logits tolerance IS BASELINE.json's bar (1e-3); measured ~5e-4 (profiles/r01l_precision_modes.txt)
ENGINES = [("simt", "fp32", 1e-5, 2e-5), ("tcgen05", "fp32", 1e-4, 1e-4), ("tcgen05", "bf16", 2e-2, 1e-2),
           ("tcgen05", "fp16", 2e-3, 1e-3), ("tcgen05", "fp32x", 1e-5, 2e-5)]


@pytest.mark.parametrize("engine,precision,tol_fm,tol_logits", ENGINES)
def test_forward_small_golden(engine, precision, tol_fm, tol_logits):
    g = load_golden("g_small_r18_img128")
    m, _ = make_model(engine, precision, img=128)
    inp = to_dev(synth.gen(2, 12, 8, seed=0, img=128))
    with torch.no_grad():
        r = m._native.forward(*inp, return_intermediates=True)
        logits = m(*inp)
    assert rel_err(t2n(r["fm"]).transpose(0, 3, 1, 2), g["fm"]) < tol_fm
    assert rel_err(t2n(r["own"]), g["own"]) < tol_fm
    assert rel_err(t2n(r["ctx"]), g["ctx"]) < max(tol_fm, 2e-5)
    assert rel_err(t2n(logits), g["logits"]) < tol_logits


@pytest.mark.parametrize("engine,precision,tol_fm,tol_logits", ENGINES)
def test_forward_config1_golden(engine, precision, tol_fm, tol_logits):
    """BASELINE.json config 1: one 1280x1280 page, N=32, K=8 - the reference's own CPU-runnable case."""
    g = load_golden("g_c1_r18_img1280")
    m, _ = make_model(engine, precision)
    inp = to_dev(synth.gen(1, 32, 8, seed=0, img=1280))
    with torch.no_grad():
        r = m._native.forward(*inp, return_intermediates=True)
    assert rel_err(t2n(r["fm"]).transpose(0, 3, 1, 2)[:, :, ::8, ::8], g["fm_sample"]) < tol_fm
    assert rel_err(t2n(r["own"]), g["own"]) < tol_fm
    assert rel_err(t2n(r["logits"]), g["logits"]) < tol_logits


@pytest.mark.parametrize("engine,precision,tol_fm,tol_logits", ENGINES)
def test_forward_ragged_pages_golden(engine, precision, tol_fm, tol_logits):
    g = load_golden("g_ragged_r18_img256")
    m, _ = make_model(engine, precision, img=256)
    inp = to_dev(synth.gen(3, 0, 24, seed=3, img=256, counts=[11, 1, 30]))
    with torch.no_grad():
        logits = m(*inp)
        vis, bb = m._get_visual_features(inp[0], inp[1]), m._get_bbox_features(inp[1])
        own = torch.cat((vis, bb, m.bn_additional_feat(inp[2])), 1)
        ctx, attn = m.gat(own, inp[3], return_attn_wts=True)      # extract_attn_wts_and_visualize.py:117-124
    assert rel_err(t2n(logits), g["logits"]) < tol_logits
    assert rel_err(t2n(vis), g["visual"]) < tol_fm and rel_err(t2n(bb), g["bbox"]) < 1e-5
    assert rel_err(t2n(ctx), g["ctx"]) < max(tol_fm, 2e-5) and np.abs(t2n(attn) - g["attn"]).max() < max(tol_fm, 1e-5)


def test_uint8_images_equal_totensor_path():
    """raw uint8 pixels in.  With the knob `stem_u8_exact` the stem forms v/255 itself == the fp32 `ToTensor`
    contract, bit for bit, on both engines (logits identical to feeding the converted fp32 images).  The tcgen05 default for
    uint8 input is the integer two-product mode (pixels 0..255 are exact in one 16-bit plane, 1/255 folded into the epilogue
    scale, no lo-plane product): within 3e-5 of the ToTensor path - the size of the three-product scheme's own error (both are
    within 2e-5 of the reference; the integer stem is the more exact of the two)."""
    from cova_b200 import ops
    g = torch.Generator().manual_seed(21)
    u8 = torch.randint(0, 256, (2, 3, 256, 256), dtype=torch.uint8, generator=g)
    f32 = u8.float().div(255)                                  # what ToTensor produces (datasets.py:41-45)
    _, bboxes, add, ci = synth.gen(2, 12, 8, seed=0, img=256)
    for engine in ("tcgen05", "simt"):
        for precision in (("fp32", "fp32x") if engine == "tcgen05" else ("fp32",)):
            m, _ = make_model(engine, precision, img=256)
            with torch.no_grad():
                b = m(f32.to(DEV), bboxes.to(DEV), add.to(DEV), ci.to(DEV))
                a = m(u8.to(DEV), bboxes.to(DEV), add.to(DEV), ci.to(DEV))
                ops.set_knob("stem_u8_exact", 1)
                try:
                    c = m(u8.to(DEV), bboxes.to(DEV), add.to(DEV), ci.to(DEV))
                finally:
                    ops.set_knob("stem_u8_exact", -1)
            assert torch.equal(c, b), (engine, precision)
            assert rel_err(t2n(a), t2n(b)) < 3e-5, (engine, precision, rel_err(t2n(a), t2n(b)))


@pytest.mark.parametrize("engine,precision,tol", [("simt", "fp32", 2e-5), ("tcgen05", "fp32", 1e-4), ("tcgen05", "fp32x", 2e-5)])
def test_forward_roi_align_variant_golden(engine, precision, tol):
    """D1 variant (the RoI op north_star names): whole forward with RoIAlign on every engine vs the live reference with
    line 58's module swapped for torchvision's RoIAlign."""
    g = load_golden("g_align_r18_img256")
    m, _ = make_model(engine, precision, img=256, roi_mode="align")
    inp = to_dev(synth.gen(2, 20, 8, seed=4, img=256))
    with torch.no_grad():
        logits = m(*inp)
        vis = m._get_visual_features(inp[0], inp[1])
    assert rel_err(t2n(vis), g["visual"]) < tol
    assert rel_err(t2n(logits), g["logits"]) < tol


# ----------------------------------------------------------------------------- the BASELINE.json configs at full shape
@pytest.mark.parametrize("engine,precision,tol_fm,tol_logits", ENGINES)
def test_forward_config2_golden(engine, precision, tol_fm, tol_logits):
    """BASELINE config 2 on the very inputs bench.py times (B=16 pages of 1280^2, N=90, K=24, seed 1): logits [1440,4],
    a strided sample of the feature map and of the RoI features against the live reference."""
    g = load_golden("g_c2_r18_b16")
    m, _ = make_model(engine, precision)
    inp = to_dev(synth.gen(16, 90, 24, seed=1))
    with torch.no_grad():
        r = m._native.forward(*inp, return_intermediates=True)
        logits = m(*inp)
    assert rel_err(t2n(r["fm"]).transpose(0, 3, 1, 2)[:, :, ::16, ::16], g["fm_sample"]) < tol_fm
    assert rel_err(t2n(r["own"][::7, :576]), g["visual_sample"]) < tol_fm
    assert logits.shape == (1440, 4) and rel_err(t2n(logits), g["logits"]) < tol_logits
    # element-wise figure beside the tensor-scale one: |a-b| / max(|b|, 1 % of the tensor scale)
    want = g["logits"].astype(np.float64)
    elt = np.abs(t2n(logits) - want) / np.maximum(np.abs(want), 1e-2 * np.abs(want).max())
    assert elt.max() < 100 * tol_logits


@pytest.mark.parametrize("engine,tol", [("simt", 2e-5), ("tcgen05", 1e-4)])
def test_forward_resnet50_ragged_golden(engine, tol):
    """ResNet-50 (D2) at 320^2 with ragged pages 17 / 2 / 40: multi-strip stem, multi-tile 1x1 GEMMs, F = 2336."""
    g = load_golden("g_r50_ragged_img320")
    m, _ = make_model(engine, img=320, backbone="resnet50")
    inp = to_dev(synth.gen(3, 0, 24, seed=14, img=320, counts=[17, 2, 40]))
    with torch.no_grad():
        r = m._native.forward(*inp, return_intermediates=True)
    assert rel_err(t2n(r["fm"]).transpose(0, 3, 1, 2)[:, :, ::4, ::4], g["fm_sample"]) < tol
    assert rel_err(t2n(r["logits"]), g["logits"]) < 2.5 * tol


def test_forward_config5_golden():
    """BASELINE config 5 shape (ResNet-50, N=300 boxes/page, K=48, 2-head GAT, 1280^2; 2 pages) on the tensor-core
    engine vs the live reference composed as SURVEY D3 defines it."""
    g = load_golden("g_c5_r50_n300_k48_h2")
    m, _ = make_model("tcgen05", backbone="resnet50", n_heads=2)
    inp = to_dev(synth.gen(2, 300, 48, seed=1))
    with torch.no_grad():
        r = m._native.forward(*inp, return_intermediates=True)
    assert rel_err(t2n(r["fm"]).transpose(0, 3, 1, 2)[:, :, ::16, ::16], g["fm_sample"]) < 1e-4
    assert r["logits"].shape == (600, 4) and rel_err(t2n(r["logits"]), g["logits"]) < 2.5e-4


@pytest.mark.parametrize("engine,tol", [("simt", 2e-5), ("tcgen05", 1e-4)])
def test_forward_resnet50_golden(engine, tol):
    """ResNet-50 truncated the same way (3 Bottlenecks, 256 channels) vs the live-reference fixture."""
    g = load_golden("g_r50_img128")
    m, _ = make_model(engine, img=128, backbone="resnet50")
    with torch.no_grad():
        r = m._native.forward(*to_dev(synth.gen(1, 16, 8, seed=5, img=128)), return_intermediates=True)
    assert rel_err(t2n(r["fm"]).transpose(0, 3, 1, 2), g["fm"]) < tol
    assert rel_err(t2n(r["logits"]), g["logits"]) < 2.5 * tol


@pytest.mark.parametrize("cin,cout,with_res,out", [(64, 64, False, "planes"), (64, 256, True, "planes"),
                                                  (256, 64, False, "planes"), (64, 256, True, "f32")])
def test_conv1x1_tcgen05_kernel(cin, cout, with_res, out):
    """ResNet-50 Bottleneck 1x1 convs on split planes vs the oracle conv (ragged M: 2 x 37 x 45 pixel rows)."""
    o = ops()
    g = torch.Generator().manual_seed(9)
    B, H, W = 2, 37, 45
    x = torch.randn(B, H, W, cin, generator=g)
    w = torch.randn(cout, cin, 1, 1, generator=g) / cin ** 0.5
    sc, sh = 0.5 + torch.rand(cout, generator=g), 0.1 * torch.randn(cout, generator=g)
    res = torch.randn(B, H, W, cout, generator=g)
    ref = torch_conv(x.permute(0, 3, 1, 2).numpy(), w.numpy(), 1, 0).transpose(0, 2, 3, 1) * sc.numpy() + sh.numpy()
    if with_res:
        ref = ref + res.numpy()
    ref = np.maximum(ref, 0)
    xp, rp = o.Planes(o.BF16X2, x.shape, DEV), o.Planes(o.BF16X2, res.shape, DEV)
    for pl, t in ((xp, x), (rp, res)):
        hi, lo = _split(t)
        pl.p0.copy_(hi); pl.p1.copy_(lo)
    y = o.conv1x1_bn_act_fwd(xp, o.pack_linear_weight(w.flatten(1).to(DEV)), sc.to(DEV), sh.to(DEV),
                             res=rp if with_res else None, relu=True, out_dtype=o.F32 if out == "f32" else o.BF16X2)
    torch.cuda.synchronize()
    assert rel_err(t2n(y.float()), ref) < 5e-5


def test_linear_tcgen05_split_plane_output():
    """GEMM epilogue writing split-bf16 planes (feeds the tensor-core 3x3 conv): hi + lo reproduces the fp32 result."""
    o = ops()
    g = torch.Generator().manual_seed(8)
    x, w = torch.randn(333, 256, generator=g), torch.randn(64, 256, generator=g) / 16
    sc, sh = 0.5 + torch.rand(64, generator=g), torch.randn(64, generator=g)
    want = np.maximum(O.linear(x.numpy(), w.numpy()) * sc.numpy() + sh.numpy(), 0)
    pl = o.linear_fwd(x.to(DEV), o.pack_linear_weight(w.to(DEV)), None, sc.to(DEV), sh.to(DEV), relu=True,
                      engine=o.ENGINE_TCGEN05, out_planes=True)
    assert rel_err(t2n(pl.float()).reshape(333, 64), want) < 5e-5


def test_forward_constructor_variants_golden():
    g = load_golden("g_noctx_nobbox_roi2x5")
    m, _ = make_model("simt", img=128, roi=(2, 5), use_context=False, hidden=0, bbhd=0)
    images, bboxes, add, _ = to_dev(synth.gen(2, 9, 0, seed=6, img=128))
    with torch.no_grad():
        out = m(images, bboxes, add, torch.empty((18, 0), dtype=torch.long, device=DEV))
    assert rel_err(t2n(out), g["logits"]) < 2e-5
    g = load_golden("g_addfeat7")
    m, _ = make_model("simt", img=128, n_add=7)
    images, bboxes, _, ci = to_dev(synth.gen(2, 9, 8, seed=7, img=128))
    with torch.no_grad():
        out = m(images, bboxes, torch.from_numpy(g["additional_feats"]).to(DEV), ci)
    assert rel_err(t2n(out), g["logits"]) < 2e-5


@pytest.mark.parametrize("engine,tol", [("tcgen05", 1e-4), ("simt", 2e-5)])
def test_forward_odd_image_size_and_empty_pages(engine, tol):
    """img_H = 200 (conv 100, map 50x50: partial stem strips/bands, partial 8x16 conv tiles), a page with ZERO boxes
    in the middle of the batch, a single-box page, and a batch with no boxes at all."""
    m, sd = make_model(engine, img=200)
    inp = synth.gen(4, 0, 8, seed=13, img=200, counts=[7, 0, 1, 5])
    want = O.cova_forward(sd, *[t.numpy() for t in inp], conv=torch_conv)
    with torch.no_grad():
        got = m(*to_dev(inp))
        none = m(inp[0].to(DEV), torch.empty((0, 5), device=DEV), torch.empty((0, 0), device=DEV),
                 torch.empty((0, 8), dtype=torch.long, device=DEV))
    assert got.shape == (13, 4) and rel_err(t2n(got), want) < tol
    assert none.shape == (0, 4)


def test_single_page_single_sm_wave():
    """B = 1 (fewer tiles than a full persistent wave in the small kernels) at full resolution vs the simt engine."""
    inp = to_dev(synth.gen(1, 90, 24, seed=4))
    a, _ = make_model("tcgen05")
    b, _ = make_model("simt")
    with torch.no_grad():
        assert rel_err(t2n(a(*inp)), t2n(b(*inp))) < 1e-4


def test_two_head_forward_vs_oracle():
    m, sd = make_model("simt", img=128, n_heads=2)
    inp = synth.gen(2, 12, 8, seed=9, img=128)
    with torch.no_grad():
        r = m._native.forward(*to_dev(inp), return_intermediates=True)
    own = t2n(r["own"])
    heads = [(sd[f"gat.heads.{i}.W_i.weight"], sd[f"gat.heads.{i}.W_j.weight"],
              sd[f"gat.heads.{i}.attention_layer.weight"], sd[f"gat.heads.{i}.attention_layer.bias"]) for i in range(2)]
    want = O.gat_multihead(own, inp[3].numpy(), heads)
    assert rel_err(t2n(r["ctx"]), want) < 2e-5
    assert rel_err(t2n(r["logits"]), O.decoder(np.concatenate([own, want], 1), sd)) < 2e-5


# ----------------------------------------------------------------------------- full BASELINE size: properties
def test_full_size_properties_config2():
    """BASELINE config 2 (B=16, 1280^2, N=90, K=24) is too big for the CPU oracle in a test, so check
    size-independent properties: (1) the two engines agree, (2) a page's logits do not depend on the rest of the
    batch (pages are independent units in eval mode), (3) RoIPool of the produced feature map is
    bit-identical to the oracle's RoIPool on a sample of boxes, (4) determinism."""
    B, N, K = 16, 90, 24
    inp = synth.gen(B, N, K, seed=1)
    dinp = to_dev(inp)
    m_tc, _ = make_model("tcgen05", "fp32")
    m_si, _ = make_model("simt")
    with torch.no_grad():
        r = m_tc._native.forward(*dinp, return_intermediates=True)
        l_tc, l_tc2, l_si = r["logits"], m_tc(*dinp), m_si(*dinp)
        p = 11
        rows = slice(p * N, (p + 1) * N)
        ci = dinp[3][rows].clone()
        ci[ci >= 0] -= p * N
        bb = dinp[1][rows].clone()
        bb[:, 0] = 0
        l_page = m_tc(dinp[0][p:p + 1], bb, dinp[2][rows], ci)
    assert torch.equal(l_tc, l_tc2)
    assert rel_err(t2n(l_tc), t2n(l_si)) < 1e-4
    assert rel_err(t2n(l_page), t2n(l_tc[rows])) < 1e-5
    sample = torch.arange(0, B * N, 37)
    fm_nchw = t2n(r["fm"]).transpose(0, 3, 1, 2)
    want = O.roi_pool(fm_nchw, inp[1][sample].numpy(), (3, 3), 0.25).reshape(len(sample), -1)
    assert np.array_equal(t2n(r["own"][sample.to(DEV), :576]), want)


# ----------------------------------------------------------------------------- training path (composite + native RoI)
def test_train_step_matches_reference_grads():
    """`train.py:45-60` semantics on the autograd path: train-mode BN (batch stats), dropout off, CE(sum)."""
    from cova_b200.models import CoVA
    g = load_golden("g_train_r18_img128")
    m = CoVA((3, 3), 128, 4, True, 384, 32, 0, 0.0, None, pretrained=False)
    m.load_state_dict(synth.make_state_dict(123), strict=True)
    m = m.to(DEV).train()
    images, bboxes, add, ci, labels = to_dev(synth.gen(2, 12, 8, seed=8, img=128, with_labels=True))
    out = m(images, bboxes, add, ci)
    loss = torch.nn.CrossEntropyLoss(reduction="sum")(out, labels)
    loss.backward()
    assert rel_err(t2n(out), g["logits"]) < 1e-4 and abs(float(loss) - float(g["loss"])) < 1e-3 * abs(float(g["loss"]))
    grads = dict(m.named_parameters())
    gscale = max(float(np.abs(g[k]).max()) for k in g if k.startswith("grad:"))
    for k in [k for k in g if k.startswith("grad:")]:
        # Two parameters have a mathematically ZERO gradient, so both implementations hold rounding noise only:
        # gat.W_i shifts every logit of a row by the same s_i (softmax-invariant where LeakyReLU is linear) and
        # decoder.1.bias is removed by the train-mode BatchNorm that follows it -> absolute floor from the
        # overall gradient scale.
        got, want = t2n(grads[k[5:]].grad), g[k]
        assert np.abs(got - want).max() < 2e-3 * np.abs(want).max() + 1e-6 * gscale, k
    sd = m.state_dict()
    for k in [k for k in g if k.startswith("buf:")]:
        assert rel_err(t2n(sd[k[4:]]), g[k]) < 1e-4, k
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    first = float(loss)
    for _ in range(5):
        opt.zero_grad()
        l2 = torch.nn.CrossEntropyLoss(reduction="sum")(m(images, bboxes, add, ci), labels)
        l2.backward()
        opt.step()
    assert float(l2) < first
    m.eval()                                   # weights changed -> the native path must rebuild its caches
    with torch.no_grad():
        a = m(images, bboxes, add, ci)
        m.engine = "simt"
        b = m(images, bboxes, add, ci)
    assert rel_err(t2n(a), t2n(b)) < 1e-4


def test_train_step_resnet50_matches_reference_grads():
    """Configs 3/4 semantics (ResNet-50 backbone, train mode, CE(sum)) against the LIVE reference: logits, loss, every
    backbone gradient, BatchNorm buffers (`train.py:45-60`).  Gradient bound 2e-3 of each tensor's scale (+ a floor
    from the global scale for the mathematically-zero ones)."""
    from cova_b200.models import CoVA
    g = load_golden("g_train_r50_img192")
    m = CoVA((3, 3), 192, 4, True, 384, 32, 0, 0.0, None, pretrained=False, backbone="resnet50")
    m.load_state_dict(synth.make_state_dict(123, backbone="resnet50"), strict=True)
    m = m.to(DEV).train()
    # input seed = the first one for which the reference's own fp32 and fp64 gradients agree (no ReLU / pooling decision
    # inside rounding noise; oracle/make_golden.py:train_fixture_r50, profiles/r02_train_grad_conditioning_r50.txt)
    images, bboxes, add, ci, labels = to_dev(synth.gen(2, 14, 8, seed=int(g["seed"]), img=192, with_labels=True))
    out = m(images, bboxes, add, ci)
    loss = torch.nn.CrossEntropyLoss(reduction="sum")(out, labels)
    loss.backward()
    assert rel_err(t2n(out), g["logits"]) < 1e-4 and abs(float(loss) - float(g["loss"])) < 1e-3 * abs(float(g["loss"]))
    grads = dict(m.named_parameters())
    gscale = max(float(np.abs(g[k]).max()) for k in g if k.startswith("grad:"))
    worst = {}
    for k in [k for k in g if k.startswith("grad:")]:
        got, want = t2n(grads[k[5:]].grad), g[k]
        if got.shape != want.shape:
            got = got[:, ::8]
        worst[k] = np.abs(got - want).max() / (np.abs(want).max() + 1e-3 * gscale)
    # Two bounds.  (1) Every tensor within 5e-3: a single ReLU / pooling decision on an element within rounding of a tie
    # still flips between this path and the reference (~1 element per million per 1e-6 of noise; measured with
    # tools/diag_train_grads.py: one element after block 0's bn2 here), and everything upstream of it moves by up to
    # ~0.3 %; the reference's own fp32-vs-fp64 gradients move by 0.3-9 % on 16 of 17 input seeds for the same reason
    # (profiles/r02_train_grad_conditioning_r50.txt).  (2) Everything not upstream of such an element - at least 70 % of
    # the tensors - agrees to 5e-5, which is what pins the kernels (wgrad / dgrad / BatchNorm backward).
    bad = {k: v for k, v in worst.items() if v > 5e-3}
    assert not bad, bad
    tight = [k for k, v in worst.items() if v <= 5e-5]
    assert len(tight) >= 0.7 * len(worst), sorted(worst.items(), key=lambda kv: -kv[1])[:12]
    sd = m.state_dict()
    for k in [k for k in g if k.startswith("buf:")]:
        assert rel_err(t2n(sd[k[4:]]), g[k]) < 1e-4, k


def test_roi_pool_backward_kernel_vs_torch_scatter():
    """Native RoIPool backward == an index_add over the arg-max indices (overlapping bins/boxes collide)."""
    o = ops()
    g = torch.Generator().manual_seed(17)
    fm = torch.randn(2, 40, 48, 64, generator=g).to(DEV)
    _, bboxes, _, _ = synth.gen(2, 30, 8, seed=17, img=160)
    bboxes = bboxes.to(DEV)
    out = torch.empty((60, 576), device=DEV)
    am = o.roi_fwd(fm, bboxes, (3, 3), 0.25, out, want_argmax=True)
    go = torch.randn(60, 576, generator=g).to(DEV)
    got = o.roi_pool_bwd(go, am, bboxes, fm.shape)
    amf = am.reshape(60, 64, 9).long()
    valid = amf >= 0
    flat = ((bboxes[:, 0].long().view(60, 1, 1) * 40 * 48 + amf.clamp_min(0)) * 64 + torch.arange(64, device=DEV).view(1, 64, 1))
    want = torch.zeros(fm.numel(), device=DEV, dtype=torch.float64)
    want.index_add_(0, flat[valid], go.reshape(60, 64, 9)[valid].double())
    assert rel_err(t2n(got).reshape(-1), want.cpu().numpy()) < 1e-5


@pytest.mark.parametrize("P", [(3, 3), (7, 7), (2, 5)])
def test_roi_align_backward_kernel_vs_torchvision(P):
    """Native RoIAlign backward (atomic scatter of the bilinear taps) vs torchvision.ops.roi_align's autograd - the op
    the reference model would call with line 58 swapped - on the adversarial boxes of g_roi (negative
    coordinates, sub-pixel boxes, boxes outside the map), C = 64."""
    import torchvision
    from cova_b200.models import _RoIAlignFn
    g = load_golden("g_roi")
    gen = torch.Generator().manual_seed(31)
    fm = torch.randn(2, 64, 40, 48, generator=gen).to(DEV)
    boxes = torch.from_numpy(g["boxes"]).to(DEV)
    go = torch.randn(boxes.shape[0], 64 * P[0] * P[1], generator=gen).to(DEV)
    a = fm.permute(0, 2, 3, 1).contiguous().requires_grad_(True)
    out = _RoIAlignFn.apply(a, boxes, P, 0.25)
    out.backward(go)
    b = fm.clone().requires_grad_(True)
    ref = torchvision.ops.roi_align(b, boxes, P, 0.25, 2, False).reshape(boxes.shape[0], -1)
    ref.backward(go)
    assert rel_err(t2n(out), t2n(ref)) < 3e-5
    assert rel_err(t2n(a.grad).transpose(0, 3, 1, 2), t2n(b.grad)) < 2e-5


def test_out_of_range_indices_are_safe():
    """A batch index outside [0, B) or a context id outside [0, T) (never produced by the reference's loader) must not
    read or write out of bounds: such a box pools to zeros and gets no gradient, such a neighbour counts as padding."""
    o = ops()
    g = torch.Generator().manual_seed(5)
    fm = torch.randn(2, 20, 24, 64, generator=g).to(DEV)
    _, bboxes, _, _ = synth.gen(2, 6, 4, seed=5, img=80)
    bad = bboxes.clone()
    bad[3, 0], bad[7, 0] = 9.0, -2.0
    for mode in ("pool", "align"):
        out_ok, out_bad = torch.empty((12, 576), device=DEV), torch.empty((12, 576), device=DEV)
        o.roi_fwd(fm, bboxes.to(DEV), (3, 3), 0.25, out_ok, mode=mode)
        am = o.roi_fwd(fm, bad.to(DEV), (3, 3), 0.25, out_bad, mode=mode, want_argmax=(mode == "pool"))
        keep = [i for i in range(12) if i not in (3, 7)]
        assert torch.equal(out_ok[keep], out_bad[keep]) and float(out_bad[[3, 7]].abs().max()) == 0.0
        go = torch.ones((12, 576), device=DEV)
        gfm = (o.roi_pool_bwd(go, am, bad.to(DEV), fm.shape) if mode == "pool"
               else o.roi_align_bwd(go, bad.to(DEV), (3, 3), 0.25, fm.shape))
        assert torch.isfinite(gfm).all()
    T, K, Hd = 10, 6, 32
    whj, s, t = torch.randn(T, Hd, generator=g).to(DEV), torch.randn(T, generator=g).to(DEV), torch.randn(T, generator=g).to(DEV)
    ci = torch.randint(-1, T, (T, K), generator=g)
    ci_bad = ci.clone()
    ci_pad = ci.clone()
    ci_bad[2, 1], ci_pad[2, 1] = T + 5, -1
    ci_bad[4, 0], ci_pad[4, 0] = 1 << 40, -1
    a, b = torch.empty((T, Hd), device=DEV), torch.empty((T, Hd), device=DEV)
    o.gat_fwd(whj, s, t, 0.1, 0.2, ci_bad.to(DEV), a)
    o.gat_fwd(whj, s, t, 0.1, 0.2, ci_pad.to(DEV), b)
    assert torch.equal(a, b)


def test_gat_backward_kernel_vs_autograd():
    """Native GAT gather backward vs PyTorch autograd of the operator formulation (arbitrary ids, padding, an
    all -1 row), for the layer's input and all four parameters."""
    from cova_b200.models import GraphAttentionLayer
    import os
    g = torch.Generator().manual_seed(23)
    h = torch.randn(50, 96, generator=g).to(DEV)
    ci = torch.randint(-1, 50, (50, 12), generator=g)
    ci[5] = -1
    ci = ci.to(DEV)
    go = torch.randn(50, 64, generator=g).to(DEV)
    layer = GraphAttentionLayer(96, 64).to(DEV)
    grads = {}
    for mode in ("native", "torch"):
        os.environ["COVA_B200_GAT_TRAIN"] = mode
        x = h.clone().requires_grad_(True)
        layer.zero_grad()
        out = layer(x, ci)
        (out * go).sum().backward()
        grads[mode] = [out.detach(), x.grad] + [p.grad.clone() for p in layer.parameters()]
    os.environ.pop("COVA_B200_GAT_TRAIN")
    for a, b in zip(grads["native"], grads["torch"]):
        assert rel_err(t2n(a), t2n(b)) < 2e-4 or float(b.abs().max()) < 1e-6


def test_inputs_rejected_loudly():
    o = ops()
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        o.roi_fwd(torch.zeros(1, 8, 8, 64), torch.zeros(1, 5), (3, 3), 0.25, torch.zeros(1, 576))
    with pytest.raises(RuntimeError, match="multiple of 64"):
        o.roi_fwd(torch.zeros(1, 8, 8, 48, device=DEV), torch.zeros(1, 5, device=DEV), (3, 3), 0.25,
                  torch.zeros(1, 432, device=DEV))
