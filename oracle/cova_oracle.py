"""ORACLE - TEST INFRASTRUCTURE ONLY.  Never imported by the product package.

CPU restatement (numpy, fp32) of the reference's per-webpage forward hot path
(`/root/reference/models.py`, commit 794e95a).  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` legs may import this module.

Pinning: the reference ships no tests or golden vectors, so this oracle is
pinned against OUTPUTS OF THE REFERENCE ITSELF, run in the build container by
`oracle/make_golden.py` (imports `/root/reference/models.py` live, one offline monkeypatch for the
pretrained-weights download) and frozen under `tests/golden/*.npz`; `tests/test_oracle.py` checks
every function below against those fixtures.  RoIPool/RoIAlign live in torchvision's compiled
`_C.so` (pinned 0.7.0 in `requirements.txt:4`, 0.26.0 installed); their published algorithm is
restated here and checked bit-exactly (RoIPool) against the installed binary in the same fixtures.

Each function cites the reference lines it follows.
"""
import numpy as np

F32 = np.float32


# ----------------------------------------------------------------------------- backbone (A2)
def conv2d_nchw(x, w, stride=1, pad=0):
    """Plain cross-correlation, NCHW x [B,Ci,H,W], OIHW w [Co,Ci,kh,kw], no bias.
    torchvision ResNet convs (`models.py:49-51` keeps conv1 + layer1); fp32 accumulate."""
    x = np.asarray(x, F32)
    w = np.asarray(w, F32)
    B, Ci, H, W = x.shape
    Co, _, kh, kw = w.shape
    Ho = (H + 2 * pad - kh) // stride + 1
    Wo = (W + 2 * pad - kw) // stride + 1
    xp = np.zeros((B, Ci, H + 2 * pad, W + 2 * pad), F32)
    xp[:, :, pad : pad + H, pad : pad + W] = x
    out = np.zeros((B, Co, Ho, Wo), F32)
    for r in range(kh):
        for s in range(kw):
            patch = xp[:, :, r : r + stride * Ho : stride, s : s + stride * Wo : stride]
            out += np.einsum("bchw,oc->bohw", patch, w[:, :, r, s], optimize=True).astype(F32)
    return out


def bn_eval(x, weight, bias, mean, var, eps=1e-5):
    """Eval-mode BatchNorm: running-stat affine (`nn.BatchNorm2d/1d`, eps 1e-5). Channel = dim 1."""
    shape = [1, -1] + [1] * (x.ndim - 2)
    scale = (np.asarray(weight, F32) / np.sqrt(np.asarray(var, F32) + F32(eps))).astype(F32)
    shift = (np.asarray(bias, F32) - np.asarray(mean, F32) * scale).astype(F32)
    return (x * scale.reshape(shape) + shift.reshape(shape)).astype(F32)


def bn_train(x, weight, bias, eps=1e-5):
    """Train-mode BatchNorm: biased batch variance over all dims but the channel dim."""
    axes = tuple(i for i in range(x.ndim) if i != 1)
    shape = [1, -1] + [1] * (x.ndim - 2)
    m = x.mean(axis=axes, dtype=np.float64)
    v = x.var(axis=axes, dtype=np.float64)
    y = (x - m.reshape(shape)) / np.sqrt(v.reshape(shape) + eps)
    return (y * np.asarray(weight, F32).reshape(shape) + np.asarray(bias, F32).reshape(shape)).astype(F32)


def maxpool3x3s2p1(x):
    """`nn.MaxPool2d(3, 2, 1)` of the ResNet stem; padding acts as -inf."""
    B, C, H, W = x.shape
    Ho, Wo = (H + 2 - 3) // 2 + 1, (W + 2 - 3) // 2 + 1
    xp = np.full((B, C, H + 2, W + 2), -np.inf, F32)
    xp[:, :, 1 : 1 + H, 1 : 1 + W] = x
    out = np.full((B, C, Ho, Wo), -np.inf, F32)
    for r in range(3):
        for s in range(3):
            out = np.maximum(out, xp[:, :, r : r + 2 * Ho : 2, s : s + 2 * Wo : 2])
    return out


def _bn(sd, prefix, x, train):
    if train:
        return bn_train(x, sd[prefix + ".weight"], sd[prefix + ".bias"])
    return bn_eval(x, sd[prefix + ".weight"], sd[prefix + ".bias"],
                   sd[prefix + ".running_mean"], sd[prefix + ".running_var"])


def backbone(images, sd, train=False, conv=conv2d_nchw):
    """`CoVA._get_visual_features` first half (`models.py:124-125`): the `[:-5]`-truncated ResNet
    (`models.py:49-51`) = conv1 7x7 s2 p3 -> bn1 -> relu -> maxpool -> layer1.
    BasicBlock (torchvision resnet.py BasicBlock.forward) when `convnet.4.0.conv3.weight` is absent,
    else Bottleneck (Bottleneck.forward, first block with 1x1 downsample)."""
    relu = lambda t: np.maximum(t, F32(0))
    x = conv(images, sd["convnet.0.weight"], 2, 3)
    x = relu(_bn(sd, "convnet.1", x, train))
    x = maxpool3x3s2p1(x)
    b = 0
    while f"convnet.4.{b}.conv1.weight" in sd:
        p = f"convnet.4.{b}"
        idt = x
        if p + ".conv3.weight" in sd:  # Bottleneck
            o = relu(_bn(sd, p + ".bn1", conv(x, sd[p + ".conv1.weight"], 1, 0), train))
            o = relu(_bn(sd, p + ".bn2", conv(o, sd[p + ".conv2.weight"], 1, 1), train))
            o = _bn(sd, p + ".bn3", conv(o, sd[p + ".conv3.weight"], 1, 0), train)
            if p + ".downsample.0.weight" in sd:
                idt = _bn(sd, p + ".downsample.1", conv(x, sd[p + ".downsample.0.weight"], 1, 0), train)
        else:  # BasicBlock
            o = relu(_bn(sd, p + ".bn1", conv(x, sd[p + ".conv1.weight"], 1, 1), train))
            o = _bn(sd, p + ".bn2", conv(o, sd[p + ".conv2.weight"], 1, 1), train)
        x = relu(o + idt)
        b += 1
    return x


# ----------------------------------------------------------------------------- RoIPool (A4)
def _c_round(v):
    """C `round()` (half away from zero) of an fp32 value, evaluated exactly in fp64."""
    v = np.float64(v)
    return int(np.sign(v) * np.floor(np.abs(v) + 0.5))


def roi_pool(fm, rois, P, spatial_scale):
    """`torchvision.ops.RoIPool(P, spatial_scale)` (`models.py:58`, applied `models.py:125-127`).
    fm [B,C,H,W] f32, rois [T,5] = [batch_idx, x1,y1,x2,y2] in image pixels -> [T,C,PH,PW].
    Semantics restated from SURVEY.md section 8 row A4 (torchvision roi_pool kernel): round-half-away
    box edges on the fp32 product, +1 extents clamped to >=1, floor/ceil bin edges on fp32
    products, clamp to the map, empty bin -> 0, else max."""
    fm = np.asarray(fm, F32)
    rois = np.asarray(rois, F32)
    PH, PW = P
    T = rois.shape[0]
    _, C, H, W = fm.shape
    s = F32(spatial_scale)
    out = np.zeros((T, C, PH, PW), F32)
    for n in range(T):
        b = int(rois[n, 0])
        sw, sh = _c_round(rois[n, 1] * s), _c_round(rois[n, 2] * s)
        ew, eh = _c_round(rois[n, 3] * s), _c_round(rois[n, 4] * s)
        rw, rh = max(ew - sw + 1, 1), max(eh - sh + 1, 1)
        bh, bw = F32(rh) / F32(PH), F32(rw) / F32(PW)
        for ph in range(PH):
            hs = min(max(int(np.floor(F32(ph) * bh)) + sh, 0), H)
            he = min(max(int(np.ceil(F32(ph + 1) * bh)) + sh, 0), H)
            for pw in range(PW):
                ws = min(max(int(np.floor(F32(pw) * bw)) + sw, 0), W)
                we = min(max(int(np.ceil(F32(pw + 1) * bw)) + sw, 0), W)
                if he <= hs or we <= ws:
                    continue
                out[n, :, ph, pw] = fm[b, :, hs:he, ws:we].max(axis=(1, 2))
    return out


# ----------------------------------------------------------------------------- RoIAlign (A4', D1)
def _bilinear(fm_b, y, x):
    """torchvision roi_align bilinear_interpolate: out-of-range (< -1 or > size) -> 0, clamp to edge."""
    C, H, W = fm_b.shape
    if y < -1.0 or y > H or x < -1.0 or x > W:
        return np.zeros(C, F32)
    y = F32(max(y, F32(0)))
    x = F32(max(x, F32(0)))
    yl, xl = int(y), int(x)
    if yl >= H - 1:
        yh = yl = H - 1
        y = F32(yl)
    else:
        yh = yl + 1
    if xl >= W - 1:
        xh = xl = W - 1
        x = F32(xl)
    else:
        xh = xl + 1
    ly, lx = F32(y - F32(yl)), F32(x - F32(xl))
    hy, hx = F32(F32(1) - ly), F32(F32(1) - lx)
    w1, w2, w3, w4 = F32(hy * hx), F32(hy * lx), F32(ly * hx), F32(ly * lx)
    return (w1 * fm_b[:, yl, xl] + w2 * fm_b[:, yl, xh] + w3 * fm_b[:, yh, xl] + w4 * fm_b[:, yh, xh]).astype(F32)


def roi_align(fm, rois, P, spatial_scale, sampling_ratio=2, aligned=False):
    """`torchvision.ops.RoIAlign(P, s, sampling_ratio=2, aligned=False)` - the north-star-named
    variant of `models.py:58` (SURVEY.md D1, row A4').  Average of g x g bilinear samples per bin."""
    fm = np.asarray(fm, F32)
    rois = np.asarray(rois, F32)
    PH, PW = P
    T = rois.shape[0]
    C = fm.shape[1]
    s = F32(spatial_scale)
    off = F32(0.5) if aligned else F32(0)
    out = np.zeros((T, C, PH, PW), F32)
    for n in range(T):
        b = int(rois[n, 0])
        x1, y1 = F32(rois[n, 1] * s - off), F32(rois[n, 2] * s - off)
        x2, y2 = F32(rois[n, 3] * s - off), F32(rois[n, 4] * s - off)
        rw, rh = F32(x2 - x1), F32(y2 - y1)
        if not aligned:
            rw, rh = max(rw, F32(1)), max(rh, F32(1))
        bh, bw = F32(rh / F32(PH)), F32(rw / F32(PW))
        gh = sampling_ratio if sampling_ratio > 0 else int(np.ceil(rh / PH))
        gw = sampling_ratio if sampling_ratio > 0 else int(np.ceil(rw / PW))
        cnt = F32(max(gh * gw, 1))
        for ph in range(PH):
            for pw in range(PW):
                acc = np.zeros(C, F32)
                for iy in range(gh):
                    y = F32(y1 + F32(ph) * bh + F32(F32(iy) + F32(0.5)) * bh / F32(gh))
                    for ix in range(gw):
                        x = F32(x1 + F32(pw) * bw + F32(F32(ix) + F32(0.5)) * bw / F32(gw))
                        acc += _bilinear(fm[b], y, x)
                out[n, :, ph, pw] = acc / cnt
    return out


# ----------------------------------------------------------------------------- positional encoder (A5)
def bbox_raw_features(bboxes):
    """`models.py:134-142`: [x1, y1, w, h, w/h] in raw pixels from [batch_idx,x1,y1,x2,y2]."""
    b = np.asarray(bboxes, F32)[:, 1:].copy()
    b[:, 2:] -= b[:, :2]
    with np.errstate(divide="ignore", invalid="ignore"):
        asp = (b[:, 2] / b[:, 3]).reshape(-1, 1)
    return np.concatenate([b, asp], axis=1).astype(F32)


def linear(x, w, b=None):
    y = np.asarray(x, F32) @ np.asarray(w, F32).T
    if b is not None:
        y = y + np.asarray(b, F32)
    return y.astype(F32)


def bbox_encoder(bboxes, sd, train=False):
    """`_get_bbox_features` + `bbox_feat_encoder` (`models.py:129-148`, `:65-70`)."""
    x = linear(bbox_raw_features(bboxes), sd["bbox_feat_encoder.0.weight"], sd["bbox_feat_encoder.0.bias"])
    x = _bn(sd, "bbox_feat_encoder.1", x, train)
    return np.maximum(x, F32(0))


# ----------------------------------------------------------------------------- GAT (A6)
def gat(h, ci, W_i, W_j, att_w, att_b, alpha=0.2, return_attn_wts=False):
    """`GraphAttentionLayer.forward` AS WRITTEN (`models.py:171-212`): pad a zero row, gather K
    neighbour rows, project with W_i / W_j, score with Linear(2H,1) on the concat, LeakyReLU,
    mask -1 -> -9e15, softmax over K, weighted sum."""
    h = np.asarray(h, F32)
    ci = np.asarray(ci, np.int64)
    N, K = ci.shape
    hp = np.concatenate([h, np.zeros((1, h.shape[1]), F32)], 0)
    h_j = hp[ci.reshape(-1)].reshape(N, K, -1)                       # :180-186 (-1 -> zero row)
    Wh_i = linear(h, W_i)                                            # :188
    Wh_j = (h_j.reshape(N * K, -1) @ np.asarray(W_j, F32).T).reshape(N, K, -1).astype(F32)  # :193
    cat = np.concatenate([np.repeat(Wh_i[:, None, :], K, 1), Wh_j], 2)
    e = (cat @ np.asarray(att_w, F32).reshape(-1) + np.asarray(att_b, F32).reshape(())).astype(F32)  # :195-199
    e = np.where(e > 0, e, F32(alpha) * e).astype(F32)               # :200
    e = np.where(ci >= 0, e, F32(-9e15)).astype(F32)                 # :202-203
    e = e - e.max(axis=1, keepdims=True)
    p = np.exp(e)
    a = (p / p.sum(axis=1, keepdims=True)).astype(F32)               # :204
    out = (a[:, :, None] * Wh_j).sum(1).astype(F32)                  # :206-208
    return (out, a) if return_attn_wts else out


def gat_multihead(h, ci, heads, alpha=0.2):
    """H independent reference layers (hidden_dim // H each), concatenated on dim 1.
    `heads` = list of (W_i, W_j, att_w, att_b)."""
    return np.concatenate([gat(h, ci, *hd, alpha=alpha) for hd in heads], axis=1)


# ----------------------------------------------------------------------------- decoder (A8)
def decoder(x, sd, train=False):
    """`models.py:82-90` with dropout inactive: Linear -> BN1d -> ReLU -> Linear (raw logits)."""
    x = linear(x, sd["decoder.1.weight"], sd["decoder.1.bias"])
    x = np.maximum(_bn(sd, "decoder.2", x, train), F32(0))
    return linear(x, sd["decoder.5.weight"], sd["decoder.5.bias"])


# ----------------------------------------------------------------------------- whole forward
def cova_forward(sd, images, bboxes, additional_feats, context_indices, roi_output_size=(3, 3),
                 roi_mode="pool", train=False, conv=conv2d_nchw, return_intermediates=False):
    """`CoVA.forward` (`models.py:94-122`), eval mode (or train mode with dropout off).
    `sd` = the model's state_dict as numpy arrays."""
    fm = backbone(images, sd, train, conv)
    scale = fm.shape[2] / images.shape[2]                            # models.py:56
    pool = roi_pool if roi_mode == "pool" else roi_align
    vis = pool(fm, bboxes, roi_output_size, scale).reshape(len(bboxes), -1)   # :125-127
    parts = [vis]
    if "bbox_feat_encoder.0.weight" in sd:
        parts.append(bbox_encoder(bboxes, sd, train))                # :108
    add = np.asarray(additional_feats, F32)
    if "bn_additional_feat.weight" in sd:
        add = _bn(sd, "bn_additional_feat", add, train)              # :109
    parts.append(add.reshape(len(bboxes), -1))
    own = np.concatenate(parts, 1).astype(F32)                       # :110
    if "gat.W_i.weight" in sd:                                       # :113-116
        ctx = gat(own, context_indices, sd["gat.W_i.weight"], sd["gat.W_j.weight"],
                  sd["gat.attention_layer.weight"], sd["gat.attention_layer.bias"])
    else:
        ctx = own[:, :0]
    comb = np.concatenate([own, ctx], 1)                             # :119
    logits = decoder(comb, sd, train)                                # :120
    if return_intermediates:
        return dict(fm=fm, visual=vis, own=own, ctx=ctx, logits=logits)
    return logits


# ----------------------------------------------------------------------------- callers either side of the forward
def ce_sum(logits, labels, ignore_index=-100):
    """`nn.CrossEntropyLoss(reduction="sum")` (`/root/reference/main.py:139`, applied `train.py:56`) and the arg-max
    hit count of `train.py:53-54`.  Returns (loss, d loss / d logits, n_correct)."""
    x = np.asarray(logits, np.float64)
    y = np.asarray(labels, np.int64)
    live = (y != ignore_index) & (y >= 0) & (y < x.shape[1])
    m = x.max(1, keepdims=True)
    lse = m[:, 0] + np.log(np.exp(x - m).sum(1))
    yy = np.where(live, y, 0)
    li = (lse - x[np.arange(len(y)), yy]) * live
    d = np.exp(x - lse[:, None])
    d[np.arange(len(y)), yy] -= 1.0
    d *= live[:, None]
    n_correct = int(((x.argmax(1) == y) & live).sum())
    return np.float32(li.sum()), d.astype(np.float32), n_correct


def adam_step(p, g, m, v, lr, beta1, beta2, eps, weight_decay, step):
    """One `torch.optim.Adam` step (`main.py:133-135`, `train.py:60`; amsgrad off) in fp32, the single-tensor
    formulation of torch/optim/adam.py: L2 decay folded into the gradient, lerp first moment, bias corrections as
    Python floats.  Returns new (p, m, v)."""
    f = np.float32
    p, g, m, v = (np.asarray(a, np.float32) for a in (p, g, m, v))
    g = g + f(weight_decay) * p
    m = m + f(1.0 - beta1) * (g - m)
    v = f(beta2) * v + f(1.0 - beta2) * g * g
    bc1, bc2 = 1.0 - beta1 ** step, 1.0 - beta2 ** step
    denom = np.sqrt(v) / f(np.sqrt(bc2)) + f(eps)
    p = p - f(lr / bc1) * (m / denom)
    return p.astype(np.float32), m.astype(np.float32), v.astype(np.float32)


def topk_hits(logits, labels, page_offsets, k=1):
    """The per-image / per-class loop of `evaluate_model` (`train.py:131-154`): for class c >= 1 the first row of the
    page labelled c (`:146`) must be among `argsort(output_img, dim=0)[n-k:]` (`:141-143`; taken as a stable ascending
    sort).  int32 [B, C]; column 0 unused, -1 where the page has no row labelled c (the reference raises there)."""
    logits, labels = np.asarray(logits, np.float32), np.asarray(labels)
    B, C = len(page_offsets) - 1, logits.shape[1]
    hits = np.zeros((B, C), np.int32)
    for b in range(B):
        r0, r1 = int(page_offsets[b]), int(page_offsets[b + 1])
        out, lab = logits[r0:r1], labels[r0:r1]
        top = np.argsort(out, axis=0, kind="stable")[max(out.shape[0] - k, 0):]
        for c in range(1, C):
            rows = np.nonzero(lab == c)[0]
            hits[b, c] = -1 if len(rows) == 0 else int(rows[0] in top[:, c])
    return hits


def build_batch(boxes_per_page, context_size, boxes_xywh=None):
    """`WebDataset.__getitem__`'s context window (`datasets.py:117-128`) and `custom_collate_fn`'s batch-index column /
    batch-global ids (`datasets.py:170-178`), from the per-page box counts.  Returns (bboxes [T,5] or None,
    context_indices int64 [T, 2*context_size])."""
    cs = context_size
    ctx, bb, seen = [], [], 0
    for page, n in enumerate(boxes_per_page):
        for i in range(n):
            c = list(range(max(0, i - cs), i)) + list(range(i + 1, min(n, i + cs + 1)))
            ctx.append([j + seen for j in c] + [-1] * (2 * cs - len(c)))
        if boxes_xywh is not None:
            b = np.asarray(boxes_xywh[seen:seen + n], np.float32).copy()
            b[:, 2:] += b[:, :2]
            bb.append(np.concatenate((np.full((n, 1), page, np.float32), b), 1))
        seen += n
    ctx = np.asarray(ctx, np.int64).reshape(seen, 2 * cs)
    return (np.concatenate(bb) if bb else None), ctx


# ----------------------------------------------------------------------------- training-mode backbone pieces (A9)
def bn_act_train(x, weight, bias, res=None, relu=True, eps=1e-5):
    """Train-mode `nn.BatchNorm2d` (+ residual) (+ ReLU) on an NCHW map, as torchvision's BasicBlock / Bottleneck run
    it under `model.train()` (`/root/reference/train.py:27`): returns (y, batch mean, biased batch variance)."""
    x = np.asarray(x, np.float64)
    m = x.mean(axis=(0, 2, 3))
    v = x.var(axis=(0, 2, 3))
    y = (x - m.reshape(1, -1, 1, 1)) / np.sqrt(v.reshape(1, -1, 1, 1) + eps)
    y = y * np.asarray(weight, np.float64).reshape(1, -1, 1, 1) + np.asarray(bias, np.float64).reshape(1, -1, 1, 1)
    if res is not None:
        y = y + np.asarray(res, np.float64)
    if relu:
        y = np.maximum(y, 0.0)
    return y.astype(F32), m.astype(F32), v.astype(F32)


def bn_act_train_backward(dy, x, weight, bias, res=None, relu=True, eps=1e-5):
    """Gradient of `bn_act_train` (what autograd computes for `train.py:59`): g = dy * [y > 0];
    dx = w / sqrt(v + eps) * (g - mean(g) - xhat * mean(g * xhat)); dres = g; dweight = sum g * xhat; dbias = sum g."""
    x = np.asarray(x, np.float64)
    g = np.asarray(dy, np.float64).copy()
    w = np.asarray(weight, np.float64).reshape(1, -1, 1, 1)
    b = np.asarray(bias, np.float64).reshape(1, -1, 1, 1)
    m = x.mean(axis=(0, 2, 3), keepdims=True)
    inv = 1.0 / np.sqrt(x.var(axis=(0, 2, 3), keepdims=True) + eps)
    xh = (x - m) * inv
    if relu:
        y = xh * w + b + (0.0 if res is None else np.asarray(res, np.float64))
        g *= y > 0
    s1 = g.mean(axis=(0, 2, 3), keepdims=True)
    s2 = (g * xh).mean(axis=(0, 2, 3), keepdims=True)
    dx = w * inv * (g - s1 - xh * s2)
    return dx.astype(F32), g.astype(F32), (g * xh).sum(axis=(0, 2, 3)).astype(F32), g.sum(axis=(0, 2, 3)).astype(F32)


def maxpool3x3s2p1_backward(x, dy):
    """Gradient of `nn.MaxPool2d(3, 2, 1)`: each output's gradient goes to the FIRST maximum of its window in
    row-major scan order (torch's max_pool2d_with_indices rule; ReLU'd maps are full of tied zeros)."""
    x, dy = np.asarray(x, F32), np.asarray(dy, F32)
    B, C, H, W = x.shape
    Ho, Wo = dy.shape[2], dy.shape[3]
    dx = np.zeros_like(x)
    for oh in range(Ho):
        for ow in range(Wo):
            h0, h1, w0, w1 = max(2 * oh - 1, 0), min(2 * oh + 2, H), max(2 * ow - 1, 0), min(2 * ow + 2, W)
            win = x[:, :, h0:h1, w0:w1].reshape(B, C, -1)
            k = win.argmax(axis=2)                        # first maximum
            hh, ww = h0 + k // (w1 - w0), w0 + k % (w1 - w0)
            bi, ci = np.meshgrid(np.arange(B), np.arange(C), indexing="ij")
            np.add.at(dx, (bi, ci, hh, ww), dy[:, :, oh, ow])
    return dx
