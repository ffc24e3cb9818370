"""Seeded synthetic inputs in the layout the reference's loader hands to ``CoVA.forward``.

Mirrors ``datasets.py:112-128`` (xywh->xyxy boxes, +-context_size pre-order window padded with -1)
and ``custom_collate_fn`` (``datasets.py:159-190``: batch-index column, batch-global context ids).
Specification: SURVEY.md section 8(d).
"""
import numpy as np
import torch


def context_window(n_boxes, context_size):
    """[n_boxes, 2*context_size] int64 window of pre-order neighbours, right-padded with -1
    (``datasets.py:117-128``)."""
    K = 2 * context_size
    ci = np.full((n_boxes, K), -1, dtype=np.int64)
    for i in range(n_boxes):
        ctx = list(range(max(0, i - context_size), i)) + list(
            range(i + 1, min(n_boxes, i + context_size + 1))
        )
        ci[i, : len(ctx)] = ctx
    return ci


def gen(B, N, K, seed=0, img=1280, with_labels=False, counts=None):
    """Returns (images [B,3,img,img] f32, bboxes [T,5] f32, additional_feats [T,0] f32,
    context_indices [T,K] i64[, labels [T] i64]).  ``counts`` (list of B ints) gives ragged pages."""
    g = torch.Generator().manual_seed(seed)
    images = torch.rand(B, 3, img, img, generator=g)
    counts = [N] * B if counts is None else list(counts)
    T = int(sum(counts))
    w = 16 + torch.rand(T, generator=g) * (min(512, img // 2) - 16)
    h = 8 + torch.rand(T, generator=g) * (min(256, img // 4) - 8)
    x1 = torch.rand(T, generator=g) * (img - w)
    y1 = torch.rand(T, generator=g) * (img - h)
    page = torch.repeat_interleave(torch.arange(B), torch.tensor(counts)).float()
    bboxes = torch.stack([page, x1, y1, x1 + w, y1 + h], 1).contiguous()
    cs = K // 2
    cis, off = [], 0
    for n in counts:
        ci = context_window(n, cs)
        ci[ci >= 0] += off
        cis.append(ci)
        off += n
    ci = torch.from_numpy(np.concatenate(cis, 0)) if T else torch.empty(0, K, dtype=torch.long)
    add = torch.empty(T, 0)
    if not with_labels:
        return images, bboxes, add, ci
    labels = torch.zeros(T, dtype=torch.long)
    off = 0
    for n in counts:
        pos = torch.randperm(n, generator=g)[:3]
        for c, p in enumerate(pos):
            labels[off + p] = c + 1
        off += n
    return images, bboxes, add, ci, labels


# ----------------------------------------------------------------------------- seeded weights
def state_dict_spec(backbone="resnet18", roi_output_size=(3, 3), hidden_dim=384, bbox_hidden_dim=32,
                    n_additional_feat=0, n_classes=4, use_context=True, n_heads=1):
    """Ordered (key, shape) list of the reference `CoVA.state_dict()` (SURVEY.md section 8(b),
    `models.py:49-90`, `:156-163`) for the given constructor arguments."""
    spec = []

    def bn(p, c):
        spec.extend([(p + ".weight", (c,)), (p + ".bias", (c,)), (p + ".running_mean", (c,)),
                     (p + ".running_var", (c,)), (p + ".num_batches_tracked", ())])

    spec.append(("convnet.0.weight", (64, 3, 7, 7)))
    bn("convnet.1", 64)
    if backbone == "resnet18":
        C = 64
        for b in range(2):
            p = f"convnet.4.{b}"
            spec.append((p + ".conv1.weight", (64, 64, 3, 3)))
            bn(p + ".bn1", 64)
            spec.append((p + ".conv2.weight", (64, 64, 3, 3)))
            bn(p + ".bn2", 64)
    elif backbone == "resnet50":
        C = 256
        for b in range(3):
            p = f"convnet.4.{b}"
            cin = 64 if b == 0 else 256
            spec.append((p + ".conv1.weight", (64, cin, 1, 1)))
            bn(p + ".bn1", 64)
            spec.append((p + ".conv2.weight", (64, 64, 3, 3)))
            bn(p + ".bn2", 64)
            spec.append((p + ".conv3.weight", (256, 64, 1, 1)))
            bn(p + ".bn3", 256)
            if b == 0:
                spec.append((p + ".downsample.0.weight", (256, 64, 1, 1)))
                bn(p + ".downsample.1", 256)
    else:
        raise ValueError(backbone)
    n_feat = C * roi_output_size[0] * roi_output_size[1] + bbox_hidden_dim + n_additional_feat
    if bbox_hidden_dim > 0:
        spec.extend([("bbox_feat_encoder.0.weight", (bbox_hidden_dim, 5)),
                     ("bbox_feat_encoder.0.bias", (bbox_hidden_dim,))])
        bn("bbox_feat_encoder.1", bbox_hidden_dim)
    if n_additional_feat > 0:
        bn("bn_additional_feat", n_additional_feat)
    if use_context:
        if n_heads == 1:
            heads = [("gat", hidden_dim)]
        else:
            heads = [(f"gat.heads.{i}", hidden_dim // n_heads) for i in range(n_heads)]
        for p, hd in heads:
            spec.extend([(p + ".W_i.weight", (hd, n_feat)), (p + ".W_j.weight", (hd, n_feat)),
                         (p + ".attention_layer.weight", (1, 2 * hd)), (p + ".attention_layer.bias", (1,))])
    n_total = n_feat + (hidden_dim if use_context else 0)
    spec.extend([("decoder.1.weight", (n_total, n_total)), ("decoder.1.bias", (n_total,))])
    bn("decoder.2", n_total)
    spec.extend([("decoder.5.weight", (n_classes, n_total)), ("decoder.5.bias", (n_classes,))])
    return spec


def make_state_dict(seed=123, **cfg):
    """Seeded, non-trivial weights for every key of `state_dict_spec(**cfg)` (random-init stands in
    for the pretrained download the reference needs at `models.py:49`; there is no network).
    BN gets non-identity affine + running stats so folded-BN paths are actually exercised."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for key, shape in state_dict_spec(**cfg):
        leaf = key.rsplit(".", 1)[1]
        if leaf == "num_batches_tracked":
            t = torch.tensor(1, dtype=torch.long)
        elif leaf == "running_var":
            t = 0.5 + torch.rand(shape, generator=g)
        elif leaf == "running_mean":
            t = 0.1 * torch.randn(shape, generator=g)
            if key.startswith("bbox_feat_encoder.1"):
                t = 20.0 * torch.randn(shape, generator=g)
        elif len(shape) == 1 and (".bn" in key or key.startswith("convnet.1.") or ".downsample.1." in key
                                  or key.startswith(("bbox_feat_encoder.1.", "decoder.2.", "bn_additional_feat."))):
            t = (0.5 + torch.rand(shape, generator=g)) if leaf == "weight" else 0.1 * torch.randn(shape, generator=g)
            if key.startswith("bbox_feat_encoder.1") and leaf == "weight":
                pass
        elif len(shape) == 4:
            fan_out = shape[0] * shape[2] * shape[3]
            t = torch.randn(shape, generator=g) * (2.0 / fan_out) ** 0.5
        elif len(shape) == 2:
            bound = 1.0 / shape[1] ** 0.5
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        else:  # linear bias
            t = (torch.rand(shape, generator=g) * 2 - 1) * 0.05
        if key.startswith("bbox_feat_encoder.1.running_var"):
            t = t * 4.0e4   # Linear(5,32) of raw pixel coordinates has O(100) outputs (models.py:134-142)
        sd[key] = t.float() if t.dtype != torch.long else t
    return sd
