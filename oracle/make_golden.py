"""ORACLE - TEST INFRASTRUCTURE ONLY.  Generates `tests/golden/*.npz` by running the UNMODIFIED
reference (`/root/reference/models.py`) live in the build container.  The reference cannot travel
to the GPU box, so its outputs are frozen here as small fixtures.

    python oracle/make_golden.py            # regenerate every fixture

The only patch applied to the reference is the one SURVEY.md section 8(c) documents: the backbone factory
is called with `weights=None` instead of downloading `resnet18-f37072fd.pth` (no network).
Weights come from `cova_b200.synth.make_state_dict` (seeded) and are pushed through the
reference's own strict `load_state_dict`, which also pins checkpoint-key compatibility.
"""
import os
import sys
import warnings

import numpy as np
import torch
import torchvision

warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
import cova_b200.synth as synth  # noqa: E402

_r18, _r50 = torchvision.models.resnet18, torchvision.models.resnet50
_backbone = {"name": "resnet18"}
torchvision.models.resnet18 = lambda pretrained=True, **kw: (
    _r18(weights=None) if _backbone["name"] == "resnet18" else _r50(weights=None))
import models as ref  # noqa: E402  (the reference)

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def build_ref(backbone="resnet18", roi=(3, 3), img=1280, use_context=True, hidden=384, bbhd=32, n_add=0,
              drop=0.2, seed=123):
    _backbone["name"] = backbone
    m = ref.CoVA(roi, img, 4, use_context, hidden, bbhd, n_add, drop, None)
    sd = synth.make_state_dict(seed, backbone=backbone, roi_output_size=roi, hidden_dim=hidden,
                               bbox_hidden_dim=bbhd, n_additional_feat=n_add, use_context=use_context)
    m.load_state_dict(sd, strict=True)
    return m, sd


def intermediates(m, inp):
    images, bboxes, add, ci = inp
    fm = m.convnet(images)
    vis = m.roi_pool(fm, bboxes).view(bboxes.shape[0], m.n_visual_feat)
    bb = m._get_bbox_features(bboxes)
    own = torch.cat((vis, bb, m.bn_additional_feat(add)), 1)
    out = dict(fm=fm, visual=vis, bbox=bb, own=own)
    if m.use_context:
        ctx, attn = m.gat(own, ci, return_attn_wts=True)
        out.update(ctx=ctx, attn=attn)
    out["logits"] = m(images, bboxes, add, ci)
    return {k: v.detach().numpy() for k, v in out.items()}


def save(name, **arrs):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrs)
    print("%-28s %8.1f KB  %s" % (name, os.path.getsize(path) / 1024, sorted(arrs)))


@torch.no_grad()
def main():
    torch.set_num_threads(8)

    # ---- state_dict key/shape pin (checkpoint compatibility)
    for bk in ("resnet18", "resnet50"):
        _backbone["name"] = bk
        m = ref.CoVA((3, 3), 256, 4, True, 384, 32, 0, 0.2, None)
        keys = list(m.state_dict().keys())
        shapes = [tuple(v.shape) for v in m.state_dict().values()]
        assert keys == [k for k, _ in synth.state_dict_spec(backbone=bk)], bk
        assert shapes == [s for _, s in synth.state_dict_spec(backbone=bk)], bk
        save("state_dict_keys_" + bk, keys=np.array(keys), shapes=np.array([str(s) for s in shapes]))
    _backbone["name"] = "resnet18"
    m = ref.CoVA((3, 3), 256, 4, True, 384, 32, 7, 0.2, None)
    assert list(m.state_dict().keys()) == [k for k, _ in synth.state_dict_spec(n_additional_feat=7)]
    m = ref.CoVA((2, 5), 256, 4, False, 0, 0, 0, 0.2, None)
    assert list(m.state_dict().keys()) == [
        k for k, _ in synth.state_dict_spec(roi_output_size=(2, 5), use_context=False, hidden_dim=0, bbox_hidden_dim=0)]

    # ---- g_small: img 128, full feature map kept (eval)
    m, _ = build_ref(img=128)
    m.eval()
    inp = synth.gen(2, 12, 8, seed=0, img=128)
    r = intermediates(m, inp)
    save("g_small_r18_img128", **r)

    # ---- g_c1: BASELINE config 1 (1280^2, B=1, N=32, K=8); fm kept as a strided sample
    m, _ = build_ref(img=1280)
    m.eval()
    inp = synth.gen(1, 32, 8, seed=0, img=1280)
    r = intermediates(m, inp)
    r["fm_sample"] = r.pop("fm")[:, :, ::8, ::8].copy()
    save("g_c1_r18_img1280", **r)

    # ---- g_ragged: ragged pages (11 / 1 / 30 boxes), K=24, img 256
    m, _ = build_ref(img=256)
    m.eval()
    inp = synth.gen(3, 0, 24, seed=3, img=256, counts=[11, 1, 30])
    r = intermediates(m, inp)
    r["fm_sample"] = r.pop("fm")[:, :, ::4, ::4].copy()
    save("g_ragged_r18_img256", **r)

    # ---- g_align: D1 variant = line 58's module swapped for RoIAlign(P, s, 2, aligned=False)
    m, _ = build_ref(img=256)
    m.eval()
    m.roi_pool = torchvision.ops.RoIAlign((3, 3), 0.25, sampling_ratio=2, aligned=False)
    inp = synth.gen(2, 20, 8, seed=4, img=256)
    r = intermediates(m, inp)
    r.pop("fm")
    save("g_align_r18_img256", **r)

    # ---- g_r50: D2 (resnet50 truncated the same way)
    m, _ = build_ref(backbone="resnet50", img=128)
    m.eval()
    inp = synth.gen(1, 16, 8, seed=5, img=128)
    r = intermediates(m, inp)
    save("g_r50_img128", **r)
    _backbone["name"] = "resnet18"

    # ---- g_variants: no context / no bbox encoder / additional feats / other roi size
    m, _ = build_ref(img=128, roi=(2, 5), use_context=False, hidden=0, bbhd=0)
    m.eval()
    inp = synth.gen(2, 9, 0, seed=6, img=128)
    inp = (inp[0], inp[1], inp[2], torch.empty(18, 0, dtype=torch.long))
    save("g_noctx_nobbox_roi2x5", logits=m(*inp).numpy())
    m, _ = build_ref(img=128, n_add=7)
    m.eval()
    inp = list(synth.gen(2, 9, 8, seed=7, img=128))
    inp[2] = torch.randn(18, 7, generator=torch.Generator().manual_seed(7))
    save("g_addfeat7", logits=m(*inp).numpy(), additional_feats=inp[2].numpy())

    # ---- g_gat: the layer alone incl. edge cases (all -1 row, partial pads, arbitrary ids) and D3 heads
    g = torch.Generator().manual_seed(11)
    torch.manual_seed(11)      # the layers below draw their weights from the global RNG: seeded, so that the file regenerates
    layer = ref.GraphAttentionLayer(96, 64)
    h = torch.randn(40, 96, generator=g)
    ci = torch.randint(-1, 40, (40, 10), generator=g)
    ci[3] = -1
    ci[7, 4:] = -1
    out, attn = layer(h, ci, return_attn_wts=True)
    heads = [ref.GraphAttentionLayer(96, 32) for _ in range(2)]
    out2 = torch.cat([hd(h, ci) for hd in heads], 1)
    save("g_gat", h=h.numpy(), ci=ci.numpy(), out=out.numpy(), attn=attn.numpy(),
         W_i=layer.W_i.weight.numpy(), W_j=layer.W_j.weight.numpy(),
         att_w=layer.attention_layer.weight.numpy(), att_b=layer.attention_layer.bias.numpy(),
         out_2head=out2.numpy(),
         **{f"h{i}_{n}": getattr(hd, n).weight.numpy() for i, hd in enumerate(heads) for n in ("W_i", "W_j")},
         **{f"h{i}_att_w": hd.attention_layer.weight.numpy() for i, hd in enumerate(heads)},
         **{f"h{i}_att_b": hd.attention_layer.bias.numpy() for i, hd in enumerate(heads)})

    # ---- g_roi: the op the reference calls (`torchvision.ops.RoIPool`, models.py:58) on adversarial boxes
    g = torch.Generator().manual_seed(12)
    fm = torch.randn(2, 8, 40, 48, generator=g)
    n = 96
    x1 = torch.rand(n, generator=g) * 220 - 20
    y1 = torch.rand(n, generator=g) * 180 - 20
    w = torch.rand(n, generator=g) * 120
    hgt = torch.rand(n, generator=g) * 90
    w[:8] = torch.rand(8, generator=g) * 2          # sub-pixel boxes
    hgt[:8] = torch.rand(8, generator=g) * 2
    boxes = torch.stack([torch.randint(0, 2, (n,), generator=g).float(), x1, y1, x1 + w, y1 + hgt], 1)
    boxes[8:24, 1:] = (boxes[8:24, 1:] / 4).round() * 4 + 2.0     # exact .5 ties after x0.25
    boxes[24:28, 1:] += 400                                       # fully outside the map -> empty bins
    res = dict(fm=fm.numpy(), boxes=boxes.numpy())
    for P in [(1, 1), (3, 3), (7, 7), (2, 5)]:
        res["pool_%dx%d" % P] = torchvision.ops.roi_pool(fm, boxes, P, 0.25).numpy()
        res["align_%dx%d" % P] = torchvision.ops.roi_align(fm, boxes, P, 0.25, 2, False).numpy()
    save("g_roi", **res)


@torch.enable_grad()
def train_fixture():
    """Train-mode forward (batch-stat BN, dropout off) + CE(sum) loss + a few gradients
    (`train.py:45-60`, `main.py:139`)."""
    m, _ = build_ref(img=128, drop=0.0)
    m.train()
    images, bboxes, add, ci, labels = synth.gen(2, 12, 8, seed=8, img=128, with_labels=True)
    out = m(images, bboxes, add, ci)
    loss = torch.nn.CrossEntropyLoss(reduction="sum")(out, labels)
    loss.backward()
    grads = {n: p.grad.numpy() for n, p in m.named_parameters()}
    keep = ["convnet.0.weight", "convnet.4.1.conv2.weight", "convnet.4.0.bn1.weight", "gat.W_i.weight",
            "gat.W_j.weight", "gat.attention_layer.weight", "bbox_feat_encoder.0.weight", "decoder.5.weight",
            "decoder.1.bias"]
    save("g_train_r18_img128", logits=out.detach().numpy(), loss=np.float32(loss.item()), labels=labels.numpy(),
         **{"grad:" + k: grads[k] for k in keep},
         **{"buf:" + k: v.numpy() for k, v in m.state_dict().items() if "running" in k})


def tail_fixture():
    """The callers either side of the forward, run live: torch's CE(sum) + Adam (what `main.py:133-139` constructs),
    the reference's own `train.evaluate_model` (`train.py:99-171`) and `datasets.custom_collate_fn`
    (`datasets.py:141-190`).  The per-item context window is the literal loop of `datasets.py:120-127` (that method
    also opens image files, which do not exist here)."""
    import datasets as ref_ds
    import train as ref_train
    g = torch.Generator().manual_seed(11)
    # ---- CE(sum) forward / gradient, arg-max hits
    logits = (torch.randn(53, 4, generator=g) * 3).requires_grad_(True)
    labels = torch.randint(0, 4, (53,), generator=g)
    labels[5] = -100                                   # torch's default ignore_index
    loss = torch.nn.CrossEntropyLoss(reduction="sum")(logits, labels)
    loss.backward()
    ce = dict(ce_logits=logits.detach().numpy(), ce_labels=labels.numpy(), ce_loss=np.float32(loss.item()),
              ce_grad=logits.grad.numpy(), ce_correct=np.int64((logits.argmax(1) == labels).sum().item()))
    # ---- Adam(lr, weight_decay): 4 steps on a 1003-element parameter with fresh gradients each step
    p = torch.nn.Parameter(torch.randn(1003, generator=g))
    opt = torch.optim.Adam([p], lr=5e-4, weight_decay=1e-3)
    grads = torch.randn(4, 1003, generator=g)
    adam = dict(adam_p0=p.detach().numpy().copy(), adam_grads=grads.numpy())
    for i in range(4):
        p.grad = grads[i].clone()
        opt.step()
        adam["adam_p%d" % (i + 1)] = p.detach().numpy().copy()
    adam["adam_m4"] = opt.state[p]["exp_avg"].numpy().copy()
    adam["adam_v4"] = opt.state[p]["exp_avg_sq"].numpy().copy()
    # ---- custom_collate_fn on ragged pages (context window: datasets.py:120-127 verbatim)
    counts, cs = [7, 1, 12, 30, 2], 4
    items, raw = [], []
    for pi, n in enumerate(counts):
        xywh = torch.rand(n, 4, generator=g) * 200 + 1
        raw.append(xywh)
        bb = xywh.clone()
        bb[:, 2:] += bb[:, :2]                         # datasets.py:114-115
        context_indices = []
        for i in range(n):
            context = list(range(max(0, i - cs), i)) + list(range(i + 1, min(n, i + cs + 1)))
            context_indices.append(context + [-1] * (2 * cs - len(context)))
        items.append((pi, torch.zeros(3, 4, 4), bb, torch.empty(n, 0), torch.LongTensor(context_indices),
                      torch.zeros(n, dtype=torch.long)))
    _, _, cb, _, cci, _ = ref_ds.custom_collate_fn(items)
    coll = dict(coll_counts=np.asarray(counts), coll_cs=np.int64(cs), coll_xywh=torch.cat(raw).numpy(),
                coll_bboxes=cb.numpy(), coll_ctx=cci.numpy())
    # ---- evaluate_model (reference, unmodified) on preset logits: 3 batches of ragged pages, k = 1 and 3
    class Fake(torch.nn.Module):
        n_classes, class_names = 4, ["BG", "price", "title", "image"]

        def __init__(self, outs):
            super().__init__()
            self.outs, self.i = outs, 0

        def forward(self, *a):
            self.i += 1
            return self.outs[self.i - 1]
    batches, outs, ev = [], [], {}
    img = 0
    for bi, cnts in enumerate([[9, 14, 5], [33], [6, 6]]):
        bb, lab = [], []
        for pi, n in enumerate(cnts):
            l = torch.zeros(n, dtype=torch.long)
            l[torch.randperm(n, generator=g)[:3]] = torch.tensor([1, 2, 3])
            bb.append(torch.cat((torch.full((n, 1), float(pi)), torch.rand(n, 4, generator=g)), 1))
            lab.append(l)
        T = sum(cnts)
        out = torch.randn(T, 4, generator=g)
        out[::3, 1] = out[0, 1]                        # ties inside a column
        batches.append((np.arange(img, img + len(cnts)), torch.zeros(len(cnts), 3, 4, 4), torch.cat(bb), torch.empty(T, 0),
                        torch.zeros(T, 0, dtype=torch.long), torch.cat(lab)))
        outs.append(out)
        img += len(cnts)
        ev["ev_bboxes%d" % bi], ev["ev_labels%d" % bi], ev["ev_logits%d" % bi] = batches[-1][2].numpy(), batches[-1][5].numpy(), out.numpy()
    for k in (1, 3):
        ia, ca = ref_train.evaluate_model(Fake(outs), batches, "cpu", k, "VAL", "/tmp/cova_golden_log.txt")
        ev["ev_img_acc_k%d" % k], ev["ev_class_acc_k%d" % k] = ia, ca
    save("g_tail", **ce, **adam, **coll, **ev)


class _RefMultiHead(torch.nn.Module):
    """H independent reference `GraphAttentionLayer(n_feat, hidden // H)` on the same inputs, outputs
    concatenated on dim 1 (state_dict keys `gat.heads.{i}.*`)."""

    def __init__(self, n_feat, hidden, n_heads):
        super().__init__()
        self.heads = torch.nn.ModuleList(ref.GraphAttentionLayer(n_feat, hidden // n_heads) for _ in range(n_heads))

    def forward(self, h, ci, return_attn_wts=False):
        outs = [hd(h, ci, return_attn_wts) for hd in self.heads]
        if return_attn_wts:
            return torch.cat([o[0] for o in outs], 1), torch.stack([o[1] for o in outs], 1)
        return torch.cat(outs, 1)


def build_ref_multihead(backbone, img, n_heads, seed=123):
    _backbone["name"] = backbone
    m = ref.CoVA((3, 3), img, 4, True, 384, 32, 0, 0.2, None)
    m.gat = _RefMultiHead(m.n_feat, 384, n_heads)
    sd = synth.make_state_dict(seed, backbone=backbone, n_heads=n_heads)
    m.load_state_dict(sd, strict=True)
    return m, sd


@torch.no_grad()
def config_fixtures():
    """The BASELINE.json configs at their full shapes, on the very inputs `bench.py` times (rank 0: seed 1)."""
    torch.set_num_threads(8)
    # ---- g_c2: config 2 = B=16 pages of 1280^2, N=90, K=24, ResNet-18, eval (the headline workload)
    m, _ = build_ref(img=1280)
    m.eval()
    inp = synth.gen(16, 90, 24, seed=1)
    fm = m.convnet(inp[0])
    vis = m.roi_pool(fm, inp[1]).view(inp[1].shape[0], m.n_visual_feat)
    logits = m(*inp)
    save("g_c2_r18_b16", logits=logits.numpy(), fm_sample=fm[:, :, ::16, ::16].numpy().copy(),
         visual_sample=vis[::7].numpy().copy())
    del fm, vis
    # ---- g_r50_ragged: ResNet-50 at 256^2 / 320^2-class sizes with ragged pages (multi-strip stem, multi-tile pw_tc)
    m, _ = build_ref(backbone="resnet50", img=320)
    m.eval()
    inp = synth.gen(3, 0, 24, seed=14, img=320, counts=[17, 2, 40])
    r = intermediates(m, inp)
    r["fm_sample"] = r.pop("fm")[:, :, ::4, ::4].copy()
    for k in ("visual", "own", "ctx", "attn", "bbox"):
        r.pop(k, None)
    save("g_r50_ragged_img320", **r)
    # ---- g_c5: config 5 shape = ResNet-50, N=300, K=48, 2-head GAT, 1280^2 (2 pages)
    m, _ = build_ref_multihead("resnet50", 1280, 2)
    m.eval()
    inp = synth.gen(2, 300, 48, seed=1)
    fm = m.convnet(inp[0])
    logits = m(*inp)
    save("g_c5_r50_n300_k48_h2", logits=logits.numpy(), fm_sample=fm[:, :, ::16, ::16].numpy().copy())
    _backbone["name"] = "resnet18"


@torch.enable_grad()
def train_fixture_r50():
    """Config 3/4 semantics at a CPU-sized shape: ResNet-50 backbone, train mode (batch-statistics BatchNorm, dropout
    off), CE(sum), every backbone gradient + a few of the head's (`train.py:45-60`, `main.py:139`).

    Conditioning: a ReLU / max-pool / RoIPool decision on an element that sits within rounding of a tie flips between
    two correct implementations (about one element per million per 1e-6 of noise), and when that element carries a large
    gradient every parameter gradient upstream moves by ~1 % (measured: tools/diag_train_grads.py, profiles/r02_*).
    The fixture therefore uses the first input seed for which the reference's OWN float32 and float64 gradients agree to
    2e-5 on every tensor - a sample with no decision inside fp32 noise - and records that deviation."""
    torch.set_num_threads(8)
    for seed in range(15, 60):
        m, _ = build_ref(backbone="resnet50", img=192, drop=0.0)
        m.train()
        images, bboxes, add, ci, labels = synth.gen(2, 14, 8, seed=seed, img=192, with_labels=True)
        out = m(images, bboxes, add, ci)
        loss = torch.nn.CrossEntropyLoss(reduction="sum")(out, labels)
        loss.backward()
        grads = {n: p.grad.numpy() for n, p in m.named_parameters()}
        bufs = {"buf:" + k: v.numpy().copy() for k, v in m.state_dict().items() if "running" in k}
        m64, _ = build_ref(backbone="resnet50", img=192, drop=0.0)
        m64 = m64.double().train()
        out64 = m64(images.double(), bboxes.double(), add.double(), ci)
        torch.nn.CrossEntropyLoss(reduction="sum")(out64, labels).backward()
        dev = max(float(np.abs(grads[n] - p.grad.numpy()).max() / max(np.abs(p.grad.numpy()).max(), 1e-30))
                  for n, p in m64.named_parameters() if n.startswith("convnet."))
        print("train_fixture_r50: seed %d, fp32-vs-fp64 reference gradients (backbone, max over tensors) %.2e" % (seed, dev))
        if dev < 2e-5:
            break
    else:
        raise RuntimeError("no well-conditioned seed found")
    keep = [k for k in grads if k.startswith("convnet.")] + ["gat.W_j.weight", "gat.attention_layer.weight",
                                                             "bbox_feat_encoder.0.weight", "decoder.5.weight"]
    save("g_train_r50_img192", logits=out.detach().numpy(), loss=np.float32(loss.item()), labels=labels.numpy(),
         seed=np.int64(seed), ref_fp32_vs_fp64=np.float64(dev),
         **{"grad:" + k: (grads[k] if grads[k].size < 400000 else grads[k][:, ::8].copy()) for k in keep}, **bufs)
    _backbone["name"] = "resnet18"


if __name__ == "__main__":
    which = sys.argv[1:] or ["main", "train", "tail", "configs", "train_r50"]
    if "main" in which:
        main()
    if "train" in which:
        train_fixture()
    if "tail" in which:
        tail_fixture()
    if "configs" in which:
        config_fixtures()
    if "train_r50" in which:
        train_fixture_r50()
