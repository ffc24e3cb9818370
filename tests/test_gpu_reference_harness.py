"""The reference's OWN harness, unmodified, driving this repo's model on the GPU (SURVEY.md section 4 tier 3, section 8(b)):
`train.train_model` (`/root/reference/train.py:9-96`, incl. the best-checkpoint `torch.save` / `load_state_dict` of
`:84,94`), `train.evaluate_model` (`:99-171`), `evaluate.evaluate` (`evaluate.py:14-84`) and `datasets.custom_collate_fn`
(`datasets.py:141-190`) are imported from `oracle/_ref/` (staged from /root/reference by `oracle/build_ref.py`; see
`oracle/ref_loader.py`), with `models` resolved to the INTEGRATION.md stub (`from cova_b200.models import *`).
The same harness is then run on the reference's own `models.CoVA` (torch / cuDNN / torchvision on the same GPU) from the
same seeded weights and data: both runs must log the same loss / accuracy trajectory and end with the same metrics."""
import os
import re
import types
import warnings

import numpy as np
import pytest
import torch

import cova_b200.synth as synth
from oracle import ref_loader

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_loader.available(), reason="oracle/_ref not staged")]
warnings.filterwarnings("ignore")
DEV = "cuda:0"
IMG, CS = 128, 4


def _items(ref, n_pages, seed):
    """Per-page items in the layout `WebDataset.__getitem__` returns (`datasets.py:86-135`): a learnable toy task -
    the three labelled boxes of a page are drawn over bright / dark / striped patches."""
    g = torch.Generator().manual_seed(seed)
    items = []
    for pid in range(n_pages):
        n = int(torch.randint(9, 16, (1,), generator=g))
        img = torch.rand(3, IMG, IMG, generator=g) * 0.2 + 0.4
        w = 12 + torch.rand(n, generator=g) * 30
        h = 10 + torch.rand(n, generator=g) * 24
        x1 = torch.rand(n, generator=g) * (IMG - w)
        y1 = torch.rand(n, generator=g) * (IMG - h)
        bb = torch.stack([x1, y1, x1 + w, y1 + h], 1)
        labels = torch.zeros(n, dtype=torch.long)
        pos = torch.randperm(n, generator=g)[:3]
        for c, p in enumerate(pos):
            labels[p] = c + 1
            xa, ya, xb, yb = [int(v) for v in bb[p]]
            img[:, ya:yb, xa:xb] = [1.0, 0.0, 0.7][c]
            if c == 2:
                img[:, ya:yb:2, xa:xb] = 0.1
        ci = torch.from_numpy(synth.context_window(n, CS))
        items.append((pid + 100 * seed, img, bb, torch.empty(n, 0), ci, labels))
    return items


def _loaders(ref, seed):
    tr = _items(ref, 8, seed)
    va = _items(ref, 4, seed + 1)
    coll = ref.datasets.custom_collate_fn
    clone = lambda its: [tuple(t.clone() if isinstance(t, torch.Tensor) else t for t in it) for it in its]   # collate edits ci in place
    train_loader = [coll(clone(tr[i:i + 4])) for i in range(0, 8, 4)]
    val_loader = [coll(clone(va[i:i + 2])) for i in range(0, 4, 2)]
    return train_loader, val_loader


def _run(ref, model, tmp, tag):
    """`main.py:133-164` with the reference's own functions."""
    opt = torch.optim.Adam(model.parameters(), lr=5e-4, weight_decay=1e-3)                 # main.py:133-135
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=100, gamma=1)                   # main.py:136
    crit = torch.nn.CrossEntropyLoss(reduction="sum").to(DEV)                              # main.py:139
    log, ckpt = str(tmp / (tag + "_log.txt")), str(tmp / (tag + "_ckpt.pth"))
    tr, va = _loaders(ref, 1)
    best = ref.train.train_model(model, tr, opt, sched, crit, 3, DEV, va, 1, log, ckpt)    # main.py:141-153
    csv = str(tmp / (tag + "_img_acc.csv"))
    class_acc, macro = ref.evaluate.evaluate(model, va, DEV, log, csv)                     # main.py:155-164
    txt = open(log).read()
    losses = [float(x) for x in re.findall(r"Loss: ([0-9.]+)", txt)]
    accs = [float(x) for x in re.findall(r"Accuracy: ([0-9.]+)%", txt)]
    return dict(best=best, class_acc=class_acc, losses=losses, accs=accs, ckpt=ckpt, csv=open(csv).read())


def test_reference_harness_unmodified_on_native_model(tmp_path):
    import cova_b200.models as native_models
    stub = types.ModuleType("models")                      # the one-line models.py
    stub.CoVA, stub.GraphAttentionLayer = native_models.CoVA, native_models.GraphAttentionLayer
    ref = ref_loader.load(models_module=stub)
    assert ref.evaluate.CoVA is native_models.CoVA         # evaluate.py:9 bound this repo's class
    sd = synth.make_state_dict(123)
    torch.manual_seed(0)
    ours = ref.evaluate.CoVA((3, 3), IMG, 4, True, 384, 32, 0, 0.0, ref.constants.Constants.CLASS_NAMES).to(DEV)
    ours.load_state_dict(sd, strict=True)
    got = _run(ref, ours, tmp_path, "native")

    ref_own = ref_loader.load()                            # the reference's own models.py, same harness
    theirs = ref_own.models.CoVA((3, 3), IMG, 4, True, 384, 32, 0, 0.0, ref.constants.Constants.CLASS_NAMES)
    theirs.load_state_dict(sd, strict=True)
    theirs = theirs.to(DEV)
    with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        try:
            want = _run(ref_own, theirs, tmp_path, "reference")
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev

    assert len(got["losses"]) == len(want["losses"]) == 3
    assert got["losses"][-1] < got["losses"][0]                                            # it learns
    for a, b in zip(got["losses"], want["losses"]):
        assert abs(a - b) <= 2e-3 * abs(b) + 2e-4, (got["losses"], want["losses"])
    assert got["accs"] == want["accs"]
    assert np.allclose(got["class_acc"], want["class_acc"]) and got["csv"] == want["csv"]
    # the checkpoint the reference's loop wrote from our model loads into the reference's model (and back)
    ck = torch.load(got["ckpt"], map_location="cpu")
    theirs.load_state_dict(ck, strict=True)
    ours.load_state_dict(torch.load(want["ckpt"], map_location="cpu"), strict=True)
    # after the restore of train.py:94 the native inference path serves the restored weights
    ours.eval()
    b = _loaders(ref, 1)[1][0]
    with torch.no_grad(), torch.backends.cudnn.flags(enabled=True, allow_tf32=False):      # the reference in plain fp32
        a = ours(b[1].to(DEV), b[2].to(DEV), b[3].to(DEV), b[4].to(DEV))
        theirs.load_state_dict(torch.load(want["ckpt"], map_location="cpu"), strict=True)
        c = theirs.eval()(b[1].to(DEV), b[2].to(DEV), b[3].to(DEV), b[4].to(DEV))
    assert float((a - c).abs().max()) < 1e-4 * float(c.abs().max())
