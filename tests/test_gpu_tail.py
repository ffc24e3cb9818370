"""GPU parity of the callers either side of the forward, through the C ABI, against
the oracle and the fixtures frozen from the live reference (`tests/golden/g_tail.npz`: torch's CE(sum) and Adam, the
reference's own `evaluate_model` and `custom_collate_fn`).  Tolerances: integer / index outputs bit-exact; CE loss and
gradient 2e-6; Adam parameters 5e-7 absolute after 4 steps."""
import warnings

import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err
from oracle import cova_oracle as O

pytestmark = pytest.mark.gpu
warnings.filterwarnings("ignore")
DEV = "cuda:0"


def ops():
    from cova_b200 import ops as _ops
    return _ops


def test_ce_sum_kernel_vs_torch_fixture():
    g = load_golden("g_tail")
    loss, dl, nc = ops().ce_sum_fwd_bwd(torch.from_numpy(g["ce_logits"]).to(DEV), torch.from_numpy(g["ce_labels"]).to(DEV),
                                        want_grad=True, want_correct=True)
    assert abs(float(loss.item()) - float(g["ce_loss"])) <= 2e-6 * abs(float(g["ce_loss"]))
    assert np.abs(dl.cpu().numpy() - g["ce_grad"]).max() < 2e-6
    assert int(nc.item()) == int(g["ce_correct"])


@pytest.mark.parametrize("T", [0, 1, 255, 257, 5760])
def test_ce_sum_kernel_vs_oracle(T):
    gen = torch.Generator().manual_seed(T)
    logits = torch.randn(T, 4, generator=gen) * 4
    labels = torch.randint(0, 4, (T,), generator=gen)
    wide = torch.zeros(T, 6)
    wide[:, :4] = logits                                  # strided rows
    loss, dl, nc = ops().ce_sum_fwd_bwd(wide.to(DEV)[:, :4], labels.to(DEV), want_grad=True, want_correct=True)
    want_loss, want_d, want_nc = O.ce_sum(logits.numpy(), labels.numpy())
    assert abs(float(loss.item()) - float(want_loss)) <= 3e-6 * max(abs(float(want_loss)), 1.0)
    if T:
        assert np.abs(dl.cpu().numpy() - want_d).max() < 2e-6
    assert int(nc.item()) == want_nc


def test_criterion_module_backward_matches_torch():
    from cova_b200.train_ops import CrossEntropyLossSum
    gen = torch.Generator().manual_seed(3)
    x = (torch.randn(300, 4, generator=gen) * 2).to(DEV).requires_grad_(True)
    w = torch.randn(4, 4, generator=gen).to(DEV).requires_grad_(True)
    y = torch.randint(0, 4, (300,), generator=gen).to(DEV)
    crit = CrossEntropyLossSum()
    (crit(x @ w, y) * 0.5).backward()
    gx, gw = x.grad.clone(), w.grad.clone()
    x.grad = w.grad = None
    (torch.nn.CrossEntropyLoss(reduction="sum")(x @ w, y) * 0.5).backward()
    assert rel_err(gx.cpu().numpy(), x.grad.cpu().numpy()) < 1e-5
    assert rel_err(gw.cpu().numpy(), w.grad.cpu().numpy()) < 1e-5
    assert int(crit.n_correct.item()) == int(((x @ w).argmax(1) == y).sum().item())


def test_adam_kernel_vs_torch_fixture():
    g = load_golden("g_tail")
    p = torch.from_numpy(g["adam_p0"].copy()).to(DEV)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for i in range(4):
        ops().adam_step(p, torch.from_numpy(g["adam_grads"][i]).to(DEV), m, v, 5e-4, 0.9, 0.999, 1e-8, 1e-3, i + 1)
        assert np.abs(p.cpu().numpy() - g["adam_p%d" % (i + 1)]).max() < 5e-7
    assert rel_err(m.cpu().numpy(), g["adam_m4"]) < 1e-6 and rel_err(v.cpu().numpy(), g["adam_v4"]) < 1e-6


def test_flat_adam_equals_torch_adam_on_model():
    """FlatAdam vs torch.optim.Adam on two copies of the real model: same gradients in, same parameters out, the
    optimizer state_dict has torch's layout, and the native inference caches notice the in-place update."""
    import copy
    import cova_b200.synth as synth
    from cova_b200.models import CoVA
    from cova_b200.train_ops import FlatAdam
    m1 = CoVA((3, 3), 128, 4, True, 384, 32, 0, 0.0, None, pretrained=False)
    m1.load_state_dict(synth.make_state_dict(123), strict=True)
    m1 = m1.to(DEV)
    m2 = copy.deepcopy(m1)
    o1 = FlatAdam(m1.parameters(), lr=5e-4, weight_decay=1e-3)
    o2 = torch.optim.Adam(m2.parameters(), lr=5e-4, weight_decay=1e-3)
    gen = torch.Generator().manual_seed(5)
    inp = [t.to(DEV) for t in synth.gen(2, 12, 8, seed=2, img=128)]
    m1.eval(); m2.eval()
    with torch.no_grad():
        before = m1(*inp).clone()
    for _ in range(3):
        o1.zero_grad(); o2.zero_grad()
        for p1, p2 in zip(m1.parameters(), m2.parameters()):
            gr = torch.randn(p1.shape, generator=gen).to(DEV) * 0.1
            p1.grad.copy_(gr)
            p2.grad = gr.clone()
        o1.step(); o2.step()
    for (n, p1), p2 in zip(m1.named_parameters(), m2.parameters()):
        assert (p1 - p2).abs().max().item() < 1e-6, n
    sd1, sd2 = o1.state_dict(), o2.state_dict()
    assert sd1["state"].keys() == sd2["state"].keys()
    assert set(sd1["state"][0].keys()) == set(sd2["state"][0].keys()) == {"step", "exp_avg", "exp_avg_sq"}
    with torch.no_grad():
        after1, after2 = m1(*inp), m2(*inp)
    assert (after1 - before).abs().max().item() > 1e-4          # the native caches were rebuilt after the raw-pointer update
    assert rel_err(after1.cpu().numpy(), after2.cpu().numpy()) < 1e-4


def test_build_batch_kernel_vs_collate_fixture():
    from cova_b200.train_ops import assemble_batch
    g = load_golden("g_tail")
    bb, ctx, off = assemble_batch(list(g["coll_counts"]), int(g["coll_cs"]), DEV, torch.from_numpy(g["coll_xywh"]))
    assert np.array_equal(ctx.cpu().numpy(), g["coll_ctx"])
    assert np.array_equal(bb.cpu().numpy(), g["coll_bboxes"])
    assert off.cpu().tolist() == [0] + list(np.cumsum(g["coll_counts"]))


@pytest.mark.parametrize("counts,cs", [([1], 12), ([230, 11, 90], 12), ([3, 3, 3, 3], 1), ([300] * 16, 24), ([5, 9], 0)])
def test_build_batch_kernel_vs_oracle(counts, cs):
    from cova_b200.train_ops import assemble_batch
    _, ctx, _ = assemble_batch(counts, cs, DEV)
    _, want = O.build_batch(counts, cs)
    assert ctx.shape == want.shape and np.array_equal(ctx.cpu().numpy(), want)


def test_topk_hits_kernel_vs_reference_evaluate_fixture():
    from cova_b200.train_ops import page_offsets_of
    g = load_golden("g_tail")
    for k in (1, 3):
        rows = []
        for bi in range(3):
            off = page_offsets_of(torch.from_numpy(g["ev_bboxes%d" % bi]).to(DEV))
            rows.append(ops().topk_hits(torch.from_numpy(g["ev_logits%d" % bi]).to(DEV),
                                        torch.from_numpy(g["ev_labels%d" % bi]).to(DEV), off, k).cpu().numpy())
        assert np.array_equal(np.concatenate(rows)[:, 1:], g["ev_img_acc_k%d" % k][:, 1:])


def test_topk_hits_kernel_vs_oracle_large_and_absent_class():
    gen = torch.Generator().manual_seed(9)
    counts = [230, 11, 90, 1, 64]
    off = np.concatenate(([0], np.cumsum(counts))).astype(np.int32)
    T = int(off[-1])
    logits = torch.randn(T, 4, generator=gen).round(decimals=1)       # many ties
    labels = torch.zeros(T, dtype=torch.long)
    for b in range(len(counts) - 2):
        rows = off[b] + torch.randperm(counts[b], generator=gen)[:3].numpy()
        labels[rows] = torch.tensor([1, 2, 3])
    for k in (1, 2, 5):
        got = ops().topk_hits(logits.to(DEV), labels.to(DEV), torch.from_numpy(off).to(DEV), k).cpu().numpy()
        want = O.topk_hits(logits.numpy(), labels.numpy(), off, k)
        assert np.array_equal(got[:, 1:], want[:, 1:])
        assert (got[3:, 1:] == -1).all()


def test_evaluate_model_drop_in_matches_reference_fixture(tmp_path):
    """`cova_b200.train_ops.evaluate_model` on the loader / logits the reference's `evaluate_model` was frozen on."""
    from cova_b200.train_ops import evaluate_model
    g = load_golden("g_tail")

    class Fake(torch.nn.Module):
        n_classes, class_names = 4, ["BG", "price", "title", "image"]

        def __init__(self):
            super().__init__()
            self.i = 0

        def forward(self, *a):
            self.i += 1
            return torch.from_numpy(g["ev_logits%d" % ((self.i - 1) % 3)]).to(DEV)
    loader, img = [], 0
    for bi in range(3):
        bb = torch.from_numpy(g["ev_bboxes%d" % bi])
        n_pages = int(bb[:, 0].max().item()) + 1
        T = bb.shape[0]
        loader.append((np.arange(img, img + n_pages), torch.zeros(n_pages, 3, 4, 4), bb, torch.empty(T, 0),
                       torch.zeros(T, 0, dtype=torch.long), torch.from_numpy(g["ev_labels%d" % bi])))
        img += n_pages
    for k in (1, 3):
        ia, ca = evaluate_model(Fake(), loader, DEV, k, "VAL", str(tmp_path / "log.txt"))
        assert np.array_equal(ia, g["ev_img_acc_k%d" % k])
        assert np.allclose(ca, g["ev_class_acc_k%d" % k])
