"""GPU: the reference's training / evaluation harness semantics against the drop-in model.

`/root/reference` does not exist on the GPU box, so this restates - call for call - what `train.py:train_model`
(:9-96) and `train.py:evaluate_model` (:99-171) do with the model: `.to(device)`, `.train()` / `.eval()`, Adam +
StepLR(gamma=1) + CrossEntropyLoss(reduction="sum") (`main.py:133-139`), `loss.backward()`, per-image top-k accuracy
from raw logits via `argsort(dim=0)`, `state_dict()` save / `load_state_dict()` restore of the best epoch."""
import copy
import warnings

import numpy as np
import pytest
import torch

import cova_b200.synth as synth

pytestmark = pytest.mark.gpu
warnings.filterwarnings("ignore")
DEV = torch.device("cuda:0")


def make_loader(n_batches, B, N, K, img, seed):
    out = []
    for i in range(n_batches):
        images, bboxes, add, ci, labels = synth.gen(B, N, K, seed=seed + i, img=img, with_labels=True)
        out.append((np.array([f"img{i}_{j}" for j in range(B)]), images, bboxes, add, ci, labels))
    return out


@torch.no_grad()
def evaluate_model(model, loader, k=1):                      # train.py:99-171
    model.eval()
    n_classes = model.n_classes
    img_acc = []
    for _, images, bboxes, add, ci, labels in loader:
        labels = labels.to(DEV)
        output = model(images.to(DEV), bboxes.to(DEV), add.to(DEV), ci.to(DEV))      # native kernels (no_grad + eval)
        batch_indices = torch.unique(bboxes[:, 0]).long()
        for index in batch_indices:                                                  # train.py:131-154
            img_indices = (bboxes[:, 0] == index).to(DEV)
            labels_img, output_img = labels[img_indices].view(-1, 1), output[img_indices]
            indexes = torch.arange(labels_img.shape[0], device=DEV).view(-1, 1)
            top_k = torch.argsort(output_img, dim=0)[output_img.shape[0] - k:]
            correct = [float(indexes[labels_img == c].view(-1)[0] in top_k[:, c]) if (labels_img == c).any() else 1.0
                       for c in range(1, n_classes)]
            img_acc.append(correct)
    return np.array(img_acc).mean(0) * 100


def test_train_eval_cycle_like_reference_harness():
    from cova_b200.models import CoVA
    torch.manual_seed(123)
    model = CoVA((3, 3), 256, 4, True, 384, 32, 0, 0.2, ["BG", "Price", "Title", "Image"], pretrained=False).to(DEV)
    assert model.n_classes == 4 and list(model.class_names)[1] == "Price"
    train_loader = make_loader(3, 4, 20, 8, 256, seed=100)
    val_loader = make_loader(1, 4, 20, 8, 256, seed=200)
    optimizer = torch.optim.Adam(model.parameters(), lr=5e-4, weight_decay=1e-3)       # main.py:133-135
    scheduler = torch.optim.lr_scheduler.StepLR(optimizer, step_size=100, gamma=1)     # main.py:136-138
    criterion = torch.nn.CrossEntropyLoss(reduction="sum").to(DEV)                     # main.py:139

    losses, best, best_sd = [], -1.0, None
    for epoch in range(4):
        model.train()                                                                  # train.py:27
        epoch_loss = 0.0
        for _, images, bboxes, add, ci, labels in train_loader:
            labels = labels.to(DEV)
            optimizer.zero_grad()
            output = model(images.to(DEV), bboxes.to(DEV), add.to(DEV), ci.to(DEV))    # train.py:47-52 (autograd path)
            assert output.shape == (labels.shape[0], 4) and output.requires_grad
            predictions = output.argmax(dim=1)
            _ = (predictions == labels).sum().item()
            loss = criterion(output, labels)
            epoch_loss += loss.item()
            loss.backward()
            optimizer.step()
        scheduler.step()
        losses.append(epoch_loss)
        acc = evaluate_model(model, val_loader).mean()                                 # train.py:72-78
        if acc >= best:
            best, best_sd = acc, copy.deepcopy(model.state_dict())                     # train.py:84 (torch.save)
    assert losses[-1] < losses[0], losses
    assert all(p.grad is not None for p in model.parameters())

    # restore the best checkpoint into a FRESH model (train.py:94, evaluate.py:198) and reproduce its logits exactly
    fresh = CoVA((3, 3), 256, 4, True, 384, 32, 0, 0.2, None, pretrained=False).to(DEV)
    fresh.load_state_dict(best_sd)
    model.load_state_dict(best_sd)
    model.eval(); fresh.eval()
    _, images, bboxes, add, ci, _ = val_loader[0]
    with torch.no_grad():
        a = model(images.to(DEV), bboxes.to(DEV), add.to(DEV), ci.to(DEV))
        b = fresh(images.to(DEV), bboxes.to(DEV), add.to(DEV), ci.to(DEV))
    assert torch.equal(a, b)
    # eval-mode native logits agree with the eval-mode autograd composite (same weights, running statistics)
    with torch.enable_grad():
        c = model(images.to(DEV), bboxes.to(DEV), add.to(DEV), ci.to(DEV))
    assert float((a - c).abs().max()) < 1e-3 * float(c.abs().max())


def test_native_tail_trains_like_torch_tail():
    """The same `train_model` loop with the native callers swapped in (`CrossEntropyLossSum` for the criterion of
    `main.py:139`, `FlatAdam` for the optimizer of `main.py:133-135`, the segmented top-k `evaluate_model`): with dropout
    off, both runs see identical batches and must follow the same loss trajectory and end at the same parameters."""
    from cova_b200.models import CoVA
    from cova_b200.train_ops import CrossEntropyLossSum, FlatAdam, evaluate_model as native_eval
    train_loader = make_loader(2, 3, 16, 8, 128, seed=300)
    val_loader = [(np.arange(3),) + b[1:] for b in make_loader(1, 3, 16, 8, 128, seed=400)]
    runs = {}
    for tail in ("torch", "native"):
        model = CoVA((3, 3), 128, 4, True, 384, 32, 0, 0.0, ["BG", "Price", "Title", "Image"], pretrained=False)
        model.load_state_dict(synth.make_state_dict(123), strict=True)
        model = model.to(DEV)
        if tail == "torch":
            optimizer = torch.optim.Adam(model.parameters(), lr=5e-4, weight_decay=1e-3)
            criterion = torch.nn.CrossEntropyLoss(reduction="sum").to(DEV)
        else:
            optimizer = FlatAdam(model.parameters(), lr=5e-4, weight_decay=1e-3)
            criterion = CrossEntropyLossSum().to(DEV)
        losses, correct = [], []
        for epoch in range(3):
            model.train()
            for _, images, bboxes, add, ci, labels in train_loader:
                labels = labels.to(DEV)
                optimizer.zero_grad()
                output = model(images.to(DEV), bboxes.to(DEV), add.to(DEV), ci.to(DEV))
                loss = criterion(output, labels)
                losses.append(loss.item())
                correct.append(int(criterion.n_correct.item()) if tail == "native"
                               else int((output.argmax(1) == labels).sum().item()))
                loss.backward()
                optimizer.step()
        if tail == "native":
            img_acc, class_acc = native_eval(model, val_loader, DEV, 1, "VAL", "/tmp/cova_b200_test_log.txt")
            assert img_acc.shape == (3, 4) and class_acc.shape == (4,)
            assert np.allclose(class_acc[1:], evaluate_model(model, [(None,) + b[1:] for b in val_loader]))
        runs[tail] = (losses, correct, {n: p.detach().clone() for n, p in model.named_parameters()})
    lt, ln = np.array(runs["torch"][0]), np.array(runs["native"][0])
    assert np.abs(lt - ln).max() < 2e-3 * np.abs(lt).max(), (lt, ln)
    assert runs["torch"][1][0] == runs["native"][1][0]
    for n, p in runs["torch"][2].items():
        q = runs["native"][2][n]
        # Adam normalises every element's step to ~lr, so an element whose gradient is pure accumulation-order noise
        # (atomics in the RoIPool / GAT backward) may walk the other way in the two runs: bound the FRACTION of such
        # elements and the mean drift instead of the maximum (6 steps of lr 5e-4 move a weight by <= 3e-3).
        d = (p - q).abs()
        assert float((d > 2e-4 + 5e-3 * float(p.abs().max())).float().mean()) < 2e-2, n   # measured up to 3.4e-3
        assert float(d.mean()) < 3e-4 + 1e-2 * float(p.abs().mean()), n     # measured 5e-5 on conv1 (non-deterministic library wgrad)
        # (exact optimizer / criterion equality is pinned separately: test_gpu_tail.py::test_flat_adam_equals_torch_adam_on_model)


def test_batched_attention_export_matches_page_at_a_time(tmp_path):
    """N4: `cova_b200.attn_export` on a batch of 3 ragged pages writes, per page, the csv the reference script
    (`extract_attn_wts_and_visualize.py:89-150`, restated here page at a time with the oracle's GAT) would write."""
    from cova_b200.attn_export import export_attention
    from cova_b200.models import CoVA
    from oracle import cova_oracle as O
    model = CoVA((3, 3), 128, 4, True, 384, 32, 0, 0.2, None, pretrained=False)
    sd = synth.make_state_dict(123)
    model.load_state_dict(sd, strict=True)
    model = model.to(DEV).eval()
    counts = [14, 5, 23]
    images, bboxes, add, ci, labels = synth.gen(3, 0, 8, seed=21, img=128, counts=counts, with_labels=True)
    files = export_attention(model, [(["a", "b", "c"], images, bboxes, add, ci, labels)], DEV, str(tmp_path))
    nsd = {k: v.numpy() for k, v in sd.items()}
    r0 = 0
    for pi, n in enumerate(counts):                     # the reference loop: batch of ONE page, page-local ids
        bb = bboxes[r0:r0 + n].clone(); bb[:, 0] = 0
        cl = ci[r0:r0 + n].clone(); cl[cl >= 0] -= r0
        lab = labels[r0:r0 + n].numpy()
        out = O.cova_forward(nsd, images[pi:pi + 1].numpy(), bb.numpy(), add[r0:r0 + n].numpy(), cl.numpy(),
                             return_intermediates=True)
        coords = bb[:, 1:].numpy().copy(); coords[:, 2:] -= coords[:, :2]
        padded = np.concatenate((coords, np.zeros((1, 4), np.float32)))
        ctx_coords = padded[cl.numpy().reshape(-1)].reshape(n, -1)
        _, attn = O.gat(out["own"], cl.numpy(), nsd["gat.W_i.weight"], nsd["gat.W_j.weight"],
                        nsd["gat.attention_layer.weight"], nsd["gat.attention_layer.bias"], return_attn_wts=True)
        want = np.concatenate((coords, lab.reshape(-1, 1).astype(np.float32), ctx_coords, attn), 1)[lab > 0]
        got = np.loadtxt(files[pi], delimiter=",", ndmin=2)
        assert got.shape == want.shape == (3, 5 + 8 * 4 + 8)
        assert np.abs(got - want).max() <= 1.01e-3       # both rounded to 3 decimals by the csv format
        r0 += n
