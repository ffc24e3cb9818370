"""CPU: the oracle (numpy restatement) against fixtures frozen from the live reference
(`oracle/make_golden.py`).  This is the pin that lets the GPU parity tests trust the oracle."""
import numpy as np
import pytest
import torch

import cova_b200.synth as synth
from conftest import load_golden, rel_err
from oracle import cova_oracle as O


def np_sd(**cfg):
    return {k: v.numpy() for k, v in synth.make_state_dict(123, **cfg).items()}


def np_inp(*a, **k):
    return [t.numpy() for t in synth.gen(*a, **k)]


def test_state_dict_spec_matches_reference_keys():
    for bk in ("resnet18", "resnet50"):
        g = load_golden("state_dict_keys_" + bk)
        spec = synth.state_dict_spec(backbone=bk)
        assert [k for k, _ in spec] == list(g["keys"])
        assert [str(tuple(s)) for _, s in spec] == list(g["shapes"])


def test_full_forward_small_r18():
    g = load_golden("g_small_r18_img128")
    r = O.cova_forward(np_sd(), *np_inp(2, 12, 8, seed=0, img=128), return_intermediates=True)
    assert rel_err(r["fm"], g["fm"]) < 2e-6
    assert rel_err(r["visual"], g["visual"]) < 2e-6
    assert rel_err(r["own"], g["own"]) < 2e-6
    assert rel_err(r["ctx"], g["ctx"]) < 5e-6
    assert rel_err(r["logits"], g["logits"]) < 5e-6


def test_roi_pool_bit_exact_on_reference_fm():
    """Given the reference's own feature map, RoIPool must reproduce its visual features exactly."""
    g = load_golden("g_small_r18_img128")
    _, bboxes, _, _ = np_inp(2, 12, 8, seed=0, img=128)
    v = O.roi_pool(g["fm"], bboxes, (3, 3), 0.25).reshape(len(bboxes), -1)
    assert np.array_equal(v, g["visual"])


@pytest.mark.parametrize("P", [(1, 1), (3, 3), (7, 7), (2, 5)])
def test_roi_pool_adversarial_boxes_bit_exact(P):
    g = load_golden("g_roi")
    out = O.roi_pool(g["fm"], g["boxes"], P, 0.25)
    assert np.array_equal(out, g["pool_%dx%d" % P])


@pytest.mark.parametrize("P", [(1, 1), (3, 3), (7, 7), (2, 5)])
def test_roi_align_adversarial_boxes(P):
    g = load_golden("g_roi")
    out = O.roi_align(g["fm"], g["boxes"], P, 0.25, 2, False)
    assert np.abs(out - g["align_%dx%d" % P]).max() < 2e-6


def test_ragged_pages_and_c1_tail_from_reference_fm():
    """Everything after the feature map, on ragged pages (11/1/30 boxes, K=24) and on BASELINE
    config 1 (1280^2, N=32, K=8) - driven from the reference's `own` so no 1280^2 conv runs on CPU."""
    for name, cfgen in (("g_ragged_r18_img256", dict(B=3, N=0, K=24, seed=3, img=256, counts=[11, 1, 30])),
                        ("g_c1_r18_img1280", dict(B=1, N=32, K=8, seed=0, img=64))):
        g = load_golden(name)
        sd = np_sd()
        ci = synth.gen(**cfgen)[3].numpy()
        ctx, attn = O.gat(g["own"], ci, sd["gat.W_i.weight"], sd["gat.W_j.weight"],
                          sd["gat.attention_layer.weight"], sd["gat.attention_layer.bias"], return_attn_wts=True)
        assert rel_err(ctx, g["ctx"]) < 5e-6 and np.abs(attn - g["attn"]).max() < 2e-6
        logits = O.decoder(np.concatenate([g["own"], ctx], 1), sd)
        assert rel_err(logits, g["logits"]) < 5e-6


def test_bbox_encoder():
    g = load_golden("g_ragged_r18_img256")
    bboxes = synth.gen(3, 0, 24, seed=3, img=256, counts=[11, 1, 30])[1].numpy()
    assert rel_err(O.bbox_encoder(bboxes, np_sd()), g["bbox"]) < 2e-6


def test_roi_align_variant_forward():
    g = load_golden("g_align_r18_img256")
    r = O.cova_forward(np_sd(), *np_inp(2, 20, 8, seed=4, img=256), roi_mode="align", return_intermediates=True)
    assert rel_err(r["visual"], g["visual"]) < 5e-6
    assert rel_err(r["logits"], g["logits"]) < 5e-6


def test_resnet50_variant():
    g = load_golden("g_r50_img128")
    r = O.cova_forward(np_sd(backbone="resnet50"), *np_inp(1, 16, 8, seed=5, img=128), return_intermediates=True)
    assert rel_err(r["fm"], g["fm"]) < 5e-6
    assert rel_err(r["logits"], g["logits"]) < 1e-5


def test_constructor_variants():
    g = load_golden("g_noctx_nobbox_roi2x5")
    sd = np_sd(roi_output_size=(2, 5), use_context=False, hidden_dim=0, bbox_hidden_dim=0)
    images, bboxes, add, _ = np_inp(2, 9, 0, seed=6, img=128)
    out = O.cova_forward(sd, images, bboxes, add, np.empty((18, 0), np.int64), roi_output_size=(2, 5))
    assert rel_err(out, g["logits"]) < 5e-6
    g = load_golden("g_addfeat7")
    images, bboxes, _, ci = np_inp(2, 9, 8, seed=7, img=128)
    out = O.cova_forward(np_sd(n_additional_feat=7), images, bboxes, g["additional_feats"], ci)
    assert rel_err(out, g["logits"]) < 5e-6


def test_gat_layer_edge_cases_and_heads():
    g = load_golden("g_gat")
    out, attn = O.gat(g["h"], g["ci"], g["W_i"], g["W_j"], g["att_w"], g["att_b"], return_attn_wts=True)
    assert np.abs(out - g["out"]).max() < 2e-6 and np.abs(attn - g["attn"]).max() < 1e-6
    assert np.all(out[3] == 0) and np.allclose(attn[3], 1.0 / g["ci"].shape[1])   # all -1 row
    heads = [(g[f"h{i}_W_i"], g[f"h{i}_W_j"], g[f"h{i}_att_w"], g[f"h{i}_att_b"]) for i in range(2)]
    assert np.abs(O.gat_multihead(g["h"], g["ci"], heads) - g["out_2head"]).max() < 2e-6


def test_train_mode_forward():
    g = load_golden("g_train_r18_img128")
    images, bboxes, add, ci, labels = synth.gen(2, 12, 8, seed=8, img=128, with_labels=True)
    out = O.cova_forward(np_sd(), images.numpy(), bboxes.numpy(), add.numpy(), ci.numpy(), train=True)
    assert rel_err(out, g["logits"]) < 2e-5
    assert np.array_equal(labels.numpy(), g["labels"])


def test_torch_port_matches_golden():
    """The CPU-baseline port (`oracle/torch_port.py`, timed by bench.py) reproduces the live reference."""
    from oracle import torch_port as TP
    for name, args, kw in (("g_small_r18_img128", (2, 12, 8), dict(seed=0, img=128)),
                           ("g_ragged_r18_img256", (3, 0, 24), dict(seed=3, img=256, counts=[11, 1, 30]))):
        out = TP.forward(synth.make_state_dict(123), *synth.gen(*args, **kw))
        assert rel_err(out.numpy(), load_golden(name)["logits"]) < 5e-6
    out = TP.forward(synth.make_state_dict(123, backbone="resnet50"), *synth.gen(1, 16, 8, seed=5, img=128))
    assert rel_err(out.numpy(), load_golden("g_r50_img128")["logits"]) < 1e-5


def test_torch_port_matches_config_goldens():
    """The port on the BASELINE configs at full shape (fixtures frozen from the live reference on the inputs bench.py
    times): config 2 (B=16, 1280^2, N=90, K=24) restricted to its first 4 pages - pages are independent in eval mode,
    which is itself asserted here - and the config-5 shape (ResNet-50, N=300, K=48, 2 heads)."""
    from oracle import torch_port as TP
    images, bboxes, add, ci = synth.gen(16, 90, 24, seed=1)
    out = TP.forward(synth.make_state_dict(123), images[:4], bboxes[:360], add[:360], ci[:360])
    assert rel_err(out.numpy(), load_golden("g_c2_r18_b16")["logits"][:360]) < 5e-6
    out = TP.forward(synth.make_state_dict(123, backbone="resnet50", n_heads=2), *synth.gen(2, 300, 48, seed=1))
    assert rel_err(out.numpy(), load_golden("g_c5_r50_n300_k48_h2")["logits"]) < 1e-5


# ----------------------------------------------------------------------------- callers either side of the forward
def test_tail_ce_sum_matches_torch():
    g = load_golden("g_tail")
    loss, grad, nc = O.ce_sum(g["ce_logits"], g["ce_labels"])
    assert abs(float(loss) - float(g["ce_loss"])) <= 2e-6 * abs(float(g["ce_loss"]))
    assert np.abs(grad - g["ce_grad"]).max() < 1e-6
    assert nc == int(g["ce_correct"])


def test_tail_adam_matches_torch():
    g = load_golden("g_tail")
    p, m, v = g["adam_p0"], np.zeros(1003, np.float32), np.zeros(1003, np.float32)
    for i in range(4):
        p, m, v = O.adam_step(p, g["adam_grads"][i], m, v, 5e-4, 0.9, 0.999, 1e-8, 1e-3, i + 1)
        assert np.abs(p - g["adam_p%d" % (i + 1)]).max() < 2e-7
    assert rel_err(m, g["adam_m4"]) < 1e-6 and rel_err(v, g["adam_v4"]) < 1e-6


def test_tail_build_batch_matches_collate():
    g = load_golden("g_tail")
    bb, ctx = O.build_batch(list(g["coll_counts"]), int(g["coll_cs"]), g["coll_xywh"])
    assert np.array_equal(ctx, g["coll_ctx"])
    assert np.array_equal(bb, g["coll_bboxes"])


def test_tail_topk_hits_matches_evaluate_model():
    g = load_golden("g_tail")
    for k in (1, 3):
        rows = []
        for bi in range(3):
            page = g["ev_bboxes%d" % bi][:, 0].astype(np.int64)
            off = np.concatenate(([0], np.cumsum(np.bincount(page))))
            rows.append(O.topk_hits(g["ev_logits%d" % bi], g["ev_labels%d" % bi], off, k))
        hits = np.concatenate(rows)
        assert np.array_equal(hits[:, 1:], g["ev_img_acc_k%d" % k][:, 1:])
        assert np.allclose(hits[:, 1:].mean(0) * 100, g["ev_class_acc_k%d" % k][1:])


# ----------------------------------------------------------------------------- training-mode backbone pieces (A9)
def _train_case(seed, shape=(2, 8, 9, 11), with_res=True):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(shape, generator=g) * 2 + 0.5
    res = torch.randn(shape, generator=g) if with_res else None
    w, b = torch.rand(shape[1], generator=g) + 0.5, torch.randn(shape[1], generator=g) * 0.2
    dy = torch.randn(shape, generator=g)
    return x, res, w, b, dy


@pytest.mark.parametrize("with_res,relu", [(True, True), (False, True), (False, False)])
def test_train_bn_act_oracle_matches_torch(with_res, relu):
    """oracle.bn_act_train / _backward against torch's own BatchNorm2d(train) + autograd (the library the reference
    calls for these ops): forward, batch statistics, every gradient."""
    x, res, w, b, dy = _train_case(3, with_res=with_res)
    xt = x.clone().requires_grad_(True)
    rt = res.clone().requires_grad_(True) if with_res else None
    bn = torch.nn.BatchNorm2d(x.shape[1]).train()
    with torch.no_grad():
        bn.weight.copy_(w); bn.bias.copy_(b)
    z = bn(xt)
    if with_res:
        z = z + rt
    if relu:
        z = torch.relu(z)
    z.backward(dy)
    y, m, v = O.bn_act_train(x.numpy(), w.numpy(), b.numpy(), None if res is None else res.numpy(), relu)
    dx, dres, dw, db = O.bn_act_train_backward(dy.numpy(), x.numpy(), w.numpy(), b.numpy(),
                                               None if res is None else res.numpy(), relu)
    assert rel_err(y, z.detach().numpy()) < 2e-6
    assert rel_err(m, x.mean((0, 2, 3)).numpy()) < 2e-6 and rel_err(v, x.var((0, 2, 3), unbiased=False).numpy()) < 2e-6
    assert rel_err(dx, xt.grad.numpy()) < 2e-5
    assert rel_err(dw, bn.weight.grad.numpy()) < 2e-5 and rel_err(db, bn.bias.grad.numpy()) < 2e-5
    if with_res:
        assert rel_err(dres, rt.grad.numpy()) < 2e-6


def test_train_maxpool_backward_oracle_matches_torch_with_ties():
    g = torch.Generator().manual_seed(5)
    x = torch.relu(torch.randn(2, 5, 9, 12, generator=g)).round(decimals=1).requires_grad_(True)   # tied zeros
    y = torch.nn.functional.max_pool2d(x, 3, 2, 1)
    dy = torch.randn(y.shape, generator=g)
    y.backward(dy)
    assert np.array_equal(O.maxpool3x3s2p1(x.detach().numpy()), y.detach().numpy())
    assert np.array_equal(O.maxpool3x3s2p1_backward(x.detach().numpy(), dy.numpy()), x.grad.numpy())
