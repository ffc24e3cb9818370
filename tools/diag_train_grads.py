"""Train-step gradients of the native path against the LIVE-reference fixtures (tests/golden/g_train_*.npz), per
parameter.   FIX=r18|r50   [COVA_B200_TRAIN_* environment switches select library pieces for comparison]"""
import os, sys, warnings
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")
import cova_b200.synth as synth
from cova_b200.models import CoVA
DEV = "cuda:0"
fix = os.environ.get("FIX", "r50")
bk, img, name, gen = {"r18": ("resnet18", 128, "g_train_r18_img128", (2, 12, 8, 8)),
                      "r50": ("resnet50", 192, "g_train_r50_img192", (2, 14, 8, 15))}[fix]
g = dict(np.load(os.path.join(ROOT, "tests", "golden", name + ".npz")))
m = CoVA((3, 3), img, 4, True, 384, 32, 0, 0.0, None, pretrained=False, backbone=bk)
m.load_state_dict(synth.make_state_dict(123, backbone=bk), strict=True)
m = m.to(DEV).train()
inp = [t.to(DEV) for t in synth.gen(gen[0], gen[1], gen[2], seed=int(g["seed"]) if "seed" in g else gen[3], img=img, with_labels=True)]
out = m(*inp[:4])
loss = torch.nn.CrossEntropyLoss(reduction="sum")(out, inp[4])
loss.backward()
n = lambda t: t.detach().float().cpu().numpy()
print("logits rel err", float(np.abs(n(out) - g["logits"]).max() / np.abs(g["logits"]).max()), "loss", float(loss), float(g["loss"]))
grads = dict(m.named_parameters())
gscale = max(float(np.abs(g[k]).max()) for k in g if k.startswith("grad:"))
for k in [k for k in g if k.startswith("grad:")]:
    got, want = n(grads[k[5:]].grad), g[k]
    if got.shape != want.shape:
        got = got[:, ::8]
    print(f"{k[5:]:36s} max|g| {np.abs(want).max():10.3e}  err/max {np.abs(got - want).max() / np.abs(want).max():9.2e}  err/gscale {np.abs(got - want).max() / gscale:9.2e}")
sd = m.state_dict()
for k in [k for k in g if k.startswith("buf:")]:
    e = np.abs(n(sd[k[4:]]) - g[k]).max() / np.abs(g[k]).max()
    if e > 1e-5:
        print("buffer", k[4:], e)
if os.environ.get("TRUTH", "1") == "1":
    # float64 "truth": the same model through the PyTorch-operator composite path in double precision
    import copy
    os.environ["COVA_B200_TRAIN_BACKBONE"] = "torch"
    os.environ["COVA_B200_GAT_TRAIN"] = "torch"
    m2 = CoVA((3, 3), img, 4, True, 384, 32, 0, 0.0, None, pretrained=False, backbone=bk)
    m2.load_state_dict(synth.make_state_dict(123, backbone=bk), strict=True)
    m2 = m2.to(DEV).double().train()
    import torchvision
    from cova_b200 import models as M
    fm = m2.convnet(inp[0].double())
    vis = torchvision.ops.roi_pool(fm, inp[1].double(), (3, 3), m2.spatial_scale).reshape(inp[1].shape[0], -1)
    own = torch.cat((vis, m2._get_bbox_features(inp[1].double()), inp[2].double()), 1)
    ctxr = m2.gat._forward_composite(own, inp[3])
    out2 = m2.decoder(torch.cat((own, ctxr), 1))
    loss2 = torch.nn.CrossEntropyLoss(reduction="sum")(out2, inp[4])
    loss2.backward()
    print("TRUTH(f64) logits vs fixture", float(np.abs(out2.detach().cpu().numpy() - g["logits"]).max() / np.abs(g["logits"]).max()))
    g2 = dict(m2.named_parameters())
    for k in [k for k in g if k.startswith("grad:")]:
        t = g2[k[5:]].grad.cpu().numpy()
        got, want = n(grads[k[5:]].grad), g[k]
        if t.shape != want.shape:
            t = t[:, ::8]; got = got[:, ::8]
        sc = np.abs(t).max()
        print(f"{k[5:]:36s} native-vs-truth {np.abs(got - t).max() / sc:9.2e}   fixture(fp32 CPU reference)-vs-truth {np.abs(want - t).max() / sc:9.2e}")
    for k in ("convnet.4.0.bn3.bias", "convnet.4.1.bn3.bias", "convnet.4.0.downsample.1.weight"):
        t, got, want = g2[k].grad.cpu().numpy(), n(grads[k].grad), g["grad:" + k]
        d = np.abs(got - t)
        idx = np.argsort(-d)[:6]
        print(k, "largest native-truth diffs at channels", idx, d[idx], "truth there", t[idx], "fixture-truth", (want - t)[idx])
