"""bf16 training mode vs the fp32-parity training path on the same weights / batch: logits, loss, per-parameter gradient
norm error and cosine.  BACKBONE=resnet18|resnet50 IMG=128 B=2"""
import os, sys, warnings
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import cova_b200.synth as synth
from cova_b200.models import CoVA
DEV = "cuda:0"
for bk in os.environ.get("BACKBONE", "resnet18,resnet50").split(","):
    img, B = int(os.environ.get("IMG", "128")), int(os.environ.get("B", "2"))
    inp = [t.to(DEV) for t in synth.gen(B, int(os.environ.get("N", "12")), int(os.environ.get("K", "8")), seed=8, img=img, with_labels=True)]
    outs = {}
    for prec in ("fp32", "bf16"):
        m = CoVA((3, 3), img, 4, True, 384, 32, 0, 0.0, None, pretrained=False, backbone=bk, precision=prec)
        m.load_state_dict(synth.make_state_dict(123, backbone=bk), strict=True)
        m = m.to(DEV).train()
        out = m(*inp[:4]); loss = torch.nn.CrossEntropyLoss(reduction="sum")(out, inp[4]); loss.backward()
        outs[prec] = (out.detach(), float(loss), {k: p.grad.detach().clone() for k, p in m.named_parameters()})
    (o32, l32, g32), (o16, l16, g16) = outs["fp32"], outs["bf16"]
    print(bk, "logits max|d|/max|ref| %.3e" % float((o16 - o32).abs().max() / o32.abs().max()), "loss", l32, l16)
    gmax = max(float(v.norm()) for v in g32.values())
    for k in g32:
        a, b = g16[k].flatten().double(), g32[k].flatten().double()
        print("  %-36s |g| %.3e  rel-norm-err %.3e  cos %.5f%s" % (k, float(b.norm()), float((a - b).norm() / (b.norm() + 1e-30)),
              float(torch.dot(a, b) / (a.norm() * b.norm() + 1e-30)), "  (negligible)" if float(b.norm()) < 1e-3 * gmax else ""))

# backbone alone under a LINEAR loss (no train-mode BatchNorm1d head): isolates the bf16 rounding of the backbone from the head's
# conditioning (the decoder's BatchNorm1d amplifies 1e-5 forward perturbations into 1-2 % of the gradients)
from cova_b200.train_backbone import feature_map_train
for bk in os.environ.get("BACKBONE", "resnet18,resnet50").split(","):
    img, B = int(os.environ.get("IMG", "128")), int(os.environ.get("B", "2"))
    images = synth.gen(B, 4, 2, seed=8, img=img)[0].to(DEV)
    res = {}
    for prec in ("fp32", "bf16"):
        m = CoVA((3, 3), img, 4, True, 384, 32, 0, 0.0, None, pretrained=False, backbone=bk, precision=prec)
        m.load_state_dict(synth.make_state_dict(123, backbone=bk), strict=True)
        m = m.to(DEV).train()
        fm = feature_map_train(m.convnet, images, prec)
        gsel = torch.Generator().manual_seed(3)
        wsel = torch.randn(fm.shape, generator=gsel).to(DEV)
        (fm * wsel).sum().backward()
        res[prec] = (fm.detach(), {k: p.grad.detach().clone() for k, p in m.convnet.named_parameters()})
    (f32, g32), (f16, g16) = res["fp32"], res["bf16"]
    print(bk, "backbone only: fm max|d|/max|ref| %.3e  rms %.3e" % (float((f16 - f32).abs().max() / f32.abs().max()),
          float((f16 - f32).norm() / f32.norm())))
    for k in g32:
        a, b = g16[k].flatten().double(), g32[k].flatten().double()
        print("  %-36s |g| %.3e  rel-norm-err %.3e  cos %.5f" % (k, float(b.norm()), float((a - b).norm() / (b.norm() + 1e-30)),
              float(torch.dot(a, b) / (a.norm() * b.norm() + 1e-30))))
