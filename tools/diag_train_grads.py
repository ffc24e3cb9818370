import os, sys, copy, warnings
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import cova_b200.synth as synth
from cova_b200.models import CoVA
DEV = "cuda:0"
img = int(os.environ.get("IMG", 128))
m1 = CoVA((3, 3), img, 4, True, 384, 32, 0, 0.0, None, pretrained=False)
m1.load_state_dict(synth.make_state_dict(123), strict=True)
m1 = m1.to(DEV).train()
m2 = copy.deepcopy(m1)
inp = [t.to(DEV) for t in synth.gen(int(os.environ.get("B", 2)), int(os.environ.get("N", 12)), int(os.environ.get("K", 8)), seed=8, img=img, with_labels=True)]
crit = torch.nn.CrossEntropyLoss(reduction="sum")
with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
    out1 = m1(*inp[:4]); crit(out1, inp[4]).backward()
    os.environ["COVA_B200_TRAIN_BACKBONE"] = "torch"
    out2 = m2(*inp[:4]); crit(out2, inp[4]).backward()
print("logits", float((out1 - out2).abs().max() / out2.abs().max()))
gmax = max(float(p.grad.abs().max()) for p in m2.parameters())
for (name, p1), p2 in zip(m1.named_parameters(), m2.parameters()):
    if os.environ.get("ALL", "0") != "1" and not name.startswith("convnet"): continue
    l2 = float((p1.grad - p2.grad).norm() / p2.grad.norm().clamp_min(1e-30))
    print(f"{name:32s} relL2 {l2:9.2e} max|g| {float(p2.grad.abs().max()):10.3e}  err/max {float((p1.grad - p2.grad).abs().max() / max(float(p2.grad.abs().max()), 1e-4 * gmax)):9.2e}")
