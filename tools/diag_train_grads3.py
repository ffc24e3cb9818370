"""Is a 1e-5 relative perturbation of the visual features enough to move the gradients by ~1 %?  (conditioning check)"""
import os, sys, copy, warnings
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import cova_b200.synth as synth
from cova_b200.models import CoVA
DEV = "cuda:0"
os.environ["COVA_B200_TRAIN_BACKBONE"] = "torch"
B, N, K, img = int(os.environ.get("B", 2)), int(os.environ.get("N", 12)), int(os.environ.get("K", 8)), int(os.environ.get("IMG", 128))
m1 = CoVA((3, 3), img, 4, True, 384, 32, 0, 0.0, None, pretrained=False)
m1.load_state_dict(synth.make_state_dict(123), strict=True)
m1 = m1.to(DEV).train()
m2 = copy.deepcopy(m1)
inp = [t.to(DEV) for t in synth.gen(B, N, K, seed=8, img=img, with_labels=True)]
crit = torch.nn.CrossEntropyLoss(reduction="sum")
eps = float(os.environ.get("EPS", 1e-5))
orig = CoVA._get_visual_features
def noisy(self, images, bboxes):
    v = orig(self, images, bboxes)
    g = torch.Generator(device=DEV).manual_seed(1)
    return v * (1 + eps * torch.randn(v.shape, device=DEV, generator=g))
with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
    out1 = m1(*inp[:4]); crit(out1, inp[4]).backward()
    CoVA._get_visual_features = noisy
    out2 = m2(*inp[:4]); crit(out2, inp[4]).backward()
print("perturbation", eps, "logits rel diff", float((out1 - out2).abs().max() / out1.abs().max()))
for (name, p1), p2 in zip(m1.named_parameters(), m2.parameters()):
    if name.startswith("convnet.4.1.conv2") or not name.startswith("convnet"):
        print(f"{name:32s} relL2 {float((p1.grad - p2.grad).norm() / p1.grad.norm().clamp_min(1e-30)):9.2e}")
