"""One Bottleneck with downsample (layer1[0] of ResNet-50) on the native training path against torch autograd in float64."""
import os, sys, copy, warnings
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")
from cova_b200 import ops
from cova_b200.models import Bottleneck
from cova_b200 import train_backbone as tb
DEV = "cuda:0"
torch.manual_seed(0)
which = os.environ.get("BLK", "down")
blk = (Bottleneck(64, 64, True) if which == "down" else Bottleneck(256)).to(DEV).train()
for m in blk.modules():
    if isinstance(m, torch.nn.BatchNorm2d):
        m.weight.data.uniform_(0.5, 1.5); m.bias.data.normal_(0, 0.2)
ref = copy.deepcopy(blk).double()
cin = 64 if which == "down" else 256
x = torch.relu(torch.randn(2, 48, 48, cin, device=DEV)).requires_grad_(True)
G = torch.randn(2, 48, 48, 256, device=DEV)
pl = ops.split_planes(x.detach(), ops.F16X2)
xp = (pl.p0, pl.p1)
o, op = tb._bn_act(tb._conv(x, xp, blk.conv1), blk.bn1, planes_for=blk.conv2)
o, op = tb._bn_act(tb._conv(o, op, blk.conv2), blk.bn2, planes_for=blk.conv3)
if blk.downsample is None:
    idt = x
else:
    idt, _ = tb._bn_act(tb._conv(x, xp, blk.downsample[0]), blk.downsample[1], relu=False)
out, _ = tb._bn_act(tb._conv(o, op, blk.conv3), blk.bn3, res=idt, planes_for=None)
(out * G).sum().backward()
xr = x.detach().double().requires_grad_(True)
outr = ref(xr.permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
(outr * G.double()).sum().backward()
e = lambda a, b: float((a.double() - b).abs().max() / b.abs().max())
print("out", e(out, outr), "dx", e(x.grad, xr.grad))
for (n, p), pr in zip(blk.named_parameters(), ref.parameters()):
    print(f"{n:28s} {e(p.grad, pr.grad):9.2e}")
