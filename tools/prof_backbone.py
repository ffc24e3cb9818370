"""Minimal launcher for ncu: stem_tc, conv3x3_tc (no residual), conv3x3_tc (residual) at B=16, two launches each."""
import os, sys, warnings
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
from cova_b200 import ops
from cova_b200.ops import BF16X2, ENGINE_TCGEN05 as TC
dev = torch.device("cuda", 0)
B = int(os.environ.get("B", 16))
for k, v in (("stem_converters", os.environ.get("NCV")), ("stem_l2_prefetch", os.environ.get("SPF")),
             ("conv_res_prefetch", os.environ.get("RPF")), ("conv_l2_prefetch", os.environ.get("CPF"))):
    if v is not None:
        ops.set_knob(k, int(v))
torch.manual_seed(0)
def planes(f32):
    p = ops.Planes(BF16X2, f32.shape, dev)
    p.p0.copy_(f32.to(torch.bfloat16)); p.p1.copy_((f32 - p.p0.float()).to(torch.bfloat16))
    return p
x, r = planes(torch.randn(B, 320, 320, 64, device=dev)), planes(torch.randn(B, 320, 320, 64, device=dev))
_, whi, wlo = ops.pack_conv_weight(torch.randn(64, 64, 3, 3, device=dev) * 0.05, simt=False, tc=True, split=True)
sc, sh = torch.rand(64, device=dev) + 0.5, torch.randn(64, device=dev)
img = torch.rand(B, 3, 1280, 1280, device=dev)
sw = ops.pack_stem_weight(torch.randn(64, 3, 7, 7, device=dev) * 0.05)
for _ in range(2):
    ops.stem_fwd(img, sw, sc, sh, out_dtype=BF16X2, engine=TC)
    ops.conv3x3_bn_act_fwd(x, whi, wlo, sc, sh, res=None, relu=True, engine=TC)
    ops.conv3x3_bn_act_fwd(x, whi, wlo, sc, sh, res=r, relu=True, engine=TC)
torch.cuda.synchronize()
