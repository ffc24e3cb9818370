"""Knob sweep + wait-cycle breakdown of the two backbone kernels (stem_tc, conv3x3_tc) at BASELINE configs[1] size
(B=16 pages of 1280x1280): CUDA events on the launching stream, inputs larger than L2.  Output -> profiles/."""
import os, sys, warnings
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
from cova_b200 import ops
from cova_b200.ops import BF16X2, F32, ENGINE_TCGEN05 as TC

dev = torch.device("cuda", 0)
B = int(os.environ.get("B", 16))
torch.manual_seed(0)


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)) * 1e3


def planes(f32):
    p = ops.Planes(BF16X2, f32.shape, dev)
    p.p0.copy_(f32.to(torch.bfloat16))
    p.p1.copy_((f32 - p.p0.float()).to(torch.bfloat16))
    return p


x = planes(torch.randn(B, 320, 320, 64, device=dev))
r = planes(torch.randn(B, 320, 320, 64, device=dev))
w = torch.randn(64, 64, 3, 3, device=dev) * 0.05
_, whi, wlo = ops.pack_conv_weight(w, simt=False, tc=True, split=True)
sc, sh = torch.rand(64, device=dev) + 0.5, torch.randn(64, device=dev)
NSM = ops.device_info()[0]
dbg = torch.zeros(8 * NSM, dtype=torch.int64, device=dev)


def conv(res, out_dtype=None):
    return ops.conv3x3_bn_act_fwd(x, whi, wlo, sc, sh, res=res, relu=True, out_dtype=out_dtype, engine=TC)


def breakdown(res):
    dbg.zero_()
    ops.debug_buffer(dbg)
    conv(res)
    torch.cuda.synchronize()
    ops.debug_buffer(None)
    d = dbg.view(NSM, 8).double().cpu().numpy()
    tot = d[:, 4].mean()
    names = ["issuer waits operands", "issuer waits accumulator", "producer waits slot", "epilogue(w2) waits accumulator"]
    return "  ".join(f"{n} {d[:, i].mean() / tot:5.1%}" for i, n in enumerate(names)) + f"  | CTA cycles {tot:,.0f}, tiles/CTA {d[:, 5].mean():.1f}"


print(f"conv3x3_tc B={B}: feed-bound floor = {B * 20 * 40 / NSM * 4032 / 1.965e3:.0f} us at 1965 MHz")
for res, tag in ((None, "no residual"), (r, "residual   ")):
    for pf in (0, 1, 2, 3, 4):
        for rpf in ((1,) if res is None else (1, 2, 3)):
            ops.set_knob("conv_l2_prefetch", pf)
            ops.set_knob("conv_res_prefetch", rpf)
            us = timeit(lambda: conv(res))
            print(f"conv {tag} l2_prefetch={pf} res_prefetch={rpf}: {us:7.1f} us   {breakdown(res)}")
us = timeit(lambda: conv(r, F32))
print(f"conv residual fp32-out (last conv) at the last knobs: {us:7.1f} us")
ops.set_knob("conv_l2_prefetch", -1); ops.set_knob("conv_res_prefetch", -1)

img = torch.rand(B, 3, 1280, 1280, device=dev)
img8 = (img * 255).to(torch.uint8)
sw = ops.pack_stem_weight(torch.randn(64, 3, 7, 7, device=dev) * 0.05)
print(f"stem_tc B={B}: feed-bound floor = {B * 640 * 5 / NSM * 1568 / 1.965e3:.0f} us")
for ncv, pf in [(4, 0), (4, 1), (4, 2), (4, 3), (4, 4), (4, 6), (8, 0), (8, 1), (8, 2), (8, 3)]:
    ops.set_knob("stem_l2_prefetch", pf)
    ops.set_knob("stem_converters", ncv)
    a = timeit(lambda: ops.stem_fwd(img, sw, sc, sh, out_dtype=BF16X2, engine=TC))
    b = timeit(lambda: ops.stem_fwd(img8, sw, sc, sh, out_dtype=BF16X2, engine=TC))
    print(f"stem converters={ncv} l2_prefetch={pf}: fp32 images {a:7.1f} us   uint8 images {b:7.1f} us")
ops.set_knob("stem_l2_prefetch", -1); ops.set_knob("stem_converters", -1)
