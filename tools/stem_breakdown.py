"""Stem kernel (cova_stem_fwd, tcgen05 engine) timing + issuer wait breakdown from the kernel's debug counters
(cova_debug_buffer): fp32 images, uint8 images in the exact 3-product LUT mode and in the integer 2-product mode.
Usage (GPU box): python tools/stem_breakdown.py   [B=16]"""
import os, sys, warnings
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
from cova_b200 import ops
from cova_b200.ops import BF16X2, F32, ENGINE_TCGEN05 as TC
exec(open(os.path.join(os.path.dirname(__file__), "sweep_common.py")).read())
img8 = (torch.rand(B, 3, 1280, 1280, device=dev) * 255).round().to(torch.uint8)
imgf = img8.float() / 255
sw = ops.pack_stem_weight(torch.randn(64, 3, 7, 7, device=dev) * 0.05)
dbg2 = torch.zeros(8 * 256, dtype=torch.int64, device=dev)
ref = None
for name, img, knob in (("fp32 images", imgf, 0), ("uint8, integer pixels (2 products)", img8, 0), ("uint8, exact v/255 LUT (3 products)", img8, 1), ("fp32 images", imgf, 0)):
    ops.set_knob("stem_u8_exact", knob)
    us = timeit(lambda: ops.stem_fwd(img, sw, sc, sh, out_dtype=BF16X2, engine=TC))
    dbg2.zero_(); ops.debug_buffer(dbg2)
    o = ops.stem_fwd(img, sw, sc, sh, out_dtype=BF16X2, engine=TC); torch.cuda.synchronize(); ops.debug_buffer(None)
    v = o.p0.float() + o.p1.float()
    if ref is None: ref = v
    d = dbg2.view(256, 8)[:144].double().cpu().numpy()
    tot = d[:, 4].mean()
    print(f"stem, {name:38s}: {us:7.1f} us  max|y - y_fp32| {(v - ref).abs().max().item():.2e}  issuer 0: waits {d[:,0].mean()/tot:5.1%}  MMA issue {d[:,2].mean()/tot:5.1%}  commits {d[:,3].mean()/tot:5.1%}  clk per conv row {tot/d[:,5].mean():.0f}")
ops.set_knob("stem_u8_exact", -1)
