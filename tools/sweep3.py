import os, sys, warnings
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
from cova_b200 import ops
from cova_b200.ops import BF16X2, F32, ENGINE_TCGEN05 as TC
exec(open(os.path.join(os.path.dirname(__file__), "sweep_common.py")).read())
import subprocess, threading, time
img = torch.rand(B, 3, 1280, 1280, device=dev)
sw = ops.pack_stem_weight(torch.randn(64, 3, 7, 7, device=dev) * 0.05)
dbg2 = torch.zeros(8 * 256, dtype=torch.int64, device=dev)
for ncv, pf in [(4, 0), (8, 0), (8, 1)]:
    ops.set_knob("stem_l2_prefetch", pf); ops.set_knob("stem_converters", ncv)
    us = timeit(lambda: ops.stem_fwd(img, sw, sc, sh, out_dtype=BF16X2, engine=TC))
    dbg2.zero_(); ops.debug_buffer(dbg2)
    ops.stem_fwd(img, sw, sc, sh, out_dtype=BF16X2, engine=TC); torch.cuda.synchronize(); ops.debug_buffer(None)
    d = dbg2.view(256, 8)[:144].double().cpu().numpy()
    tot = d[:, 4].mean()
    print(f"stem converters={ncv} pf={pf}: {us:7.1f} us  issuer waits rows {d[:,0].mean()/tot:5.1%}  waits accumulator {d[:,1].mean()/tot:5.1%}  CTA cycles {tot:,.0f} max {d[:,4].max():,.0f}  tiles/CTA {d[:,5].mean():.1f}  clk/tile {tot/d[:,5].mean():.0f}")
# power / clocks under a sustained conv loop
samples = []
def poll():
    for _ in range(12):
        samples.append(subprocess.run(["nvidia-smi", "--query-gpu=power.draw,power.limit,clocks.sm,clocks.max.sm,temperature.gpu,clocks_throttle_reasons.active", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip())
        time.sleep(0.25)
for name, fn in (("conv residual", lambda: conv(r)), ("conv no-res", lambda: conv(None)), ("stem", lambda: ops.stem_fwd(img, sw, sc, sh, out_dtype=BF16X2, engine=TC))):
    samples.clear()
    th = threading.Thread(target=poll); th.start()
    t0 = time.time()
    while time.time() - t0 < 3.2:
        for _ in range(50): fn()
        torch.cuda.synchronize()
    th.join()
    print(name, "|", " || ".join(samples[2::3]))
