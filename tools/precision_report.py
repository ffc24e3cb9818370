"""Measured error of every precision mode against the fixtures frozen from the live reference (max|a-b| / max|b|).
Output -> profiles/."""
import os, sys, warnings
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
warnings.filterwarnings("ignore")
import cova_b200.synth as synth
from conftest import load_golden, rel_err
from cova_b200.models import CoVA
DEV = "cuda:0"
cases = [("g_small_r18_img128", 128, lambda: synth.gen(2, 12, 8, seed=0, img=128)),
         ("g_ragged_r18_img256", 256, lambda: synth.gen(3, 0, 24, seed=3, img=256, counts=[11, 1, 30])),
         ("g_c1_r18_img1280", 1280, lambda: synth.gen(1, 32, 8, seed=0, img=1280))]
print("%-22s %-8s %-6s %10s %10s %10s" % ("fixture", "engine", "prec", "fm", "own", "logits"))
for name, img, mk in cases:
    g = load_golden(name)
    for engine, prec in (("simt", "fp32"), ("tcgen05", "fp32x"), ("tcgen05", "fp32"), ("tcgen05", "fp16"), ("tcgen05", "bf16")):
        import io, contextlib
        with contextlib.redirect_stdout(io.StringIO()):
            m = CoVA((3, 3), img, 4, True, 384, 32, 0, 0.2, None, pretrained=False, engine=engine, precision=prec)
        m.load_state_dict(synth.make_state_dict(123), strict=True)
        m = m.to(DEV).eval()
        with torch.no_grad():
            r = m._native.forward(*[t.to(DEV) for t in mk()], return_intermediates=True)
        own = rel_err(r["own"].cpu().numpy(), g["own"]) if "own" in g else float("nan")
        fm = r["fm"].cpu().numpy().transpose(0, 3, 1, 2)
        fme = float("nan")
        if "fm" in g and g["fm"].shape == fm.shape:
            fme = rel_err(fm, g["fm"])
        elif "fm_sample" in g and fm[:, :, ::8, ::8].shape == g["fm_sample"].shape:
            fme = rel_err(fm[:, :, ::8, ::8], g["fm_sample"])
        print("%-22s %-8s %-6s %10.2e %10.2e %10.2e" % (name, engine, prec, fme, own, rel_err(r["logits"].cpu().numpy(), g["logits"])))
