import os, sys, warnings
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
from cova_b200 import ops
from cova_b200.ops import BF16X2, F32, ENGINE_TCGEN05 as TC
exec(open(os.path.join(os.path.dirname(__file__), "sweep_common.py")).read())
for mode in (0, 1, 2):
    for rpf in (1, 2):
        ops.set_knob("conv_res_load", mode); ops.set_knob("conv_res_prefetch", rpf)
        us = timeit(lambda: conv(r))
        print(f"conv residual load_mode={mode} res_prefetch={rpf}: {us:7.1f} us   {breakdown(r)}")
