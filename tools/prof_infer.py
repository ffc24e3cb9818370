"""One config-2 inference step (B=16, 1280^2, N=90, K=24, fp32-parity mode) between cudaProfilerStart / Stop, for
`ncu --profile-from-start off --set full ...` (profiles/r02_step_full.md, profiles/traffic_r02.json)."""
import os, sys, warnings
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import bench
dev = torch.device("cuda", 0)
cfg = bench.get_config(int(os.environ.get("CONFIG", "2")))
import cova_b200.synth as synth
m = bench.build_model(dev, cfg).eval()
inp = [t.to(dev) for t in synth.gen(cfg["B"], cfg["N"], cfg["K"], seed=1)]
with torch.no_grad():
    for _ in range(3):
        m(*inp)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    m(*inp)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
