"""Distribution of the end-to-end step time over repetitions (is the e2e variance ours or the host's?)."""
import os, sys, time, warnings
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import bench
import cova_b200.synth as synth
from cova_b200.pipeline import prefetch
dev = torch.device("cuda", 0)
model = bench.build_model(dev)
inp = synth.gen(16, 90, 24, seed=1)
pinned = [t.pin_memory() for t in inp]
pinned_u8 = [(inp[0] * 255).round().to(torch.uint8).pin_memory()] + pinned[1:]
host_out = torch.empty((1440, 4)).pin_memory()
dimg = torch.empty_like(inp[0], device=dev)


def rep(host, steps=20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    with torch.no_grad():
        for d in prefetch((host for _ in range(steps)), dev):
            host_out.copy_(model(*d), non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def h2d(steps=20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(steps):
        dimg.copy_(pinned[0], non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    return pinned[0].numel() * 4 / (e0.elapsed_time(e1) / steps) / 1e6


rep(pinned, 3); rep(pinned_u8, 3)
for r in range(3):
    print("round", r, "raw H2D GB/s:", " ".join("%.1f" % h2d() for _ in range(6)))
    print("round", r, "e2e fp32 ms/step:", " ".join("%.2f" % rep(pinned) for _ in range(6)))
    print("round", r, "e2e u8   ms/step:", " ".join("%.2f" % rep(pinned_u8) for _ in range(6)))
    time.sleep(1.0)
