"""Diagnose the e2e pipeline: H2D alone, compute alone (fp32 / uint8 resident), pipelined, and host-side launch cost."""
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

def timeit(fn, n=20):
    fn(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n): fn()
    e1.record(); t_cpu = (time.perf_counter() - t0) / n
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, t_cpu * 1e3

for name, host in (("fp32", pinned), ("uint8", pinned_u8)):
    d = [t.to(dev) for t in host]
    with torch.no_grad():
        g, c = timeit(lambda: model(*d))
    print(f"{name}: compute resident  gpu {g:.3f} ms/step, host-side issue {c:.3f} ms/step")
    g, c = timeit(lambda: [t.to(dev, non_blocking=True) for t in host])
    print(f"{name}: H2D only          gpu {g:.3f} ms/step ({sum(t.numel()*t.element_size() for t in host)/g/1e6:.1f} GB/s)")
    def pipe(k=20):
        with torch.no_grad():
            for dd in prefetch((host for _ in range(k)), dev):
                host_out.copy_(model(*dd), non_blocking=True)
    torch.cuda.synchronize(); pipe(3); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record(); pipe(20); e1.record(); tc = time.perf_counter() - t0
    torch.cuda.synchronize()
    print(f"{name}: pipelined e2e     gpu {e0.elapsed_time(e1)/20:.3f} ms/step, host loop returned after {tc/20*1e3:.3f} ms/step")
