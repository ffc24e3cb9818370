"""ResNet-50 tensor-core path at full size: run it repeatedly (optionally under compute-sanitizer) and report."""
import os, sys, warnings
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
os.environ["COVA_B200_BACKBONE"] = "resnet50"
import bench
import cova_b200.synth as synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
n = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda", 0)
model = bench.build_model(dev)
inp = [t.to(dev) for t in synth.gen(B, 90, 24, seed=1)]
with torch.no_grad():
    ref = None
    for i in range(n):
        out = model(*inp)
        torch.cuda.synchronize()
        if ref is None:
            ref = out.clone()
        print(i, "ok", float(out.abs().max()), "same" if torch.equal(out, ref) else "DIFFERENT", flush=True)
print("peak mem GB", torch.cuda.max_memory_allocated() / 1e9)
