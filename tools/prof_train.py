import os, sys, warnings
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import cova_b200.synth as synth
from cova_b200.models import CoVA
from cova_b200.train_ops import CrossEntropyLossSum, FlatAdam
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda", 0)
bk = os.environ.get("BACKBONE", "resnet18")
B = int(os.environ.get("B", "16"))
m = CoVA((3, 3), 1280, 4, True, 384, 32, 0, 0.2, None, pretrained=False, backbone=bk, precision=os.environ.get("PRECISION", "fp32"))
m.load_state_dict(synth.make_state_dict(123, backbone=bk), strict=True)
m = m.to(dev).train()
opt = FlatAdam(m.parameters(), lr=5e-4, weight_decay=1e-3)
crit = CrossEntropyLossSum().to(dev)
inp = [t.to(dev) for t in synth.gen(B, 90, 24, seed=1, with_labels=True)]
def step():
    opt.zero_grad(); loss = crit(m(*inp[:4]), inp[4]); loss.backward(); opt.step()
for _ in range(2): step()
torch.cuda.synchronize()
if os.environ.get("TIME") == "1":         # CUDA-event time of 10 steps (no profiler)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): step()
    e1.record(); torch.cuda.synchronize()
    print(f"{bk} B={B} precision={os.environ.get('PRECISION', 'fp32')}: {e0.elapsed_time(e1) / 10:.3f} ms per train step (fwd + CE + bwd + Adam)")
    sys.exit(0)
if os.environ.get("NCU") == "1":          # ncu --profile-from-start off: one step between cudaProfilerStart / Stop
    torch.cuda.cudart().cudaProfilerStart(); step(); torch.cuda.synchronize(); torch.cuda.cudart().cudaProfilerStop()
    sys.exit(0)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=int(os.environ.get("ROWS", "25")), max_name_column_width=90))
