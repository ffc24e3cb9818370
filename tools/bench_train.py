"""Train-step timing (BASELINE.json configs 3-4 shapes) of the CURRENT training path: library convs / BatchNorm through
autograd + the native tail (RoIPool fwd/bwd, GAT gather fwd/bwd, CE(sum) fwd+bwd, flat Adam).  Side measurement for
BASELINE.md row 3 - not the headline metric.   BACKBONE=resnet18|resnet50  B=16  TAIL=native|torch"""
import os, sys, time, warnings
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import cova_b200.synth as synth
from cova_b200.models import CoVA
from cova_b200.train_ops import CrossEntropyLossSum, FlatAdam
import torch.distributed as dist
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:                      # data parallel (config 4 shape): one all-reduce(SUM) of the flat gradient bucket per step
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
bk = os.environ.get("BACKBONE", "resnet50")
for B in [int(x) for x in os.environ.get("B", "16").split(",")]:
    for tail in os.environ.get("TAIL", "native,torch").split(","):
        torch.cuda.empty_cache()
        m = CoVA((3, 3), 1280, 4, True, 384, 32, 0, 0.2, None, pretrained=False, backbone=bk)
        m.load_state_dict(synth.make_state_dict(123, backbone=bk), strict=True)
        m = m.to(dev).train()
        opt = (FlatAdam if tail == "native" else torch.optim.Adam)(m.parameters(), lr=5e-4, weight_decay=1e-3)
        crit = (CrossEntropyLossSum() if tail == "native" else torch.nn.CrossEntropyLoss(reduction="sum")).to(dev)
        inp = [t.to(dev) for t in synth.gen(B, 90, 24, seed=1 + rank, with_labels=True)]

        def step():
            opt.zero_grad()
            loss = crit(m(*inp[:4]), inp[4])
            loss.backward()
            if world > 1:
                if tail == "native":
                    opt.allreduce_grads()
                else:
                    for p in m.parameters():
                        dist.all_reduce(p.grad, op=dist.ReduceOp.SUM)
            opt.step()
            return loss
        try:
            for _ in range(2):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                loss = step()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            if world > 1:
                t = torch.tensor([ms], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
            if rank == 0:
                print(f"train step {bk} x{world} GPUs B={B}/GPU N=90 K=24 fp32 tail={tail}: {ms:8.1f} ms/step = {world * B / ms * 1e3:7.1f} pages/s, "
                      f"peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB, loss {float(loss):.3f}")
        except torch.cuda.OutOfMemoryError:
            print(f"train step {bk} B={B}: out of memory")
        del m, opt, inp
        torch.cuda.reset_peak_memory_stats()
