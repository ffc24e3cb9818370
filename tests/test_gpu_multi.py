"""GPU, 2 ranks over NCCL (skipped on a 1-GPU box): SURVEY.md 8(e).
 * inference: each rank's forward over its page shard (`dist.shard_batch`) equals the matching rows of the
   single-process forward over the whole batch - pages are independent units, no data-path collective;
 * training: the all-reduced (SUM) flat gradient bucket equals the sum of the per-shard gradients computed in one
   process, and every rank ends up with the same bucket (the loss is CrossEntropyLoss(reduction="sum"), main.py:139)."""
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu

_WORKER = r'''
import os, sys, warnings, torch, torch.distributed as dist
warnings.filterwarnings("ignore")
sys.path.insert(0, sys.argv[1])
import cova_b200.synth as synth
from cova_b200.dist import FlatGradBucket, shard_batch
from cova_b200.models import CoVA
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
def model():
    m = CoVA((3, 3), 256, 4, True, 384, 32, 0, 0.0, None, pretrained=False)
    m.load_state_dict(synth.make_state_dict(123), strict=True)
    return m.to(dev)
batch = synth.gen(4, 0, 8, seed=31, img=256, counts=[9, 14, 5, 12], with_labels=True)
# ---- inference: shard forward == rows of the full forward
m = model().eval()
with torch.no_grad():
    full = m(*[t.to(dev) for t in batch[:4]])
    sh = shard_batch(*batch, rank=rank, world=world)
    part = m(*[t.to(dev) for t in sh[:4]])
r0 = 0 if rank == 0 else 23
ok_inf = torch.allclose(part, full[r0:r0 + part.shape[0]], rtol=1e-5, atol=1e-5)
# ---- training: BN in eval mode (per-replica batch statistics differ by construction, SURVEY 8(e)), grads summed
crit = torch.nn.CrossEntropyLoss(reduction="sum")
m = model().eval()
bucket = FlatGradBucket(m)
bucket.zero()
with torch.enable_grad():
    crit(m(*[t.to(dev) for t in sh[:4]]), sh[4].to(dev)).backward()
bucket.allreduce_sum()
got = bucket.flat.clone()
ref = model().eval()
rb = FlatGradBucket(ref)
rb.zero()
with torch.enable_grad():
    for r in range(world):
        s = shard_batch(*batch, rank=r, world=world)
        crit(ref(*[t.to(dev) for t in s[:4]]), s[4].to(dev)).backward()
err = float((got - rb.flat).abs().max() / rb.flat.abs().max())
gathered = [torch.empty_like(got) for _ in range(world)]
dist.all_gather(gathered, got)
same = all(torch.equal(g, gathered[0]) for g in gathered)
print(f"RANK {rank} inference_ok={ok_inf} grad_rel_err={err:.2e} identical_across_ranks={same} n={got.numel()}", flush=True)
assert ok_inf and err < 1e-4 and same and got.numel() == 1616485
# ---- native tail, data parallel: shard gradient -> all-reduce(SUM) of FlatAdam's bucket -> ONE Adam launch per rank;
#      every rank must hold the parameters a single process gets from the summed gradient
from cova_b200.train_ops import CrossEntropyLossSum, FlatAdam
m = model().eval()
opt = FlatAdam(m.parameters(), lr=5e-4, weight_decay=1e-3)
ncrit = CrossEntropyLossSum()
opt.zero_grad()
with torch.enable_grad():
    ncrit(m(*[t.to(dev) for t in sh[:4]]), sh[4].to(dev)).backward()
opt.allreduce_grads()
opt.step()
ref = model().eval()
ropt = torch.optim.Adam(ref.parameters(), lr=5e-4, weight_decay=1e-3)
ropt.zero_grad()
with torch.enable_grad():
    for r in range(world):
        s = shard_batch(*batch, rank=r, world=world)
        crit(ref(*[t.to(dev) for t in s[:4]]), s[4].to(dev)).backward()
ropt.step()
flat = opt._flat[0]["p"][:opt._flat[0]["n"]]
ref_flat = torch.cat([p.detach().reshape(-1) for p in ref.parameters()])
moved = float((ref_flat - torch.cat([p.detach().reshape(-1) for p in model().parameters()])).abs().max())
bad = float(((flat - ref_flat).abs() > 1e-4).float().mean())          # elements whose gradient is atomics-order noise
gathered = [torch.empty_like(flat) for _ in range(world)]
dist.all_gather(gathered, flat.contiguous())
same_p = all(torch.equal(g, gathered[0]) for g in gathered)
print(f"RANK {rank} flat_adam moved={moved:.2e} frac_off={bad:.2e} params_identical_across_ranks={same_p}", flush=True)
assert moved > 1e-4 and bad < 2e-3 and same_p
dist.destroy_process_group()
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_nccl_sharding_and_grad_allreduce(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29741", str(script), ROOT],
                       capture_output=True, text=True, timeout=600)
    print(r.stdout[-2000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("inference_ok=True") == 2
    assert r.stdout.count("params_identical_across_ranks=True") == 2
