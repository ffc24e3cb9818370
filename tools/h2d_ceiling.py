"""Host->device ceiling of the box: cudaMemcpyAsync-only microbenchmark under torchrun.  Every rank owns a
pinned 315 MB buffer (= the fp32 images of one config-2 step) and a device buffer; phase 1: rank 0 copies alone, phase 2: ranks
0..k-1 copy concurrently for k = 1, 2, 4, 8.  Prints the aggregate GB/s per k (rank 0).  No kernels, no collectives in the
timed regions (a barrier brackets them).
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/h2d_ceiling.py"""
import os, time
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n = 16 * 3 * 1280 * 1280
host = torch.empty(n, dtype=torch.float32).pin_memory()
host.fill_(1.0)
host_u8 = torch.empty(n, dtype=torch.uint8).pin_memory()
d32, d8 = torch.empty(n, dtype=torch.float32, device=dev), torch.empty(n, dtype=torch.uint8, device=dev)
REPS = 20


def run(active, src, dst):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if active:
        for _ in range(REPS):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt if active else 0.0], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


for name, src, dst in (("fp32 images (315 MB/step/GPU)", host, d32), ("uint8 images (79 MB/step/GPU)", host_u8, d8)):
    for k in [k for k in (1, 2, 4, 8) if k <= world]:
        run(rank < k, src, dst)                                   # warm-up
        dt = run(rank < k, src, dst)
        if rank == 0:
            gb = k * REPS * src.numel() * src.element_size() / 1e9
            print(f"{name}: {k} GPU(s) copying concurrently: {gb / dt:7.1f} GB/s aggregate, {gb / dt / k:6.1f} GB/s per GPU, "
                  f"= {k * REPS * 16 / dt:9.0f} pages/s of input")
if world > 1:
    dist.destroy_process_group()
