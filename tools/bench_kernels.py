"""Micro-benchmarks of the HBM-bound tail kernels (RoIPool, RoIAlign, GAT gather) at BASELINE config 2 and config 5
shapes: CUDA events on the launching stream, L2 flushed between iterations (256 MB scratch write), algorithmic
bytes per SURVEY.md 8(d).  Output -> profiles/."""
import json, os, sys, warnings
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import bench
import cova_b200.synth as synth
from cova_b200 import ops

dev = torch.device("cuda", 0)
HBM = bench.peaks()["hbm"]
scratch = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        scratch.zero_()                      # flush L2 (126 MB)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def timeit_rot(fns, iters=24):
    """Same, but instead of dirtying L2 with a 256 MB write (whose write-back then competes with a 10-us kernel for
    DRAM), rotate over input sets whose total size exceeds L2 ("inputs larger than L2"): every launch sees cold, clean
    lines - the condition inside the real step, where the previous kernel's output is what sits in L2."""
    for f in fns:
        f()
    ts = []
    for i in range(iters):
        f = fns[i % len(fns)]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def report(name, ms, bytes_alg):
    gbs = bytes_alg / ms / 1e6
    print(f"{name:58s} {ms*1e3:9.1f} us   {bytes_alg/1e6:9.1f} MB algorithmic   {gbs:8.0f} GB/s = {gbs/HBM:5.1%} of measured HBM copy peak")


for (B, N, K, C, tag) in [(16, 90, 24, 64, "config 2 (B=16, N=90, K=24, C=64)"), (16, 300, 48, 64, "config 5 per GPU (B=16, N=300, K=48, C=64)")]:
    _, bboxes, _, ci = synth.gen(B, N, K, seed=1, img=64)          # boxes/ids only (tiny images)
    bboxes = bboxes.clone(); g = torch.Generator().manual_seed(1)
    T = B * N
    w = 16 + torch.rand(T, generator=g) * (512 - 16); h = 8 + torch.rand(T, generator=g) * (256 - 8)
    x1 = torch.rand(T, generator=g) * (1280 - w); y1 = torch.rand(T, generator=g) * (1280 - h)
    bboxes[:, 1], bboxes[:, 2], bboxes[:, 3], bboxes[:, 4] = x1, y1, x1 + w, y1 + h
    fm = torch.randn(B, 320, 320, C, device=dev)
    bb = bboxes.to(dev)
    out = torch.empty((T, C * 9 + 416), device=dev)
    for rs in (0, 1):
        ops.set_knob("roi_rowsplit", rs)
        ms = timeit(lambda: ops.roi_fwd(fm, bb, (3, 3), 0.25, out))
        report(f"RoIPool rowsplit={rs}  {tag}", ms, bench.roi_bytes(bboxes, C=C))
    ops.set_knob("roi_rowsplit", -1)
    ms = timeit(lambda: ops.roi_fwd(fm, bb, (3, 3), 0.25, out, mode="align"))
    report(f"RoIAlign  {tag}", ms, T * 9 * 4 * 4 * C * 4 + T * C * 9 * 4)
    fms = [fm] + [torch.randn_like(fm) for _ in range(2)]            # 3 x 419 MB feature maps > 126 MB L2
    ms = timeit_rot([(lambda f=f: ops.roi_fwd(f, bb, (3, 3), 0.25, out, mode="align")) for f in fms])
    report(f"RoIAlign (rotating inputs)  {tag}", ms, T * 9 * 4 * 4 * C * 4 + T * C * 9 * 4)
    ms = timeit_rot([(lambda f=f: ops.roi_fwd(f, bb, (3, 3), 0.25, out)) for f in fms])
    report(f"RoIPool  (rotating inputs)  {tag}", ms, bench.roi_bytes(bboxes, C=C))
    del fms
    for heads in (1, 2):
        Hd = 384 // heads
        ext = torch.randn(T, Hd + 4, device=dev)
        cid = ci.to(dev)
        o = torch.empty((T, 992), device=dev)
        ext = torch.randn(T, heads * Hd + 4, device=dev)
        ms = timeit(lambda: ops.gat_multihead_fwd(ext, Hd, [0.1] * heads, 0.2, cid, o[:, 608:]))
        report(f"GAT gather {heads} head(s) x {Hd} (one launch)  {tag}", ms, heads * (T * K * Hd * 4 + T * Hd * 4 + T * K * 8))
        n_sets = max(3, int(160e6 // (ext.numel() * 4)) + 1)             # ext copies totalling > L2
        exts = [torch.randn_like(ext) for _ in range(n_sets)]
        outs = [torch.empty((T, 992), device=dev) for _ in range(n_sets)]
        cids = [cid.clone() for _ in range(n_sets)]
        ms = timeit_rot([(lambda e=e, oo=oo, c=c: ops.gat_multihead_fwd(e, Hd, [0.1] * heads, 0.2, c, oo[:, 608:]))
                         for e, oo, c in zip(exts, outs, cids)], iters=2 * n_sets)
        report(f"GAT gather {heads} head(s) (rotating inputs)  {tag}", ms, heads * (T * K * Hd * 4 + T * Hd * 4 + T * K * 8))
        del exts, outs, cids
