#!/usr/bin/env python
"""bench.py - webpages/sec of the CoVA per-webpage hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--config 2|3|4|5] [--no-others]

A step = one pass of the hot path over one batch of synthetic pages.  `--config` selects the BASELINE.json config whose
shape the headline `value` is measured on (default 2 = configs[1], the config the metric is quoted on):

  2  batch=16 pages/GPU of 1280x1280, N=90 boxes, K=24, ResNet-18 backbone, inference (forward)
  3  batch=64 pages/GPU, N=90, K=24, ResNet-50 backbone, TRAIN step: forward + CE(sum) + backward + gradient
     all-reduce(SUM, NCCL, when N > 1) + Adam                       (`main.py:133-139`, `train.py:45-60`)
  4  batch=32 pages/GPU (= 256 over 8 GPUs), otherwise as 3
  5  batch=16 pages/GPU, N=300 boxes, K=48, 2-head GAT, ResNet-50 backbone, inference

N>1 = weak scaling: every rank runs the batch on its own GPU; inference has no data-path collective (pages are
independent), the train configs all-reduce one flat fp32 gradient bucket per step.  value = pages of all ranks /
max-over-ranks device time.  The default run (config 2) also measures configs 4, 5 and - when it fits - 3 after the headline
and reports them under "other_configs" (so the driver's N=1 and 1..8 scaling runs see the train step and its all-reduce);
the headline fields are always config 2's.  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import time
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

IMG = 1280
CONFIGS = {
    2: dict(B=16, N=90, K=24, heads=1, backbone="resnet18", mode="infer",
            workload="configs[1]: batch=16 synthetic 1280x1280 pages per GPU, N=90 boxes, K=24, ResNet-18 backbone, inference"),
    3: dict(B=64, N=90, K=24, heads=1, backbone="resnet50", mode="train", precision="bf16",
            workload="configs[2]: batch=64 synthetic 1280x1280 pages per GPU, N=90, K=24, ResNet-50 backbone bf16, train step "
                     "(forward + CE(sum) + backward + gradient all-reduce + Adam)"),
    4: dict(B=32, N=90, K=24, heads=1, backbone="resnet50", mode="train", precision="bf16",
            workload="configs[3]: batch=256 over 8 GPUs = 32 synthetic 1280x1280 pages per GPU, N=90, K=24, ResNet-50 backbone, "
                     "train step with the NCCL gradient all-reduce(SUM)"),
    5: dict(B=16, N=300, K=48, heads=2, backbone="resnet50", mode="infer",
            workload="configs[4] stress: batch=128 over 8 GPUs = 16 synthetic 1280x1280 pages per GPU, N=300 boxes, K=48, "
                     "2-head GAT, ResNet-50 backbone, inference"),
}
TRAIN_DTYPES = {"bf16": "bf16 maps / gradient maps, one bf16 tensor-core product (fp32 accumulate); fp32 statistics, parameters, weight gradients",
                "fp32": "fp16x3 forward / dgrad / wgrad convolutions (split-fp16, fp32 accumulate; fp32-parity), fp32 maps"}
DTYPES = {"fp32": "bf16x3 (split-bf16, fp32 accumulate; fp32-parity)", "bf16": "bf16",
          "fp16": "fp16 (one product, fp32 accumulate)", "fp32x": "fp16x3 (split-fp16, fp32 accumulate; fp32-parity)"}


def get_config(cid):
    """The selected config with the COVA_B200_* side-experiment overrides applied (a run with any override says so)."""
    c = dict(CONFIGS[cid])
    c["id"] = cid
    ov = {"B": "COVA_B200_B", "N": "COVA_B200_N", "K": "COVA_B200_K", "heads": "COVA_B200_HEADS"}
    c["overridden"] = False
    for k, env in ov.items():
        if env in os.environ:
            c[k] = int(os.environ[env]); c["overridden"] = True
    if "COVA_B200_BACKBONE" in os.environ:
        c["backbone"] = os.environ["COVA_B200_BACKBONE"]; c["overridden"] = True
    if c["overridden"]:
        c["workload"] = "side experiment (COVA_B200_* overrides on config %d), %s" % (cid, c["mode"])
    return c


def peaks():
    """Measured roofline denominators (driver-written MEASURED_PEAKS.json) or the profiling guide's fallback."""
    fallback = dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if not os.path.exists(p):
        return fallback
    try:
        d = json.load(open(p))

        def pick(*names):
            for n in names:
                if n in d and isinstance(d[n], (int, float)):
                    return float(d[n])
            for k, v in d.items():                      # tolerate renamed keys: match on substrings
                if isinstance(v, (int, float)) and all(t in k.lower() for t in names[0].split("_")[:1]):
                    return float(v)
            return None
        hbm = pick("hbm_gbs", "hbm_gb_s", "hbm_GBs", "hbm")
        burst = pick("bf16_tflops", "bf16_tflops_burst", "bf16")
        sust = pick("bf16_tflops_sustained", "bf16_sustained_tflops") or burst
        if hbm and burst:
            return dict(hbm=hbm, tf_burst=burst, tf_sust=sust, src="measured")
    except Exception:
        pass
    return fallback


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, dev_index):
        self.idx, self.proc = dev_index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out = self.proc.communicate(timeout=5)[0]
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_model(dev, cfg, precision=None):
    import cova_b200.synth as synth
    from cova_b200.models import CoVA
    m = CoVA((3, 3), IMG, 4, True, 384, 32, 0, 0.2, None, pretrained=False, backbone=cfg["backbone"], n_heads=cfg["heads"],
             engine=os.environ.get("COVA_B200_ENGINE", "tcgen05"),
             precision=precision or os.environ.get("COVA_B200_PRECISION") or cfg.get("precision", "fp32"))
    m.load_state_dict(synth.make_state_dict(123, backbone=cfg["backbone"], n_heads=cfg["heads"]), strict=True)
    return m.to(dev)


def roi_bytes(bboxes, C=64, P=3, scale=0.25, Hf=320, Wf=320):
    """Algorithmic RoIPool bytes: sum of crop areas x C x 4 read + T x C x P^2 x 4 written."""
    b = bboxes.numpy().astype(np.float32)
    rnd = lambda v: np.sign(v) * np.floor(np.abs(v.astype(np.float64)) + 0.5)
    sw, sh, ew, eh = (rnd(b[:, i] * np.float32(scale)) for i in (1, 2, 3, 4))
    cw = np.clip(ew + 1, 0, Wf) - np.clip(sw, 0, Wf)
    ch = np.clip(eh + 1, 0, Hf) - np.clip(sh, 0, Hf)
    return float((np.maximum(cw, 0) * np.maximum(ch, 0)).sum() * C * 4 + len(b) * C * P * P * 4)


# ---------------------------------------------------------------------------------------------------------------------
# Algorithmic work of one ABI call, computed from the call's own arguments (positions: cova_b200/_lib.py SIGNATURES):
# ("tensor", FLOP) for the contraction kernels, ("hbm", bytes) for the streaming ones (SURVEY 8(d) per-unit figures).
def _alg_table(roi_b):
    i = int
    es = lambda code: 2 if int(code) == 1 else 4
    return {
        "cova_stem_fwd": lambda a: ("tensor", 2.0 * 64 * 147 * i(a[2]) * (i(a[3]) // 2) * (i(a[4]) // 2)),
        "cova_stem_conv_raw_fwd": lambda a: ("tensor", 2.0 * 64 * 147 * i(a[2]) * (i(a[3]) // 2) * (i(a[4]) // 2)),
        "cova_conv3x3_bn_act_fwd": lambda a: ("tensor", 2.0 * 9 * i(a[6]) * i(a[7]) * i(a[3]) * i(a[4]) * i(a[5])),
        "cova_conv3x3_wgrad": lambda a: ("tensor", 2.0 * 9 * 64 * 64 * i(a[4]) * i(a[5]) * i(a[6])),
        "cova_stem_wgrad": lambda a: ("tensor", 2.0 * 64 * 147 * i(a[2]) * (i(a[3]) // 2) * (i(a[4]) // 2)),
        "cova_conv1x1_wgrad": lambda a: ("hbm", (2.0 if i(a[7]) == 1 else 4.0) * i(a[4]) * (i(a[5]) + i(a[6]))),
        # 1x1 convs over split planes: 4 B/elt in (hi+lo), 4 B/elt out, + residual planes
        "cova_conv1x1_bn_act_fwd": lambda a: ("hbm", float(i(a[2])) * (4 * i(a[3]) + 4 * i(a[4]) + (4 * i(a[4]) if a[8] else 0))),
        "cova_roi_fwd": lambda a: ("hbm", roi_b * i(a[4]) / 64.0) if i(a[10]) == 0 else
                                  ("hbm", float(i(a[6])) * i(a[7]) * i(a[8]) * 16 * i(a[4]) * 4 + float(i(a[6])) * i(a[4]) * i(a[7]) * i(a[8]) * 4),
        "cova_gat_fwd": lambda a: ("hbm", float(i(a[8])) * i(a[9]) * i(a[10]) * 4 + float(i(a[8])) * i(a[10]) * 4 + float(i(a[8])) * i(a[9]) * 8),
        "cova_gat_multihead_fwd": lambda a: ("hbm", float(i(a[7])) * i(a[8]) * i(a[2]) * i(a[3]) * 4 + float(i(a[7])) * i(a[2]) * i(a[3]) * 4
                                             + float(i(a[7])) * i(a[8]) * 8),
        "cova_linear_fwd": lambda a: ("tensor", 2.0 * i(a[2]) * i(a[3]) * i(a[5])),
        "cova_bbox_enc_fwd": lambda a: ("hbm", float(i(a[1])) * (20 + 4 * i(a[6]))),
        "cova_bn_train_stats": lambda a: ("hbm", 4.0 * i(a[1]) * i(a[2])),
        "cova_bn_act_fwd": lambda a: ("hbm", 4.0 * i(a[1]) * i(a[2]) * (2 + (1 if a[7] else 0) + (1 if a[10] else 0) + (0.0625 if a[13] else 0))),
        "cova_bn_act_bwd": lambda a: ("hbm", 4.0 * i(a[3]) * i(a[4]) * (2 * (2 + (1 if a[2] else 0)) + 1 + (1 if a[12] else 0))),
        # two passes over (x, dy [, res]) + the scaled split planes of dx (4 B/elt) [+ dres]
        "cova_bn_act_bwd_planes": lambda a: ("hbm", 4.0 * i(a[3]) * i(a[4]) * (2 * (2 + (0.0625 if a[20] else (1 if a[2] else 0))) + 1 + (1 if a[17] else 0))),
        "cova_maxpool3x3s2_fwd": lambda a: ("hbm", 4.0 * i(a[1]) * i(a[2]) * i(a[3]) * i(a[4]) * (1 + 0.25 * (1.25 + (1 if a[7] else 0)))),
        "cova_maxpool3x3s2_bwd": lambda a: ("hbm", 4.0 * i(a[2]) * i(a[3]) * i(a[4]) * i(a[5]) * (1 + 0.25 * 1.25)),
        # typed (bf16 training mode) passes: element sizes from the dtype codes (0 = fp32, 1 = bf16)
        "cova_bn_train_stats_t": lambda a: ("hbm", float(es(a[1])) * i(a[2]) * i(a[3])),
        "cova_bn_act_fwd_t": lambda a: ("hbm", float(i(a[2])) * i(a[3]) * (es(a[1]) * (1 + (1 if a[8] else 0)) + es(a[11]) + (0.125 if a[12] else 0))),
        "cova_bn_act_bwd_t": lambda a: ("hbm", float(i(a[5])) * i(a[6]) * (2 * (es(a[1]) + es(a[4]) + (0.125 if a[17] else (es(a[4]) if a[3] else 0)))
                                                                         + es(a[4]) * (1 + (1 if a[14] else 0)))),
        "cova_maxpool3x3s2_fwd_t": lambda a: ("hbm", float(es(a[1])) * i(a[2]) * i(a[3]) * i(a[4]) * i(a[5]) * 1.25 + 0.25 * i(a[2]) * i(a[3]) * i(a[4]) * i(a[5])),
        "cova_maxpool3x3s2_bwd_t": lambda a: ("hbm", float(es(a[2])) * i(a[3]) * i(a[4]) * i(a[5]) * i(a[6]) * 1.25 + 0.25 * i(a[3]) * i(a[4]) * i(a[5]) * i(a[6])),
        "cova_stem_conv_raw_fwd_bf16": lambda a: ("tensor", 2.0 * 64 * 147 * i(a[2]) * (i(a[3]) // 2) * (i(a[4]) // 2)),
        # 1x1 convolutions of the training path: planes in (4 B/elt split, 2 B/elt bf16), rows out (fp32 / bf16)
        "cova_conv1x1_raw_fwd": lambda a: ("hbm", float(i(a[3])) * (i(a[4]) + i(a[5])) * (2 if i(a[2]) == 1 else 4)),
        "cova_split_planes": lambda a: ("hbm", 8.0 * i(a[1])),
        "cova_split_planes_scaled": lambda a: ("hbm", 12.0 * i(a[1])),
        "cova_adam_step": lambda a: ("hbm", 28.0 * i(a[4])),
    }


def time_stages(step_fn, iters, roi_b):
    """Per-ABI-call CUDA-event timing on the launching stream, inside bench.py (not under a profiler).
    Returns {name: (ms per step, launches per step, kind, algorithmic work per step)} in first-call order."""
    from cova_b200 import ops
    alg = _alg_table(roi_b)
    names, evs, works = [], [], []
    orig = ops._call

    def timed(name, *args):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        orig(name, *args)
        b.record()
        names.append(name); evs.append((a, b))
        try:
            works.append(alg[name](args) if name in alg else (None, None))
        except Exception:
            works.append((None, None))

    ops._call = timed
    try:
        for _ in range(iters):
            step_fn()
        torch.cuda.synchronize()
    finally:
        ops._call = orig
    per, order = {}, []
    for n, (a, b), (kind, w) in zip(names, evs, works):
        if n not in per:
            per[n] = [0.0, 0, kind, 0.0]; order.append(n)
        per[n][0] += a.elapsed_time(b); per[n][1] += 1
        if w is not None:
            per[n][3] += w
    return {n: (per[n][0] / iters, per[n][1] // iters, per[n][2], per[n][3] / iters) for n in order}


def kernel_table(stages, pk, exec_factor, traffic):
    kernels = {}
    for n, (msn, cnt, kind, work) in stages.items():
        d = {"ms_per_step": round(msn, 4), "launches_per_step": cnt, "bound": kind}
        if kind is not None and msn > 0:
            ach = work / (msn / 1e3) / (1e12 if kind == "tensor" else 1e9)
            peak = pk["tf_sust"] if kind == "tensor" else pk["hbm"]
            d.update({"achieved": round(ach, 2), "unit": "TFLOP/s" if kind == "tensor" else "GB/s", "frac": round(ach / peak, 4),
                      "traffic": traffic.get(n)})
            if kind == "tensor" and n in exec_factor:
                d["executed"] = round(ach * exec_factor[n], 2)
                d["executed_frac"] = round(ach * exec_factor[n] / peak, 4)
            if kind == "hbm" and traffic.get(n) and cnt:
                # algorithmic bytes count every crop / neighbour row as read from HBM; overlapping boxes and neighbour windows
                # are served by L2 / shared memory, so the DRAM traffic of the same launch (ncu) is reported beside it
                d["dram_gb_s"] = round(traffic[n] / (msn / cnt / 1e3) / 1e9, 1)
                d["dram_frac"] = round(d["dram_gb_s"] / peak, 4)
            if n == "cova_gat_fwd":
                d["note"] = ("L2 / shared-memory resident gather: unique bytes are T*(H+4)*4 (2.2 MB), every neighbour row is staged once "
                             "per CTA; ~3 us of the event-timed figure is launch latency of a ~11.5 us kernel (ncu duration)")
            if n == "cova_roi_fwd":
                d["note"] = "crops of nested / overlapping boxes hit L2: the DRAM-side floor is the feature map once (see dram_gb_s)"
        kernels[n] = d
    return kernels


def roofline_of(stages, pk, exec_factor, traffic):
    """`roofline` object for the dominant kernel (by measured time) of a step."""
    known = {n: v for n, v in stages.items() if v[2] is not None}
    if not known:
        return None
    dom = max(known, key=lambda n: known[n][0])
    msn, n_l, kind, work = known[dom]
    ach = (work / n_l) / (msn / n_l / 1e3) / (1e12 if kind == "tensor" else 1e9)
    peak = pk["tf_sust"] if kind == "tensor" else pk["hbm"]
    r = {"kernel": dom, "bound": "tensor" if kind == "tensor" else "hbm", "achieved": round(ach, 3), "peak": peak,
         "unit": "TFLOP/s" if kind == "tensor" else "GB/s", "frac": round(ach / peak, 4), "traffic": traffic.get(dom),
         "algorithmic_per_launch": work / n_l, "launches_per_step": n_l,
         "peak_source": pk["src"] + (" (sustained bf16: kernel timed inside the step)" if kind == "tensor" else " (HBM copy)"),
         "share_of_step": round(msn / sum(v[0] for v in stages.values()), 3)}
    if kind == "tensor" and dom in exec_factor:
        r["executed"] = round(ach * exec_factor[dom], 3)
        r["executed_frac"] = round(ach * exec_factor[dom] / peak, 4)
        r["executed_note"] = ("bf16/fp16 tensor FLOP/s issued: %.2f MMA FLOPs per algorithmic fp32 FLOP (split 3-product "
                              "fp32-parity scheme); power-capped like the cuBLAS peak" % exec_factor[dom])
    return r


# ------------------------------------------------------------------------------------------------ reference (CPU) arm
def _ref_model(cfg, train):
    """The reference's own `models.CoVA` from oracle/_ref (kind "reference") or None when it is not staged."""
    try:
        from oracle import ref_loader
        if not ref_loader.available():
            return None
        import cova_b200.synth as synth
        sd = synth.make_state_dict(123, backbone=cfg["backbone"], n_heads=cfg["heads"])
        m = ref_loader.build_model(cfg["backbone"], IMG, cfg["heads"], 0.2, sd)
        return m.train() if train else m.eval()
    except Exception as e:   # pragma: no cover
        print("bench: reference model unavailable (%s), timing the oracle port" % e, file=sys.stderr)
        return None


def reference_step_fn(cfg, pages, threads):
    """One step of the reference's CPU implementation of the selected config on `pages` pages: the reference's own
    models.py (oracle/_ref) when staged, else the oracle port (torch ATen on CPU, operator for operator)."""
    import cova_b200.synth as synth
    torch.set_num_threads(threads)
    train = cfg["mode"] == "train"
    inp = synth.gen(pages, cfg["N"], cfg["K"], seed=1, with_labels=train)
    m = _ref_model(cfg, train)
    if m is not None:
        if train:                                                     # main.py:133-139 + train.py:45-60
            opt = torch.optim.Adam(m.parameters(), lr=5e-4, weight_decay=1e-3)
            crit = torch.nn.CrossEntropyLoss(reduction="sum")

            def step():
                opt.zero_grad()
                out = m(*inp[:4])
                loss = crit(out, inp[4])
                loss.backward()
                opt.step()
                return float(loss.item())
        else:
            def step():
                with torch.no_grad():
                    return m(*inp)
        return step, "reference"
    from oracle import torch_port
    sd = synth.make_state_dict(123, backbone=cfg["backbone"], n_heads=cfg["heads"])
    if train:
        raise RuntimeError("the oracle port has no train step; stage oracle/_ref (python oracle/build_ref.py)")
    return (lambda: torch_port.forward(sd, *inp)), "port"


def cpu_baseline(cfg, threads, budget_s=12.0):
    """The reference on the box's host cores for a bounded sample of the workload (rank 0, N=1 only)."""
    pages = 4 if cfg["mode"] == "infer" else 2
    step, kind = reference_step_fn(cfg, pages, threads)
    step()                                   # warm-up
    best, reps, t_all = 1e30, 0, time.perf_counter()
    while reps < 3 and (reps == 0 or time.perf_counter() - t_all < budget_s):
        t0 = time.perf_counter()
        step()
        best = min(best, time.perf_counter() - t0)
        reps += 1
    what = "train steps" if cfg["mode"] == "train" else "forwards"
    return pages / best, kind, (f"{pages} pages of the workload (N={cfg['N']}, K={cfg['K']}, {cfg['backbone']}), best of {reps} "
                                f"{what} after 1 warm-up, {threads} threads")


def run_reference(args, rank, cfg):
    """`--impl reference`: the reference's CPU implementation of the path on the box's host cores, all threads.
    Config 2 runs the same batch (16 pages) and the same --steps / --warmup as the native arm; the train configs run a
    bounded sample (2 pages per step) because one 64-page ResNet-50 train step takes minutes on a CPU."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    pages = cfg["B"] if cfg["mode"] == "infer" else 2
    step, kind = reference_step_fn(cfg, pages, threads)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    v = pages * args.steps / dt
    sample = f"{pages} pages/step x {args.steps} steps of the workload on {threads} threads ({cfg['mode']})"
    print(json.dumps({
        "impl": "reference", "metric": "webpages/sec", "value": v, "unit": "pages/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["workload"], "config_id": cfg["id"], "pages_per_step": pages,
                   "implementation": "the reference's own models.py (oracle/_ref) on torch CPU" if kind == "reference"
                   else "oracle/torch_port.py (the reference forward, operator for operator, on torch CPU)"},
        "cpu_baseline": {"value": v, "unit": "pages/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": "pages/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------ native arm
class Runner:
    """One config on this rank's GPU: resident step, end-to-end step (host buffers), per-kernel breakdown."""

    def __init__(self, cfg, dev, rank, world, dist):
        import cova_b200.synth as synth
        self.cfg, self.dev, self.rank, self.world, self.dist = cfg, dev, rank, world, dist
        self.train = cfg["mode"] == "train"
        self.model = build_model(dev, cfg)
        self.model.train() if self.train else self.model.eval()
        self.inp = synth.gen(cfg["B"], cfg["N"], cfg["K"], seed=1 + rank, with_labels=self.train)
        self.dinp = [t.to(dev) for t in self.inp]
        self.pinned = [t.pin_memory() for t in self.inp]
        self.T = cfg["B"] * cfg["N"]
        self.logits_host = torch.empty((self.T, 4), dtype=torch.float32).pin_memory()
        self.loss_host = torch.empty(1, dtype=torch.float32).pin_memory()
        if self.train:
            from cova_b200.train_ops import CrossEntropyLossSum, FlatAdam
            self.crit = CrossEntropyLossSum()
            self.opt = FlatAdam(self.model.parameters(), lr=5e-4, weight_decay=1e-3)       # main.py:133-135

    # -- steps
    def _train_on(self, d):
        """`train.py:45-60` with the native tail; the gradient all-reduce(SUM) runs on the flat bucket when world > 1."""
        self.opt.zero_grad()
        out = self.model(d[0], d[1], d[2], d[3])
        loss = self.crit(out, d[4])
        loss.backward()
        self.opt.allreduce_grads()
        self.opt.step()
        return loss

    def step_resident(self):
        if self.train:
            return self._train_on(self.dinp)
        with torch.no_grad():
            return self.model(*self.dinp)

    def local_step_no_collective(self):
        """Warm-up probe: the whole step without the all-reduce (used to agree across ranks that the config fits)."""
        if self.train:
            self.opt.zero_grad()
            loss = self.crit(self.model(*self.dinp[:4]), self.dinp[4])
            loss.backward()
            self.opt.step()
            return loss
        return self.step_resident()

    def run_e2e(self, steps, host=None):
        """Public-API path with HOST buffers: every step copies its own pinned inputs to the device
        (cova_b200.pipeline.prefetch: side-stream H2D one batch ahead of the compute) and reads its result back
        (the logits for inference, the loss for a train step - `train.py:57`)."""
        from cova_b200.pipeline import prefetch
        host = self.pinned if host is None else host
        if self.train:
            for d in prefetch((host for _ in range(steps)), self.dev):
                self.loss_host.copy_(self._train_on(d).detach().reshape(1), non_blocking=True)
        else:
            with torch.no_grad():
                for d in prefetch((host for _ in range(steps)), self.dev):
                    self.logits_host.copy_(self.model(*d), non_blocking=True)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def timed(self, fn, steps, whole=False):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        e0.record()
        if whole:
            fn(steps)
        else:
            for _ in range(steps):
                fn()
        e1.record()
        self.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(ms, op=self.dist.ReduceOp.MAX)
        return float(ms.item())

    def h2d_bytes(self, host=None):
        return sum(t.numel() * t.element_size() for t in (self.pinned if host is None else host))

    def exec_factors(self):
        m = self.model
        if m.engine != "tcgen05":
            return {}
        split = (m.precision in ("fp32", "fp32x") or self.train or m.backbone == "resnet50") and not (self.train and m.precision == "bf16")
        f = 3.0 if split else 1.0
        return {"cova_conv3x3_bn_act_fwd": f, "cova_stem_fwd": f * 224.0 / 147.0, "cova_stem_conv_raw_fwd": f * 224.0 / 147.0,
                "cova_stem_conv_raw_fwd_bf16": 224.0 / 147.0,
                "cova_linear_fwd": 3.0, "cova_conv3x3_wgrad": 3.5, "cova_stem_wgrad": 14 * 2.0 * 128 * 32 * 16 / (16 * 2.0 * 64 * 147)}


def measure(cfg, dev, rank, world, dist, args, sample_clocks, full):
    """Times one config.  `full` = the headline treatment (three e2e repetitions, uint8 e2e, fp16 side mode)."""
    from cova_b200 import ops
    r = Runner(cfg, dev, rank, world, dist)
    steps = args.steps if full else max(3, min(args.steps, 10 if cfg["mode"] == "infer" else 5))
    warm = args.warmup if full else 3
    # agree across ranks that the config fits before any collective is issued inside a step
    ok = torch.ones(1, device=dev)
    try:
        r.local_step_no_collective()
        torch.cuda.synchronize()
    except Exception as e:
        ok.zero_()
        err = "%s: %s" % (type(e).__name__, str(e).splitlines()[0][:160])
    if world > 1:
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if float(ok.item()) == 0.0:
        del r
        torch.cuda.empty_cache()
        return {"skipped": locals().get("err", "another rank could not run this config")}
    for _ in range(warm):
        r.step_resident()
    sampler = ClockSampler(dev.index)
    if rank == 0 and sample_clocks:
        sampler.start()
    ops.launch_count = 0
    ms = r.timed(r.step_resident, steps)
    launches = ops.launch_count
    clocks = sampler.stop() if (rank == 0 and sample_clocks) else None
    r.run_e2e(3)
    # each repetition times exactly `steps` steps; the best of the repetitions is reported (a PCIe / host hiccup in one
    # repetition was observed to triple a 40 ms measurement) and all are listed under "reps_ms_per_step"
    reps_e2e = [r.timed(r.run_e2e, steps, whole=True) for _ in range(3 if full else 1)]
    ms_e2e = min(reps_e2e)
    out = dict(cfg=cfg, steps=steps, warmup=warm, ms=ms, launches=launches, clocks=clocks, ms_e2e=ms_e2e, reps_e2e=reps_e2e,
               h2d=r.h2d_bytes(), d2h=(4 if r.train else r.logits_host.numel() * 4), model=r.model)
    # SURVEY 8(f) N1 (optional input format): the same pages as raw uint8 pixels, converted v/255 inside the stem
    pinned_u8 = [(r.inp[0] * 255).round().to(torch.uint8).pin_memory()] + r.pinned[1:]
    r.run_e2e(2, pinned_u8)
    reps_u8 = [r.timed(lambda k: r.run_e2e(k, pinned_u8), steps, whole=True) for _ in range(3 if full else 1)]
    out.update(ms_e2e_u8=min(reps_u8), reps_u8=reps_u8, h2d_u8=r.h2d_bytes(pinned_u8))
    out["ms_train_fp32"] = None
    if r.train and r.model.precision == "bf16" and world == 1:
        # side measurement (single rank only: no collective inside a try block): the same train step in the fp32-parity mode
        torch.cuda.empty_cache()
        r32 = Runner(dict(cfg, precision="fp32"), dev, rank, world, dist)
        try:
            for _ in range(2):
                r32.step_resident()
            n32 = 3
            out["ms_train_fp32"] = r32.timed(r32.step_resident, n32) / n32
        except Exception as e:   # pragma: no cover  (memory: the fp32 maps of B=64 need ~100 GB)
            out["ms_train_fp32"] = None
            out["train_fp32_error"] = "%s: %s" % (type(e).__name__, str(e).splitlines()[0][:120])
        del r32
        torch.cuda.empty_cache()
    out["sustained"] = None
    if full and not r.train:
        # side measurement: the same resident step over >= 1.5 s.  The K-step region above is a few tens of milliseconds, in which
        # the SM clock has not yet come down to what the 1000 W cap allows under the tensor-core convolutions.
        n_sus = max(steps, int(np.ceil(1500.0 / max(ms / steps, 1e-3))))
        s2 = ClockSampler(dev.index)
        if rank == 0 and sample_clocks:
            s2.start()
        ms_sus = r.timed(r.step_resident, n_sus)
        out["sustained"] = {"steps": n_sus, "ms": ms_sus, "clocks": s2.stop() if (rank == 0 and sample_clocks) else None}
    out["ms_fp16"] = None
    if full and not r.train and r.model.engine == "tcgen05" and r.model.precision == "fp32" and cfg["backbone"] == "resnet18":
        # side measurement (not the headline): the one-product fp16 mode on the same inputs - logits within 5-7e-4 of the
        # live-reference fixtures, i.e. inside the 1e-3 bar but without the 50x margin of the fp32-parity mode
        m16 = build_model(dev, cfg, precision="fp16").eval()

        def step16():
            with torch.no_grad():
                return m16(*r.dinp)
        for _ in range(warm):
            step16()
        out["ms_fp16"] = r.timed(step16, steps)
        del m16
    if rank == 0:
        roi_b = roi_bytes(r.inp[1])
        # rank 0 alone times the stages: a train step must not issue its all-reduce here (the other ranks have moved on)
        stage_step = r.local_step_no_collective if (r.train and world > 1) else r.step_resident
        out["stages"] = time_stages(stage_step, max(2, min(steps, 10 if not r.train else 3)), roi_b)
        out["exec_factor"] = r.exec_factors()
    out["precision"] = r.model.precision
    out["engine"] = r.model.engine
    out["pages_per_gpu"] = cfg["B"]
    del r
    torch.cuda.empty_cache()
    return out


def summarize_other(res, world, pk):
    """Compact entry of a non-headline config for the "other_configs" object."""
    if "skipped" in res:
        return res
    cfg = res["cfg"]
    pages = cfg["B"] * world * res["steps"]
    d = {"workload": cfg["workload"], "mode": cfg["mode"], "value": pages / (res["ms"] / 1e3), "unit": "pages/s",
         "ms_per_step": res["ms"] / res["steps"], "steps": res["steps"], "warmup": res["warmup"],
         "pages_per_gpu_per_step": cfg["B"], "n_gpus": world, "gpu_launches": res["launches"],
         "collective": ("NCCL all-reduce(SUM) of the flat fp32 gradient bucket inside every step" if (cfg["mode"] == "train" and world > 1)
                        else "none (single rank)" if cfg["mode"] == "train" else "none (pages are independent)"),
         "dtype": (TRAIN_DTYPES.get(res["precision"], TRAIN_DTYPES["fp32"]) if cfg["mode"] == "train"
                   else DTYPES.get(res["precision"], res["precision"])),
         "e2e": {"value": pages / (res["ms_e2e"] / 1e3), "unit": "pages/s", "h2d_bytes_per_step": res["h2d"],
                 "d2h_bytes_per_step": res["d2h"], "ms_per_step": res["ms_e2e"] / res["steps"]},
         "e2e_uint8_images": {"value": pages / (res["ms_e2e_u8"] / 1e3), "unit": "pages/s", "h2d_bytes_per_step": res["h2d_u8"]}}
    if res.get("ms_train_fp32"):
        d["fp32_parity_mode"] = {"value": cfg["B"] * world / (res["ms_train_fp32"] / 1e3), "unit": "pages/s", "ms_per_step": res["ms_train_fp32"],
                                 "note": "same train step with fp32 maps and split-fp16 three-product convolutions (every gradient within "
                                         "1.2e-5 of the live-reference fixture); side measurement, 3 steps"}
    if "stages" in res:
        d["roofline"] = roofline_of(res["stages"], pk, res["exec_factor"], {})
        tot = sum(v[0] for v in res["stages"].values())
        d["native_kernel_ms_per_step"] = round(tot, 3)
        d["top_kernels"] = {n: round(v[0], 3) for n, v in sorted(res["stages"].items(), key=lambda kv: -kv[1][0])[:6]}
    return d


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS))
    ap.add_argument("--no-others", action="store_true", help="measure only the selected config")
    ap.add_argument("--skip-cpu", action="store_true", help="side experiments: do not time the CPU reference")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = get_config(args.config)
    if args.impl == "reference":
        return run_reference(args, rank, cfg)

    t_start = time.perf_counter()
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    from cova_b200.pipeline import bind_to_gpu_numa
    numa_cores = bind_to_gpu_numa(local) if world > 1 else None     # pinned buffers on the GPU's own NUMA node
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))   # a wedged collective fails fast

    head = measure(cfg, dev, rank, world, dist, args, sample_clocks=True, full=True)
    if "skipped" in head:
        if rank == 0:
            print(json.dumps({"metric": "webpages/sec", "value": None, "unit": "pages/s", "n_gpus": world,
                              "config": {"workload": cfg["workload"]}, "error": head["skipped"]}))
        if world > 1:
            dist.destroy_process_group()
        return
    others = {}
    if not args.no_others and not cfg["overridden"]:
        # the remaining configs, smallest memory footprint first; each is skipped (on every rank alike) once the run has
        # used its time budget, so the default invocation still ends within minutes
        for cid in [c for c in (5, 4, 3) if c != args.config] + ([2] if args.config != 2 else []):
            go = torch.tensor([1.0 if time.perf_counter() - t_start < 170.0 else 0.0], device=dev)
            if world > 1:
                dist.all_reduce(go, op=dist.ReduceOp.MIN)
            if float(go.item()) == 0.0:
                others[str(cid)] = {"skipped": "time budget of the default run used up"}
                continue
            others[str(cid)] = measure(get_config(cid), dev, rank, world, dist, args, sample_clocks=False, full=False)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    model = head["model"]
    steps = head["steps"]
    pages = cfg["B"] * world * steps
    value, e2e = pages / (head["ms"] / 1e3), pages / (head["ms_e2e"] / 1e3)
    traffic = {}
    tp = os.path.join(ROOT, "profiles", "traffic_r02e.json")    # dram__bytes per launch from the committed ncu capture
    if not os.path.exists(tp):
        tp = os.path.join(ROOT, "profiles", "traffic_r01.json")
    if os.path.exists(tp) and args.config == 2 and model.precision in ("fp32", "fp32x") and model.engine == "tcgen05" and not cfg["overridden"]:
        traffic = {k: v["dram_bytes_per_launch"] for k, v in json.load(open(tp)).items() if not k.startswith("_")}
    # Tensor-core FLOPs actually ISSUED per algorithmic FLOP: the fp32-parity mode runs 3 bf16 products per fp32
    # product (x*w ~ hi*Whi + lo*Whi + hi*Wlo); the stem additionally pads K from 147 to 7 rows x 32 (sliding-window
    # operand).  `executed` = algorithmic rate x this factor = what the tensor pipe delivers, measured against the same
    # sustained cuBLAS bf16 peak (both run at the 1000 W power cap: profiles/r01k_power_clocks.txt).
    kernels = kernel_table(head["stages"], pk, head["exec_factor"], traffic)
    roofline = roofline_of(head["stages"], pk, head["exec_factor"], traffic)

    cores = os.cpu_count() or 1
    if args.skip_cpu:
        cpu_v, cpu_kind, sample = None, "port", "skipped (--skip-cpu)"
    else:
        cpu_v, cpu_kind, sample = cpu_baseline(cfg, cores)
    agg_h2d = lambda ms, nbytes: round(nbytes * world / (ms / steps / 1e3) / 1e9, 2)
    line = {
        "metric": "webpages/sec", "value": value, "unit": "pages/s", "n_gpus": world, "steps": steps,
        "warmup": head["warmup"], "ms_per_step": head["ms"] / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None,
        "dtype": (TRAIN_DTYPES.get(model.precision, TRAIN_DTYPES["fp32"]) if cfg["mode"] == "train" else
                  DTYPES.get(model.precision, model.precision)) if model.engine == "tcgen05" else "f32",
        "data": "synthetic",
        "config": {"workload": cfg["workload"], "config_id": cfg["id"], "mode": cfg["mode"],
                   "pages_per_gpu_per_step": cfg["B"], "engine": model.engine, "precision": model.precision,
                   "backbone": model.backbone, "boxes_per_page": cfg["N"], "neighbours": cfg["K"], "gat_heads": cfg["heads"],
                   "headline": not cfg["overridden"] and args.config == 2,
                   "parity": "logits of this very batch vs the live reference: tests/golden/g_c2_r18_b16.npz "
                             "(test_forward_config2_golden, max|a-b|/max|b| <= 1e-4; bar 1e-3)" if args.config == 2 else
                             "tests/test_gpu_parity.py (live-reference fixtures of this config's shape)",
                   "cpu_affinity": None if numa_cores is None else "%d cores local to the GPU (NVML)" % len(numa_cores),
                   "l2": "inputs larger than L2 (%d MB of images per step vs 126 MB L2); no flush needed" % (cfg["B"] * 3 * IMG * IMG * 4 >> 20)},
        "clocks": head["clocks"], "gpu_launches": head["launches"],
        "e2e": {"value": e2e, "unit": "pages/s", "h2d_bytes_per_step": head["h2d"], "d2h_bytes_per_step": head["d2h"],
                "ms_per_step": head["ms_e2e"] / steps, "reps_ms_per_step": [round(x / steps, 4) for x in head["reps_e2e"]],
                "aggregate_h2d_gb_s": agg_h2d(head["ms_e2e"], head["h2d"]),
                "note": "fp32 NCHW images = the reference's input contract; bound by the host->device path: %.1f GB/s per GPU here "
                        "(PCIe Gen5 x16), and by the VM's ~180 GB/s aggregate at N >= 4 (profiles/r01l_host_topology_8gpu.txt)"
                        % (head["h2d"] / (head["ms_e2e"] / steps / 1e3) / 1e9)},
        "e2e_uint8_images": {"value": pages / (head["ms_e2e_u8"] / 1e3), "unit": "pages/s", "h2d_bytes_per_step": head["h2d_u8"],
                             "ms_per_step": head["ms_e2e_u8"] / steps,
                             "reps_ms_per_step": [round(x / steps, 4) for x in head["reps_u8"]],
                             "aggregate_h2d_gb_s": agg_h2d(head["ms_e2e_u8"], head["h2d_u8"]),
                             "note": "optional input format (SURVEY 8(f) N1; the PNGs of datasets.py:96-97 are uint8): raw "
                                     "pixels as integers in the stem kernel, 1/255 in its epilogue scale (two products): within one ulp of the split-bf16 stem output of the ToTensor path (tests/test_gpu_parity.py::test_uint8_images_equal_totensor_path); knob stem_u8_exact=1 gives the bit-identical v/255 path"},
        "fp32_parity_mode": None if not head.get("ms_train_fp32") else {
            "value": cfg["B"] * world / (head["ms_train_fp32"] / 1e3), "unit": "pages/s", "ms_per_step": head["ms_train_fp32"],
            "note": "same train step with fp32 maps and split-fp16 three-product convolutions; side measurement, 3 steps"},
        "sustained": None if not head.get("sustained") else {
            "value": cfg["B"] * world * head["sustained"]["steps"] / (head["sustained"]["ms"] / 1e3), "unit": "pages/s",
            "ms_per_step": head["sustained"]["ms"] / head["sustained"]["steps"], "steps": head["sustained"]["steps"],
            "clocks": head["sustained"]["clocks"],
            "note": "the same resident step timed over >= 1.5 s (side measurement): under the tensor-core convolutions the board "
                    "reaches its 1000 W cap and the SM clock settles below what the K-step region above sees"},
        "throughput_mode_fp16": None if head["ms_fp16"] is None else {
            "value": pages / (head["ms_fp16"] / 1e3), "unit": "pages/s", "ms_per_step": head["ms_fp16"] / steps,
            "note": "precision='fp16' (one fp16 product per MMA): max rel. error of the logits vs the live-reference "
                    "fixtures 5-7e-4 (bar 1e-3); side measurement, not the headline"},
        "roofline": roofline, "kernels": kernels,
        "cpu_baseline": {"value": cpu_v, "unit": "pages/s", "cores": cores, "kind": cpu_kind, "sample": sample},
        "other_configs": {k: summarize_other(v, world, pk) for k, v in others.items()},
        "wall_s": round(time.perf_counter() - t_start, 1),
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
