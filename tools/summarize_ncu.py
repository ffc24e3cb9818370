#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into tracked text files under profiles/.
  python tools/summarize_ncu.py launches gpurun_out/launches.csv profiles/r01_launches.md
  python tools/summarize_ncu.py full gpurun_out/prof.ncu-rep profiles/r01_conv_tc.md"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "launch__shared_mem_per_block_dynamic",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__average_warp"]


def launches(src, dst):
    rows = list(csv.reader(open(src)))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hi]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) > vi:
            agg.setdefault(r[ki].split("(")[0], []).append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    with open(dst, "w") as f:
        f.write("ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n\n")
        f.write("| kernel | launches | mean us | share |\n|---|---|---|---|\n")
        for n, v in agg.items():
            f.write(f"| `{n[:90]}` | {len(v)} | {sum(v) / len(v) / 1e3:.1f} | {sum(v) / tot:.3f} |\n")
    print(open(dst).read())


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"ncu --set full --clock-control none, source {src}\n\n")
        for r in rows[2:]:
            f.write("## " + r[hdr.index("Kernel Name")][:100] + "\n\n| metric | value | unit |\n|---|---|---|\n")
            for i, h in enumerate(hdr):
                if any(h.startswith(k) for k in KEYS):
                    f.write(f"| {h} | {r[i]} | {units[i]} |\n")
            f.write("\n")
    print(open(dst).read()[:3000])


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
