#!/usr/bin/env python
"""profiles/traffic_rNN.json from an `ncu --set full` capture of ONE bench step: dram__bytes_read.sum +
dram__bytes_write.sum per launch, averaged per ABI entry point (what bench.py reports as `roofline.traffic`).
  python tools/make_traffic.py gpurun_out/prof_step.ncu-rep profiles/traffic_r01.json"""
import csv, io, json, subprocess, sys

KERNEL_TO_ABI = [("stem_tc_kernel", "cova_stem_fwd"), ("conv3x3_tc_kernel", "cova_conv3x3_bn_act_fwd"),
                 ("roi_pool_kernel", "cova_roi_fwd"), ("roi_align_kernel", "cova_roi_fwd"), ("gat_fwd_kernel", "cova_gat_fwd"),
                 ("linear_tc_kernel", "cova_linear_fwd"), ("bbox_enc_kernel", "cova_bbox_enc_fwd")]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}

src, dst = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
ki, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
acc = {}
for r in rows[2:]:
    for pat, abi in KERNEL_TO_ABI:
        if pat in r[ki]:
            b = float(r[ri]) * UNIT[units[ri]] + float(r[wi]) * UNIT[units[wi]]
            acc.setdefault(abi, []).append(b)
res = {abi: {"dram_bytes_per_launch": sum(v) / len(v), "launches_captured": len(v)} for abi, v in acc.items()}
res["_source"] = ("ncu --set full --clock-control none of one bench.py step (%s; B=16, 1280^2, N=90, K=24, fp32-parity mode); "
                  "dram__bytes_read.sum + dram__bytes_write.sum per launch" % src)
json.dump(res, open(dst, "w"), indent=1)
print(json.dumps(res, indent=1))
