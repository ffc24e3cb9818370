#!/bin/bash
mkdir -p gpurun_out
NCV=8 SPF=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv3x3_tc|stem_tc" -s 3 -c 3 -o gpurun_out/prof_backbone_r01k python tools/prof_backbone.py > gpurun_out/ncu_backbone.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/ncu_backbone.log; ls -la gpurun_out/*.ncu-rep
