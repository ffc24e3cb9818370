#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/sweep_backbone.py > gpurun_out/sweep_backbone.txt 2>&1; echo "sweep rc=$?"
cat gpurun_out/sweep_backbone.txt | cut -c1-260
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "stem or conv or forward" -p no:cacheprovider 2>&1 | tail -3
