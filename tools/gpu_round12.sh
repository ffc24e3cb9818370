#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -25
timeout 300 python tools/precision_report.py > gpurun_out/precision_modes.txt 2>&1; cat gpurun_out/precision_modes.txt
COVA_B200_PRECISION=fp16 timeout 300 python bench.py --skip-cpu > gpurun_out/bench_fp16.log 2>&1; tail -1 gpurun_out/bench_fp16.log | cut -c1-3000
