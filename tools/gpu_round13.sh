#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_harness.py -m gpu -q --tb=short -p no:cacheprovider 2>&1 | grep -v "^E    +" | tail -12
COVA_B200_PRECISION=fp16 timeout 300 python bench.py --skip-cpu > gpurun_out/bench_fp16.log 2>&1; tail -1 gpurun_out/bench_fp16.log | cut -c1-3500
timeout 300 python bench.py > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log | cut -c1-3500
