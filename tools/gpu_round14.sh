#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider -k "roi or forward or train" 2>&1 | grep -v "^E    +" | tail -8
timeout 300 python tools/bench_kernels.py 2>&1 | grep -v Model > gpurun_out/kernel_microbench.txt; cat gpurun_out/kernel_microbench.txt
