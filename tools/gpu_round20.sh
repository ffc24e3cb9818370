#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_harness.py -m gpu -q --tb=short -p no:cacheprovider 2>&1 | grep -v "^E    +" | tail -6
timeout 300 python tools/bench_kernels.py 2>&1 | grep -E "GAT" | grep -v rotating
timeout 300 python bench.py --skip-cpu > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], {k: v['ms_per_step'] for k, v in d['kernels'].items()}, d['e2e']['value'], d['e2e_uint8_images']['value'])"
