#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider -k "stem_tcgen05_kernel or conv3x3_tcgen05_kernel or forward_" 2>&1 | grep -v "^E    +" | tail -12
timeout 300 python tools/precision_report.py > gpurun_out/precision_modes.txt 2>&1; cat gpurun_out/precision_modes.txt
COVA_B200_PRECISION=fp32x timeout 300 python bench.py --skip-cpu 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], {k: v['ms_per_step'] for k, v in d['kernels'].items()})"
