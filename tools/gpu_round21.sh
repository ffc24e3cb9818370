#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider -k "stem or forward or uint8" 2>&1 | grep -v "^E    +" | tail -4
timeout 300 python tools/sweep3.py 2>&1 | grep "^stem conv"
timeout 300 python bench.py --skip-cpu 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], {k: v['ms_per_step'] for k, v in d['kernels'].items()}, d['e2e']['reps_ms_per_step'], d['e2e_uint8_images']['reps_ms_per_step'], d['throughput_mode_fp16']['value'])"
