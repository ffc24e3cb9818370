#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_train_backbone.py -m gpu -q --tb=short -p no:cacheprovider -k "stem or forward or uint8 or full_size or train" 2>&1 | grep -E "^E  |passed|failed" | head -5
timeout 300 python tools/sweep3.py 2>&1 | grep "^stem conv"
timeout 300 python bench.py --skip-cpu 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], {k: v['ms_per_step'] for k, v in d['kernels'].items()})"
