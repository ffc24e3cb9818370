#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_train_backbone.py -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | grep -v "^E    +" | tail -15
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_harness.py -m gpu -q --tb=short -p no:cacheprovider -k "train or harness or tail" 2>&1 | grep -v "^E    +" | tail -6
BACKBONE=resnet18 B=16 TAIL=native timeout 300 python tools/bench_train.py 2>&1 | grep "train step"
