#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_train_backbone.py tests/test_gpu_harness.py -m gpu -q --tb=short -p no:cacheprovider -k "train or harness or backbone or conv3x3_functions" 2>&1 | grep -E "^E  |passed|failed|^tests/" | grep -v "^E    +" | head -8
BACKBONE=resnet18 B=16 TAIL=native timeout 300 python tools/bench_train.py 2>&1 | grep "train step"
COVA_B200_TRAIN_DGRAD=library BACKBONE=resnet18 B=16 TAIL=native timeout 300 python tools/bench_train.py 2>&1 | grep "train step" | sed "s/^/dgrad=library /"
python tools/diag_train_grads2.py 2>&1 | grep -E "native logits|grad:convnet" | head -4
