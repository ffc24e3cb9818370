#!/bin/bash
for i in 1 2; do timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_harness.py tests/test_gpu_train_backbone.py tests/test_gpu_multi.py -m gpu -q --tb=short -p no:cacheprovider -k "train or harness or tail or backbone" 2>&1 | grep -v "^E    +" | tail -6; done
BACKBONE=resnet18 B=16 TAIL=native timeout 300 python tools/bench_train.py 2>&1 | grep "train step"
COVA_B200_TRAIN_CONV=cudnn BACKBONE=resnet18 B=16 TAIL=native timeout 300 python tools/bench_train.py 2>&1 | grep "train step" | sed "s/^/conv=cudnn /"
COVA_B200_TRAIN_TF32=1 BACKBONE=resnet18 B=16 TAIL=native timeout 300 python tools/bench_train.py 2>&1 | grep "train step" | sed "s/^/tf32=1 /"
BACKBONE=resnet50 B=16 TAIL=native timeout 300 python tools/bench_train.py 2>&1 | grep "train step"
