#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_train_backbone.py -m gpu -q --tb=short -p no:cacheprovider 2>&1 | grep -v "^E    +" | tail -12
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_harness.py -m gpu -q --tb=short -p no:cacheprovider -k "train or harness or tail" 2>&1 | grep -v "^E    +" | tail -8
for conv in cudnn tcgen05; do for tf in 0 1; do
COVA_B200_TRAIN_CONV=$conv COVA_B200_TRAIN_TF32=$tf BACKBONE=resnet18 B=16 TAIL=native timeout 300 python tools/bench_train.py 2>&1 | grep "train step" | sed "s/^/conv=$conv tf32=$tf  /"
done; done
COVA_B200_TRAIN_BACKBONE=torch BACKBONE=resnet18 B=16 TAIL=torch timeout 300 python tools/bench_train.py 2>&1 | grep "train step" | sed "s/^/all-library path  /"
