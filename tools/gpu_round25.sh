#!/bin/bash
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
BACKBONE=resnet18 B=16 TAIL=native timeout 300 $T tools/bench_train.py 2>&1 | grep "train step"
COVA_B200_TRAIN_TF32=1 BACKBONE=resnet18 B=16 TAIL=native timeout 300 $T tools/bench_train.py 2>&1 | grep "train step" | sed "s/^/tf32=1 /"
BACKBONE=resnet50 B=16 TAIL=native timeout 300 python tools/bench_train.py 2>&1 | grep "train step"
COVA_B200_TRAIN_TF32=1 BACKBONE=resnet50 B=16,32 TAIL=native timeout 300 python tools/bench_train.py 2>&1 | grep "train step" | sed "s/^/tf32=1 /"
