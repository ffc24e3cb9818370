#!/bin/bash
mkdir -p gpurun_out
BACKBONE=resnet18 B=16 TAIL=native timeout 600 ncu --set full --clock-control none -k regex:"bn_reduce|bn_act|maxpool|split_planes|ce_sum|adam_kernel" -s 40 -c 24 -o gpurun_out/prof_train_r01n python tools/bench_train.py > gpurun_out/ncu_train.log 2>&1; tail -2 gpurun_out/ncu_train.log | cut -c1-200
