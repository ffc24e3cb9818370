#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for n in 4 8; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 3 --skip-cpu > gpurun_out/bench_${n}gpu.log 2>&1; tail -1 gpurun_out/bench_${n}gpu.log | cut -c1-260
done
