#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -s --tb=short -p no:cacheprovider 2>&1 | grep -E "RANK|passed|failed|Error|error" | head -12 | tee gpurun_out/two_rank_nccl_test.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 --skip-cpu > gpurun_out/bench_2gpu.log 2>&1; tail -1 gpurun_out/bench_2gpu.log | cut -c1-700
