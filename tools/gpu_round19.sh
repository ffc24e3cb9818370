#!/bin/bash
mkdir -p gpurun_out
(lscpu | grep -E "Socket|NUMA|^CPU\(s\)|Model name"; nvidia-smi topo -m | head -14) > gpurun_out/topo.txt 2>&1; cat gpurun_out/topo.txt | cut -c1-200
n=8
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 20 --warmup 3 --skip-cpu > gpurun_out/bench_${n}gpu_numa.log 2>&1; tail -1 gpurun_out/bench_${n}gpu_numa.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['value'], d['e2e'], d['e2e_uint8_images'], d['config'].get('cpu_affinity'))"
