#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log; tail -2 gpurun_out/bench.log | cut -c1-400
timeout 300 python tools/bench_kernels.py 2>&1 | grep -v Model > gpurun_out/kernel_microbench.txt; cat gpurun_out/kernel_microbench.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --skip-cpu > gpurun_out/ncu_launch.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"roi_pool|gat_fwd|ce_sum|adam|topk|build_batch" -c 12 -o gpurun_out/prof_tail_r01l python -m pytest tests/test_gpu_tail.py tests/test_gpu_parity.py -q -x -p no:cacheprovider -k "full_size or adam_kernel_vs_torch or ce_sum_kernel_vs_oracle or topk_hits_kernel_vs_oracle or build_batch_kernel_vs_oracle" > gpurun_out/ncu_tail.log 2>&1; tail -3 gpurun_out/ncu_tail.log
