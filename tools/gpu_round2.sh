#!/bin/bash
mkdir -p gpurun_out
timeout 120 tools/bin/probe_umma_noswz > gpurun_out/probe_noswz.log 2>&1; echo "probe rc=$?" >> gpurun_out/probe_noswz.log
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
timeout 400 python bench.py > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
COVA_B200_PRECISION=bf16 timeout 300 python bench.py > gpurun_out/bench_bf16.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc -s 8 -c 2 -o gpurun_out/prof_conv_tc_v2 python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_full.log 2>&1
cat gpurun_out/probe_noswz.log; tail -8 gpurun_out/pytest.log; tail -2 gpurun_out/bench.log | cut -c1-600
