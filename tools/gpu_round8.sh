#!/bin/bash
mkdir -p gpurun_out
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
rc=$?; echo "smoke rc=$rc" >> gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log
if [ $rc -ne 0 ]; then echo "SMOKE FAILED - stopping"; exit 1; fi
timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
grep -E "^E  |FAILED|passed|failed" gpurun_out/pytest.log | cut -c1-250 | head -24
timeout 300 python bench.py > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
tail -2 gpurun_out/bench.log | cut -c1-200
COVA_B200_BACKBONE=resnet50 COVA_B200_N=300 COVA_B200_K=48 COVA_B200_HEADS=2 timeout 300 python bench.py --steps 5 --skip-cpu > gpurun_out/bench_config5_shape.log 2>&1
tail -1 gpurun_out/bench_config5_shape.log | cut -c1-200
timeout 300 python tools/bench_kernels.py 2>&1 | grep -v Model > gpurun_out/kernel_microbench.txt; cat gpurun_out/kernel_microbench.txt
