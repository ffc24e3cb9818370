#!/bin/bash
mkdir -p gpurun_out
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
rc=$?; echo "smoke rc=$rc" >> gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log
if [ $rc -ne 0 ]; then echo "SMOKE FAILED - stopping"; exit 1; fi
timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
grep -E "^E  |FAILED|passed|failed" gpurun_out/pytest.log | cut -c1-250 | head -24
timeout 300 python bench.py > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
tail -2 gpurun_out/bench.log | cut -c1-200
COVA_B200_BACKBONE=resnet50 timeout 300 python bench.py --steps 5 > gpurun_out/bench_r50.log 2>&1
timeout 200 python tools/diag_e2e.py 2>&1 | grep -v Model > gpurun_out/diag_e2e.log; cat gpurun_out/diag_e2e.log
