#!/bin/bash
mkdir -p gpurun_out
timeout 120 tools/bin/probe_mma_rate > gpurun_out/probe_mma_rate.log 2>&1; echo "probe rc=$?" >> gpurun_out/probe_mma_rate.log
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
timeout 400 python bench.py > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
cat gpurun_out/probe_mma_rate.log; grep -E "^E  |FAILED|passed|failed" gpurun_out/pytest.log | cut -c1-300 | head -20; tail -2 gpurun_out/bench.log | cut -c1-300
