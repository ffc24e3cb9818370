#!/bin/bash
mkdir -p gpurun_out
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
rc=$?; echo "smoke rc=$rc" >> gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log
if [ $rc -ne 0 ]; then echo "SMOKE FAILED - stopping"; exit 1; fi
timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
grep -E "^E  |FAILED|passed|failed" gpurun_out/pytest.log | cut -c1-250 | head -24
timeout 300 python bench.py > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
tail -2 gpurun_out/bench.log | cut -c1-200
# memcheck + racecheck on the small-shape kernel tests (race detection / sanitizers)
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests -m gpu -q -p no:cacheprovider -k "stem_tcgen05_kernel or conv3x3_tcgen05_kernel or linear_tcgen05 or roi_pool_adversarial or gat_layer_golden or forward_small_golden" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/sanitizer_memcheck.log
tail -5 gpurun_out/sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests -m gpu -q -p no:cacheprovider -k "roi_pool_adversarial or gat_layer_golden or stem_kernel or conv3x3_simt or linear_kernel" > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/sanitizer_racecheck.log
tail -5 gpurun_out/sanitizer_racecheck.log
