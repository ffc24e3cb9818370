#!/bin/bash
mkdir -p gpurun_out
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
rc=$?; echo "smoke rc=$rc" >> gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log
if [ $rc -ne 0 ]; then echo "SMOKE FAILED - stopping"; exit 1; fi
timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
grep -E "^E  |FAILED|passed|failed" gpurun_out/pytest.log | cut -c1-250 | head -24
timeout 300 python bench.py > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
tail -2 gpurun_out/bench.log | cut -c1-200
COVA_B200_PRECISION=bf16 timeout 300 python bench.py > gpurun_out/bench_bf16.log 2>&1
COVA_B200_BACKBONE=resnet50 timeout 300 python bench.py --steps 5 > gpurun_out/bench_r50.log 2>&1
timeout 300 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/bench_reference.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_launch.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"conv3x3_tc|stem_tc" -s 5 -c 3 -o gpurun_out/prof_tc python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_full.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"roi_pool|gat_fwd|linear_tc" -s 5 -c 5 -o gpurun_out/prof_tail python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_full2.log 2>&1
