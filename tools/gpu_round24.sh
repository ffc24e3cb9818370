#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider 2>&1 | grep -v "^E    +" | tail -5 | tee gpurun_out/pytest_tail.txt
timeout 300 python bench.py > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench.log | cut -c1-300
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | cut -c1-400
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --skip-cpu > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"stem_tc|conv3x3_tc|roi_pool|gat_fwd|linear_tc|bbox_enc" -s 33 -c 11 -o gpurun_out/prof_step_r01m python bench.py --steps 3 --warmup 3 --skip-cpu > gpurun_out/ncu_step.log 2>&1; tail -2 gpurun_out/ncu_step.log | cut -c1-200
