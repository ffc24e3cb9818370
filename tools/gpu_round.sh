#!/bin/bash
# One GPU-box session: smoke, descriptor probe, parity tests, bench, ncu launch list + full capture.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 120 tools/bin/probe_umma > gpurun_out/probe.log 2>&1; echo "probe rc=$?" >> gpurun_out/probe.log
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
timeout 400 python bench.py > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
COVA_B200_PRECISION=bf16 timeout 300 python bench.py > gpurun_out/bench_bf16.log 2>&1
COVA_B200_ENGINE=simt timeout 300 python bench.py --steps 5 > gpurun_out/bench_simt.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_launch.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc -s 8 -c 2 -o gpurun_out/prof_conv_tc python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_full.log 2>&1
tail -5 gpurun_out/smoke.log; tail -15 gpurun_out/pytest.log; tail -3 gpurun_out/bench.log
