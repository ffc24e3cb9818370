#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep smoke
for i in 1 2; do timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider 2>&1 | grep -E "^E  |passed|failed" | grep -v "^E    +" | head -5; done
timeout 300 python bench.py > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e_uint8_images']['value'], d['throughput_mode_fp16']['value'], d['roofline']['frac'], d['roofline']['executed_frac'], d['cpu_baseline']['value'])"
