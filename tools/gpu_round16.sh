#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_tail.py tests/test_gpu_parity.py -q -x -p no:cacheprovider -k "tail or roi or stem_tcgen05_kernel or conv3x3_tcgen05_kernel or build_batch or topk or ce_sum or adam_kernel" > gpurun_out/sanitizer_memcheck.txt 2>&1; echo "memcheck rc=$?"; tail -6 gpurun_out/sanitizer_memcheck.txt
