#!/bin/bash
# Quick GPU check of a kernel change: unit + e2e parity, then single-launch kernel rates at 1M.  Usage: bash tools/gpu_quick.sh <tag>
tag=${1:-q}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_units.py tests/test_gpu_e2e.py tests/test_gpu_configs.py -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/quick_tests_$tag.txt
SIGOPS_MAX_CHUNKS=1 timeout 300 python tools/prof_run.py 1048576 3 time 2>&1 | tee gpurun_out/quick_rates_$tag.txt
