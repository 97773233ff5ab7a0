#!/bin/bash
# One full ncu capture per fused kernel (303,104 signatures = 4 waves, single launch each).  Usage: bash tools/gpu_ncu.sh <tag> [n]
tag=${1:-r02}
n=${2:-303104}
mkdir -p gpurun_out
SIGOPS_MAX_CHUNKS=1 SIGOPS_TAIL_SPLIT=0 timeout 1500 ncu --set full --clock-control none --import-source on \
  -k regex:'ecrecover_kernel|ed25519_verify_kernel' -c 3 -f -o gpurun_out/prof_$tag python tools/prof_run.py $n 1 2>&1 | tail -3
ls -la gpurun_out/prof_$tag.ncu-rep
