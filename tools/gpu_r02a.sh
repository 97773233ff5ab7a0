#!/bin/bash
# Round-2 first GPU session (1 GPU): parity tests, smoke, bench, and the geometry / pass-time probes.
tag=${1:-r02a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu_$tag.txt; nproc >> gpurun_out/gpu_$tag.txt
grep -m1 "model name" /proc/cpuinfo >> gpurun_out/gpu_$tag.txt
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/tests_$tag.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke_$tag.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.log; tail -25 gpurun_out/bench_$tag.log
# kernel span of shards of one to a few waves: balanced geometry vs round 1's (full blocks, +/- separate tail launch)
for n in 37888 75776 87381 131072 174763 262144 349525; do
  for mode in "SIGOPS_BALANCED=1" "SIGOPS_BALANCED=0 SIGOPS_TAIL_SPLIT=1" "SIGOPS_BALANCED=0 SIGOPS_TAIL_SPLIT=0"; do
    echo "n=$n $mode: $(env $mode SIGOPS_MAX_CHUNKS=1 timeout 120 python tools/prof_run.py $n 3 time 2>&1 | tr '\n' ' ')"
  done
done 2>&1 | tee gpurun_out/geometry_$tag.txt
ls -la gpurun_out | tail -12
