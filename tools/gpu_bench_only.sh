#!/bin/bash
# bench (both arms) + ncu launch list + one full capture per kernel; no tests.  Usage: bash tools/gpu_bench_only.sh <tag>
tag=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu_$tag.txt; nproc >> gpurun_out/gpu_$tag.txt
grep -m1 "model name" /proc/cpuinfo >> gpurun_out/gpu_$tag.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.log; tail -3 gpurun_out/bench_$tag.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 3 --no-cpu --pool 32768 > gpurun_out/bench_under_ncu_$tag.json 2>/dev/null
SIGOPS_MAX_CHUNKS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ecrecover_kernel|ed25519_verify_kernel' -c 3 -f -o gpurun_out/prof_$tag python tools/prof_run.py 303104 1 2>&1 | tail -2
ls -la gpurun_out
