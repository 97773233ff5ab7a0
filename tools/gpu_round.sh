#!/bin/bash
# One GPU session: parity tests, smoke, bench (both arms), ncu launch list and one full capture per kernel.
# Usage (under gpurun): bash tools/gpu_round.sh <tag>
tag=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu_$tag.txt; nproc >> gpurun_out/gpu_$tag.txt
grep -m1 "model name" /proc/cpuinfo >> gpurun_out/gpu_$tag.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.log; tail -c 600 gpurun_out/bench_ref_$tag.json
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.log; tail -5 gpurun_out/bench_$tag.log; cat gpurun_out/bench_$tag.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 3 --no-cpu --pool 32768 > gpurun_out/bench_under_ncu_$tag.json 2>/dev/null
SIGOPS_MAX_CHUNKS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ecrecover_kernel|ed25519_verify_kernel' -c 3 -f -o gpurun_out/prof_$tag python tools/prof_run.py 303104 1 2>&1 | tail -3
ls -la gpurun_out
