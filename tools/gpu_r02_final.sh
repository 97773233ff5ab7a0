#!/bin/bash
# Round-2 evidence session on ONE GPU: parity tests, smoke, both bench arms, ncu launch list, full ncu captures (thread and
# lane-group kernels), compute-sanitizer (memcheck / racecheck / synccheck / initcheck), latency sweep, queue sweep, soak.
# Usage (under gpurun): bash tools/gpu_r02_final.sh <tag>
tag=${1:-r02}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu_$tag.txt; nproc >> gpurun_out/gpu_$tag.txt
grep -m1 "model name" /proc/cpuinfo >> gpurun_out/gpu_$tag.txt
cat wgpu-sigops_b200/libsigops.srchash >> gpurun_out/gpu_$tag.txt
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/tests_$tag.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke_$tag.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.log; tail -c 400 gpurun_out/bench_ref_$tag.json; echo
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.log; tail -22 gpurun_out/bench_$tag.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-strong --pool 32768 > gpurun_out/bench_under_ncu_$tag.json 2>/dev/null
SIGOPS_MAX_CHUNKS=1 SIGOPS_TAIL_SPLIT=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ecrecover_kernel|ed25519_verify_kernel' -c 3 -f -o gpurun_out/prof_$tag python tools/prof_run.py 303104 1 2>&1 | tail -2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'group_kernel' -c 3 -f -o gpurun_out/prof_group_$tag python tools/prof_run.py 1024 1 2>&1 | tail -2
for tool in memcheck racecheck synccheck initcheck; do
  echo "== compute-sanitizer --tool $tool (lane-group kernels, 2048 signatures; thread kernels, 20000)"
  timeout 900 compute-sanitizer --tool $tool python tools/prof_run.py 2048 1 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|ok" | tail -5
  SIGOPS_LANEGROUP=0 timeout 900 compute-sanitizer --tool $tool python tools/prof_run.py 20000 1 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|ok" | tail -5
done 2>&1 | tee gpurun_out/sanitizer_$tag.txt
timeout 600 python tools/latency_sweep.py 40 2>&1 | tail -26 | tee gpurun_out/latency_$tag.txt
timeout 600 python tools/queue_bench.py > gpurun_out/queue_sweep_$tag.json 2> gpurun_out/queue_sweep_$tag.log; tail -3 gpurun_out/queue_sweep_$tag.log
timeout 600 python tools/soak.py 30 700 2>&1 | tail -4 | tee gpurun_out/soak_$tag.txt
ls -la gpurun_out | tail -25
