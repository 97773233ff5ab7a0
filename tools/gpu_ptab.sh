#!/bin/bash
# Positional fixed-base tables: parity at the default width, then init time, single-launch kernel rates and small-request
# latency per window width.  Usage: bash tools/gpu_ptab.sh <tag> [widths...]
tag=${1:-p}; shift
widths=${@:-"16 18 20 22 24"}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_units.py tests/test_gpu_e2e.py tests/test_gpu_lanegroup.py -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/ptab_tests_$tag.txt
for w in $widths; do
  echo "== SIGOPS_GWIN=$w" | tee -a gpurun_out/ptab_rates_$tag.txt
  SIGOPS_GWIN=$w timeout 300 python - <<PY 2>&1 | tee -a gpurun_out/ptab_rates_$tag.txt
import time, sys, subprocess
sys.path[:0] = ["."]
import wgpu_sigops_b200 as w
lib = w.load()
import torch; torch.cuda.init(); torch.zeros(1, device="cuda")   # CUDA context first, so that the time below is the tables'
t0 = time.perf_counter(); n = lib.sigops_num_devices(); rc = lib.sigops_init(None, 0)
print("init rc", rc, "devices", n, "%.3f s" % (time.perf_counter() - t0), lib.sigops_last_error())
print(subprocess.run("nvidia-smi --query-gpu=memory.used --format=csv,noheader", shell=True, capture_output=True, text=True).stdout.strip())
PY
  SIGOPS_GWIN=$w SIGOPS_MAX_CHUNKS=1 timeout 300 python tools/prof_run.py 1048576 3 time 2>&1 | tee -a gpurun_out/ptab_rates_$tag.txt
  SIGOPS_GWIN=$w LAT_SIZES=21,1024,1365 LAT_MODES=group LAT_OUT=gpurun_out/ptab_latency_${tag}_w$w.json timeout 300 python tools/latency_sweep.py 2>&1 | tail -4 | tee -a gpurun_out/ptab_rates_$tag.txt
done
