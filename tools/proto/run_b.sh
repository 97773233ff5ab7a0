mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_lanegroup.py tests/test_gpu_units.py -x -q -m gpu 2>&1 | tail -3
for w in 20 22; do
echo "== w=$w batch 8"; SIGOPS_GWIN=$w SIGOPS_MAX_CHUNKS=1 timeout 300 python tools/prof_run.py 1048576 3 time
echo "== w=$w batch 16"; SIGOPS_LIB=$PWD/tools/proto/bin/libsigops_b16.so SIGOPS_GWIN=$w SIGOPS_MAX_CHUNKS=1 timeout 300 python tools/prof_run.py 1048576 3 time
done
echo "== 4M batch 8";  SIGOPS_MAX_CHUNKS=1 timeout 300 python tools/prof_run.py 4194304 2 time
echo "== 4M batch 16"; SIGOPS_LIB=$PWD/tools/proto/bin/libsigops_b16.so SIGOPS_MAX_CHUNKS=1 timeout 300 python tools/prof_run.py 4194304 2 time
for w in 20 22; do
SIGOPS_GWIN=$w LAT_SIZES=21,1024,1365,4736 LAT_MODES=group LAT_OUT=gpurun_out/ptab_latency_b_w$w.json timeout 300 python tools/latency_sweep.py 2>&1 | tail -4
done
