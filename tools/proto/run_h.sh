timeout 900 python -m pytest tests/test_gpu_lanegroup.py tests/test_gpu_host_layer.py tests/test_service_queue.py -x -q -m gpu 2>&1 | tail -3
LAT_SIZES=21,256,1024,1365,2048,3072,4096,4736 LAT_MODES=group,thread LAT_OUT=/tmp/lat.json timeout 300 python tools/latency_sweep.py 40 2>&1 | tail -16
for m in 0 1000000; do echo "== SIGOPS_GROUP_COLD_MIN=$m"; SIGOPS_GROUP_COLD_MIN=$m LAT_SIZES=1024,1536,2048,2560,3072,4096,4736 LAT_MODES=group LAT_OUT=/tmp/lat.json timeout 300 python tools/latency_sweep.py 40 2>&1 | tail -7; done
