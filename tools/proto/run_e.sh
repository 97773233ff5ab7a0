timeout 900 python -m pytest tests/test_gpu_lanegroup.py tests/test_gpu_units.py tests/test_gpu_fuzz.py -x -q -m gpu 2>&1 | tail -3
LAT_SIZES=21,256,1024,1365,1408 LAT_MODES=group LAT_OUT=/tmp/lat.json timeout 300 python tools/latency_sweep.py 40 2>&1 | tail -5
LAT_SIZES=2048,4736 LAT_MODES=group_forced LAT_OUT=/tmp/lat.json timeout 300 python tools/latency_sweep.py 40 2>&1 | tail -2
