LAT_SIZES=1408,2048,2560,3072,3584,4096,4736 LAT_MODES=group_forced,thread LAT_OUT=/tmp/lat.json timeout 400 python tools/latency_sweep.py 30 2>&1 | tail -14
