for lib in wgpu-sigops_b200/libsigops.so tools/proto/bin/libsigops_gcold.so; do
  echo "== $lib"
  SIGOPS_LIB=$PWD/$lib LAT_SIZES=21,1024,1365,4736 LAT_MODES=group LAT_OUT=/tmp/lat.json timeout 300 python tools/latency_sweep.py 40 2>&1 | tail -4
done
echo "== repeat"
for lib in wgpu-sigops_b200/libsigops.so tools/proto/bin/libsigops_gcold.so; do
  echo "== $lib"
  SIGOPS_LIB=$PWD/$lib LAT_SIZES=21,1024,1365,4736 LAT_MODES=group LAT_OUT=/tmp/lat.json timeout 300 python tools/latency_sweep.py 40 2>&1 | tail -4
done
