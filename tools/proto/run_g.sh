for lib in wgpu-sigops_b200/libsigops.so tools/proto/bin/libsigops_r1cold.so; do
echo "== $lib"
SIGOPS_LIB=$PWD/$lib LAT_SIZES=21,1024,1408,2048,3072,4096,4736 LAT_MODES=group_forced LAT_OUT=/tmp/lat.json timeout 400 python tools/latency_sweep.py 30 2>&1 | tail -7 | sed "s/'k1'.*'r1'/'r1'/; s/'ed'.*//"
done
