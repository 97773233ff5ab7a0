cat wgpu-sigops_b200/libsigops.srchash
echo "== lane-group kernels forced, out-of-line products (default switch)"
SIGOPS_FORCE_LANEGROUP=1 timeout 600 python tools/soak.py 40 2000 2>&1 | tail -2
echo "== lane-group kernels forced, inlined products"
SIGOPS_FORCE_LANEGROUP=1 SIGOPS_GROUP_COLD_MIN=1000000000 timeout 600 python tools/soak.py 40 3000 2>&1 | tail -2
echo "== rejecting paths, lane-group kernels forced (both flavours by size)"
SIGOPS_FORCE_LANEGROUP=1 timeout 300 python tools/fuzz_soak.py 100 20000 2>&1 | tail -1
SIGOPS_FORCE_LANEGROUP=1 SIGOPS_GROUP_COLD_MIN=0 timeout 300 python tools/fuzz_soak.py 60 3000 2>&1 | tail -1
