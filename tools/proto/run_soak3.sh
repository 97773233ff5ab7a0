cat wgpu-sigops_b200/libsigops.srchash
timeout 600 python tools/soak.py 120 5000 2>&1 | tail -2
timeout 300 python tools/fuzz_soak.py 150 2>&1 | tail -1
