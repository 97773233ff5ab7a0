cat wgpu-sigops_b200/libsigops.srchash
timeout 900 python tools/soak.py 150 1000 2>&1 | tail -3 | tee gpurun_out/soak_r02_2.txt
timeout 400 python tools/fuzz_soak.py 200 2>&1 | tail -3 | tee gpurun_out/fuzz_soak_r02.txt
SIGOPS_FORCE_LANEGROUP=1 timeout 300 python tools/fuzz_soak.py 90 20000 2>&1 | tail -3 | tee gpurun_out/fuzz_soak_group_r02.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-strong > gpurun_out/bench_r02b.json 2> gpurun_out/bench_r02b.log; tail -c 600 gpurun_out/bench_r02b.json
