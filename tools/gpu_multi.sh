#!/bin/bash
# Multi-GPU session: the multi-device tests, then bench.py under torchrun exactly as the driver launches it.
# Usage (under gpurun --gpus N): bash tools/gpu_multi.sh <N> <tag>
N=${1:-2}
tag=${2:-r02}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > gpurun_out/gpus_$tag.txt
timeout 900 python -m pytest tests/test_gpu_host_layer.py tests/test_gpu_configs.py tests/test_service_queue.py -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/tests_multi_$tag.txt
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${N}gpu_$tag.json 2> gpurun_out/bench_${N}gpu_$tag.log
grep -E "strong|sweep|sigs/s|Error|error|Traceback" gpurun_out/bench_${N}gpu_$tag.log | tail -30
tail -c 300 gpurun_out/bench_${N}gpu_$tag.json
