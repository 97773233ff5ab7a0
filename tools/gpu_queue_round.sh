#!/bin/bash
# Streaming-mode session: queue parity tests, the queue sweep (graphs on / off), then the whole GPU suite.
# Usage (under gpurun): bash tools/gpu_queue_round.sh <tag>
tag=${1:-r01q}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_service_queue.py -m gpu -x -q 2>&1 | tail -15
timeout 600 python tools/queue_bench.py > gpurun_out/queue_sweep_$tag.json 2> gpurun_out/queue_sweep_$tag.log; tail -14 gpurun_out/queue_sweep_$tag.log
SIGOPS_QUEUE_GRAPHS=0 timeout 300 python tools/queue_bench.py --curves secp256k1 --sizes 256,1024 > gpurun_out/queue_sweep_nograph_$tag.json 2> gpurun_out/queue_sweep_nograph_$tag.log; tail -3 gpurun_out/queue_sweep_nograph_$tag.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
