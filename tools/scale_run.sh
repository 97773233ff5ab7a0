#!/bin/bash
# 1/2/4/8-GPU weak-scaling runs of bench.py on one box (one process per GPU).  Usage (under gpurun --gpus 8): bash tools/scale_run.sh <tag>
tag=${1:-r01}
mkdir -p gpurun_out
for n in 1 2 4 8; do
  if [ $n -eq 1 ]; then
    timeout 400 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu > gpurun_out/scale_${tag}_n$n.json 2> gpurun_out/scale_${tag}_n$n.log
  else
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/scale_${tag}_n$n.json 2> gpurun_out/scale_${tag}_n$n.log
  fi
  grep "rank 0.*sigs" gpurun_out/scale_${tag}_n$n.log | head -3
done
