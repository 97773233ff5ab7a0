#!/bin/bash
# Kernel span of shards of one to a few waves with and without the separate tail launch (SIGOPS_TAIL_SPLIT=0 disables it).
# Usage (under gpurun): bash tools/tail_probe.sh
for n in 87381 113664 131072 189440 262144 349525 1048576; do
  for split in 1 0; do
    echo "n=$n split=$split: $(SIGOPS_TAIL_SPLIT=$split SIGOPS_MAX_CHUNKS=1 timeout 120 python tools/prof_run.py $n 3 time | tr '\n' ' ' | sed -e 's/kernel span; wall [0-9.]* Msig\/s (pageable host buffers)//g')"
  done
done
