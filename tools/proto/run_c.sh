for n in 87381 131072 174763 262144 349525 524288 1048576; do
  for mode in "SIGOPS_EVEN_PASSES=0" "SIGOPS_EVEN_PASSES=1" "SIGOPS_TAIL_SPLIT=0"; do
    echo "n=$n $mode: $(env $mode SIGOPS_MAX_CHUNKS=1 timeout 120 python tools/prof_run.py $n 3 time 2>&1 | sed 's/ kernel span.*//' | tr '\n' ' ')"
  done
done
timeout 900 python -m pytest tests/test_gpu_host_layer.py -x -q -m gpu -k "window" 2>&1 | tail -3
