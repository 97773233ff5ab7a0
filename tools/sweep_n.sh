for mc in 1 8; do echo "chunks<=$mc"; SIGOPS_MAX_CHUNKS=$mc python tools/prof_run.py 1048576 3 time 2>&1 | grep ok; done
python bench.py --steps 5 --warmup 3 --no-cpu 2>&1 | grep -E "rank 0"
