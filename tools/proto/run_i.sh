for lib in wgpu-sigops_b200/libsigops.so tools/proto/bin/libsigops_dblonly.so tools/proto/bin/libsigops_u2.so; do
echo "== $lib"; SIGOPS_LIB=$PWD/$lib SIGOPS_MAX_CHUNKS=1 timeout 300 python tools/prof_run.py 1048576 3 time 2>&1 | sed 's/ kernel span.*//'
done
