#!/usr/bin/env python3
"""Probe of the streaming mode's concurrency: submit `depth` requests back to back, then wait for all; prints the host
time of the submits and the time until the last result (device concurrency shows as t_all ~ one request's latency)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import coracle  # noqa: E402
import wgpu_sigops_b200 as w  # noqa: E402

curve = sys.argv[1] if len(sys.argv) > 1 else "secp256k1"
pool = bench.make_batch(curve, 1 << 14, 1 << 14, 3, coracle.host_threads())
for req_n in (256, 1024, 4096):
    for depth in (1, 2, 4, 8, 16, 32):
        with w.service.SigQueue(curve, req_n, depth) as q:
            for s in range(depth):
                q.sigs(s)[:req_n] = pool[0][:req_n]
                q.msgs(s)[:req_n] = pool[1][:req_n]
                if pool[2] is not None:
                    q.pks(s)[:req_n] = pool[2][:req_n]
            best = None
            for rep in range(5):
                t0 = time.perf_counter()
                for s in range(depth):
                    q.submit(s, req_n)
                t1 = time.perf_counter()
                for s in range(depth):
                    q.wait(s)
                t2 = time.perf_counter()
                cur = (t2 - t0, t1 - t0, q.last_device_ms)
                best = cur if best is None or cur[0] < best[0] else best
            print(f"{curve} n={req_n} depth={depth}: submits {best[1]*1e3:.3f} ms, all done {best[0]*1e3:.3f} ms, "
                  f"last request on device {best[2]:.3f} ms -> {depth*req_n/best[0]/1e6:.2f} M sigs/s", flush=True)
