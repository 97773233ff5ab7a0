#!/usr/bin/env python3
"""Streaming service mode sweep: request size x requests in flight through `service.SigQueue`, next to
the blocking entry point called back to back with the same request size.  Every result is checked against the expected
values.  Output: JSON on stdout (-> profiles/r01_queue_sweep.json).

    python tools/queue_bench.py [--curves secp256k1,secp256r1,ed25519] [--sizes 256,1024,4096,16384] [--depths 1,2,4,8,16]
"""
import argparse
import ctypes
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import bench  # noqa: E402
import coracle  # noqa: E402
import wgpu_sigops_b200 as w  # noqa: E402


def blocking(lib, curve, req_n, n_requests, pool):
    sigs, msgs, pks, exp = pool
    entry = {"secp256k1": lib.sigops_secp256k1_ecrecover, "secp256r1": lib.sigops_secp256r1_ecrecover}.get(curve)
    out = np.zeros((req_n, 64), dtype=np.uint8)
    st = np.zeros(req_n, dtype=np.uint8)
    lat = []
    t0 = time.perf_counter()
    for i in range(n_requests):
        a = (i * req_n) % (sigs.shape[0] - req_n + 1)
        t = time.perf_counter()
        if entry is not None:
            rc = entry(sigs[a:a + req_n].ctypes.data, msgs[a:a + req_n].ctypes.data, req_n, out.ctypes.data, st.ctypes.data)
            ok = np.array_equal(out, exp[a:a + req_n]) and not st.any()
        else:
            rc = lib.sigops_ed25519_ecverify(sigs[a:a + req_n].ctypes.data, msgs[a:a + req_n].ctypes.data,
                                             pks[a:a + req_n].ctypes.data, req_n, st.ctypes.data)
            ok = st.all()
        lat.append(time.perf_counter() - t)
        assert rc == 0 and ok
    dt = time.perf_counter() - t0
    lat.sort()
    return {"sigs_per_s": n_requests * req_n / dt, "latency_ms_p50": lat[len(lat) // 2] * 1e3}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--curves", default="secp256k1,secp256r1,ed25519")
    ap.add_argument("--sizes", default="256,1024,4096,16384")
    ap.add_argument("--depths", default="1,2,4,8,16,32")
    ap.add_argument("--requests-per-slot", type=int, default=40)
    args = ap.parse_args()
    lib = w.load()
    ids = (ctypes.c_int * 1)(0)
    assert lib.sigops_init(ids, 1) == 0, lib.sigops_last_error()
    threads = coracle.host_threads()
    res = {"config": "requests through service.SigQueue on one GPU: host submit -> wait per request, H2D + fused kernel + D2H "
                     "inside; `blocking` = the same requests through the blocking C entry point, one at a time",
           "graphs": os.environ.get("SIGOPS_QUEUE_GRAPHS", "1") != "0",
           "cuda_device_max_connections": os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS"), "curves": {}}
    for curve in args.curves.split(","):
        pool = bench.make_batch(curve, 1 << 16, 1 << 16, 0x51600003, threads)
        rows = []
        for req_n in [int(x) for x in args.sizes.split(",")]:
            row = {"request_sigs": req_n, "blocking": blocking(lib, curve, req_n, 40, pool), "queue": []}
            for d in [int(x) for x in args.depths.split(",")]:
                r = bench.queue_throughput(w, curve, req_n, d, max(40, args.requests_per_slot * d), pool)
                row["queue"].append({k: r[k] for k in ("depth", "sigs_per_s", "latency_ms_p50", "latency_ms_p99")})
            rows.append(row)
            print(curve, req_n, "blocking %.2f M/s" % (row["blocking"]["sigs_per_s"] / 1e6),
                  " ".join("d%d: %.2f M/s (%.2f ms)" % (q["depth"], q["sigs_per_s"] / 1e6, q["latency_ms_p50"]) for q in row["queue"]),
                  file=sys.stderr, flush=True)
        res["curves"][curve] = rows
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
