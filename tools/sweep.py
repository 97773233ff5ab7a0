#!/usr/bin/env python3
"""BASELINE config 5: mixed k1 / r1 / ed25519 batch-size sweep (64 ... 16M signatures in x4 steps) through the host C ABI
with pinned buffers, on every visible GPU of the box (the library shards each call).  Per size: three back-to-back
calls (n/3 signatures each), latency p50 / p99 per call over the repetitions and aggregate signatures/s.
   python tools/sweep.py [max_n] [tag]        -> gpurun_out/sweep_<tag>.json"""
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import coracle  # noqa: E402  (input synthesis and expected values only)
import wgpu_sigops_b200 as w  # noqa: E402


def pinned(lib, a):
    ptr = lib.sigops_host_alloc(a.nbytes)
    buf = np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_uint8)), shape=(a.nbytes,))
    buf[:] = a.reshape(-1)
    return ptr


def main():
    max_n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
    tag = sys.argv[2] if len(sys.argv) > 2 else "r01"
    lib = w.load()
    ngpu = lib.sigops_num_devices()
    pool = 65536
    k1 = coracle.gen_ecdsa(0, pool, seed=31)
    r1 = coracle.gen_ecdsa(1, pool, seed=32, low_s=False)
    ed = coracle.gen_ed25519(pool, seed=33)
    rows = []
    n = 64
    while n <= max_n:
        k = max(1, n // 3)

        def tile(a):
            return np.ascontiguousarray(np.tile(a, ((k + pool - 1) // pool,) + (1,) * (a.ndim - 1))[:k])

        bufs = {"k1": [pinned(lib, tile(x)) for x in k1[:2]], "r1": [pinned(lib, tile(x)) for x in r1[:2]],
                "ed": [pinned(lib, tile(x)) for x in ed]}
        out = lib.sigops_host_alloc(k * 64)
        st = lib.sigops_host_alloc(k)
        exp_k1, exp_r1 = tile(k1[2]), tile(r1[2])

        def call(which):
            if which == "k1":
                rc = lib.sigops_secp256k1_ecrecover(bufs["k1"][0], bufs["k1"][1], k, out, st)
            elif which == "r1":
                rc = lib.sigops_secp256r1_ecrecover(bufs["r1"][0], bufs["r1"][1], k, out, st)
            else:
                rc = lib.sigops_ed25519_ecverify(bufs["ed"][0], bufs["ed"][1], bufs["ed"][2], k, out)
            assert rc == 0, lib.sigops_last_error()

        def view(nbytes):
            return np.ctypeslib.as_array(ctypes.cast(out, ctypes.POINTER(ctypes.c_uint8)), shape=(nbytes,))

        # parity of each call once, then timing
        call("k1"); assert (view(k * 64).reshape(-1, 64) == exp_k1).all()
        call("r1"); assert (view(k * 64).reshape(-1, 64) == exp_r1).all()
        call("ed"); assert view(k).all()
        reps = 30 if n <= 65536 else 10 if n <= (1 << 22) else 4
        lat = {"k1": [], "r1": [], "ed": []}
        t_all = time.perf_counter()
        for _ in range(reps):
            for which in ("k1", "r1", "ed"):
                t0 = time.perf_counter()
                call(which)
                lat[which].append((time.perf_counter() - t0) * 1e3)
        total = time.perf_counter() - t_all
        row = {"n_total": 3 * k, "n_per_call": k, "reps": reps, "sigs_per_s": 3 * k * reps / total}
        for which in lat:
            a = np.array(lat[which])
            row[which] = {"p50_ms": float(np.percentile(a, 50)), "p99_ms": float(np.percentile(a, 99)),
                          "sigs_per_s_p50": k / (float(np.percentile(a, 50)) * 1e-3)}
        rows.append(row)
        print(json.dumps(row), flush=True)
        for p in bufs["k1"] + bufs["r1"] + bufs["ed"] + [out, st]:
            lib.sigops_host_free(p)
        n *= 4
    res = {"config": "BASELINE config 5: mixed k1/r1/ed25519 sweep, host C ABI, pinned buffers, H2D + kernels + D2H inside each call",
           "n_gpus": ngpu, "rows": rows}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"sweep_{tag}.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
