#!/usr/bin/env python3
"""Short driver for ncu captures: runs each fused kernel `reps` times on n signatures through the host C ABI.
   python tools/prof_run.py [n] [reps] [curves]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import coracle  # noqa: E402  (input synthesis only)
import wgpu_sigops_b200 as w  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    curves = sys.argv[3].split(",") if len(sys.argv) > 3 else ["secp256k1", "secp256r1", "ed25519"]
    timing = curves == ["time"]
    if timing:
        curves = ["secp256k1", "secp256r1", "ed25519"]
    import ctypes

    def ktime():
        h, k, d = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
        w.load().sigops_last_timing(ctypes.byref(h), ctypes.byref(k), ctypes.byref(d))
        return k.value
    pool = min(n, 32768)
    import numpy as np

    def tile(a):
        return np.ascontiguousarray(np.tile(a, ((n + pool - 1) // pool, 1))[:n])

    for c in curves:
        if c == "ed25519":
            s, m, p = coracle.gen_ed25519(pool)
            ts, tm, tp = tile(s), tile(m), tile(p)
            for _ in range(reps):
                t0 = time.perf_counter()
                v = w.ed25519_eddsa.ecverify_array(ts, tm, tp)
                wall = time.perf_counter() - t0
            assert v.all()
        else:
            cid = 0 if c == "secp256k1" else 1
            s, m, e = coracle.gen_ecdsa(cid, pool)
            mod = w.secp256k1_ecdsa if cid == 0 else w.secp256r1_ecdsa
            ts, tm = tile(s), tile(m)
            for _ in range(reps):
                t0 = time.perf_counter()
                out, st = mod.ecrecover_with_status(ts, tm)
                wall = time.perf_counter() - t0
            assert (out == tile(e)).all() and not st.any()
        print(c, "ok", ("%.3f ms %.2f Msig/s kernel span; wall %.2f Msig/s (pageable host buffers)" % (ktime(), n / ktime() / 1e3, n / wall / 1e6)) if timing else "", flush=True)


if __name__ == "__main__":
    main()
