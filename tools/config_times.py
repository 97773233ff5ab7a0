#!/usr/bin/env python3
"""BASELINE.json configurations 1-3 timed through the blocking C ABI (pinned host buffers, H2D + kernel + D2H inside), with
the kernel span of the same call next to it and every output row compared with the oracle-labelled batch.
   python tools/config_times.py [reps]      (under gpurun; one GPU)   -> gpurun_out/config_times.json"""
import ctypes
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np  # noqa: E402

import batches  # noqa: E402
import wgpu_sigops_b200 as w  # noqa: E402


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    lib = w.load()
    one = (ctypes.c_int * 1)(0)
    assert lib.sigops_init(one, 1) == 0, lib.sigops_last_error()

    def pin(a):
        p = lib.sigops_host_alloc(max(1, a.nbytes))
        v = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint8)), shape=(max(1, a.nbytes),))
        v[: a.nbytes] = a.reshape(-1)
        return p, v

    configs = [
        ("1. secp256k1 ecrecover, n = 1,024, all valid", "k1", batches.ecdsa_batch(0, 1024, edge_every=0, seed=11)),
        ("2. ed25519 ecverify, n = 65,536, 25 % edge classes", "ed", batches.ed25519_batch(65536, edge_every=4, seed=12)),
        ("3. secp256r1 ecrecover, n = 65,536, high-s mix, 2 % invalid / edge rows", "r1",
         batches.ecdsa_batch(1, 65536, edge_every=50, seed=13, mix_high_s=True)),
    ]
    rows = []
    for name, kind, b in configs:
        n = b[0].shape[0]
        ins = [pin(a) for a in (b[:3] if kind == "ed" else b[:2])]
        po, vo = pin(np.zeros(n * 64, np.uint8))
        pt, vt = pin(np.zeros(n, np.uint8))

        def call():
            if kind == "k1":
                return lib.sigops_secp256k1_ecrecover(ins[0][0], ins[1][0], n, po, pt)
            if kind == "r1":
                return lib.sigops_secp256r1_ecrecover(ins[0][0], ins[1][0], n, po, pt)
            return lib.sigops_ed25519_ecverify(ins[0][0], ins[1][0], ins[2][0], n, po)

        for _ in range(3):
            assert call() == 0, lib.sigops_last_error()
        ts, ks = [], []
        for _ in range(reps):
            t0 = time.perf_counter()
            call()
            ts.append(time.perf_counter() - t0)
            h, k, d = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
            lib.sigops_last_timing(ctypes.byref(h), ctypes.byref(k), ctypes.byref(d))
            ks.append(k.value)
        if kind == "ed":
            assert (vo[:n] == b[3]).all(), name
            accepted = int(b[3].sum())
        else:
            assert (vo[: n * 64].reshape(n, 64) == b[2]).all() and (vt[:n] == b[3]).all(), name
            accepted = int(n - b[3].astype(bool).sum())
        ts.sort()
        ks.sort()
        p50, k50 = ts[len(ts) // 2], ks[len(ks) // 2]
        rows.append({"config": name, "n": n, "accepted_rows": accepted, "e2e_ms_p50": round(p50 * 1e3, 4), "e2e_ms_p99": round(ts[-1] * 1e3, 4),
                     "e2e_sigs_per_s": n / p50, "kernel_ms_p50": round(k50, 4), "kernel_sigs_per_s": n / k50 * 1e3,
                     "parity": "every row bit-exact vs the oracle-labelled batch", "host_buffers": "pinned (sigops_host_alloc)"})
        print(rows[-1], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "config_times.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
