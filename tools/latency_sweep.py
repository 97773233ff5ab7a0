#!/usr/bin/env python3
"""p50 latency of the blocking C-ABI calls for small batches (pinned buffers), lane-group kernels on / off.
   python tools/latency_sweep.py [reps]            (under gpurun)
   LAT_SIZES=21,1024  LAT_MODES=group,thread  LAT_OUT=<file>  narrow the sweep."""
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
    assert lib.sigops_num_devices() >= 1
    nmax = 40000
    k1 = batches.ecdsa_batch(0, nmax, edge_every=1000, seed=1)
    r1 = batches.ecdsa_batch(1, nmax, edge_every=1000, seed=2, mix_high_s=True)
    ed = batches.ed25519_batch(nmax, edge_every=100, seed=3)

    def pin(a):
        p = lib.sigops_host_alloc(a.nbytes)
        v = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint8)), shape=(a.nbytes,))
        v[:] = a.reshape(-1)
        return p, v

    P = {"k1": [pin(a) for a in k1[:2]], "r1": [pin(a) for a in r1[:2]], "ed": [pin(a) for a in ed[:3]]}
    po, vo = pin(np.zeros(nmax * 64, np.uint8))
    pt, vt = pin(np.zeros(nmax, np.uint8))
    res = {}
    sizes = [int(x) for x in os.environ.get("LAT_SIZES", "21,256,1024,1365,2048,4736,9472,16384").split(",")]
    modes = os.environ.get("LAT_MODES", "group,thread,group_forced").split(",")
    for mode, env in (("group", {}), ("thread", {"SIGOPS_LANEGROUP": "0"}), ("group_forced", {"SIGOPS_FORCE_LANEGROUP": "1"})):
        if mode not in modes:
            continue
        for k, v in env.items():
            os.environ[k] = v
        for n in sizes:
            row = {}
            for c in ("k1", "r1", "ed"):
                def call():
                    if c == "k1":
                        return lib.sigops_secp256k1_ecrecover(P[c][0][0], P[c][1][0], n, po, pt)
                    if c == "r1":
                        return lib.sigops_secp256r1_ecrecover(P[c][0][0], P[c][1][0], n, po, pt)
                    return lib.sigops_ed25519_ecverify(P[c][0][0], P[c][1][0], P[c][2][0], n, po)
                for _ in range(3):
                    assert call() == 0, lib.sigops_last_error()
                ts = []
                for _ in range(reps):
                    t0 = time.perf_counter()
                    call()
                    ts.append(time.perf_counter() - t0)
                src = k1 if c == "k1" else r1 if c == "r1" else ed
                if c == "ed":
                    assert (vo[:n] == src[3][:n]).all(), (mode, c, n)
                else:
                    assert (vo[: n * 64].reshape(n, 64) == src[2][:n]).all() and (vt[:n] == src[3][:n]).all(), (mode, c, n)
                ts.sort()
                row[c] = round(ts[len(ts) // 2] * 1e3, 4)
                h, k, dd = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
                lib.sigops_last_timing(ctypes.byref(h), ctypes.byref(k), ctypes.byref(dd))
                row[c + "_kernel"] = round(k.value, 4)
            res.setdefault(mode, {})[n] = row
            print(mode, n, row, flush=True)
        for k in env:
            del os.environ[k]
    json.dump(res, open(os.environ.get("LAT_OUT", os.path.join(ROOT, "gpurun_out", "latency_sweep.json")), "w"), indent=1)


if __name__ == "__main__":
    main()
