#!/usr/bin/env python3
"""Quick look: IMAD peak micro-benchmark + per-curve kernel time through the host API (pool of valid signatures tiled
to n).  Development tool; bench.py is the measured contract."""
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import sigops_oracle as o  # noqa: E402
import wgpu_sigops_b200 as w  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
    lib = w.load()
    res = {"n": n}
    ops, ms = ctypes.c_double(), ctypes.c_double()
    names = {0: "imad", 1: "imad_wide", 2: "imad_wide_x_chain", 3: "iadd", 4: "imad_wide_plus_iadd", 5: "dfma", 6: "imad_hi"}
    for kind, name in names.items():
        rc = lib.sigops_imad_peak(kind, 4096, ctypes.byref(ops), ctypes.byref(ms))
        assert rc == 0, lib.sigops_last_error()
        res["peak_" + name] = {"ops_per_sec": ops.value, "ms": ms.value}
        print(name, "%.3e ops/s" % ops.value, "%.2f ms" % ms.value, flush=True)
    pool = 1024
    t0 = time.time()
    k1 = [o.gen_ecdsa_valid(o.K1, i) for i in range(pool)]
    r1 = [o.gen_ecdsa_valid(o.R1, i) for i in range(pool)]
    ed = [o.gen_ed25519_valid(i) for i in range(256)]
    print("pool gen %.1fs" % (time.time() - t0), flush=True)

    def tile(items, j, width):
        a = np.frombuffer(b"".join(x[j] for x in items), dtype=np.uint8).reshape(-1, width)
        return np.ascontiguousarray(np.tile(a, ((n + len(items) - 1) // len(items), 1))[:n])

    h2d, ker, d2h = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
    for name, items, mod in (("secp256k1", k1, w.secp256k1_ecdsa), ("secp256r1", r1, w.secp256r1_ecdsa)):
        sigs, msgs, exp = tile(items, 0, 64), tile(items, 1, 32), tile(items, 2, 64)
        for rep in range(3):
            t0 = time.time()
            out, st = mod.ecrecover_with_status(sigs, msgs)
            wall = time.time() - t0
            lib.sigops_last_timing(ctypes.byref(h2d), ctypes.byref(ker), ctypes.byref(d2h))
        ok = bool((out == exp).all() and not st.any())
        res[name] = {"kernel_ms": ker.value, "h2d_ms": h2d.value, "d2h_ms": d2h.value, "wall_ms": wall * 1e3,
                     "sigs_per_s_kernel": n / (ker.value * 1e-3), "parity": ok}
        print(name, res[name], flush=True)
    sigs, msgs, pks = tile(ed, 0, 64), tile(ed, 1, 32), tile(ed, 2, 32)
    for rep in range(3):
        t0 = time.time()
        v = w.ed25519_eddsa.ecverify_array(sigs, msgs, pks)
        wall = time.time() - t0
        lib.sigops_last_timing(ctypes.byref(h2d), ctypes.byref(ker), ctypes.byref(d2h))
    res["ed25519"] = {"kernel_ms": ker.value, "h2d_ms": h2d.value, "d2h_ms": d2h.value, "wall_ms": wall * 1e3,
                      "sigs_per_s_kernel": n / (ker.value * 1e-3), "parity": bool(v.all())}
    print("ed25519", res["ed25519"], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "quick_bench.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
