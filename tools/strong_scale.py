#!/usr/bin/env python3
"""BASELINE config 4 as the drop-in API sees it: ONE 1,048,576-signature secp256k1 block per call, sharded by the
library over 1 / 2 / 4 / 8 GPUs of the box (single process, pinned host buffers, H2D + kernels + D2H inside the call).
   python tools/strong_scale.py [n]   -> gpurun_out/strong_scale.json"""
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


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
    lib = w.load()
    sigs, msgs, pks = coracle.gen_ecdsa(0, n, seed=61)
    import torch

    ngpu = torch.cuda.device_count()
    rows = []
    for g in (1, 2, 4, 8):
        if g > ngpu:
            break
        lib.sigops_shutdown()
        ids = (ctypes.c_int * g)(*range(g))
        assert lib.sigops_init(ids, g) == 0, lib.sigops_last_error()

        def pin(a):
            p = lib.sigops_host_alloc(a.nbytes)
            np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint8)), shape=(a.nbytes,))[:] = a.reshape(-1)
            return p

        ps, pm = pin(sigs), pin(msgs)
        po, pt = lib.sigops_host_alloc(n * 64), lib.sigops_host_alloc(n)
        for _ in range(3):
            assert lib.sigops_secp256k1_ecrecover(ps, pm, n, po, pt) == 0, lib.sigops_last_error()
        got = np.ctypeslib.as_array(ctypes.cast(po, ctypes.POINTER(ctypes.c_uint8)), shape=(n * 64,)).reshape(-1, 64)
        assert (got == pks).all()
        lat = []
        for _ in range(20):
            t0 = time.perf_counter()
            lib.sigops_secp256k1_ecrecover(ps, pm, n, po, pt)
            lat.append((time.perf_counter() - t0) * 1e3)
        p50 = float(np.percentile(lat, 50))
        rows.append({"gpus": g, "p50_ms": p50, "p99_ms": float(np.percentile(lat, 99)), "sigs_per_s": n / (p50 * 1e-3)})
        print(rows[-1], flush=True)
        for p in (ps, pm, po, pt):
            lib.sigops_host_free(p)
    json.dump({"what": f"one {n}-signature secp256k1 block per call, sharded over the first g GPUs by the library (strong scaling)",
               "rows": rows}, open(os.path.join(ROOT, "gpurun_out", "strong_scale.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
