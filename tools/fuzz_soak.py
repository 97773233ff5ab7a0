#!/usr/bin/env python3
"""Rejecting-path soak: batches of mostly-invalid rows (tests/fuzz_cases.py) per curve, a fresh seed per round, through
the host C ABI; every status / key / verdict is compared with the C oracle (test infrastructure) row by row.
   python tools/fuzz_soak.py [seconds [rows_per_batch]]   -> one line per batch, summary at the end"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import coracle  # noqa: E402
import fuzz_cases  # noqa: E402
import wgpu_sigops_b200 as w  # noqa: E402


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 200_000
    total = bad = rejected = 0
    t0 = time.time()
    r = 1  # seed 0 is the pytest run
    while time.time() - t0 < budget:
        for cid, mod in ((0, w.secp256k1_ecdsa), (1, w.secp256r1_ecdsa)):
            sigs, msgs = fuzz_cases.ecdsa_batch(cid, n, seed=r)
            exp_out, exp_st = coracle.ecrecover(cid, sigs, msgs)
            out, st = mod.ecrecover_with_status(sigs, msgs)
            nb = int(((st != exp_st) | (out != exp_out).any(axis=1)).sum())
            bad += nb
            total += n
            rejected += int(exp_st.astype(bool).sum())
            print(f"round {r} curve {cid}: {n} rows, {int(exp_st.astype(bool).sum())} rejected, {nb} mismatches", flush=True)
        sigs, msgs, pks = fuzz_cases.ed25519_batch(n, seed=r)
        exp = coracle.ecverify_ed25519(sigs, msgs, pks)
        got = w.ed25519_eddsa.ecverify_array(sigs, msgs, pks)
        nb = int((got != exp).sum())
        bad += nb
        total += n
        rejected += int((exp == 0).sum())
        print(f"round {r} ed25519: {n} rows, {int((exp == 0).sum())} rejected, {nb} mismatches", flush=True)
        r += 1
    print(f"FUZZ SOAK: {total} rows ({rejected} rejected by the oracle), {bad} mismatches, {time.time() - t0:.0f} s", flush=True)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
