#!/usr/bin/env python3
"""Parity soak: `rounds` batches of 1,048,576 UNIQUE valid signatures per curve, each from a fresh generator seed, through the
host C ABI; every recovered key / verdict is compared with the signer's key (known by construction from the generator --
test infrastructure).  The once-in-2^31 fix-up paths of the field arithmetic are expected about three times per
1M-signature secp256k1 batch (6.4e9 field operations), so this is also their at-scale check.
   python tools/soak.py [rounds [first_round]]   -> one line per batch, summary at the end (first_round offsets the seeds
   so that a later soak extends an earlier one instead of repeating it)"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import coracle  # noqa: E402
import wgpu_sigops_b200 as w  # noqa: E402


def main():
    rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    n = 1 << 20
    threads = coracle.host_threads()
    total = bad = 0
    t0 = time.time()
    for r in range(first, first + rounds):
        seed = 0x50A40000 + r
        for cid, mod in ((0, w.secp256k1_ecdsa), (1, w.secp256r1_ecdsa)):
            s, m, pk = coracle.gen_ecdsa(cid, n, seed=seed, low_s=(r % 2 == 0), threads=threads)
            out, st = mod.ecrecover_with_status(s, m)
            nb = int((out != pk).any(axis=1).sum()) + int(st.astype(bool).sum())
            bad += nb
            total += n
            print(f"round {r} curve {cid}: {n} signatures, {nb} mismatches", flush=True)
        s, m, pk = coracle.gen_ed25519(n, seed=seed, threads=threads)
        v = w.ed25519_eddsa.ecverify_array(s, m, pk)
        nb = int((v != 1).sum())
        bad += nb
        total += n
        print(f"round {r} ed25519: {n} signatures, {nb} mismatches", flush=True)
    print(f"SOAK: {total} signatures, {bad} mismatches, {time.time() - t0:.0f} s", flush=True)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
