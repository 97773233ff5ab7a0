"""Synthetic batches for the BASELINE.json configurations, with expected outputs.

Valid signatures come from the C oracle's generator; edge classes from the Python oracle's corpus
(`ecdsa_edge_cases`, `ed25519_edge_cases`) plus random corruptions.  Expected outputs for corrupted / edge rows are
the C oracle's (itself pinned to the Python oracle on these classes by tests/test_oracle.py); for untouched valid rows
they are the signer's public key / True by construction."""
import numpy as np

import coracle
import sigops_oracle as o

ED_L = o.ED_L


def ecdsa_batch(curve_id: int, n: int, edge_every: int, seed: int = 0x51600002, mix_high_s: bool = False):
    """Returns (sigs, msgs, expected_pubkeys, expected_status, n_edge_rows)."""
    c = (o.K1, o.R1)[curve_id]
    sigs, msgs, pks = coracle.gen_ecdsa(curve_id, n, seed=seed, low_s=True)
    if mix_high_s:  # every other row produced without low-s normalisation (p256 / libsecp256k1 recover accept both)
        s2, m2, p2 = coracle.gen_ecdsa(curve_id, n, seed=seed + 1, low_s=False)
        sigs[1::2], msgs[1::2], pks[1::2] = s2[1::2], m2[1::2], p2[1::2]
    status = np.zeros(n, dtype=np.uint8)
    edge_rows = np.arange(edge_every - 1, n, edge_every) if edge_every else np.zeros(0, dtype=np.int64)
    if len(edge_rows):
        corpus = o.ecdsa_edge_cases(c)
        rng = np.random.default_rng(seed)
        for k, i in enumerate(edge_rows):
            kind = k % (len(corpus) + 3)
            if kind < len(corpus):
                _, sg, m = corpus[kind]
                sigs[i] = np.frombuffer(sg, dtype=np.uint8)
                msgs[i] = np.frombuffer(m, dtype=np.uint8)
            elif kind == len(corpus):  # corrupt r: usually not on the curve, else another key
                sigs[i, rng.integers(0, 32)] ^= 1 << rng.integers(0, 8)
            elif kind == len(corpus) + 1:  # corrupt s / parity: valid signature of another key
                sigs[i, rng.integers(32, 64)] ^= 1 << rng.integers(0, 8)
            else:  # corrupt the message
                msgs[i, rng.integers(0, 32)] ^= 1 << rng.integers(0, 8)
        e_out, e_st = coracle.ecrecover(curve_id, sigs[edge_rows], msgs[edge_rows])
        pks[edge_rows] = e_out
        status[edge_rows] = e_st
    return sigs, msgs, pks, status, len(edge_rows)


def ed25519_batch(n: int, edge_every: int, seed: int = 0x51600002):
    """Returns (sigs, msgs, pks, expected_valid, n_edge_rows)."""
    sigs, msgs, pks = coracle.gen_ed25519(n, seed=seed)
    valid = np.ones(n, dtype=np.uint8)
    edge_rows = np.arange(edge_every - 1, n, edge_every) if edge_every else np.zeros(0, dtype=np.int64)
    if len(edge_rows):
        corpus = o.ed25519_edge_cases()
        rng = np.random.default_rng(seed)
        for k, i in enumerate(edge_rows):
            kind = k % (len(corpus) + 6)
            if kind < len(corpus):
                _, sg, m, pk = corpus[kind]
                sigs[i] = np.frombuffer(sg, dtype=np.uint8)
                msgs[i] = np.frombuffer(m, dtype=np.uint8)
                pks[i] = np.frombuffer(pk, dtype=np.uint8)
            elif kind == len(corpus):  # bit flip in R
                sigs[i, rng.integers(0, 32)] ^= 1 << rng.integers(0, 8)
            elif kind == len(corpus) + 1:  # bit flip in s
                sigs[i, rng.integers(32, 64)] ^= 1 << rng.integers(0, 8)
            elif kind == len(corpus) + 2:  # bit flip in A
                pks[i, rng.integers(0, 32)] ^= 1 << rng.integers(0, 8)
            elif kind == len(corpus) + 3:  # bit flip in M
                msgs[i, rng.integers(0, 32)] ^= 1 << rng.integers(0, 8)
            elif kind == len(corpus) + 4:  # s + L: same residue, non-canonical encoding -> reject
                s = int.from_bytes(sigs[i, 32:].tobytes(), "little") + ED_L
                sigs[i, 32:] = np.frombuffer(s.to_bytes(32, "little"), dtype=np.uint8)
            else:  # top bits of s set
                sigs[i, 63] |= 0xE0
        valid[edge_rows] = coracle.ecverify_ed25519(sigs[edge_rows], msgs[edge_rows], pks[edge_rows])
    return sigs, msgs, pks, valid, len(edge_rows)
