"""Parity at the BASELINE.json configurations (full sizes), through the reference-shaped API over the C ABI.

config 1  secp256k1 ecrecover, 1,024 random signatures                  (the reference's own benchmark size)
config 2  ed25519 ecverify, 65,536 signatures, 25% dalek edge classes   (invalid, non-canonical, small order, ...)
config 3  secp256r1 ecrecover, 65,536 signatures, high-s mix, 2% invalid
config 4  secp256k1 ecrecover, 1,048,576-signature block, 0.1% edge classes
config 5  mixed k1 / r1 / ed25519 batch-size sweep
Bit-exact comparison of every row; at 1M rows additionally the size-independent properties (idempotence, permutation
equivariance, recover(sign(d)) == d*G for every untouched row)."""
import numpy as np
import pytest

import batches
import coracle

pytestmark = pytest.mark.gpu


def test_config1_k1_1024(sigops):
    sigs, msgs, pks = coracle.gen_ecdsa(0, 1024, seed=2)  # reference benchmark: seed 2, n = 1024, check = true
    table = sigops.precompute.secp256k1_bases(13)
    got = sigops.secp256k1_ecdsa.ecrecover([r.tobytes() for r in sigs], [r.tobytes() for r in msgs], table, 13)
    assert got == [r.tobytes() for r in pks]
    o_out, o_st = coracle.ecrecover(0, sigs, msgs)
    assert (o_out == pks).all() and not o_st.any()


def test_config2_ed25519_65536_edge_classes(sigops):
    sigs, msgs, pks, exp, n_edge = batches.ed25519_batch(65536, edge_every=4)
    assert n_edge == 16384 and 0 < int(exp[3::4].sum()) < n_edge  # both outcomes among the edge rows
    got = sigops.ed25519_eddsa.ecverify_array(sigs, msgs, pks)
    bad = np.nonzero(got != exp)[0]
    assert len(bad) == 0, bad[:10]
    full = coracle.ecverify_ed25519(sigs, msgs, pks)  # the whole batch through the oracle as well
    assert (full == got).all()


def test_config3_r1_65536(sigops):
    sigs, msgs, pks, st, n_edge = batches.ecdsa_batch(1, 65536, edge_every=50, mix_high_s=True)
    assert n_edge == 1310 and 0 < int(st.sum()) < n_edge
    out, got_st = sigops.secp256r1_ecdsa.ecrecover_with_status(sigs, msgs)
    assert (got_st == st).all() and (out == pks).all()
    o_out, o_st = coracle.ecrecover(1, sigs, msgs)
    assert (o_out == out).all() and (o_st == got_st).all()


def test_config4_k1_1m_block(sigops):
    n = 1 << 20
    sigs, msgs, pks, st, n_edge = batches.ecdsa_batch(0, n, edge_every=1000)
    assert n_edge == 1048
    out, got_st = sigops.secp256k1_ecdsa.ecrecover_with_status(sigs, msgs)
    assert (got_st == st).all()
    assert (out == pks).all()  # every untouched row: recover(sign_d(m)) == d*G; edge rows: the oracle's answer
    # oracle on a 65,536-row random subset
    idx = np.random.default_rng(4).choice(n, 65536, replace=False)
    o_out, o_st = coracle.ecrecover(0, sigs[idx], msgs[idx])
    assert (o_out == out[idx]).all() and (o_st == got_st[idx]).all()
    # idempotence and permutation equivariance (no cross-row state, no index arithmetic slip at scale)
    out2, st2 = sigops.secp256k1_ecdsa.ecrecover_with_status(sigs, msgs)
    assert (out2 == out).all() and (st2 == got_st).all()
    perm = np.random.default_rng(5).permutation(n)
    out3, st3 = sigops.secp256k1_ecdsa.ecrecover_with_status(sigs[perm], msgs[perm])
    assert (out3 == out[perm]).all() and (st3 == got_st[perm]).all()


@pytest.mark.parametrize("n", [64, 1000, 4096, 65536, 262144])
def test_config5_mixed_sweep(sigops, n):
    k = n // 3 + 1
    s, m, pk, st, _ = batches.ecdsa_batch(0, k, edge_every=97, seed=11)
    out, got = sigops.secp256k1_ecdsa.ecrecover_with_status(s, m)
    assert (out == pk).all() and (got == st).all()
    s, m, pk, st, _ = batches.ecdsa_batch(1, k, edge_every=89, seed=12, mix_high_s=True)
    out, got = sigops.secp256r1_ecdsa.ecrecover_with_status(s, m)
    assert (out == pk).all() and (got == st).all()
    s, m, pk, v, _ = batches.ed25519_batch(k, edge_every=5, seed=13)
    assert (sigops.ed25519_eddsa.ecverify_array(s, m, pk) == v).all()


@pytest.mark.parametrize("n", [75776 + 1, 75776 + 40000, 2 * 75776 + 56832, 3 * 75776 + 5])
def test_tail_launch_boundaries(sigops, n):
    """Shards of w.f waves (one wave = 148 SMs x 512 threads on a B200) with f <= 3/4 run as a main launch of whole waves
    plus a separate, smaller launch for the tail: rows on both sides of the seam, edge rows included, must be exact, and the
    result must not depend on the split (SIGOPS_TAIL_SPLIT=0 is the single-launch path)."""
    import os

    s, m, pk, st, _ = batches.ecdsa_batch(0, n, edge_every=101, seed=21)
    out, got = sigops.secp256k1_ecdsa.ecrecover_with_status(s, m)
    assert (out == pk).all() and (got == st).all()
    s2, m2, pk2, v2, _ = batches.ed25519_batch(n, edge_every=7, seed=22)
    got_v = sigops.ed25519_eddsa.ecverify_array(s2, m2, pk2)
    assert (got_v == v2).all()
    os.environ["SIGOPS_TAIL_SPLIT"] = "0"
    try:
        out0, got0 = sigops.secp256k1_ecdsa.ecrecover_with_status(s, m)
        v0 = sigops.ed25519_eddsa.ecverify_array(s2, m2, pk2)
    finally:
        del os.environ["SIGOPS_TAIL_SPLIT"]
    assert (out0 == out).all() and (got0 == got).all() and (v0 == got_v).all()


def test_multi_device_sharding_matches_single(sigops):
    """When the box has more than one GPU the host entry point shards the batch; the result must not depend on it."""
    lib = sigops.load()
    if lib.sigops_num_devices() < 2:
        pytest.skip("one visible GPU")
    s, m, pk, st, _ = batches.ecdsa_batch(0, 100003, edge_every=101, seed=21)
    out, got = sigops.secp256k1_ecdsa.ecrecover_with_status(s, m)
    assert (out == pk).all() and (got == st).all()
    s, m, p, v, _ = batches.ed25519_batch(100003, edge_every=7, seed=22)
    assert (sigops.ed25519_eddsa.ecverify_array(s, m, p) == v).all()


def test_sub_shards_cover_large_batches(sigops, monkeypatch):
    """Batches beyond the device-buffer bound are processed in sub-shards (2^24 signatures in production; forced down
    to 10,000 here): results identical to the single-pass ones, ragged last sub-shard included."""
    s, m, pk, st, _ = batches.ecdsa_batch(0, 34567, edge_every=113, seed=41)
    monkeypatch.setenv("SIGOPS_MAX_SUBSHARD", "10000")
    out, got = sigops.secp256k1_ecdsa.ecrecover_with_status(s, m)
    assert (out == pk).all() and (got == st).all()
    s2, m2, p2, v, _ = batches.ed25519_batch(23456, edge_every=9, seed=42)
    assert (sigops.ed25519_eddsa.ecverify_array(s2, m2, p2) == v).all()


def test_concurrent_callers_are_serialised(sigops):
    """The entry points are blocking and thread-safe (one internal lock): four threads calling different operations at
    once all get the right answers (the reference serialises its GPU tests with serial_test instead,
    src/tests/secp256k1_ecdsa.rs:11-13)."""
    import threading

    k1 = batches.ecdsa_batch(0, 20011, edge_every=101, seed=51)
    r1 = batches.ecdsa_batch(1, 15013, edge_every=103, seed=52)
    ed = batches.ed25519_batch(17017, edge_every=7, seed=53)
    errors = []

    def run_k1():
        for _ in range(3):
            out, st = sigops.secp256k1_ecdsa.ecrecover_with_status(k1[0], k1[1])
            if not ((out == k1[2]).all() and (st == k1[3]).all()):
                errors.append("k1")

    def run_r1():
        for _ in range(3):
            out, st = sigops.secp256r1_ecdsa.ecrecover_with_status(r1[0], r1[1])
            if not ((out == r1[2]).all() and (st == r1[3]).all()):
                errors.append("r1")

    def run_ed():
        for _ in range(3):
            if not (sigops.ed25519_eddsa.ecverify_array(ed[0], ed[1], ed[2]) == ed[3]).all():
                errors.append("ed")

    threads = [threading.Thread(target=f) for f in (run_k1, run_r1, run_ed, run_k1)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
