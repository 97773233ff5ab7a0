"""message hashing (`Message::new` = SHA-256) before and address derivation (SHA-256(X || Y)) after
recovery, on the device.  Expected values: hashlib + the oracle's recovery."""
import hashlib
import random

import numpy as np
import pytest

import coracle
import sigops_oracle as o
from simlib import load_hostsim

LENGTHS = [0, 1, 3, 31, 32, 33, 54, 55, 56, 57, 63, 64, 65, 119, 120, 127, 128, 129, 1000, 5000]


def _msgs(seed=1):
    rng = random.Random(seed)
    return [bytes(rng.getrandbits(8) for _ in range(ln)) for ln in LENGTHS]


def test_hostsim_sha256(sim_units):
    lib = load_hostsim()
    msgs = _msgs()
    offs = np.zeros(len(msgs) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(m) for m in msgs], dtype=np.uint64)
    out = np.zeros((len(msgs), 32), dtype=np.uint8)
    lib.hostsim_sha256(b"".join(msgs), offs.ctypes.data, len(msgs), out.ctypes.data)
    for m, d in zip(msgs, out):
        assert d.tobytes() == hashlib.sha256(m).digest(), len(m)
    pk = bytes(range(64))
    w = sim_units.run_words("SHA256_64", np.frombuffer(pk, dtype=np.uint32))[0]
    assert w.tobytes() == hashlib.sha256(pk).digest()


@pytest.mark.gpu
def test_gpu_sha256_batch(sigops, gpu_units):
    msgs = _msgs() * 50
    out = sigops.pipeline.sha256_batch(msgs)
    for m, d in zip(msgs, out):
        assert d.tobytes() == hashlib.sha256(m).digest(), len(m)
    pk = bytes(range(64))
    assert gpu_units.run_words("SHA256_64", np.frombuffer(pk, dtype=np.uint32))[0].tobytes() == hashlib.sha256(pk).digest()
    assert sigops.pipeline.sha256_batch([]).shape == (0, 32)


@pytest.mark.gpu
@pytest.mark.parametrize("curve,cid", [("secp256k1", 0), ("secp256r1", 1)])
def test_gpu_raw_messages_to_addresses(sigops, curve, cid):
    """Raw transaction bytes -> SHA-256 -> recover -> SHA-256(X || Y): equals hashing on the host around the oracle."""
    c = (o.K1, o.R1)[cid]
    rng = random.Random(7 + cid)
    n = 3000
    raw, sigs, want_pk = [], [], []
    for i in range(n):
        m = bytes(rng.getrandbits(8) for _ in range(rng.randrange(0, 200)))
        z = hashlib.sha256(m).digest()
        d = rng.randrange(1, c.n)
        k = rng.randrange(1, c.n)
        if i % 40 == 0:  # cheap to generate: reuse the oracle's signer on a few rows, the C generator elsewhere
            sg = o.ecdsa_sign(c, d, z, k, low_s=True)
            Q = o.sw_mul(c, d, (c.gx, c.gy))
            pk = Q[0].to_bytes(32, "big") + Q[1].to_bytes(32, "big")
        else:
            sg, pk = None, None
        raw.append(m), sigs.append(sg), want_pk.append(pk)
    # the remaining rows: signatures from the C generator are over ITS messages, so build those rows the other way round:
    # take generator (sig, prehash) pairs and treat them as prehashed input in the second half of the test
    idx = [i for i in range(n) if sigs[i] is not None]
    addr, pks, st = sigops.pipeline.ecrecover_addresses(curve, [sigs[i] for i in idx], [raw[i] for i in idx])
    assert not st.any()
    for j, i in enumerate(idx):
        assert pks[j].tobytes() == want_pk[i]
        assert addr[j].tobytes() == hashlib.sha256(want_pk[i]).digest()
    # corrupt one message byte: recovery yields a different key (or fails), address follows the oracle
    bad = [raw[i] + b"x" for i in idx]
    addr2, pks2, st2 = sigops.pipeline.ecrecover_addresses(curve, [sigs[i] for i in idx], bad)
    for j, i in enumerate(idx):
        exp = o.ecrecover(c, sigs[i], hashlib.sha256(bad[j]).digest())
        if exp is None:
            assert st2[j] == 1 and not pks2[j].any() and not addr2[j].any()
        else:
            assert st2[j] == 0 and pks2[j].tobytes() == exp and addr2[j].tobytes() == hashlib.sha256(exp).digest()
    # prehashed mode on a larger generated batch, incl. rejected rows (zero address)
    g_sigs, g_msgs, g_pks = coracle.gen_ecdsa(cid, 20000, seed=77, low_s=(cid == 0))
    g_sigs = g_sigs.copy()
    g_sigs[::97, 5] ^= 0x40  # corrupt r on some rows
    e_out, e_st = coracle.ecrecover(cid, g_sigs, g_msgs)
    addr3, pks3, st3 = sigops.pipeline.ecrecover_addresses(curve, g_sigs, g_msgs, prehashed=True)
    assert (pks3 == e_out).all() and (st3 == e_st).all() and 0 < int(e_st.sum()) < 20000
    for i in range(0, 20000, 137):
        want = bytes(32) if e_st[i] else hashlib.sha256(e_out[i].tobytes()).digest()
        assert addr3[i].tobytes() == want
    bad_rows = np.nonzero(e_st)[0]
    assert not addr3[bad_rows].any()
