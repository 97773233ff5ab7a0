"""Differential fuzzing at scale: mostly-invalid inputs (random bytes, values hugging 0 / n / p / 2^255) through the CUDA
path against the C oracle, row by row.  Complements the configuration tests, whose rows are mostly valid signatures."""
import numpy as np
import pytest

import coracle
import sigops_oracle as o

pytestmark = pytest.mark.gpu
N = 200_000


def _be(x):
    return np.frombuffer(int(x).to_bytes(32, "big"), dtype=np.uint8)


@pytest.mark.parametrize("curve,cid", [("secp256k1", 0), ("secp256r1", 1)])
def test_fuzz_ecrecover(sigops, curve, cid):
    c = (o.K1, o.R1)[cid]
    rng = np.random.default_rng(100 + cid)
    sigs, msgs, _ = coracle.gen_ecdsa(cid, N, seed=900 + cid, low_s=False)
    sigs, msgs = sigs.copy(), msgs.copy()
    q = N // 4
    sigs[:q] = rng.integers(0, 256, size=(q, 64), dtype=np.uint8)  # random r (half of them off the curve), s, parity
    msgs[:q] = rng.integers(0, 256, size=(q, 32), dtype=np.uint8)
    sigs[q:2 * q, 32:] = rng.integers(0, 256, size=(q, 32), dtype=np.uint8)  # valid r, random s / parity
    special = [0, 1, 2, 3, c.n - 2, c.n - 1, c.n, c.n + 1, c.p - 1, c.p, c.p + 1, 2**255 - 1, 2**255, 2**256 - 1,
               2**128, 2**224, 2**192 + 2**96, c.n // 2, c.n // 2 + 1, 7, c.gx]
    for i in range(2 * q, 3 * q):  # special values in r, s or z
        v = special[i % len(special)]
        where = (i // len(special)) % 3
        if where == 0:
            sigs[i, :32] = _be(v)
        elif where == 1:
            sigs[i, 32:] = _be(v % 2**255)
            sigs[i, 32] |= (i & 1) << 7
        else:
            msgs[i] = _be(v)
    exp_out, exp_st = coracle.ecrecover(cid, sigs, msgs)
    mod = sigops.secp256k1_ecdsa if cid == 0 else sigops.secp256r1_ecdsa
    out, st = mod.ecrecover_with_status(sigs, msgs)
    bad = np.nonzero((st != exp_st) | (out != exp_out).any(axis=1))[0]
    assert len(bad) == 0, (bad[:5], len(bad))
    assert 0.1 < exp_st.mean() < 0.6  # a healthy share of rejections


def test_fuzz_ed25519(sigops):
    rng = np.random.default_rng(200)
    sigs, msgs, pks = coracle.gen_ed25519(N, seed=910)
    sigs, msgs, pks = sigs.copy(), msgs.copy(), pks.copy()
    q = N // 4
    sigs[:q] = rng.integers(0, 256, size=(q, 64), dtype=np.uint8)
    pks[:q] = rng.integers(0, 256, size=(q, 32), dtype=np.uint8)
    pks[q:2 * q] = rng.integers(0, 256, size=(q, 32), dtype=np.uint8)  # valid signature under a random (maybe invalid) key
    sigs[2 * q:3 * q, 32:] = rng.integers(0, 256, size=(q, 32), dtype=np.uint8)  # random s (mostly non-canonical)
    sigs[2 * q:3 * q:2, 63] &= 0x0F  # ... half of them canonical-range
    special = [0, 1, o.ED_P - 1, o.ED_P, o.ED_P + 1, 2**255 - 1, 2**255 - 19 + 2**255, 2**256 - 1, o.ED_L, o.ED_L - 1]
    for i in range(3 * q, min(N, 3 * q + 3000)):  # special y in the key / R, special s
        v = special[i % len(special)]
        le = np.frombuffer(int(v % 2**256).to_bytes(32, "little"), dtype=np.uint8)
        where = (i // len(special)) % 3
        if where == 0:
            pks[i] = le
        elif where == 1:
            sigs[i, :32] = le
        else:
            sigs[i, 32:] = le
    exp = coracle.ecverify_ed25519(sigs, msgs, pks)
    got = sigops.ed25519_eddsa.ecverify_array(sigs, msgs, pks)
    bad = np.nonzero(got != exp)[0]
    assert len(bad) == 0, (bad[:5], len(bad))
    assert 0.15 < exp.mean() < 0.4
