"""Differential fuzzing at scale: mostly-invalid inputs (random bytes, values hugging 0 / n / p / 2^255, single bit
flips) through the CUDA path against the C oracle, row by row.  Complements the configuration tests, whose rows are
mostly valid signatures.  tools/fuzz_soak.py runs the same builders over many seeds."""
import numpy as np
import pytest

import coracle
import fuzz_cases

pytestmark = pytest.mark.gpu
N = 200_000


@pytest.mark.parametrize("curve,cid", [("secp256k1", 0), ("secp256r1", 1)])
def test_fuzz_ecrecover(sigops, curve, cid):
    sigs, msgs = fuzz_cases.ecdsa_batch(cid, N, seed=0)
    exp_out, exp_st = coracle.ecrecover(cid, sigs, msgs)
    mod = sigops.secp256k1_ecdsa if cid == 0 else sigops.secp256r1_ecdsa
    out, st = mod.ecrecover_with_status(sigs, msgs)
    bad = np.nonzero((st != exp_st) | (out != exp_out).any(axis=1))[0]
    assert len(bad) == 0, (bad[:5], len(bad))
    assert 0.1 < exp_st.mean() < 0.6  # a healthy share of rejections


def test_fuzz_ed25519(sigops):
    sigs, msgs, pks = fuzz_cases.ed25519_batch(N, seed=0)
    exp = coracle.ecverify_ed25519(sigs, msgs, pks)
    got = sigops.ed25519_eddsa.ecverify_array(sigs, msgs, pks)
    bad = np.nonzero(got != exp)[0]
    assert len(bad) == 0, (bad[:5], len(bad))
    assert 0.15 < exp.mean() < 0.4
