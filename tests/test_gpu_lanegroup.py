"""The lane-group kernels (csrc/group.cuh: several cooperating warps per 32 signatures, complete projective / four-way
Edwards formulas) on the GPU.  By default the host picks them for small batches; SIGOPS_FORCE_LANEGROUP=1 routes EVERY size
through them, so the suites that define parity for the one-thread-per-signature kernels -- unit shims, golden fixtures, the
edge corpus, the BASELINE configurations, fuzz rows -- are repeated here on the group kernels and must be bit-exact too.
Replaces nothing in the reference: its only strategy is one invocation per signature
(src/wgsl/main/secp256k1_ecdsa_main_0.wgsl:21-30)."""
import json
import os

import numpy as np
import pytest

import batches
import coracle
import fuzz_cases
import sigops_oracle as o
import unit_checks as uc

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["products_inlined", "products_out_of_line"])
def forced(monkeypatch, request):
    """Every size through the lane-group kernels, once per flavour (the host switches from inlined to out-of-line field
    products at SIGOPS_GROUP_COLD_MIN signatures; secp256r1 always runs the out-of-line one)."""
    monkeypatch.setenv("SIGOPS_FORCE_LANEGROUP", "1")
    monkeypatch.setenv("SIGOPS_GROUP_COLD_MIN", "0" if request.param == "products_out_of_line" else "1000000000")


def test_group_unit_shims(gpu_units):
    uc.check_group_curves(gpu_units, n=64)


@pytest.mark.parametrize("curve", ["secp256k1", "secp256r1"])
def test_group_ecrecover_edge_cases(sigops, forced, curve):
    c = o.K1 if curve == "secp256k1" else o.R1
    mod = sigops.secp256k1_ecdsa if curve == "secp256k1" else sigops.secp256r1_ecdsa
    cases = uc.ecdsa_cases(c, nvalid=300)
    out, st = mod.ecrecover_with_status([x[1] for x in cases], [x[2] for x in cases])
    uc.check_ecrecover_against_oracle(c, cases, out, st)
    for n in (1, 10):  # the reference's E2E sizes
        sigs, msgs, pks = zip(*[o.gen_ecdsa_valid(c, 1000 + i) for i in range(n)])
        assert mod.ecrecover_single_shader(list(sigs), list(msgs), 13) == list(pks)


def test_group_ed25519_edge_cases(sigops, forced):
    cases = uc.ed_cases(nvalid=300)
    valid = sigops.ed25519_eddsa.ecverify_array([x[1] for x in cases], [x[2] for x in cases], [x[3] for x in cases])
    uc.check_ed_against_oracle(cases, valid)


def test_group_golden_fixtures(sigops, forced):
    gdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    for name, mod in (("secp256k1", sigops.secp256k1_ecdsa), ("secp256r1", sigops.secp256r1_ecdsa)):
        rows = json.load(open(os.path.join(gdir, f"{name}_ecrecover.json")))["cases"]
        out, st = mod.ecrecover_with_status([bytes.fromhex(r["sig"]) for r in rows], [bytes.fromhex(r["msg"]) for r in rows])
        for r, ob, sb in zip(rows, out, st):
            if r["pubkey"] is None:
                assert sb == 1 and not ob.any(), (name, r["label"])
            else:
                assert sb == 0 and ob.tobytes().hex() == r["pubkey"], (name, r["label"])
    rows = json.load(open(os.path.join(gdir, "ed25519_ecverify.json")))["cases"]
    got = sigops.ed25519_eddsa.ecverify_array(*[[bytes.fromhex(r[k]) for r in rows] for k in ("sig", "msg", "pk")])
    for r, v in zip(rows, got):
        assert bool(v) == r["valid"], r["label"]


@pytest.mark.parametrize("n", [1, 31, 32, 33, 1024, 4736, 4737, 20011, 65536])
def test_group_config_batches(sigops, forced, n):
    """BASELINE configurations 1-3 and ragged sizes around the 32-signature block, edge rows included, all through the
    group kernels (grid-stride over blocks beyond the resident set)."""
    s, m, pk, st, _ = batches.ecdsa_batch(0, n, edge_every=97, seed=31)
    out, got = sigops.secp256k1_ecdsa.ecrecover_with_status(s, m)
    assert (out == pk).all() and (got == st).all()
    s, m, pk, st, _ = batches.ecdsa_batch(1, n, edge_every=89, seed=32, mix_high_s=True)
    out, got = sigops.secp256r1_ecdsa.ecrecover_with_status(s, m)
    assert (out == pk).all() and (got == st).all()
    s, m, pk, v, _ = batches.ed25519_batch(n, edge_every=4, seed=33)
    assert (sigops.ed25519_eddsa.ecverify_array(s, m, pk) == v).all()


def test_group_fuzz(sigops, forced):
    n = 100000
    for cid, mod in ((0, sigops.secp256k1_ecdsa), (1, sigops.secp256r1_ecdsa)):
        sigs, msgs = fuzz_cases.ecdsa_batch(cid, n, seed=13)
        exp_out, exp_st = coracle.ecrecover(cid, sigs, msgs)
        out, st = mod.ecrecover_with_status(sigs, msgs)
        assert not (st != exp_st).any() and not (out != exp_out).any()
    sigs, msgs, pks = fuzz_cases.ed25519_batch(n, seed=13)
    exp = coracle.ecverify_ed25519(sigs, msgs, pks)
    assert not (sigops.ed25519_eddsa.ecverify_array(sigs, msgs, pks) != exp).any()


def test_default_selection_and_agreement(sigops, monkeypatch):
    """Without the override the host uses the group kernels up to one block per SM and the one-thread-per-signature
    kernels above; SIGOPS_LANEGROUP=0 disables them.  Same bytes either way, and the queue mode rides on the same choice."""
    n = 3000
    s, m, pk, st, _ = batches.ecdsa_batch(0, n, edge_every=53, seed=41)
    out_a, st_a = sigops.secp256k1_ecdsa.ecrecover_with_status(s, m)
    monkeypatch.setenv("SIGOPS_LANEGROUP", "0")
    out_b, st_b = sigops.secp256k1_ecdsa.ecrecover_with_status(s, m)
    monkeypatch.delenv("SIGOPS_LANEGROUP")
    assert (out_a == pk).all() and (st_a == st).all() and (out_b == out_a).all() and (st_b == st_a).all()
    with sigops.service.SigQueue("secp256k1", n, depth=2) as q:
        for slot in range(2):
            q.sigs(slot)[:n] = s
            q.msgs(slot)[:n] = m
            q.submit(slot, n)
        for slot in range(2):
            out, got = q.wait(slot)
            assert (out == pk).all() and (got == st).all()
