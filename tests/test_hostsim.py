"""CPU-only logic tests: the CUDA headers compiled for the host (portable bodies instead of inline PTX) against the
oracle.  These catch algorithmic mistakes without a GPU; the PTX bodies themselves are covered by tests/test_gpu_*.py."""
import numpy as np

import sigops_oracle as o
import unit_checks as uc
from simlib import load_hostsim


def test_field_k1(sim_units):
    uc.check_field(sim_units, "K1", o.K1.p, True)


def test_field_ed25519(sim_units):
    uc.check_field(sim_units, "ED", o.ED_P, True)


def test_field_r1(sim_units):
    uc.check_field(sim_units, "R1", o.R1.p, True)


def test_raw_fixup_paths(sim_units):
    uc.check_raw_fixups(sim_units)


def test_wide_products(sim_units):
    uc.check_wide(sim_units)


def test_addition_chains(sim_units):
    uc.check_chains(sim_units)


def test_scalar_fields(sim_units):
    uc.check_scalar(sim_units)


def test_sha512(sim_units):
    uc.check_sha512(sim_units)


def test_glv(sim_units):
    uc.check_glv(sim_units)


def test_curves(sim_units):
    uc.check_curves(sim_units)


def test_ecrecover_logic():
    lib = load_hostsim()
    for cid, c in ((0, o.K1), (1, o.R1)):
        cases = uc.ecdsa_cases(c)
        n = len(cases)
        out = np.zeros((n, 64), dtype=np.uint8)
        st = np.zeros(n, dtype=np.uint8)
        lib.hostsim_ecrecover(cid, b"".join(x[1] for x in cases), b"".join(x[2] for x in cases), n, out.ctypes.data, st.ctypes.data)
        uc.check_ecrecover_against_oracle(c, cases, out, st)


def test_ed25519_logic():
    lib = load_hostsim()
    cases = uc.ed_cases()
    n = len(cases)
    out = np.zeros(n, dtype=np.uint8)
    lib.hostsim_ed25519_verify(b"".join(x[1] for x in cases), b"".join(x[2] for x in cases), b"".join(x[3] for x in cases), n, out.ctypes.data)
    uc.check_ed_against_oracle(cases, out)


def test_fuzz_builders_through_host_simulation():
    """The mostly-invalid batches of tests/fuzz_cases.py (the GPU fuzz tests and tools/fuzz_soak.py use them at scale),
    small, through the host-compiled kernel logic against the C oracle: rejections, bit flips and special values."""
    import coracle
    import fuzz_cases

    lib = load_hostsim()
    n = 4000
    for cid in (0, 1):
        sigs, msgs = fuzz_cases.ecdsa_batch(cid, n, seed=7)
        exp_out, exp_st = coracle.ecrecover(cid, sigs, msgs)
        out = np.zeros((n, 64), dtype=np.uint8)
        st = np.zeros(n, dtype=np.uint8)
        lib.hostsim_ecrecover(cid, sigs.tobytes(), msgs.tobytes(), n, out.ctypes.data, st.ctypes.data)
        assert not (st != exp_st).any() and not (out != exp_out).any()
        assert 0.1 < exp_st.mean() < 0.6
    sigs, msgs, pks = fuzz_cases.ed25519_batch(n, seed=7)
    exp = coracle.ecverify_ed25519(sigs, msgs, pks)
    got = np.zeros(n, dtype=np.uint8)
    lib.hostsim_ed25519_verify(sigs.tobytes(), msgs.tobytes(), pks.tobytes(), n, got.ctypes.data)
    assert not (got != exp).any()
    assert 0.05 < exp.mean() < 0.4


def test_lanegroup_units(sim_units):
    """group.cuh through a thread-per-role emulation (pthread barrier = the block barrier, shared arrays = shared memory)."""
    uc.check_group_curves(sim_units, n=10)


def test_lanegroup_ecrecover_and_verify_logic():
    """The lane-group programs end to end on the whole edge corpus and on fuzz rows, against the oracle."""
    import coracle
    import fuzz_cases

    lib = load_hostsim()
    for cid, c in ((0, o.K1), (1, o.R1)):
        cases = uc.ecdsa_cases(c, nvalid=12)
        n = len(cases)
        out = np.zeros((n, 64), dtype=np.uint8)
        st = np.zeros(n, dtype=np.uint8)
        lib.hostsim_group_ecrecover(cid, b"".join(x[1] for x in cases), b"".join(x[2] for x in cases), n, out.ctypes.data, st.ctypes.data)
        uc.check_ecrecover_against_oracle(c, cases, out, st)
        sigs, msgs = fuzz_cases.ecdsa_batch(cid, 600, seed=9)
        exp_out, exp_st = coracle.ecrecover(cid, sigs, msgs)
        out = np.zeros((600, 64), dtype=np.uint8)
        st = np.zeros(600, dtype=np.uint8)
        lib.hostsim_group_ecrecover(cid, sigs.tobytes(), msgs.tobytes(), 600, out.ctypes.data, st.ctypes.data)
        assert not (st != exp_st).any() and not (out != exp_out).any()
    cases = uc.ed_cases(nvalid=12)
    n = len(cases)
    got = np.zeros(n, dtype=np.uint8)
    lib.hostsim_group_ed25519_verify(b"".join(x[1] for x in cases), b"".join(x[2] for x in cases), b"".join(x[3] for x in cases), n, got.ctypes.data)
    uc.check_ed_against_oracle(cases, got)
    sigs, msgs, pks = fuzz_cases.ed25519_batch(600, seed=9)
    exp = coracle.ecverify_ed25519(sigs, msgs, pks)
    got = np.zeros(600, dtype=np.uint8)
    lib.hostsim_group_ed25519_verify(sigs.tobytes(), msgs.tobytes(), pks.tobytes(), 600, got.ctypes.data)
    assert not (got != exp).any()


def test_positional_fixed_base_tables(sim_units):
    """The fixed-base halves (u1*G, s*B) through positional tables of several window widths: the digit arithmetic (offset
    recoding, carries into the top window, the most negative digit) is width-dependent, the kernels take the width at run
    time, and the GPU library runs a width the CPU cannot generate in reasonable time."""
    lib = load_hostsim()
    try:
        for w in (4, 5, 9, 12):
            assert lib.hostsim_set_gwin(w) == 0
            uc.check_fixed_base(sim_units, w)
    finally:
        assert lib.hostsim_set_gwin(7) == 0
    uc.check_fixed_base(sim_units, 7)
