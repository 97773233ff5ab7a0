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
