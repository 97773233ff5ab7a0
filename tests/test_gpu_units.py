"""GPU unit tests: every device function through the `sigops_test_unit` shim kernel (C ABI), against Python integers /
the oracle.  Mirrors the reference's single-invocation shader tests (src/wgsl/tests/*.wgsl driven by src/tests/*.rs)."""
import pytest

import sigops_oracle as o
import unit_checks as uc

pytestmark = pytest.mark.gpu


def test_field_k1(gpu_units):
    uc.check_field(gpu_units, "K1", o.K1.p, True, nrand=20000)


def test_field_ed25519(gpu_units):
    uc.check_field(gpu_units, "ED", o.ED_P, True, nrand=20000)


def test_field_r1(gpu_units):
    uc.check_field(gpu_units, "R1", o.R1.p, True, nrand=20000)


def test_raw_fixup_paths(gpu_units):
    uc.check_raw_fixups(gpu_units, nrand=4000)


def test_wide_products(gpu_units):
    uc.check_wide(gpu_units, nrand=20000)


def test_addition_chains(gpu_units):
    uc.check_chains(gpu_units, n=300)


def test_scalar_fields(gpu_units):
    uc.check_scalar(gpu_units, n=5000, ninv=300)


def test_sha512(gpu_units):
    uc.check_sha512(gpu_units, n=500)


def test_glv(gpu_units):
    uc.check_glv(gpu_units, n=20000)


def test_curves(gpu_units):
    uc.check_curves(gpu_units, n=40)


def test_positional_fixed_base_tables(gpu_units):
    """u1*G and s*B through the device-generated positional tables (csrc/ptab.h) at the library's window width: edge
    scalars of that width (every window at an extreme digit, bits on both sides of every window boundary) and random ones."""
    import os

    w = int(os.environ.get("SIGOPS_GWIN", "20"))
    uc.check_fixed_base(gpu_units, w)
    if w != 20:
        uc.check_fixed_base(gpu_units, 20)  # scalars built for another width are ordinary inputs: one more sample
