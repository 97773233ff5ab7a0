"""CPU-only checks of the drop-in boundary: libsigops.so loads, exports every symbol include/sigops.h declares, refuses to
compute without a CUDA device (no CPU fallback), and the host-only entry points (precompute_bases, plan_shards) match
the reference's conventions."""
import ctypes
import os
import re

import numpy as np
import pytest

import sigops_oracle as o

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def test_every_declared_symbol_is_exported(sigops):
    hdr = open(os.path.join(ROOT, "include", "sigops.h")).read()
    declared = set(re.findall(r"\b(sigops_\w+)\s*\(", hdr))
    assert len(declared) >= 19
    lib = sigops.load()
    for sym in sorted(declared):
        assert getattr(lib, sym) is not None, sym
    from wgpu_sigops_b200 import _lib

    assert declared == set(_lib.SYMBOLS)


def test_header_cites_reference_interfaces():
    hdr = open(os.path.join(ROOT, "include", "sigops.h")).read()
    for cite in ("src/secp256k1_ecdsa.rs:61-66", "src/secp256r1_ecdsa.rs:62-67", "src/ed25519_eddsa.rs:67-73",
                 "src/precompute.rs:12,36-69", "src/lib.rs:12-14", "src/gpu.rs"):
        assert cite in hdr, cite


@pytest.mark.skipif(_has_gpu(), reason="checks the no-device behaviour")
def test_no_cpu_fallback(sigops):
    """Without a CUDA device every compute entry point fails loudly (nonzero rc -> ShaderFailureError)."""
    sig, msg, _ = o.gen_ecdsa_valid(o.K1, 1)
    with pytest.raises(sigops.ShaderFailureError) as e:
        sigops.secp256k1_ecdsa.ecrecover_single_shader([sig], [msg], 13)
    assert "CUDA" in str(e.value)
    with pytest.raises(sigops.ShaderFailureError):
        sigops.secp256r1_ecdsa.ecrecover_single_shader([sig], [msg], 13)
    s, m, pk = o.gen_ed25519_valid(1)
    with pytest.raises(sigops.ShaderFailureError):
        sigops.ed25519_eddsa.ecverify_single([s], [m], [pk], 13)
    lib = sigops.load()
    out = np.zeros(8, dtype=np.uint32)
    assert lib.sigops_test_unit(0, np.zeros(16, dtype=np.uint32).ctypes.data, 1, out.ctypes.data) != 0


def test_empty_batch_and_argument_checks(sigops):
    # n == 0 -> Ok(vec![]) without touching the device (src/secp256k1_ecdsa.rs:71-73)
    assert sigops.secp256k1_ecdsa.ecrecover([], [], None, 13) == []
    assert sigops.secp256r1_ecdsa.ecrecover([], [], None, 13) == []
    assert sigops.ed25519_eddsa.ecverify([], [], [], None, 13) == []
    sig, msg, _ = o.gen_ecdsa_valid(o.K1, 1)
    with pytest.raises(AssertionError):  # assert!(signatures.len() == messages.len()), src/secp256k1_ecdsa.rs:21
        sigops.secp256k1_ecdsa.ecrecover([sig, sig], [msg], None, 13)
    with pytest.raises(ValueError):
        sigops.secp256k1_ecdsa.ecrecover([sig], [msg], None, 16)  # mont_mul supports 11..15 only
    with pytest.raises(ValueError):
        sigops.secp256k1_ecdsa.ecrecover([sig], [msg], [0] * 639, 13)  # table of the wrong length
    with pytest.raises(ValueError):
        sigops.secp256k1_ecdsa.ecrecover([sig[:63]], [msg], None, 13)


@pytest.mark.parametrize("log_limb_size", [11, 12, 13, 14, 15])
def test_precompute_bases_match_oracle(sigops, log_limb_size):
    """precompute::*_bases (src/precompute.rs:36-69): same limbs as the restated table builder."""
    for name, fn in (("secp256k1", sigops.precompute.secp256k1_bases), ("secp256r1", sigops.precompute.secp256r1_bases),
                     ("ed25519", sigops.precompute.ed25519_bases)):
        got = fn(log_limb_size)
        assert got == o.precompute_bases(name, log_limb_size), name
    if log_limb_size == 13:
        assert len(sigops.precompute.secp256k1_bases(13)) == 640 and len(sigops.precompute.ed25519_bases(13)) == 960
    assert sigops.precompute.WINDOW_SIZE == 4


def test_precompute_bases_rejects_bad_limb_size(sigops):
    for bad in (10, 16, 0):
        with pytest.raises(sigops.ShaderFailureError):
            sigops.precompute.secp256k1_bases(bad)


def test_plan_shards(sigops):
    """Contiguous, disjoint, covering; small batches use fewer devices."""
    lib = sigops.load()
    for n in (1, 63, 4096, 4097, 65536, 1 << 20, (1 << 20) + 7, 16777216):
        for g in (1, 2, 4, 8):
            bounds = (ctypes.c_size_t * (g + 1))()
            used = ctypes.c_int()
            assert lib.sigops_plan_shards(n, g, bounds, ctypes.byref(used)) == 0
            u = used.value
            assert 1 <= u <= g and u == min(g, max(1, -(-n // 4096)))
            b = list(bounds)[: u + 1]
            assert b[0] == 0 and b[-1] == n and all(b[i] <= b[i + 1] for i in range(u))
            assert [b[i] for i in range(u + 1)] == [i * n // u for i in range(u + 1)]
            sizes = [b[i + 1] - b[i] for i in range(u)]
            assert max(sizes) - min(sizes) <= 1
