"""Shared host-side plumbing of the three batch entry points: flatten, validate, call the C ABI, unflatten.

Mirrors `init` in the reference (src/secp256k1_ecdsa.rs:11-59, src/ed25519_eddsa.rs:12-65) minus the zero padding to
a power of two and the workgroup-grid lookup, which the CUDA engine does not need.
"""
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib

MAX_SIGNATURES = 256 * 256 * 256 * 64  # src/secp256k1_ecdsa.rs:22


class LengthMismatch(AssertionError, ValueError):
    """The reference `assert!`s on these (src/secp256k1_ecdsa.rs:21-22, src/ed25519_eddsa.rs:22-24): callers that expect
    the panic catch AssertionError; it is raised explicitly so that `python -O` cannot strip the check."""


def check_compat_args(table_limbs: Optional[Sequence[int]], log_limb_size: int, expected_len_at_13: int) -> None:
    """`table_limbs` / `log_limb_size` are accepted for compatibility and validated, then ignored: the engine's own
    32-bit tables are baked into the library (BASELINE north_star)."""
    if not 11 <= int(log_limb_size) <= 15:  # src/wgsl/mont.wgsl:12,37 supports 11..15 only
        raise ValueError("log_limb_size must be in 11..=15")
    if table_limbs is not None:
        num_limbs = 256 // log_limb_size
        while num_limbs * log_limb_size <= 256:
            num_limbs += 1
        expected = expected_len_at_13 // 20 * num_limbs
        if len(table_limbs) != expected:
            raise ValueError(f"table_limbs has {len(table_limbs)} limbs, expected {expected}")


def flatten(items: Sequence[bytes], width: int, what: str) -> np.ndarray:
    if isinstance(items, np.ndarray):
        # no silent re-rowing or value casts: the array must already be n x width bytes
        if items.dtype != np.uint8 or items.ndim != 2 or items.shape[1] != width:
            raise ValueError(f"{what} array must have dtype uint8 and shape (n, {width}), got {items.dtype} {items.shape}")
        return np.ascontiguousarray(items)
    for it in items:
        if len(it) != width:
            raise ValueError(f"{what} must be {width} bytes")
    return np.frombuffer(b"".join(bytes(i) for i in items), dtype=np.uint8).reshape(-1, width).copy()


def ecrecover(entry: str, signatures, messages) -> Tuple[np.ndarray, np.ndarray]:
    sigs = flatten(signatures, 64, "signature")
    msgs = flatten(messages, 32, "message")
    n = sigs.shape[0]
    # the reference panics here (assert!, src/secp256k1_ecdsa.rs:21-22); an exception that survives `python -O`
    if n != msgs.shape[0]:
        raise LengthMismatch("signatures and messages differ in length")
    if n > MAX_SIGNATURES:
        raise LengthMismatch("more than 2^30 signatures")
    out = np.zeros((n, 64), dtype=np.uint8)
    status = np.zeros(n, dtype=np.uint8)
    if n == 0:
        return out, status  # src/secp256k1_ecdsa.rs:71-73
    lib = _lib.load()
    _lib.check(getattr(lib, entry)(sigs.ctypes.data, msgs.ctypes.data, n, out.ctypes.data, status.ctypes.data))
    return out, status


def ecverify(signatures, messages, verifying_keys) -> np.ndarray:
    sigs = flatten(signatures, 64, "signature")
    msgs = flatten(messages, 32, "message")
    pks = flatten(verifying_keys, 32, "verifying key")
    n = sigs.shape[0]
    if not (n == msgs.shape[0] == pks.shape[0]):
        raise LengthMismatch("signatures, messages and verifying keys differ in length")
    if n > MAX_SIGNATURES:
        raise LengthMismatch("more than 2^30 signatures")
    out = np.zeros(n, dtype=np.uint8)
    if n == 0:
        return out
    lib = _lib.load()
    _lib.check(lib.sigops_ed25519_ecverify(sigs.ctypes.data, msgs.ctypes.data, pks.ctypes.data, n, out.ctypes.data))
    return out
