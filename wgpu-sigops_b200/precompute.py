"""Fixed-base tables in the reference's limb format -- mirror of src/precompute.rs (CPU only, compatibility)."""
import ctypes
from typing import List

import numpy as np

from . import _lib

WINDOW_SIZE = 4  # src/precompute.rs:12


def _bases(curve: int, log_limb_size: int) -> List[int]:
    lib = _lib.load()
    n = ctypes.c_size_t(0)
    _lib.check(lib.sigops_precompute_bases(curve, log_limb_size, None, ctypes.byref(n)))
    out = np.zeros(n.value, dtype=np.uint32)
    _lib.check(lib.sigops_precompute_bases(curve, log_limb_size, out.ctypes.data, ctypes.byref(n)))
    return [int(x) for x in out]


def secp256k1_bases(log_limb_size: int) -> List[int]:
    """`precompute::secp256k1_bases` (src/precompute.rs:36-43)."""
    return _bases(0, log_limb_size)


def secp256r1_bases(log_limb_size: int) -> List[int]:
    """`precompute::secp256r1_bases` (src/precompute.rs:45-52)."""
    return _bases(1, log_limb_size)


def ed25519_bases(log_limb_size: int) -> List[int]:
    """`precompute::ed25519_bases` (src/precompute.rs:54-69)."""
    return _bases(2, log_limb_size)
