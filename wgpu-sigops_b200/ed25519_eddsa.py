"""ed25519 EdDSA verification -- mirror of src/ed25519_eddsa.rs (dalek `verify` semantics)."""
from typing import List, Optional, Sequence

import numpy as np

from . import _batch


def ecverify(signatures, messages, verifying_keys, table_limbs: Optional[Sequence[int]], log_limb_size: int) -> List[bool]:
    """`ed25519_eddsa::ecverify` (src/ed25519_eddsa.rs:67-73): Vec<bool>."""
    _batch.check_compat_args(table_limbs, log_limb_size, 960)
    return [bool(v) for v in _batch.ecverify(signatures, messages, verifying_keys)]


def ecverify_single(signatures, messages, verifying_keys, log_limb_size: int) -> List[bool]:
    """`ed25519_eddsa::ecverify_single` (src/ed25519_eddsa.rs:259-264)."""
    return ecverify(signatures, messages, verifying_keys, None, log_limb_size)


def ecverify_array(signatures, messages, verifying_keys) -> np.ndarray:
    """Extension: uint8 array of 0/1 without the list conversion."""
    return _batch.ecverify(signatures, messages, verifying_keys)


def ecverify_msgs(signatures, messages: Sequence[bytes], verifying_keys, strict: bool = False) -> np.ndarray:
    """Extension: messages of any length; `strict=True` applies ed25519-dalek `verify_strict`
    (what fuel_crypto::ed25519::verify uses): R must decompress, A and R must not have small order."""
    from . import _lib

    sigs = _batch.flatten(signatures, 64, "signature")
    pks = _batch.flatten(verifying_keys, 32, "verifying key")
    n = sigs.shape[0]
    if not (n == pks.shape[0] == len(messages)):
        raise _batch.LengthMismatch("signatures, messages and verifying keys differ in length")
    out = np.zeros(n, dtype=np.uint8)
    if n == 0:
        return out
    offs = np.zeros(n + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(m) for m in messages], dtype=np.uint64)
    blob = np.frombuffer(b"".join(bytes(m) for m in messages) or b"\0", dtype=np.uint8)
    lib = _lib.load()
    _lib.check(lib.sigops_ed25519_ecverify_msgs(sigs.ctypes.data, blob.ctypes.data, offs.ctypes.data, pks.ctypes.data, n,
                                                1 if strict else 0, out.ctypes.data))
    return out


def ecverify_strict(signatures, messages: Sequence[bytes], verifying_keys) -> List[bool]:
    """`fuel_crypto::ed25519::verify` semantics (dalek `verify_strict`), arbitrary-length messages."""
    return [bool(v) for v in ecverify_msgs(signatures, messages, verifying_keys, strict=True)]
