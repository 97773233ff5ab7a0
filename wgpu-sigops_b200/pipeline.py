"""The steps on either side of recovery, on the device (not part of the reference's API).

`sha256_batch` is `fuel_crypto::Message::new` for a whole block (the reference's callers do it on the host before every
`ecrecover`, src/tests/secp256k1_ecdsa.rs:21-22); `ecrecover_addresses` chains SHA-256(message) -> recover ->
SHA-256(X || Y) (the Fuel address of the signer) without a host pass in between."""
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _batch, _lib


def _blob(messages: Sequence[bytes]):
    offs = np.zeros(len(messages) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(m) for m in messages], dtype=np.uint64)
    blob = np.frombuffer(b"".join(bytes(m) for m in messages) or b"\0", dtype=np.uint8)
    return blob, offs


def sha256_batch(messages: Sequence[bytes]) -> np.ndarray:
    """n x 32 uint8: SHA-256 of each message."""
    n = len(messages)
    out = np.zeros((n, 32), dtype=np.uint8)
    if n:
        blob, offs = _blob(messages)
        _lib.check(_lib.load().sigops_sha256_batch(blob.ctypes.data, offs.ctypes.data, n, out.ctypes.data))
    return out


def ecrecover_addresses(curve: str, signatures, messages, prehashed: bool = False) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """(addresses n x 32, public keys n x 64, status n).  `messages`: raw byte strings (hashed on the device), or
    32-byte prehashes when `prehashed`."""
    cid = {"secp256k1": 0, "secp256r1": 1}[curve]
    sigs = _batch.flatten(signatures, 64, "signature")
    n = sigs.shape[0]
    addr = np.zeros((n, 32), dtype=np.uint8)
    pks = np.zeros((n, 64), dtype=np.uint8)
    st = np.zeros(n, dtype=np.uint8)
    if n == 0:
        return addr, pks, st
    lib = _lib.load()
    if prehashed:
        msgs = _batch.flatten(messages, 32, "message")
        if msgs.shape[0] != n:
            raise _batch.LengthMismatch("signatures and messages differ in length")
        _lib.check(lib.sigops_ecrecover_addresses(cid, sigs.ctypes.data, msgs.ctypes.data, None, n, addr.ctypes.data,
                                                  pks.ctypes.data, st.ctypes.data))
    else:
        if len(messages) != n:
            raise _batch.LengthMismatch("signatures and messages differ in length")
        blob, offs = _blob(messages)
        _lib.check(lib.sigops_ecrecover_addresses(cid, sigs.ctypes.data, blob.ctypes.data, offs.ctypes.data, n,
                                                  addr.ctypes.data, pks.ctypes.data, st.ctypes.data))
    return addr, pks, st
