"""secp256k1 ECDSA public-key recovery -- mirror of src/secp256k1_ecdsa.rs."""
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _batch


def ecrecover(signatures, messages, table_limbs: Optional[Sequence[int]], log_limb_size: int) -> List[bytes]:
    """`secp256k1_ecdsa::ecrecover` (src/secp256k1_ecdsa.rs:61-66): one 64-byte X||Y per signature; 64 zero bytes for
    a signature the CPU library would reject (the reference has no per-signature error channel)."""
    _batch.check_compat_args(table_limbs, log_limb_size, 640)
    out, _ = _batch.ecrecover("sigops_secp256k1_ecrecover", signatures, messages)
    return [bytes(r) for r in out]


def ecrecover_single_shader(signatures, messages, log_limb_size: int) -> List[bytes]:
    """`secp256k1_ecdsa::ecrecover_single_shader` (src/secp256k1_ecdsa.rs:215-219): same engine, same result."""
    return ecrecover(signatures, messages, None, log_limb_size)


def ecrecover_with_status(signatures, messages) -> Tuple[np.ndarray, np.ndarray]:
    """Extension: (n x 64 uint8 public keys, n uint8 status) with status 0 = recovered, 1 = invalid signature."""
    return _batch.ecrecover("sigops_secp256k1_ecrecover", signatures, messages)
