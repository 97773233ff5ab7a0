"""secp256r1 (P-256) ECDSA public-key recovery -- mirror of src/secp256r1_ecdsa.rs."""
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _batch


def ecrecover(signatures, messages, table_limbs: Optional[Sequence[int]], log_limb_size: int) -> List[bytes]:
    """`secp256r1_ecdsa::ecrecover` (src/secp256r1_ecdsa.rs:62-67)."""
    _batch.check_compat_args(table_limbs, log_limb_size, 640)
    out, _ = _batch.ecrecover("sigops_secp256r1_ecrecover", signatures, messages)
    return [bytes(r) for r in out]


def ecrecover_single_shader(signatures, messages, log_limb_size: int) -> List[bytes]:
    """`secp256r1_ecdsa::ecrecover_single_shader` (src/secp256r1_ecdsa.rs:216-220)."""
    return ecrecover(signatures, messages, None, log_limb_size)


def ecrecover_with_status(signatures, messages) -> Tuple[np.ndarray, np.ndarray]:
    return _batch.ecrecover("sigops_secp256r1_ecrecover", signatures, messages)
