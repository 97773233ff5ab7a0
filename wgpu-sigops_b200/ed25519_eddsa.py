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
