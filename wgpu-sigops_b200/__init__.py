"""b200-sigops: Python mirror of the wgpu-sigops public API over libsigops.so (CUDA, sm_100a).

Module and function names follow the reference crate so the parity tests read like its own tests:

    secp256k1_ecdsa.ecrecover / ecrecover_single_shader   (src/secp256k1_ecdsa.rs:61-66,215-219)
    secp256r1_ecdsa.ecrecover / ecrecover_single_shader   (src/secp256r1_ecdsa.rs:62-67,216-220)
    ed25519_eddsa.ecverify / ecverify_single               (src/ed25519_eddsa.rs:67-73,259-264)
    precompute.{secp256k1_bases, secp256r1_bases, ed25519_bases}, WINDOW_SIZE   (src/precompute.rs:12,36-69)
    ShaderFailureError                                      (src/lib.rs:12-14)

The package directory is named `wgpu-sigops_b200`; import it as `wgpu_sigops_b200` (see the loader module of that
name at the repository root).
"""
from . import ed25519_eddsa, pipeline, precompute, secp256k1_ecdsa, secp256r1_ecdsa, service  # noqa: F401
from ._lib import ShaderFailureError, load  # noqa: F401
