"""ctypes loader for libsigops.so (the C ABI in include/sigops.h).

There is no fallback: if the library is missing or cannot be loaded the import fails loudly, and every compute
entry point raises ShaderFailureError when the C call returns nonzero (e.g. no CUDA device).
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# SIGOPS_LIB selects an experimental build of the same library (tools/variants.py); default = the product
LIB_PATH = os.environ.get("SIGOPS_LIB") or os.path.join(HERE, "libsigops.so")

# every symbol include/sigops.h declares
SYMBOLS = [
    "sigops_init", "sigops_shutdown", "sigops_num_devices", "sigops_last_error",
    "sigops_secp256k1_ecrecover", "sigops_secp256r1_ecrecover", "sigops_ed25519_ecverify",
    "sigops_precompute_bases", "sigops_last_timing", "sigops_kernel_launches",
    "sigops_host_alloc", "sigops_host_free",
    "sigops_secp256k1_ecrecover_device", "sigops_secp256r1_ecrecover_device", "sigops_ed25519_ecverify_device",
    "sigops_test_unit", "sigops_test_unit_shape", "sigops_imad_peak", "sigops_plan_shards",
    "sigops_ed25519_ecverify_msgs", "sigops_sha256_batch", "sigops_ecrecover_addresses",
    "sigops_queue_create", "sigops_queue_destroy", "sigops_queue_buffers", "sigops_queue_submit", "sigops_queue_poll",
    "sigops_queue_wait", "sigops_queue_info", "sigops_queue_submit_device", "sigops_queue_slot_device",
    "sigops_batch_on_devices",
]

_lib = None


class ShaderFailureError(Exception):
    """Mirror of the reference's `ShaderFailureError` (src/lib.rs:12-14): the device path did not complete."""


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback."
        )
    # many small requests in flight (service.SigQueue) need more than the default 8 hardware work queues; must be set
    # before the process creates its CUDA context (the library sets the same default in sigops_init)
    os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
    lib = ctypes.CDLL(LIB_PATH)
    c = ctypes
    vp, sz, i32, u32p = c.c_void_p, c.c_size_t, c.c_int, c.POINTER(c.c_uint32)
    lib.sigops_init.argtypes = [c.POINTER(c.c_int), i32]
    lib.sigops_last_error.restype = c.c_char_p
    lib.sigops_secp256k1_ecrecover.argtypes = [vp, vp, sz, vp, vp]
    lib.sigops_secp256r1_ecrecover.argtypes = [vp, vp, sz, vp, vp]
    lib.sigops_ed25519_ecverify.argtypes = [vp, vp, vp, sz, vp]
    lib.sigops_ed25519_ecverify_msgs.argtypes = [vp, vp, vp, vp, sz, c.c_uint32, vp]
    lib.sigops_sha256_batch.argtypes = [vp, vp, sz, vp]
    lib.sigops_ecrecover_addresses.argtypes = [i32, vp, vp, vp, sz, vp, vp, vp]
    lib.sigops_precompute_bases.argtypes = [i32, c.c_uint32, vp, c.POINTER(sz)]
    lib.sigops_last_timing.argtypes = [c.POINTER(c.c_double)] * 3
    lib.sigops_kernel_launches.restype = c.c_uint64
    lib.sigops_host_alloc.argtypes = [sz]
    lib.sigops_host_alloc.restype = vp
    lib.sigops_host_free.argtypes = [vp]
    lib.sigops_host_free.restype = None
    lib.sigops_secp256k1_ecrecover_device.argtypes = [vp, vp, sz, vp, vp, vp]
    lib.sigops_secp256r1_ecrecover_device.argtypes = [vp, vp, sz, vp, vp, vp]
    lib.sigops_ed25519_ecverify_device.argtypes = [vp, vp, vp, sz, vp, vp]
    lib.sigops_test_unit.argtypes = [i32, vp, sz, vp]
    lib.sigops_test_unit_shape.argtypes = [i32, c.POINTER(i32), c.POINTER(i32)]
    lib.sigops_plan_shards.argtypes = [sz, i32, c.POINTER(sz), c.POINTER(i32)]
    lib.sigops_imad_peak.argtypes = [i32, i32, c.POINTER(c.c_double), c.POINTER(c.c_double)]
    pvp = c.POINTER(vp)
    lib.sigops_queue_create.argtypes = [i32, i32, sz, i32, pvp]
    lib.sigops_queue_destroy.argtypes = [vp]
    lib.sigops_queue_buffers.argtypes = [vp, i32, pvp, pvp, pvp, pvp, pvp]
    lib.sigops_queue_submit.argtypes = [vp, i32, sz]
    lib.sigops_queue_poll.argtypes = [vp, i32, c.POINTER(i32)]
    lib.sigops_queue_wait.argtypes = [vp, i32, c.POINTER(sz), c.POINTER(c.c_double)]
    lib.sigops_queue_submit_device.argtypes = [vp, i32, vp, vp, vp, sz, i32, vp]
    lib.sigops_queue_slot_device.argtypes = [vp, i32]
    lib.sigops_batch_on_devices.argtypes = [i32, c.POINTER(i32), i32, vp, vp, vp, sz, vp, vp]
    lib.sigops_queue_info.argtypes = [vp, c.POINTER(i32), c.POINTER(i32), c.POINTER(sz), c.POINTER(i32),
                                      c.POINTER(c.c_uint64), c.POINTER(c.c_uint64)]
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        raise ShaderFailureError(load().sigops_last_error().decode() or f"sigops error {rc}")
