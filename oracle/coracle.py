"""ctypes wrapper of oracle/sigops_oracle.c (TEST INFRASTRUCTURE ONLY -- see that file's header).

Imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs.  The library
is (re)built with `-march=native` on the machine it runs on (a stamp records the CPU model), because the build
container and the GPU box may have different host CPUs.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "liboracle.so")
STAMP = SO + ".cpu"
_lib = None


def _cpu_model() -> str:
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def build(force: bool = False) -> str:
    srcs = [os.path.join(HERE, f) for f in ("sigops_oracle.c", "k1_fast.c", "openssl_ref.c", "Makefile")]
    cpu = _cpu_model()
    fresh = (
        os.path.exists(SO)
        and os.path.exists(STAMP)
        and open(STAMP).read() == cpu
        and all(os.path.getmtime(SO) >= os.path.getmtime(f) for f in srcs)
    )
    if force or not fresh:
        subprocess.check_call(["make", "-s", "-B", "-C", HERE])
        with open(STAMP, "w") as f:
            f.write(cpu)
    return SO


def load() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(build())
        vp, sz, i32, u64 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_uint64
        lib.oracle_ecrecover.argtypes = [i32, vp, vp, sz, vp, vp, i32]
        lib.oracle_ed25519_verify.argtypes = [vp, vp, vp, sz, vp, i32]
        lib.oracle_gen_ecdsa.argtypes = [i32, u64, sz, i32, vp, vp, vp, i32]
        lib.oracle_gen_ed25519.argtypes = [u64, sz, vp, vp, vp, i32]
        lib.oracle_sha512.argtypes = [vp, sz, vp]
        lib.oracle_ed25519_verify_msgs.argtypes = [vp, vp, vp, vp, sz, i32, vp]
        lib.oracle_ed25519_sign.argtypes = [vp, vp, sz, vp, vp]
        lib.oracle_k1_ecrecover_fast.argtypes = [vp, vp, sz, vp, vp, i32]
        _lib = lib
    return _lib


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def _u8(a, width):
    if isinstance(a, (bytes, bytearray)):
        a = np.frombuffer(bytes(a), dtype=np.uint8)
    elif not isinstance(a, np.ndarray):
        a = np.frombuffer(b"".join(bytes(x) for x in a), dtype=np.uint8)
    return np.ascontiguousarray(a, dtype=np.uint8).reshape(-1, width)


def ecrecover(curve: int, sigs, msgs, threads: int = 0):
    """curve 0 = secp256k1, 1 = secp256r1.  Returns (n x 64 uint8 X||Y, n uint8 status)."""
    sigs, msgs = _u8(sigs, 64), _u8(msgs, 32)
    n = sigs.shape[0]
    assert msgs.shape[0] == n
    out = np.zeros((n, 64), dtype=np.uint8)
    st = np.zeros(n, dtype=np.uint8)
    if n:
        rc = load().oracle_ecrecover(curve, sigs.ctypes.data, msgs.ctypes.data, n, out.ctypes.data, st.ctypes.data,
                                     threads or host_threads())
        assert rc == 0
    return out, st


def k1_ecrecover_fast(sigs, msgs, threads: int = 0):
    """secp256k1 recovery by oracle/k1_fast.c (special-prime field, GLV, wNAF: the algorithm class of libsecp256k1) -- the
    timed CPU baseline of bench.py; pinned to `ecrecover(0, ...)` by tests/test_oracle.py."""
    sigs, msgs = _u8(sigs, 64), _u8(msgs, 32)
    n = sigs.shape[0]
    assert msgs.shape[0] == n
    out = np.zeros((n, 64), dtype=np.uint8)
    st = np.zeros(n, dtype=np.uint8)
    if n:
        rc = load().oracle_k1_ecrecover_fast(sigs.ctypes.data, msgs.ctypes.data, n, out.ctypes.data, st.ctypes.data,
                                             threads or host_threads())
        assert rc == 0
    return out, st


_ossl = None


def openssl_ref():
    """libosslref.so (oracle/openssl_ref.c: OpenSSL 3 verify legs) or None when libcrypto could not be linked."""
    global _ossl
    if _ossl is None:
        build()
        path = os.path.join(HERE, "_build", "libosslref.so")
        if not os.path.exists(path):
            _ossl = False
        else:
            try:
                lib = ctypes.CDLL(path)
                vp, sz, i32 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int
                lib.openssl_ecdsa_verify.argtypes = [i32, vp, vp, vp, sz, vp, i32]
                lib.openssl_ed25519_verify.argtypes = [vp, vp, vp, sz, vp, i32]
                _ossl = lib
            except OSError:
                _ossl = False
    return _ossl or None


def openssl_ecdsa_verify(curve: int, sigs, msgs, pubkeys, threads: int = 0):
    """OpenSSL ECDSA verify of Fuel-encoded signatures against the given 64-byte keys; returns uint8 verdicts."""
    lib = openssl_ref()
    sigs, msgs, pks = _u8(sigs, 64), _u8(msgs, 32), _u8(pubkeys, 64)
    ok = np.zeros(sigs.shape[0], dtype=np.uint8)
    if sigs.shape[0]:
        lib.openssl_ecdsa_verify(curve, sigs.ctypes.data, msgs.ctypes.data, pks.ctypes.data, sigs.shape[0], ok.ctypes.data,
                                 threads or host_threads())
    return ok


def openssl_ed25519_verify(sigs, msgs, pks, threads: int = 0):
    lib = openssl_ref()
    sigs, msgs, pks = _u8(sigs, 64), _u8(msgs, 32), _u8(pks, 32)
    ok = np.zeros(sigs.shape[0], dtype=np.uint8)
    if sigs.shape[0]:
        lib.openssl_ed25519_verify(sigs.ctypes.data, msgs.ctypes.data, pks.ctypes.data, sigs.shape[0], ok.ctypes.data,
                                   threads or host_threads())
    return ok


def ecverify_ed25519(sigs, msgs, pks, threads: int = 0):
    sigs, msgs, pks = _u8(sigs, 64), _u8(msgs, 32), _u8(pks, 32)
    n = sigs.shape[0]
    assert msgs.shape[0] == n == pks.shape[0]
    out = np.zeros(n, dtype=np.uint8)
    if n:
        rc = load().oracle_ed25519_verify(sigs.ctypes.data, msgs.ctypes.data, pks.ctypes.data, n, out.ctypes.data,
                                          threads or host_threads())
        assert rc == 0
    return out


def gen_ecdsa(curve: int, n: int, seed: int = 0x51600002, low_s: bool = True, threads: int = 0):
    """n valid Fuel-encoded signatures: (sigs n x 64, msgs n x 32, signer public keys n x 64)."""
    sigs = np.zeros((n, 64), dtype=np.uint8)
    msgs = np.zeros((n, 32), dtype=np.uint8)
    pks = np.zeros((n, 64), dtype=np.uint8)
    if n:
        rc = load().oracle_gen_ecdsa(curve, seed, n, int(low_s), sigs.ctypes.data, msgs.ctypes.data, pks.ctypes.data,
                                     threads or host_threads())
        assert rc == 0
    return sigs, msgs, pks


def gen_ed25519(n: int, seed: int = 0x51600002, threads: int = 0):
    """n valid RFC 8032 signatures over 32-byte messages: (sigs n x 64, msgs n x 32, pks n x 32)."""
    sigs = np.zeros((n, 64), dtype=np.uint8)
    msgs = np.zeros((n, 32), dtype=np.uint8)
    pks = np.zeros((n, 32), dtype=np.uint8)
    if n:
        rc = load().oracle_gen_ed25519(seed, n, sigs.ctypes.data, msgs.ctypes.data, pks.ctypes.data,
                                       threads or host_threads())
        assert rc == 0
    return sigs, msgs, pks


def sha512(msg: bytes) -> bytes:
    out = np.zeros(64, dtype=np.uint8)
    buf = np.frombuffer(msg, dtype=np.uint8) if msg else np.zeros(1, dtype=np.uint8)
    assert load().oracle_sha512(buf.ctypes.data, len(msg), out.ctypes.data) == 0
    return out.tobytes()


def _blob(messages):
    offs = np.zeros(len(messages) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(m) for m in messages], dtype=np.uint64)
    blob = np.frombuffer(b"".join(bytes(m) for m in messages) or b"\0", dtype=np.uint8)
    return blob, offs


def ecverify_ed25519_msgs(sigs, messages, pks, strict: bool = False):
    """dalek `verify` / `verify_strict` over messages of any length (list of bytes)."""
    sigs, pks = _u8(sigs, 64), _u8(pks, 32)
    n = sigs.shape[0]
    assert n == pks.shape[0] == len(messages)
    out = np.zeros(n, dtype=np.uint8)
    if n:
        blob, offs = _blob(messages)
        rc = load().oracle_ed25519_verify_msgs(sigs.ctypes.data, blob.ctypes.data, offs.ctypes.data, pks.ctypes.data, n,
                                               int(strict), out.ctypes.data)
        assert rc == 0
    return out


def ed25519_sign(seed: bytes, msg: bytes):
    """RFC 8032 signature of an arbitrary-length message: (sig 64 B, pk 32 B)."""
    sig = np.zeros(64, dtype=np.uint8)
    pk = np.zeros(32, dtype=np.uint8)
    s = np.frombuffer(seed, dtype=np.uint8)
    m = np.frombuffer(msg, dtype=np.uint8) if msg else np.zeros(1, dtype=np.uint8)
    assert load().oracle_ed25519_sign(s.ctypes.data, m.ctypes.data, len(msg), sig.ctypes.data, pk.ctypes.data) == 0
    return sig.tobytes(), pk.tobytes()
