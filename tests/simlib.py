"""Shared helpers for driving the unit-op shims (host simulation and the CUDA library expose the same op table)."""
import ctypes
import os
import random
import re
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

# op ids from include/sigops.h
_hdr = open(os.path.join(ROOT, "include", "sigops.h")).read()
OPS = {m.group(1): int(m.group(2)) for m in re.finditer(r"SIGOPS_UNIT_(\w+) = (\d+)", _hdr)}


def build_hostsim() -> str:
    d = os.path.join(ROOT, "tests", "hostsim")
    so = os.path.join(d, "libhostsim.so")
    srcs = [os.path.join(d, "hostsim.cpp")] + [
        os.path.join(ROOT, "wgpu-sigops_b200", "csrc", f)
        for f in os.listdir(os.path.join(ROOT, "wgpu-sigops_b200", "csrc"))
        if f.endswith(".cuh")
    ]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-pthread", "-o", so, srcs[0]])
    return so


def load_hostsim():
    lib = ctypes.CDLL(build_hostsim())
    lib.hostsim_unit.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
    lib.hostsim_unit_shape.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]
    lib.hostsim_ecrecover.argtypes = [ctypes.c_int, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]
    lib.hostsim_ed25519_verify.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_void_p]
    lib.hostsim_ed25519_verify_msgs.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
    lib.hostsim_sha256.argtypes = [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
    lib.hostsim_group_unit.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
    lib.hostsim_group_ecrecover.argtypes = [ctypes.c_int, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]
    lib.hostsim_group_ed25519_verify.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_void_p]
    lib.hostsim_set_gwin.argtypes = [ctypes.c_int]
    return lib


def to_words(x: int, n: int = 8):
    return [(x >> (32 * i)) & 0xFFFFFFFF for i in range(n)]


def from_words(ws) -> int:
    return sum(int(w) << (32 * i) for i, w in enumerate(ws))


class UnitRunner:
    """run(op_name, items) where each item is a list of ints (each int = one 8-word operand unless widths given)."""

    def __init__(self, unit_fn, shape_fn):
        self.unit_fn, self.shape_fn = unit_fn, shape_fn

    def shape(self, op):
        a, b = ctypes.c_int(), ctypes.c_int()
        self.shape_fn(op, ctypes.byref(a), ctypes.byref(b))
        return a.value, b.value

    def run_words(self, op_name, in_words: np.ndarray) -> np.ndarray:
        op = OPS[op_name]
        in_w, out_w = self.shape(op)
        in_words = np.ascontiguousarray(in_words, dtype=np.uint32).reshape(-1, in_w)
        n = in_words.shape[0]
        out = np.zeros((n, out_w), dtype=np.uint32)
        rc = self.unit_fn(op, in_words.ctypes.data, n, out.ctypes.data)
        assert rc == 0, rc
        return out

    def run(self, op_name, items, widths=None):
        """items: list of tuples of ints; widths: words per operand (default 8 each)."""
        rows = []
        for it in items:
            row = []
            for j, v in enumerate(it):
                row += to_words(v, widths[j] if widths else 8)
            rows.append(row)
        return self.run_words(op_name, np.array(rows, dtype=np.uint32))


def edge_values(p: int):
    """operands that stress the weak-reduction paths: around 0, p, 2^256."""
    vals = [0, 1, 2, p - 2, p - 1, p, p + 1, 2**256 - 1, 2**256 - 2, 2**255, 2**255 - 1, 2**255 - 19, 2**255 - 18,
            2**256 - 38, 2**256 - 39, 2**224, 2**192, 2**96, 2**32, 2**32 - 1, (1 << 256) - (1 << 32) - 977 + 5]
    return [v for v in vals if 0 <= v < 2**256]


def rand256(rng: random.Random) -> int:
    r = rng.random()
    if r < 0.1:
        return rng.getrandbits(256) | (((1 << 200) - 1) << 56)  # many ones in the top
    if r < 0.2:
        return rng.getrandbits(64)
    if r < 0.3:
        return rng.getrandbits(256) & ~(((1 << 128) - 1) << 64)
    return rng.getrandbits(256)
