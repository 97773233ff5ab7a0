"""Build libsigops.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo snapshot).

One translation unit per kernel family (csrc/kern_*.cu) plus the host layer (csrc/sigops.cu), compiled in parallel and
linked into one shared library.  `-Xptxas -v` output of every unit is kept in build.log (registers / spills per kernel).
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsigops.so")
OBJDIR = os.path.join(HERE, "build")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]


def units():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def headers():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cuh", ".h"))] + [
        os.path.join(os.path.dirname(HERE), "include", "sigops.h")
    ]


def sources():
    return [os.path.join(CSRC, f) for f in units()] + headers()


def source_hash() -> str:
    """Hash of every source the library is built from; profiles/ captures are tagged with it (bench.py refuses
    `executed_frac` from a capture of different sources)."""
    h = hashlib.sha256()
    for s in sources():
        h.update(os.path.basename(s).encode())
        h.update(open(s, "rb").read())
    return h.hexdigest()[:16]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in sources())


def build(force: bool = False, verbose: bool = False, out: str = None, extra=(), jobs: int = 0) -> str:
    """Default: the product library.  `out` / `extra` build an experimental variant (tools/variants.py) elsewhere."""
    if out is None and not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    target = out or LIB
    objdir = OBJDIR if out is None else out + ".obj"
    os.makedirs(objdir, exist_ok=True)
    hdr_t = max(os.path.getmtime(h) for h in headers())
    log = (out + ".log") if out else os.path.join(HERE, "build.log")

    def compile_one(u):
        src = os.path.join(CSRC, u)
        obj = os.path.join(objdir, u[:-3] + ".o")
        ulog = obj + ".log"
        if (not force and out is None and os.path.exists(obj) and os.path.exists(ulog)
                and os.path.getmtime(obj) > max(hdr_t, os.path.getmtime(src))):
            return 0, ulog
        cmd = [nvcc] + NVCC_FLAGS + list(extra) + ["-c", "-o", obj, src]
        with open(ulog, "w") as f:
            f.write("$ " + " ".join(cmd) + "\n")
            f.flush()
            rc = subprocess.call(cmd, stdout=f, stderr=subprocess.STDOUT)
        return rc, ulog

    us = units()
    with ThreadPoolExecutor(max_workers=jobs or min(len(us), os.cpu_count() or 4)) as ex:
        results = list(ex.map(compile_one, us))
    with open(log, "w") as f:
        for (rc, ulog), u in zip(results, us):
            f.write("==== %s (rc %d)\n" % (u, rc))
            f.write(open(ulog).read())
    if any(rc for rc, _ in results):
        sys.stderr.write(open(log).read()[-6000:])
        raise RuntimeError("nvcc failed building libsigops.so (see %s)" % log)
    objs = [os.path.join(objdir, u[:-3] + ".o") for u in us]
    cmd = [nvcc, "-shared", "-o", target] + objs + ["-lcudart", "-lpthread"]
    with open(log, "a") as f:
        f.write("$ " + " ".join(cmd) + "\n")
        f.flush()
        rc = subprocess.call(cmd, stdout=f, stderr=subprocess.STDOUT)
    if rc != 0:
        sys.stderr.write(open(log).read()[-4000:])
        raise RuntimeError("link failed building libsigops.so (see %s)" % log)
    if out is None:
        with open(os.path.join(HERE, "libsigops.srchash"), "w") as f:
            f.write(source_hash() + "\n")
    if verbose:
        sys.stdout.write(open(log).read())
    return target


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="--quiet" not in sys.argv)
    print(LIB)
