"""A COMPILED consumer of include/sigops.h: tests/c_abi/abi_consumer.c is built with `gcc -std=c99 -Wall -Wextra -Werror
-pedantic` against the header, linked to libsigops.so and run as its own process -- the stand-in for the Rust shim crate
(rust/src/*.rs), which cannot be compiled in this image.  It makes the calls the shim makes (src/secp256k1_ecdsa.rs:61-66,
203-212; src/ed25519_eddsa.rs:67-73; src/precompute.rs:36-69) with malloc'ed buffers, out_status = NULL, n = 0, 1 and
100,003, the capacity protocol of sigops_precompute_bases, and compares with the golden fixtures of tests/golden/."""
import json
import os
import struct
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
LIBDIR = os.path.join(ROOT, "wgpu-sigops_b200")


def _has_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="module")
def consumer(tmp_path_factory, sigops):
    """(binary, fixture_dir): the consumer compiled against the header, the goldens flattened to binary."""
    d = tmp_path_factory.mktemp("c_abi")
    exe = str(d / "abi_consumer")
    subprocess.check_call([
        "gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-O1", "-I", os.path.join(ROOT, "include"),
        os.path.join(ROOT, "tests", "c_abi", "abi_consumer.c"), "-o", exe, "-L", LIBDIR, "-lsigops", "-Wl,-rpath," + LIBDIR])
    for name in ("secp256k1", "secp256r1"):
        cases = json.load(open(os.path.join(GOLD, name + "_ecrecover.json")))["cases"]
        assert cases[0]["pubkey"], "row 0 must be a valid signature (the consumer's n = 1 case)"
        with open(d / (name + ".bin"), "wb") as f:
            f.write(struct.pack("<I", len(cases)))
            for c in cases:
                pk = bytes.fromhex(c["pubkey"]) if c["pubkey"] else bytes(64)
                f.write(bytes.fromhex(c["sig"]) + bytes.fromhex(c["msg"]) + pk + bytes([0 if c["pubkey"] else 1]))
    cases = json.load(open(os.path.join(GOLD, "ed25519_ecverify.json")))["cases"]
    with open(d / "ed25519.bin", "wb") as f:
        f.write(struct.pack("<I", len(cases)))
        for c in cases:
            f.write(bytes.fromhex(c["sig"]) + bytes.fromhex(c["msg"]) + bytes.fromhex(c["pk"]) + bytes([1 if c["valid"] else 0]))
    bases = json.load(open(os.path.join(GOLD, "precompute_bases_13.json")))
    for k, v in bases.items():
        with open(d / ("bases_%s.bin" % k), "wb") as f:
            f.write(struct.pack("<I", len(v)) + struct.pack("<%dI" % len(v), *v))
    return exe, str(d)


def _run(consumer, mode):
    exe, d = consumer
    p = subprocess.run([exe, d, mode], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "abi_consumer %s: ok" % mode in p.stdout


@pytest.mark.skipif(_has_gpu(), reason="checks the no-device behaviour")
def test_c_consumer_host_only(consumer):
    """Header compiles as strict C99, the library links, the host-only entry points match the goldens and every compute
    call fails loudly without a device."""
    _run(consumer, "host")


@pytest.mark.gpu
def test_c_consumer_on_gpu(consumer):
    _run(consumer, "gpu")
