"""Round-2 groundwork, CPU only: the lane-group field product prototype (tools/proto/lanegroup_mul.cpp, four emulated
lanes per 256-bit product; not part of libsigops) against its own schoolbook twin and against Python integers."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P_K1 = 2**256 - 2**32 - 977
P_38 = 2**256 - 38


def test_lanegroup_product_prototype(tmp_path):
    exe = str(tmp_path / "lanegroup_mul")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tools", "proto", "lanegroup_mul.cpp")])
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    assert " 0 mismatches" in out, out
    lines = subprocess.run([exe, "--dump", "2000"], capture_output=True, text=True, check=True).stdout.split("\n")
    rows = [ln.split() for ln in lines if ln]
    assert len(rows) == 2000
    for a, b, r1, r2 in rows:
        a, b = int(a, 16), int(b, 16)
        assert int(r1, 16) == a * b % P_K1
        assert int(r2, 16) == a * b % P_38
