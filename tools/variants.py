#!/usr/bin/env python3
"""Build experimental variants of libsigops.so (different -D / -maxrregcount) side by side and, on a GPU box, time
each with tools/quick_bench-style runs.  Development tool.
   python tools/variants.py build name=flag,flag ...     (here, no GPU)
   python tools/variants.py run [n]                       (under gpurun)"""
import importlib.util
import json
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VDIR = os.path.join(ROOT, "wgpu-sigops_b200", "variants")


def build(specs):
    spec = importlib.util.spec_from_file_location("_b", os.path.join(ROOT, "wgpu-sigops_b200", "build.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    os.makedirs(VDIR, exist_ok=True)

    def one(s):
        name, _, flags = s.partition("=")
        out = os.path.join(VDIR, name + ".so")
        b.build(out=out, extra=[f for f in flags.split(",") if f])
        regs = subprocess.run("grep -A2 -E 'Compiling entry function .*(ecrecover|ed25519_verify)' %s.log | grep -E 'Used|spill'" % out,
                              shell=True, capture_output=True, text=True).stdout
        return name, regs

    with ThreadPoolExecutor(max_workers=4) as ex:
        for name, regs in ex.map(one, specs):
            print("==", name)
            print(regs)


def run(n):
    res = {}
    for f in sorted(os.listdir(VDIR)):
        if not f.endswith(".so"):
            continue
        env = dict(os.environ, SIGOPS_LIB=os.path.join(VDIR, f), SIGOPS_MAX_CHUNKS=os.environ.get("SIGOPS_MAX_CHUNKS", "1"))
        p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "prof_run.py"), str(n), "3", "time"], env=env,
                           capture_output=True, text=True, timeout=600)
        print("==", f, p.stdout.strip().replace("\n", " | "), p.stderr[-300:] if p.returncode else "", flush=True)
        res[f] = p.stdout
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "variants.json"), "w"), indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build(sys.argv[2:])
    else:
        run(int(sys.argv[2]) if len(sys.argv) > 2 else 262144)
