#!/usr/bin/env python3
"""Pipe-rate micro-benchmarks (sigops_imad_peak): prints thread-level operations per second for each kind."""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import wgpu_sigops_b200 as w
lib = w.load()
names = {0: "imad (mad.lo.u32)", 1: "imad_wide (mad.wide.u32, independent)", 2: "imad_wide_x (mad.lo.cc/madc.hi.cc chains)",
         3: "iadd3", 4: "imad_wide + iadd3 1:1 (wide counted)", 5: "dfma (fma.rn.f64)", 6: "imad_hi (mad.hi.u32)"}
res = {}
ops, ms = ctypes.c_double(), ctypes.c_double()
for k, nm in names.items():
    assert lib.sigops_imad_peak(k, 8192, ctypes.byref(ops), ctypes.byref(ms)) == 0, lib.sigops_last_error()
    res[nm] = ops.value
    print("%-50s %.3e /s   %.2f ms" % (nm, ops.value, ms.value), flush=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "peaks.json"), "w"), indent=1)
