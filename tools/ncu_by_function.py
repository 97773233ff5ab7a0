#!/usr/bin/env python3
"""Where the time goes, by DEVICE FUNCTION: aggregates the cuda,sass source page of an ncu capture (per CUDA source line:
warp-stall samples and executed instructions; inlined code is attributed to the line it came from) by the function that
encloses each source line.
   ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > /tmp/src.csv
   python tools/ncu_by_function.py /tmp/src.csv <signatures per profiled launch>"""
import collections
import csv
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEF = re.compile(r"^\s*(?:template\s*<[^>]*>\s*)?(?:static\s+)?(?:SG_HD|SG_D|SG_CALL|SG_COLD_\w+|__device__|__global__|__host__)[^;(]*?\b(\w+)\s*\(")


def function_map(path):
    """line number -> name of the last function definition that starts at or before it (good enough for flat headers)."""
    names, cur, struct = {}, "?", ""
    try:
        lines = open(path).read().split("\n")
    except OSError:
        return names
    for i, l in enumerate(lines, 1):
        m = re.match(r"^struct (\w+)", l)
        if m:
            struct = m.group(1)
        if l.startswith("}"):
            struct = struct if not l.startswith("};") else ""
        m = DEF.match(l)
        if m and not l.strip().startswith("//"):
            cur = (struct + "::" if struct and l.startswith("    ") else "") + m.group(1)
        names[i] = cur
    return names


def main():
    path, nsig = sys.argv[1], int(sys.argv[2])
    warps = nsig / 32
    fmap, kernel, fpath, hdr = {}, None, None, None
    agg = collections.defaultdict(lambda: collections.defaultdict(lambda: [0, 0]))  # kernel -> func -> [samples, instrs]
    for row in csv.reader(open(path)):
        if not row:
            continue
        if row[0] == "File Path":
            fpath = row[1]
            if fpath not in fmap:
                fmap[fpath] = function_map(fpath)
            continue
        if row[0] == "Function Name":
            kernel = re.sub(r"\(.*", "", row[1]).replace("void sigops::", "").replace("sigops::", "")
            continue
        if row[0] == "Line No":
            hdr = {n: i for i, n in enumerate(row)}
            continue
        if hdr is None or not row[0].isdigit():
            continue
        try:
            samples = int(row[hdr["# Samples"]] or 0)
            instrs = int(row[hdr["Instructions Executed"]] or 0)
        except (ValueError, IndexError):
            continue
        fn = fmap.get(fpath, {}).get(int(row[0]), "?")
        a = agg[kernel][os.path.basename(fpath) + ":" + fn]
        a[0] += samples
        a[1] += instrs
    for k, funcs in agg.items():
        ts = sum(v[0] for v in funcs.values()) or 1
        ti = sum(v[1] for v in funcs.values()) or 1
        print(f"== {k}: {ti / warps:.0f} instructions per signature-warp")
        print("   samples%  instr/sig-warp  instr%   function (file:enclosing definition of the source line)")
        for fn, (s, i) in sorted(funcs.items(), key=lambda x: -x[1][0])[:22]:
            print(f"   {100 * s / ts:7.1f}  {i / warps:14.0f}  {100 * i / ti:6.1f}   {fn}")


if __name__ == "__main__":
    main()
