#!/usr/bin/env python3
"""Dynamic instruction mix and stall attribution per kernel from `ncu --page source --csv --print-source sass` output.
   python tools/ncu_mix.py <source.csv> [n_signatures]"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    nsig = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
    kern, hdr = None, None
    seen = set()
    data = collections.OrderedDict()
    for row in csv.reader(open(path)):
        if not row:
            continue
        if row[0] == "Kernel Name":
            kern = row[1].split("(")[0].replace("void ", "").replace("sigops::", "")
            if kern in seen:
                kern = None  # later launches of the same kernel: skip
            else:
                seen.add(kern)
                data[kern] = []
            continue
        if row[0] == "Address":
            hdr = row
            continue
        if kern is None or hdr is None:
            continue
        data[kern].append(dict(zip(hdr, row)))
    for kern, rows in data.items():
        ops = collections.Counter()
        stall = collections.Counter()
        samples = collections.Counter()
        tot = 0
        for r in rows:
            n = int(r["Instructions Executed"] or 0)
            t = re.sub(r"^@!?U?P\d+\s+", "", r["Source"].strip())
            op = t.split()[0].rstrip(";") if t else "?"
            ops[op] += n
            tot += n
            samples[op] += int(r["# Samples"] or 0)
            for k, v in r.items():
                if k.startswith("stall_") and "Not Issued" not in k and v not in ("", "0"):
                    stall[k] += int(v)
        warps = nsig / 32
        print(f"== {kern}: {tot} warp-instructions = {tot / warps:.0f} per signature-warp; SASS lines {len(rows)}")
        for op, n in ops.most_common(22):
            print(f"   {op:22s} {n / warps:10.0f} /sig  {100 * n / tot:5.1f}%   samples {100 * samples[op] / max(1, sum(samples.values())):5.1f}%")
        s = sum(stall.values())
        print("   stalls: " + ", ".join(f"{k[6:]} {100 * v / s:.1f}%" for k, v in stall.most_common(8)))


if __name__ == "__main__":
    main()
