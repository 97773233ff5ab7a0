#!/usr/bin/env python3
"""Dynamic instruction mix and stall attribution per kernel from `ncu --page source --csv --print-source sass` output.
   python tools/ncu_mix.py <source.csv> [n_signatures] [--json profiles/executed_macs.json --hash <source hash> --capture <name>]

With --json the wide multiply-accumulates (IMAD.WIDE.U32[.X]) executed per signature are written per curve, tagged with
the hash of the library sources the capture was taken from (wgpu-sigops_b200/libsigops.srchash on the GPU box = build.py's
source_hash()); bench.py reports `executed_frac` only when that hash matches the sources it runs."""
import json
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    nsig = int(sys.argv[2]) if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else 262144
    opt = {sys.argv[i][2:]: sys.argv[i + 1] for i in range(2, len(sys.argv) - 1) if sys.argv[i].startswith("--")}
    macs = {}
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
        curve = "secp256k1" if "CurveK1" in kern else "secp256r1" if "CurveR1" in kern else "ed25519" if "ed25519_verify_kernel" in kern else None
        if curve and curve not in macs:
            macs[curve] = round(sum(n for op, n in ops.items() if op.startswith("IMAD.WIDE")) / warps)
        print(f"== {kern}: {tot} warp-instructions = {tot / warps:.0f} per signature-warp; SASS lines {len(rows)}")
        for op, n in ops.most_common(22):
            print(f"   {op:22s} {n / warps:10.0f} /sig  {100 * n / tot:5.1f}%   samples {100 * samples[op] / max(1, sum(samples.values())):5.1f}%")
        s = sum(stall.values())
        print("   stalls: " + ", ".join(f"{k[6:]} {100 * v / s:.1f}%" for k, v in stall.most_common(8)))
    if "json" in opt:
        json.dump({"wide_macs_per_signature": macs, "source_hash": opt.get("hash"), "capture": opt.get("capture", path),
                   "signatures_per_launch": nsig,
                   "how": "ncu --set full --import-source on; source page (sass): sum of executed IMAD.WIDE.U32[.X] warp "
                          "instructions / (signatures / 32)"}, open(opt["json"], "w"), indent=1)


if __name__ == "__main__":
    main()
