#!/usr/bin/env python3
"""Generates the fixtures in this directory from oracle/sigops_oracle.py (the spec oracle):

  secp256k1_ecrecover.json / secp256r1_ecrecover.json : {label, sig, msg, pubkey | null}
  ed25519_ecverify.json                               : {label, sig, msg, pk, valid}
  precompute_bases_13.json                            : the three compatibility tables at log_limb_size = 13

The reference (Rust + wgpu) cannot run in this image, so the expected values come from the oracle, which is itself
pinned to the reference's golden vectors, RFC 8032 and OpenSSL by tests/test_oracle.py.  Each file records the
classes the reference never tests (invalid, high-s, non-canonical, small-order: SURVEY.md section 4).
Run:  python tests/golden/make_golden.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import sigops_oracle as o  # noqa: E402
import unit_checks as uc  # noqa: E402


def main():
    for c in (o.K1, o.R1):
        rows = []
        for lab, sig, msg in uc.ecdsa_cases(c, nvalid=16):
            pk = o.ecrecover(c, sig, msg)
            rows.append({"label": lab, "sig": sig.hex(), "msg": msg.hex(), "pubkey": pk.hex() if pk else None})
        json.dump({"source": "oracle/sigops_oracle.py", "cases": rows}, open(os.path.join(HERE, f"{c.name}_ecrecover.json"), "w"), indent=0)
    rows = []
    for lab, sig, msg, pk in uc.ed_cases(nvalid=16):
        rows.append({"label": lab, "sig": sig.hex(), "msg": msg.hex(), "pk": pk.hex(), "valid": o.ecverify_ed25519(sig, msg, pk)})
    json.dump({"source": "oracle/sigops_oracle.py", "cases": rows}, open(os.path.join(HERE, "ed25519_ecverify.json"), "w"), indent=0)
    json.dump({k: o.precompute_bases(k, 13) for k in ("secp256k1", "secp256r1", "ed25519")},
              open(os.path.join(HERE, "precompute_bases_13.json"), "w"))


if __name__ == "__main__":
    main()
