#!/usr/bin/env python3
"""Driver for compute-sanitizer over the streaming mode: 12 requests of 96 signatures of each curve through a 4-slot queue
(several requests in flight on separate streams with separate scratch), results checked against the C oracle."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np  # noqa: E402

import batches  # noqa: E402
import wgpu_sigops_b200 as w  # noqa: E402

for curve in ("secp256k1", "secp256r1", "ed25519"):
    n, req = 12 * 96, 96
    if curve == "ed25519":
        sigs, msgs, pks, want, _ = batches.ed25519_batch(n, edge_every=5, seed=77)
        reqs = [(sigs[i:i + req], msgs[i:i + req], pks[i:i + req]) for i in range(0, n, req)]
    else:
        cid = 0 if curve == "secp256k1" else 1
        sigs, msgs, want, st, _ = batches.ecdsa_batch(cid, n, edge_every=7, seed=77, mix_high_s=True)
        reqs = [(sigs[i:i + req], msgs[i:i + req]) for i in range(0, n, req)]
    got = list(w.service.run_stream(curve, reqs, max_batch=req, depth=4))
    out = np.concatenate([g[0].reshape(req, -1) for g in got])
    assert np.array_equal(out.reshape(want.shape), want), curve
    print(curve, "queue ok", flush=True)
