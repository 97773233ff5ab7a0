"""Groundwork OUTSIDE the library (tools/proto/ed_halfscalar.py, plain Python on the oracle's point arithmetic): the
half-size-scalar form of the ed25519 check gives dalek's verdict on every input provided that (i) the scalar congruence is
taken mod 8L, (ii) the multiplier of the defect is odd and (iii) R is required to be canonical.  DESIGN.md section 9 lists it
as kernel-side headroom that is identified but not built; nothing here is on the product path."""
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools", "proto"))
import ed_halfscalar as hs  # noqa: E402
import fuzz_cases  # noqa: E402
import sigops_oracle as o  # noqa: E402


def _mod_l_vector(k):
    """condition (i) dropped: a short vector of the lattice mod L (second coordinate made odd)"""
    a, b = (o.ED_L, 0), (k % o.ED_L, 1)
    while True:
        if a[0] ** 2 + a[1] ** 2 > b[0] ** 2 + b[1] ** 2:
            a, b = b, a
        m = hs._round_div(a[0] * b[0] + a[1] * b[1], a[0] ** 2 + a[1] ** 2)
        if m == 0:
            break
        b = (b[0] - m * a[0], b[1] - m * a[1])
    return next(v for v in (a, b, (a[0] + b[0], a[1] + b[1])) if v[1] & 1)


def _even_vector(k):
    """condition (ii) dropped"""
    v1, v2 = hs.short_odd_vector(k)
    return 2 * v1, 2 * v2


def test_short_vectors():
    rng = random.Random(3)
    ks = [0, 1, 2, o.ED_L - 1, o.ED_L - 2, (o.ED_L + 1) // 2, 2**252, 2**128, 2**127 + 1]
    c = (-8 * o.ED_L) % 10  # 10 k = c (mod 8L) for k = (8L + c) / 10 < L: the SHORTEST vector (c, 10) has an even multiplier
    ks.append((8 * o.ED_L + c) // 10)
    ks += [rng.randrange(o.ED_L) for _ in range(3000)]
    worst = 0
    for k in ks:
        v1, v2 = hs.short_odd_vector(k)
        assert (v1 - v2 * k) % (8 * o.ED_L) == 0 and v2 & 1 and 0 < abs(v2) < o.ED_L
        worst = max(worst, max(abs(v1), abs(v2)).bit_length())
    assert worst <= 254  # the crafted k has no short odd vector at all: the device loop needs its full-length fallback for it
    st = hs.stats(4000)
    assert st["median_bits"] <= 129 and st["share_above_143_bits"] == 0.0


def test_edge_corpus_verdicts_equal_dalek():
    cases = o.ed25519_edge_cases()
    exp = [o.ecverify_ed25519(c[1], c[2], c[3]) for c in cases]
    got = [hs.verify_halfscalar(c[1], c[2], c[3]) for c in cases]
    assert got == exp and 10 < sum(exp) < len(cases)
    # the corpus tells a congruence mod L apart from the one mod 8L (keys with a torsion component)
    wrong = [hs.verify_halfscalar(c[1], c[2], c[3], vector=_mod_l_vector) for c in cases]
    assert wrong != exp


def test_torsion_defects_are_rejected():
    cases = hs.torsion_defect_cases(3)
    assert len(cases) == 21
    for _, sig, msg, pk in cases:
        assert not o.ecverify_ed25519(sig, msg, pk) and not hs.verify_halfscalar(sig, msg, pk)
    # ... and an even multiplier would let the order-2 defect through
    assert any(hs.verify_halfscalar(sig, msg, pk, vector=_even_vector) for _, sig, msg, pk in cases)


def test_fuzz_rows_verdicts_equal_dalek():
    sigs, msgs, pks = fuzz_cases.ed25519_batch(400, seed=21)
    n_ok = 0
    for i in range(400):
        a, b, c = sigs[i].tobytes(), msgs[i].tobytes(), pks[i].tobytes()
        e = o.ecverify_ed25519(a, b, c)
        n_ok += e
        assert hs.verify_halfscalar(a, b, c) == e, i
    assert 10 < n_ok < 390
