#!/usr/bin/env python3
"""PROTOTYPE, outside the library (groundwork, not product code): the half-size-scalar ed25519 check of Antipa et al. arranged so
that its verdict equals ed25519-dalek 2.1.1 `verify` (cofactorless, byte compare) on EVERY input, with plain Python integers on
top of the oracle's point arithmetic.  DESIGN.md section 9 costs what it would buy on the device (~128 shared doublings instead
of 252); this file pins down the three conditions that make it exact and measures how long the scalars get.

dalek accepts  <=>  A decompresses, s < L, and compress([s]B - [k]A) == R_bytes with k = SHA-512(R || A || M) mod L
               <=>  ... and R_bytes is the CANONICAL encoding of a curve point R with  D := [s]B - [k]A - R = identity.

Half-size form: pick (v1, v2) != 0 with  v1 = v2 * k  (mod 8 L)  and both about 128 bits, then test
               [v2 s mod L] B + [v1] (-A) + [v2] (-R) == identity                                        (*)
  (i)   the congruence is mod 8 L, not mod L: A may carry a torsion component and dalek multiplies by the INTEGER k < L, so
        [v1]A = [v2 k]A needs v1 - v2 k to kill the whole group (order 8 L); [v2 s mod L]B = [v2 s]B because B has order L;
        then (*) is exactly [v2] D = identity;
  (ii)  v2 is odd (and 0 < |v2| < L): the order of D divides 8 L, so [v2]D = identity <=> D = identity.  With an even v2 a
        defect of order 2 (an R shifted by a torsion point) would be accepted;
  (iii) R_bytes must be canonical (y < p, and not x = 0 with the sign bit set) and on the curve: the byte compare rejects
        everything else, whatever the group equation says.
The lattice {(a, b): a = b k mod 8L} has determinant 8 L ~ 2^255; Lagrange reduction gives a basis (b1, b2) with |b1| |b2| <~
1.16 * 8 L.  At least one of b1, b2, b1 + b2 has an odd second coordinate (the lattice contains (k, 1)); the shortest such vector
is used.  `stats()` reports the distribution of its length: a device loop of 35 signed 4-bit windows (140 doublings) covers
|v| < 2^143 and would need the full-length fallback with probability ~2^-30 per signature -- which still has to exist, because
an adversary can grind k."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import sigops_oracle as o  # noqa: E402  (prototype on top of the test oracle; never imported by the product)

N8L = 8 * o.ED_L


def _round_div(a: int, b: int) -> int:
    """nearest integer to a / b (b > 0)"""
    return (2 * a + b) // (2 * b)


def short_odd_vector(k: int):
    """(v1, v2) with v1 = v2 k (mod 8L), v2 odd, max(|v1|, |v2|) as small as the reduced basis allows."""
    a, b = (N8L, 0), (k % N8L, 1)
    while True:  # Lagrange / Gauss reduction
        if a[0] * a[0] + a[1] * a[1] > b[0] * b[0] + b[1] * b[1]:
            a, b = b, a
        na = a[0] * a[0] + a[1] * a[1]
        if na == 0:
            break
        m = _round_div(a[0] * b[0] + a[1] * b[1], na)
        if m == 0:
            break
        b = (b[0] - m * a[0], b[1] - m * a[1])
    cands = [a, b, (a[0] + b[0], a[1] + b[1]), (a[0] - b[0], a[1] - b[1])]
    cands = [v for v in cands if v[1] & 1]
    assert cands, "the lattice contains (k, 1): some candidate must have an odd second coordinate"
    v = min(cands, key=lambda v: max(abs(v[0]), abs(v[1])))
    assert (v[0] - v[1] * k) % N8L == 0 and v[1] % 2 == 1 and 0 < abs(v[1]) < o.ED_L
    return v


def decompress_canonical(b: bytes):
    """The point R_bytes encodes if it is the canonical encoding of a curve point (what `compress` can output), else None."""
    y = int.from_bytes(b, "little") & (2**255 - 1)
    if y >= o.ED_P:
        return None
    P = o.ed_decompress(b)
    if P is None:
        return None
    if P[0] == 0 and (b[31] >> 7):
        return None
    assert o.ed_compress(P) == b
    return P


def is_identity(P) -> bool:
    return P[0] % o.ED_P == 0 and (P[1] - P[2]) % o.ED_P == 0


def signed_mul(v: int, P):
    return o.ed_mul(v, P) if v >= 0 else o.ed_mul(-v, o.ed_neg(P))


def verify_halfscalar(sig: bytes, msg: bytes, pk: bytes, vector=None) -> bool:
    """`vector`: k -> (v1, v2); default short_odd_vector.  The tests pass deliberately wrong choices (even v2, congruence mod L
    only) to show that the corpus tells them apart."""
    A = o.ed_decompress(pk)
    if A is None:
        return False
    s = int.from_bytes(sig[32:], "little")
    if s >= o.ED_L:
        return False
    R = decompress_canonical(sig[:32])
    if R is None:
        return False
    k = o.ed_challenge(sig[:32], pk, msg)
    v1, v2 = (vector or short_odd_vector)(k)
    acc = o.ed_mul(v2 * s % o.ED_L, o.ED_B)
    acc = o.ed_add(acc, signed_mul(v1, o.ed_neg(A)))
    acc = o.ed_add(acc, signed_mul(v2, o.ed_neg(R)))
    return is_identity(acc)


def torsion_defect_cases(count: int = 8):
    """Signatures whose defect D = [s]B - [k]A - R is a NON-ZERO torsion point (R shifted by each small-order point): dalek
    rejects all of them; a cofactored check, or (*) with an even v2 for the order-2 defect, would accept."""
    import hashlib

    out = []
    small = [o.ed_decompress(b) for b in o.ed_small_order_points()]
    small = [T for T in small if T is not None and not is_identity(T)]
    for i in range(count):
        seed = hashlib.sha256(b"halfscalar-torsion-%d" % i).digest()
        a, prefix, pk = o.ed25519_expand(seed)
        msg = hashlib.sha256(seed).digest()
        r = int.from_bytes(hashlib.sha512(prefix + msg).digest(), "little") % o.ED_L
        for j, T in enumerate(small):
            Rb = o.ed_compress(o.ed_add(o.ed_mul(r, o.ED_B), T))
            k = o.ed_challenge(Rb, pk, msg)
            s = (r + k * a) % o.ED_L
            out.append(("defect_torsion_%d_%d" % (i, j), Rb + s.to_bytes(32, "little"), msg, pk))
    return out


def stats(n: int = 20000, seed: int = 7):
    import random

    rng = random.Random(seed)
    bits = []
    for _ in range(n):
        v1, v2 = short_odd_vector(rng.randrange(o.ED_L))
        bits.append(max(abs(v1), abs(v2)).bit_length())
    bits.sort()
    return {"n": n, "median_bits": bits[n // 2], "p99_bits": bits[int(n * 0.99)], "max_bits": bits[-1],
            "share_above_131_bits": sum(b > 131 for b in bits) / n, "share_above_143_bits": sum(b > 143 for b in bits) / n}


if __name__ == "__main__":
    print(stats())
