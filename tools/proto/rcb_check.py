#!/usr/bin/env python3
"""Numerical check (development tool) of the complete projective formulas of Renes-Costello-Batina 2015 as used by the
lane-group kernel (csrc/group.cuh): Algorithm 8 / 9 (a = 0: mixed addition, doubling) and Algorithm 5 / 6 (a = -3), against
plain affine arithmetic, including the exceptional inputs (P = Q, P = -Q, P = identity (0 : 1 : 0))."""
import os
import random
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "oracle"))
import sigops_oracle as o  # noqa: E402


def aff_add(c, P, Q):
    p = c.p
    if P is None:
        return Q
    if Q is None:
        return P
    if P[0] == Q[0]:
        if (P[1] + Q[1]) % p == 0:
            return None
        lam = (3 * P[0] * P[0] + c.a) * pow(2 * P[1], -1, p) % p
    else:
        lam = (Q[1] - P[1]) * pow(Q[0] - P[0], -1, p) % p
    x = (lam * lam - P[0] - Q[0]) % p
    return (x, (lam * (P[0] - x) - P[1]) % p)


def to_aff(c, P):
    X, Y, Z = P
    if Z % c.p == 0:
        return None
    zi = pow(Z, -1, c.p)
    return (X * zi % c.p, Y * zi % c.p)


def dbl_a0(c, P):  # Algorithm 9
    p, b3 = c.p, 3 * c.b % c.p
    X, Y, Z = P
    t0 = Y * Y % p; Z3 = 8 * t0 % p; t1 = Y * Z % p; t2 = Z * Z % p
    t2 = b3 * t2 % p; X3 = t2 * Z3 % p; Y3 = (t0 + t2) % p; Z3 = t1 * Z3 % p
    t2 = 3 * t2 % p; t0 = (t0 - t2) % p; Y3 = t0 * Y3 % p; Y3 = (X3 + Y3) % p
    t1 = X * Y % p; X3 = t0 * t1 % p; X3 = 2 * X3 % p
    return (X3, Y3, Z3)


def madd_a0(c, P, Q):  # Algorithm 8
    p, b3 = c.p, 3 * c.b % c.p
    X1, Y1, Z1 = P
    X2, Y2 = Q
    t0 = X1 * X2 % p; t1 = Y1 * Y2 % p; t3 = (X2 + Y2) * (X1 + Y1) % p
    t3 = (t3 - t0 - t1) % p
    t4 = (Y2 * Z1 + Y1) % p
    Y3 = (X2 * Z1 + X1) % p
    t0 = 3 * t0 % p
    t2 = b3 * Z1 % p; Z3 = (t1 + t2) % p; t1 = (t1 - t2) % p
    Y3 = b3 * Y3 % p
    X3 = t4 * Y3 % p; t2 = t3 * t1 % p; X3 = (t2 - X3) % p
    Y3 = Y3 * t0 % p; t1 = t1 * Z3 % p; Y3 = (t1 + Y3) % p
    t0 = t0 * t3 % p; Z3 = Z3 * t4 % p; Z3 = (Z3 + t0) % p
    return (X3, Y3, Z3)


def dbl_am3(c, P):  # Algorithm 6
    p, b = c.p, c.b
    X, Y, Z = P
    t0 = X * X % p; t1 = Y * Y % p; t2 = Z * Z % p; t3 = 2 * X * Y % p
    Z3 = 2 * X * Z % p
    Y3 = (b * t2 - Z3) % p
    Y3 = 3 * Y3 % p
    X3 = (t1 - Y3) % p; Y3 = (t1 + Y3) % p
    Y3 = X3 * Y3 % p; X3 = X3 * t3 % p
    t2 = 3 * t2 % p
    Z3 = (b * Z3 - t2 - t0) % p
    Z3 = 3 * Z3 % p
    t0 = (3 * t0 - t2) % p
    t0 = t0 * Z3 % p; Y3 = (Y3 + t0) % p
    t0 = 2 * Y * Z % p
    Z3 = t0 * Z3 % p; X3 = (X3 - Z3) % p
    Z3 = 4 * t0 * t1 % p
    return (X3, Y3, Z3)


def madd_am3(c, P, Q):  # Algorithm 5
    p, b = c.p, c.b
    X1, Y1, Z1 = P
    X2, Y2 = Q
    t0 = X1 * X2 % p; t1 = Y1 * Y2 % p; t3 = (X2 + Y2) * (X1 + Y1) % p
    t3 = (t3 - t0 - t1) % p
    t4 = (Y2 * Z1 + Y1) % p
    Y3 = (X2 * Z1 + X1) % p
    Z3 = b * Z1 % p
    X3 = (Y3 - Z3) % p
    X3 = 3 * X3 % p
    Z3 = (t1 - X3) % p; X3 = (t1 + X3) % p
    Y3 = b * Y3 % p
    t2 = 3 * Z1 % p
    Y3 = (Y3 - t2 - t0) % p
    Y3 = 3 * Y3 % p
    t0 = (3 * t0 - t2) % p
    t1 = t4 * Y3 % p; t2 = t0 * Y3 % p
    Y3 = X3 * Z3 % p; Y3 = (Y3 + t2) % p
    X3 = t3 * X3 % p; X3 = (X3 - t1) % p
    Z3 = t4 * Z3 % p; t1 = t3 * t0 % p; Z3 = (Z3 + t1) % p
    return (X3, Y3, Z3)


def main():
    rng = random.Random(7)
    for c, dbl, madd in ((o.K1, dbl_a0, madd_a0), (o.R1, dbl_am3, madd_am3)):
        G = (c.gx, c.gy)

        def mul(k):
            R, A = None, G
            while k:
                if k & 1:
                    R = aff_add(c, R, A)
                A = aff_add(c, A, A)
                k >>= 1
            return R

        ident = (0, 1, 0)
        assert to_aff(c, dbl(c, ident)) is None
        for it in range(200):
            k1, k2 = rng.randrange(1, c.n), rng.randrange(1, c.n)
            if it % 10 == 0:
                k2 = k1  # P == Q
            if it % 10 == 1:
                k2 = c.n - k1  # P == -Q
            P, Q = mul(k1), mul(k2)
            z = rng.randrange(1, c.p)
            Pj = (P[0] * z % c.p, P[1] * z % c.p, z)
            assert to_aff(c, dbl(c, Pj)) == aff_add(c, P, P), (c.name, "dbl")
            assert to_aff(c, madd(c, Pj, Q)) == aff_add(c, P, Q), (c.name, "madd", it)
            assert to_aff(c, madd(c, ident, Q)) == Q, (c.name, "madd identity")
        print(c.name, "ok: complete doubling / mixed addition incl. P = Q, P = -Q, identity")


if __name__ == "__main__":
    main()
