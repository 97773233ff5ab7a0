/*
 * Fast secp256k1 public-key recovery on the CPU -- TEST / BENCHMARK INFRASTRUCTURE ONLY (see sigops_oracle.c's header: nothing
 * under oracle/ is ever linked, loaded or called by the product).
 *
 * Purpose: an HONEST timed CPU baseline for bench.py (`cpu_baseline`, `--impl reference`).  The reference's CPU column is
 * `fuel_crypto::Signature::recover` -> libsecp256k1 (src/benchmarks/secp256k1_ecdsa.rs:117-122), which cannot be built here
 * (no rustc, no vendored secp256k1-sys).  The checker in sigops_oracle.c deliberately uses one generic Montgomery field for
 * all six moduli and no endomorphism, which makes it 3-4x slower per core than libsecp256k1.  This file restates the same
 * decision procedure (src/curve_algos/secp256k1_ecdsa.rs:66-120) with the algorithm class libsecp256k1
 * itself uses: the 2^256 - 2^32 - 977 field with a fold reduction, the GLV endomorphism (constants as carried by the reference
 * at src/curve_algos/secp256k1_curve.rs:47-68), an interleaved wNAF Strauss pass over four ~128-bit streams (width 5 on R and
 * lambda*R, width 12 on precomputed affine tables of G and lambda*G), and variable-time modular inverses.
 * It is pinned to oracle_ecrecover (the checker) on random inputs and on the whole edge corpus by tests/test_oracle.py.
 */
#include <pthread.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef uint64_t u64;
typedef unsigned __int128 u128;

typedef struct {
    u64 v[4];
} fe;

static const fe FP = {{0xFFFFFFFEFFFFFC2FULL, 0xFFFFFFFFFFFFFFFFULL, 0xFFFFFFFFFFFFFFFFULL, 0xFFFFFFFFFFFFFFFFULL}};
static const fe FN = {{0xBFD25E8CD0364141ULL, 0xBAAEDCE6AF48A03BULL, 0xFFFFFFFFFFFFFFFEULL, 0xFFFFFFFFFFFFFFFFULL}};
#define PC 0x1000003D1ULL /* 2^256 mod p */
/* 2^256 mod n (129 bits) */
static const u64 NC[3] = {0x402DA1732FC9BEBFULL, 0x4551231950B75FC4ULL, 1ULL};
static const fe BETA = {{0xC1396C28719501EEULL, 0x9CF0497512F58995ULL, 0x6E64479EAC3434E9ULL, 0x7AE96A2B657C0710ULL}};
static const fe G1C = {{0xE893209A45DBB031ULL, 0x3DAA8A1471E8CA7FULL, 0xE86C90E49284EB15ULL, 0x3086D221A7D46BCDULL}};
static const fe G2C = {{0x1571B4AE8AC47F71ULL, 0x221208AC9DF506C6ULL, 0x6F547FA90ABFE4C4ULL, 0xE4437ED6010E8828ULL}};
static const fe A1C = {{0xE86C90E49284EB15ULL, 0x3086D221A7D46BCDULL, 0, 0}};          /* a1 = b2 */
static const fe A2C = {{0x57C1108D9D44CFD8ULL, 0x14CA50F7A8E2F3F6ULL, 1, 0}};
static const fe MB1C = {{0x6F547FA90ABFE4C3ULL, 0xE4437ED6010E8828ULL, 0, 0}};         /* -b1 */
static const fe GX = {{0x59F2815B16F81798ULL, 0x029BFCDB2DCE28D9ULL, 0x55A06295CE870B07ULL, 0x79BE667EF9DCBBACULL}};
static const fe GY = {{0x9C47D08FFB10D4B8ULL, 0xFD17B448A6855419ULL, 0x5DA4FBFC0E1108A8ULL, 0x483ADA7726A3C465ULL}};

/* ------------------------------------------------------------------------------------------------ 256-bit helpers */
static inline int fe_is_zero(const fe* a) { return (a->v[0] | a->v[1] | a->v[2] | a->v[3]) == 0; }
static inline int fe_eq(const fe* a, const fe* b) {
    return ((a->v[0] ^ b->v[0]) | (a->v[1] ^ b->v[1]) | (a->v[2] ^ b->v[2]) | (a->v[3] ^ b->v[3])) == 0;
}
static inline int fe_gte(const fe* a, const fe* b) {
    for (int i = 3; i >= 0; i--) {
        if (a->v[i] > b->v[i]) return 1;
        if (a->v[i] < b->v[i]) return 0;
    }
    return 1;
}
static inline u64 add4(fe* r, const fe* a, const fe* b) {
    u128 t = 0;
    for (int i = 0; i < 4; i++) {
        t += (u128)a->v[i] + b->v[i];
        r->v[i] = (u64)t;
        t >>= 64;
    }
    return (u64)t;
}
static inline u64 sub4(fe* r, const fe* a, const fe* b) {
    u64 bw = 0;
    for (int i = 0; i < 4; i++) {
        u128 t = (u128)a->v[i] - b->v[i] - bw;
        r->v[i] = (u64)t;
        bw = (u64)(t >> 64) & 1;
    }
    return bw;
}
static void fe_from_be(fe* r, const uint8_t* b) {
    for (int i = 0; i < 4; i++) {
        u64 w = 0;
        for (int j = 0; j < 8; j++) w = (w << 8) | b[8 * (3 - i) + j];
        r->v[i] = w;
    }
}
static void fe_to_be(uint8_t* b, const fe* a) {
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 8; j++) b[8 * (3 - i) + j] = (uint8_t)(a->v[i] >> (56 - 8 * j));
}
/* t[0..8) = a * b */
static inline void mul4x4(u64* t, const fe* a, const fe* b) {
    u64 r[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        u64 cy = 0;
        for (int j = 0; j < 4; j++) {
            u128 x = (u128)a->v[i] * b->v[j] + r[i + j] + cy;
            r[i + j] = (u64)x;
            cy = (u64)(x >> 64);
        }
        r[i + 4] = cy;
    }
    memcpy(t, r, sizeof r);
}

/* ------------------------------------------------------------------------------------------------ field mod p */
static inline void fp_reduce(fe* r, const u64* t) {
    u64 s[4];
    u128 acc = (u128)t[4] * PC + t[0];
    s[0] = (u64)acc;
    acc >>= 64;
    acc += (u128)t[5] * PC + t[1];
    s[1] = (u64)acc;
    acc >>= 64;
    acc += (u128)t[6] * PC + t[2];
    s[2] = (u64)acc;
    acc >>= 64;
    acc += (u128)t[7] * PC + t[3];
    s[3] = (u64)acc;
    u64 top = (u64)(acc >> 64); /* < 2^34 */
    acc = (u128)top * PC + s[0];
    s[0] = (u64)acc;
    acc >>= 64;
    acc += s[1];
    s[1] = (u64)acc;
    acc >>= 64;
    acc += s[2];
    s[2] = (u64)acc;
    acc >>= 64;
    acc += s[3];
    s[3] = (u64)acc;
    if ((u64)(acc >> 64)) { /* wrapped once more: the value is tiny, adding 2^256 mod p cannot carry again */
        acc = (u128)s[0] + PC;
        s[0] = (u64)acc;
        acc >>= 64;
        acc += s[1];
        s[1] = (u64)acc;
        acc >>= 64;
        acc += s[2];
        s[2] = (u64)acc;
        s[3] += (u64)(acc >> 64);
    }
    fe x = {{s[0], s[1], s[2], s[3]}}, y;
    if (!sub4(&y, &x, &FP)) x = y; /* canonical */
    *r = x;
}
static inline void fp_mul(fe* r, const fe* a, const fe* b) {
    u64 t[8];
    mul4x4(t, a, b);
    fp_reduce(r, t);
}
static inline void fp_sqr(fe* r, const fe* a) {
    /* 6 cross products doubled + 4 squares */
    const u64 a0 = a->v[0], a1 = a->v[1], a2 = a->v[2], a3 = a->v[3];
    u64 t[8];
    u128 x;
    u64 c;
    x = (u128)a0 * a1;
    t[1] = (u64)x;
    c = (u64)(x >> 64);
    x = (u128)a0 * a2 + c;
    t[2] = (u64)x;
    c = (u64)(x >> 64);
    x = (u128)a0 * a3 + c;
    t[3] = (u64)x;
    t[4] = (u64)(x >> 64);
    x = (u128)a1 * a2 + t[3];
    t[3] = (u64)x;
    c = (u64)(x >> 64);
    x = (u128)a1 * a3 + t[4] + c;
    t[4] = (u64)x;
    t[5] = (u64)(x >> 64);
    x = (u128)a2 * a3 + t[5];
    t[5] = (u64)x;
    t[6] = (u64)(x >> 64);
    t[7] = t[6] >> 63;
    t[6] = (t[6] << 1) | (t[5] >> 63);
    t[5] = (t[5] << 1) | (t[4] >> 63);
    t[4] = (t[4] << 1) | (t[3] >> 63);
    t[3] = (t[3] << 1) | (t[2] >> 63);
    t[2] = (t[2] << 1) | (t[1] >> 63);
    t[1] = t[1] << 1;
    x = (u128)a0 * a0;
    t[0] = (u64)x;
    x = (x >> 64) + t[1];
    t[1] = (u64)x;
    c = (u64)(x >> 64);
    x = (u128)a1 * a1 + t[2] + c;
    t[2] = (u64)x;
    x = (x >> 64) + t[3];
    t[3] = (u64)x;
    c = (u64)(x >> 64);
    x = (u128)a2 * a2 + t[4] + c;
    t[4] = (u64)x;
    x = (x >> 64) + t[5];
    t[5] = (u64)x;
    c = (u64)(x >> 64);
    x = (u128)a3 * a3 + t[6] + c;
    t[6] = (u64)x;
    t[7] += (u64)(x >> 64);
    fp_reduce(r, t);
}
static inline void fp_add(fe* r, const fe* a, const fe* b) {
    fe t, u;
    u64 cy = add4(&t, a, b);
    u64 bw = sub4(&u, &t, &FP);
    *r = (cy || !bw) ? u : t;
}
static inline void fp_sub(fe* r, const fe* a, const fe* b) {
    fe t, u;
    u64 bw = sub4(&t, a, b);
    add4(&u, &t, &FP);
    *r = bw ? u : t;
}
static inline void fp_neg(fe* r, const fe* a) {
    if (fe_is_zero(a))
        *r = *a;
    else
        sub4(r, &FP, a);
}
static inline void fp_dbl(fe* r, const fe* a) { fp_add(r, a, a); }
static void fp_sqr_n(fe* r, const fe* a, int n) {
    *r = *a;
    for (int i = 0; i < n; i++) fp_sqr(r, r);
}
/* a^((p+1)/4): candidate square root (the chain libsecp256k1 uses: 254 squarings + 13 multiplications) */
static void fp_sqrt(fe* r, const fe* a) {
    fe x2, x3, x6, x9, x11, x22, x44, x88, x176, x220, x223, t;
    fp_sqr(&t, a);
    fp_mul(&x2, &t, a);
    fp_sqr(&t, &x2);
    fp_mul(&x3, &t, a);
    fp_sqr_n(&t, &x3, 3);
    fp_mul(&x6, &t, &x3);
    fp_sqr_n(&t, &x6, 3);
    fp_mul(&x9, &t, &x3);
    fp_sqr_n(&t, &x9, 2);
    fp_mul(&x11, &t, &x2);
    fp_sqr_n(&t, &x11, 11);
    fp_mul(&x22, &t, &x11);
    fp_sqr_n(&t, &x22, 22);
    fp_mul(&x44, &t, &x22);
    fp_sqr_n(&t, &x44, 44);
    fp_mul(&x88, &t, &x44);
    fp_sqr_n(&t, &x88, 88);
    fp_mul(&x176, &t, &x88);
    fp_sqr_n(&t, &x176, 44);
    fp_mul(&x220, &t, &x44);
    fp_sqr_n(&t, &x220, 3);
    fp_mul(&x223, &t, &x3);
    fp_sqr_n(&t, &x223, 23);
    fp_mul(&t, &t, &x22);
    fp_sqr_n(&t, &t, 6);
    fp_mul(&t, &t, &x2);
    fp_sqr(&t, &t);
    fp_sqr(r, &t);
}

/* variable-time modular inverse (binary extended Euclid) for an odd modulus m; a in [1, m) */
static void modinv(fe* r, const fe* a, const fe* m) {
    fe u = *a, v = *m, x1 = {{1, 0, 0, 0}}, x2 = {{0, 0, 0, 0}};
    const fe one = {{1, 0, 0, 0}};
    while (!fe_eq(&u, &one) && !fe_eq(&v, &one)) {
        while (!(u.v[0] & 1)) {
            u.v[0] = (u.v[0] >> 1) | (u.v[1] << 63);
            u.v[1] = (u.v[1] >> 1) | (u.v[2] << 63);
            u.v[2] = (u.v[2] >> 1) | (u.v[3] << 63);
            u.v[3] >>= 1;
            u64 cy = 0;
            if (x1.v[0] & 1) cy = add4(&x1, &x1, m);
            x1.v[0] = (x1.v[0] >> 1) | (x1.v[1] << 63);
            x1.v[1] = (x1.v[1] >> 1) | (x1.v[2] << 63);
            x1.v[2] = (x1.v[2] >> 1) | (x1.v[3] << 63);
            x1.v[3] = (x1.v[3] >> 1) | (cy << 63);
        }
        while (!(v.v[0] & 1)) {
            v.v[0] = (v.v[0] >> 1) | (v.v[1] << 63);
            v.v[1] = (v.v[1] >> 1) | (v.v[2] << 63);
            v.v[2] = (v.v[2] >> 1) | (v.v[3] << 63);
            v.v[3] >>= 1;
            u64 cy = 0;
            if (x2.v[0] & 1) cy = add4(&x2, &x2, m);
            x2.v[0] = (x2.v[0] >> 1) | (x2.v[1] << 63);
            x2.v[1] = (x2.v[1] >> 1) | (x2.v[2] << 63);
            x2.v[2] = (x2.v[2] >> 1) | (x2.v[3] << 63);
            x2.v[3] = (x2.v[3] >> 1) | (cy << 63);
        }
        if (fe_gte(&u, &v)) {
            sub4(&u, &u, &v);
            if (sub4(&x1, &x1, &x2)) add4(&x1, &x1, m);
        } else {
            sub4(&v, &v, &u);
            if (sub4(&x2, &x2, &x1)) add4(&x2, &x2, m);
        }
    }
    *r = fe_eq(&u, &one) ? x1 : x2;
}

/* ------------------------------------------------------------------------------------------------ scalars mod n */
/* r = t mod n for a 512-bit t: fold the high half with 2^256 = NC (mod n) three times, then conditional subtractions */
static void sc_reduce512(fe* r, const u64* t) {
    u64 m[7]; /* lo + hi * NC  (< 2^386) */
    {
        u64 p[7] = {0, 0, 0, 0, 0, 0, 0};
        for (int i = 0; i < 4; i++) {
            u64 cy = 0;
            for (int j = 0; j < 3; j++) {
                u128 x = (u128)t[4 + i] * NC[j] + p[i + j] + cy;
                p[i + j] = (u64)x;
                cy = (u64)(x >> 64);
            }
            p[i + 3] += cy;
        }
        u128 acc = 0;
        for (int i = 0; i < 7; i++) {
            acc += (u128)p[i] + (i < 4 ? t[i] : 0);
            m[i] = (u64)acc;
            acc >>= 64;
        }
    }
    u64 q[5]; /* m[0..4) + m[4..7) * NC  (< 2^260) */
    {
        u64 p[6] = {0, 0, 0, 0, 0, 0};
        for (int i = 0; i < 3; i++) {
            u64 cy = 0;
            for (int j = 0; j < 3; j++) {
                u128 x = (u128)m[4 + i] * NC[j] + p[i + j] + cy;
                p[i + j] = (u64)x;
                cy = (u64)(x >> 64);
            }
            p[i + 3] += cy;
        }
        u128 acc = 0;
        for (int i = 0; i < 5; i++) {
            acc += (u128)p[i] + (i < 4 ? m[i] : 0);
            q[i] = (u64)acc;
            acc >>= 64;
        }
    }
    fe x;
    u64 top;
    {
        u128 acc = (u128)q[4] * NC[0] + q[0];
        x.v[0] = (u64)acc;
        acc >>= 64;
        acc += (u128)q[4] * NC[1] + q[1];
        x.v[1] = (u64)acc;
        acc >>= 64;
        acc += (u128)q[4] * NC[2] + q[2];
        x.v[2] = (u64)acc;
        acc >>= 64;
        acc += q[3];
        x.v[3] = (u64)acc;
        top = (u64)(acc >> 64);
    }
    fe y;
    while (top || fe_gte(&x, &FN)) {
        u64 bw = sub4(&y, &x, &FN);
        top -= bw;
        x = y;
    }
    *r = x;
}
static void sc_mul(fe* r, const fe* a, const fe* b) {
    u64 t[8];
    mul4x4(t, a, b);
    sc_reduce512(r, t);
}

/* ------------------------------------------------------------------------------------------------ group */
typedef struct {
    fe X, Y, Z;
    int inf;
} jac;
typedef struct {
    fe x, y;
} aff;

static void jac_dbl(jac* P) { /* dbl-2009-l, a = 0 */
    if (P->inf) return;
    fe A, B, C, D, E, F, t;
    fp_sqr(&A, &P->X);
    fp_sqr(&B, &P->Y);
    fp_sqr(&C, &B);
    fp_add(&t, &P->X, &B);
    fp_sqr(&t, &t);
    fp_sub(&t, &t, &A);
    fp_sub(&t, &t, &C);
    fp_dbl(&D, &t);
    fp_dbl(&E, &A);
    fp_add(&E, &E, &A);
    fp_sqr(&F, &E);
    fp_mul(&P->Z, &P->Y, &P->Z);
    fp_dbl(&P->Z, &P->Z);
    fp_dbl(&t, &D);
    fp_sub(&P->X, &F, &t);
    fp_sub(&t, &D, &P->X);
    fp_mul(&t, &E, &t);
    fp_dbl(&C, &C);
    fp_dbl(&C, &C);
    fp_dbl(&C, &C);
    fp_sub(&P->Y, &t, &C);
}
static void jac_add_tail(jac* P, const fe* U1, const fe* S1, const fe* H, const fe* r, const fe* Zm) {
    fe HH, HHH, V, t;
    fp_sqr(&HH, H);
    fp_mul(&HHH, H, &HH);
    fp_mul(&V, U1, &HH);
    fp_sqr(&t, r);
    fp_sub(&t, &t, &HHH);
    fp_sub(&t, &t, &V);
    fp_sub(&P->X, &t, &V);
    fp_sub(&t, &V, &P->X);
    fp_mul(&t, r, &t);
    fp_mul(&HHH, S1, &HHH);
    fp_sub(&P->Y, &t, &HHH);
    fp_mul(&P->Z, Zm, H);
}
static void jac_madd(jac* P, const fe* x2, const fe* y2) { /* affine second operand, never infinity */
    if (P->inf) {
        P->X = *x2;
        P->Y = *y2;
        P->Z = (fe){{1, 0, 0, 0}};
        P->inf = 0;
        return;
    }
    fe Z1Z1, U2, S2, H, r;
    fp_sqr(&Z1Z1, &P->Z);
    fp_mul(&U2, x2, &Z1Z1);
    fp_mul(&S2, &P->Z, &Z1Z1);
    fp_mul(&S2, y2, &S2);
    fp_sub(&H, &U2, &P->X);
    fp_sub(&r, &S2, &P->Y);
    if (fe_is_zero(&H)) {
        if (fe_is_zero(&r))
            jac_dbl(P);
        else
            P->inf = 1;
        return;
    }
    fe U1 = P->X, S1 = P->Y, Zm = P->Z;
    jac_add_tail(P, &U1, &S1, &H, &r, &Zm);
}
static void jac_add(jac* P, const jac* Q) { /* Q never infinity */
    if (P->inf) {
        *P = *Q;
        return;
    }
    fe Z1Z1, Z2Z2, U1, U2, S1, S2, H, r, Zm;
    fp_sqr(&Z1Z1, &P->Z);
    fp_sqr(&Z2Z2, &Q->Z);
    fp_mul(&U1, &P->X, &Z2Z2);
    fp_mul(&U2, &Q->X, &Z1Z1);
    fp_mul(&S1, &Q->Z, &Z2Z2);
    fp_mul(&S1, &P->Y, &S1);
    fp_mul(&S2, &P->Z, &Z1Z1);
    fp_mul(&S2, &Q->Y, &S2);
    fp_sub(&H, &U2, &U1);
    fp_sub(&r, &S2, &S1);
    if (fe_is_zero(&H)) {
        if (fe_is_zero(&r))
            jac_dbl(P);
        else
            P->inf = 1;
        return;
    }
    fp_mul(&Zm, &P->Z, &Q->Z);
    jac_add_tail(P, &U1, &S1, &H, &r, &Zm);
}

/* ---- GLV split: k = k1 + k2*lambda (mod n), |k1|, |k2| < 2^128 + small; mirrors csrc/scalar.cuh k1_glv_split ---- */
typedef struct {
    u64 k1[3], k2[3];
    int neg1, neg2;
} glv;
static void neg256(fe* r, const fe* a) {
    fe z = {{0, 0, 0, 0}};
    sub4(r, &z, a);
}
static void glv_split(glv* o, const fe* k) {
    u64 t[8];
    fe c1 = {{0, 0, 0, 0}}, c2 = {{0, 0, 0, 0}};
    mul4x4(t, k, &G1C);
    { /* round(k*g1 / 2^384): add 2^383 */
        u128 x = (u128)t[5] + 0x8000000000000000ULL;
        u64 cy = (u64)(x >> 64);
        x = (u128)t[6] + cy;
        c1.v[0] = (u64)x;
        c1.v[1] = t[7] + (u64)(x >> 64);
    }
    mul4x4(t, k, &G2C);
    {
        u128 x = (u128)t[5] + 0x8000000000000000ULL;
        u64 cy = (u64)(x >> 64);
        x = (u128)t[6] + cy;
        c2.v[0] = (u64)x;
        c2.v[1] = t[7] + (u64)(x >> 64);
    }
    u64 p1[8], p2[8];
    fe r1, r2, a, b;
    mul4x4(p1, &c1, &A1C);
    mul4x4(p2, &c2, &A2C);
    memcpy(&a, p1, 32);
    memcpy(&b, p2, 32);
    sub4(&r1, k, &a);
    sub4(&r1, &r1, &b);
    mul4x4(p1, &c1, &MB1C);
    mul4x4(p2, &c2, &A1C);
    memcpy(&a, p1, 32);
    memcpy(&b, p2, 32);
    sub4(&r2, &a, &b);
    o->neg1 = (int)(r1.v[3] >> 63);
    o->neg2 = (int)(r2.v[3] >> 63);
    if (o->neg1) neg256(&r1, &r1);
    if (o->neg2) neg256(&r2, &r2);
    for (int i = 0; i < 3; i++) {
        o->k1[i] = r1.v[i];
        o->k2[i] = r2.v[i];
    }
}

/* width-w NAF of a magnitude below 2^130 (3 limbs); digits odd in (-2^(w-1), 2^(w-1)); returns the length */
#define NAF_LEN 132
static int wnaf(int16_t* naf, const u64* k, int w) {
    u64 t[3] = {k[0], k[1], k[2]};
    int len = 0;
    memset(naf, 0, NAF_LEN * sizeof(int16_t));
    for (int i = 0; i < NAF_LEN; i++) {
        if (t[0] & 1) {
            int d = (int)(t[0] & ((1u << w) - 1));
            if (d >= (1 << (w - 1))) d -= 1 << w;
            naf[i] = (int16_t)d;
            len = i + 1;
            if (d >= 0) {
                u64 bw = (u64)d;
                for (int j = 0; j < 3 && bw; j++) {
                    u64 o = t[j];
                    t[j] -= bw;
                    bw = o < bw;
                }
            } else {
                u64 cy = (u64)(-d);
                for (int j = 0; j < 3 && cy; j++) {
                    t[j] += cy;
                    cy = t[j] < cy;
                }
            }
        }
        t[0] = (t[0] >> 1) | (t[1] << 63);
        t[1] = (t[1] >> 1) | (t[2] << 63);
        t[2] >>= 1;
    }
    return len;
}

#define GW 12
#define GTAB (1 << (GW - 2)) /* odd multiples 1, 3, ..., 2^(GW-1) - 1 */
#define PW 5
#define PTAB (1 << (PW - 2))
static aff g_tab[GTAB], g_tab_lam[GTAB];
static pthread_once_t g_once = PTHREAD_ONCE_INIT;

static void init_tables(void) {
    jac* J = malloc(sizeof(jac) * GTAB);
    jac P = {GX, GY, {{1, 0, 0, 0}}, 0}, G2 = P;
    jac_dbl(&G2);
    for (int i = 0; i < GTAB; i++) {
        J[i] = P;
        jac_add(&P, &G2);
    }
    for (int i = 0; i < GTAB; i++) {
        fe zi, zi2;
        modinv(&zi, &J[i].Z, &FP);
        fp_sqr(&zi2, &zi);
        fp_mul(&g_tab[i].x, &J[i].X, &zi2);
        fp_mul(&zi2, &zi2, &zi);
        fp_mul(&g_tab[i].y, &J[i].Y, &zi2);
        fp_mul(&g_tab_lam[i].x, &g_tab[i].x, &BETA);
        g_tab_lam[i].y = g_tab[i].y;
    }
    free(J);
}

/* SURVEY.md appendix B.  Returns 0 and writes X||Y, or 1 (invalid) and writes 64 zero bytes. */
static int recover_one(const uint8_t* sig, const uint8_t* msg, uint8_t* out) {
    memset(out, 0, 64);
    uint8_t sb[32];
    memcpy(sb, sig + 32, 32);
    const int parity = sb[0] >> 7;
    sb[0] &= 0x7f;
    fe r, s, z;
    fe_from_be(&r, sig);
    fe_from_be(&s, sb);
    fe_from_be(&z, msg);
    if (fe_is_zero(&r) || fe_is_zero(&s) || fe_gte(&r, &FN) || fe_gte(&s, &FN)) return 1;
    if (fe_gte(&z, &FN)) sub4(&z, &z, &FN);
    /* lift x = r */
    fe t, y, y2;
    const fe seven = {{7, 0, 0, 0}};
    fp_sqr(&t, &r);
    fp_mul(&t, &t, &r);
    fp_add(&t, &t, &seven);
    fp_sqrt(&y, &t);
    fp_sqr(&y2, &y);
    if (!fe_eq(&y2, &t)) return 1;
    if ((int)(y.v[0] & 1) != parity) fp_neg(&y, &y);
    /* u1 = -z/r, u2 = s/r */
    fe rinv, u1, u2;
    modinv(&rinv, &r, &FN);
    sc_mul(&u2, &rinv, &s);
    sc_mul(&u1, &rinv, &z);
    if (!fe_is_zero(&u1)) sub4(&u1, &FN, &u1);
    glv sr, sg;
    glv_split(&sr, &u2);
    glv_split(&sg, &u1);
    int16_t n0[NAF_LEN], n1[NAF_LEN], n2[NAF_LEN], n3[NAF_LEN];
    int len = wnaf(n0, sr.k1, PW), l;
    if ((l = wnaf(n1, sr.k2, PW)) > len) len = l;
    if ((l = wnaf(n2, sg.k1, GW)) > len) len = l;
    if ((l = wnaf(n3, sg.k2, GW)) > len) len = l;
    /* odd multiples of R (Jacobian) */
    jac ptab[PTAB], R2;
    ptab[0] = (jac){r, y, {{1, 0, 0, 0}}, 0};
    R2 = ptab[0];
    jac_dbl(&R2);
    for (int i = 1; i < PTAB; i++) {
        ptab[i] = ptab[i - 1];
        jac_add(&ptab[i], &R2);
    }
    jac Q;
    memset(&Q, 0, sizeof Q);
    Q.inf = 1;
    for (int i = len - 1; i >= 0; i--) {
        jac_dbl(&Q);
        if (n0[i]) {
            int d = n0[i];
            jac T = ptab[(d < 0 ? -d : d) >> 1];
            if ((d < 0) != sr.neg1) fp_neg(&T.Y, &T.Y);
            jac_add(&Q, &T);
        }
        if (n1[i]) {
            int d = n1[i];
            jac T = ptab[(d < 0 ? -d : d) >> 1];
            fp_mul(&T.X, &T.X, &BETA);
            if ((d < 0) != sr.neg2) fp_neg(&T.Y, &T.Y);
            jac_add(&Q, &T);
        }
        if (n2[i]) {
            int d = n2[i];
            const aff* a = &g_tab[(d < 0 ? -d : d) >> 1];
            fe yy = a->y;
            if ((d < 0) != sg.neg1) fp_neg(&yy, &yy);
            jac_madd(&Q, &a->x, &yy);
        }
        if (n3[i]) {
            int d = n3[i];
            const aff* a = &g_tab_lam[(d < 0 ? -d : d) >> 1];
            fe yy = a->y;
            if ((d < 0) != sg.neg2) fp_neg(&yy, &yy);
            jac_madd(&Q, &a->x, &yy);
        }
    }
    if (Q.inf) return 1;
    fe zi, zi2, ax, ay;
    modinv(&zi, &Q.Z, &FP);
    fp_sqr(&zi2, &zi);
    fp_mul(&ax, &Q.X, &zi2);
    fp_mul(&zi2, &zi2, &zi);
    fp_mul(&ay, &Q.Y, &zi2);
    fe_to_be(out, &ax);
    fe_to_be(out + 32, &ay);
    return 0;
}

typedef struct {
    const uint8_t *sigs, *msgs;
    uint8_t *out, *status;
    size_t lo, hi;
} job;
static void* worker(void* p) {
    job* j = p;
    for (size_t i = j->lo; i < j->hi; i++) j->status[i] = (uint8_t)recover_one(j->sigs + 64 * i, j->msgs + 32 * i, j->out + 64 * i);
    return NULL;
}

int oracle_k1_ecrecover_fast(const uint8_t* sigs, const uint8_t* msgs, size_t n, uint8_t* out, uint8_t* status, int threads) {
    pthread_once(&g_once, init_tables);
    if (threads < 1) threads = 1;
    if ((size_t)threads > n) threads = n ? (int)n : 1;
    pthread_t* th = malloc(sizeof(pthread_t) * threads);
    job* jobs = malloc(sizeof(job) * threads);
    for (int t = 0; t < threads; t++) {
        jobs[t] = (job){sigs, msgs, out, status, n * t / threads, n * (t + 1) / threads};
        if (t > 0) pthread_create(&th[t], NULL, worker, &jobs[t]);
    }
    worker(&jobs[0]);
    for (int t = 1; t < threads; t++) pthread_join(th[t], NULL);
    free(th);
    free(jobs);
    return 0;
}
