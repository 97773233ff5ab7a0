/*
 * CPU oracle in C for the wgpu-sigops hot path -- TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may load this library
 * (oracle/_build/liboracle.so); the product (libsigops.so) never links, loads or calls it.
 *
 * It is the fast twin of oracle/sigops_oracle.py (same decision procedures) so that
 * 65,536- and 1,048,576-signature batches can be labelled in seconds, and it is the timed "port" CPU baseline.
 * What it restates (paths relative to the reference repository root):
 *   - oracle_ecrecover: `arkworks_recover` skeleton, src/curve_algos/secp256k1_ecdsa.rs:66-120 and
 *     src/curve_algos/secp256r1_ecdsa.rs:63-117, Fuel signature decoding src/tests/mod.rs:151-163 /
 *     src/wgsl/signature.wgsl:6-21, with the validity rules of the CPU libraries the reference's tests call
 *     (fuel-crypto 0.49.0 -> libsecp256k1 `secp256k1_ecdsa_recover`; p256 0.13.2 / ecdsa 0.16.9
 *     `recover_from_prehash`; call sites src/tests/secp256k1_ecdsa.rs:28-33, src/tests/secp256r1_ecdsa.rs:29-35).
 *     The scalar multiplication is an interleaved wNAF Strauss-Shamir pass in Jacobian coordinates -- the algorithm
 *     class libsecp256k1 itself uses (without its endomorphism) -- not the reference's WGSL double-and-add.
 *   - oracle_ed25519_verify: ed25519-dalek 2.1.1 `VerifyingKey::verify` (non-strict), restated in the reference at
 *     src/curve_algos/ed25519_eddsa.rs:49-184 (decompress via sqrt_ratio_i, SHA-512(R||A||M) mod L,
 *     [s]B + [k](-A), compress, byte compare), asserted at src/tests/ed25519_eddsa.rs:26.
 *   - oracle_gen_*: deterministic synthetic signatures (the role of `gen_test_data`,
 *     src/benchmarks/secp256k1_ecdsa.rs:149-172): random keys, 32-byte messages, valid signatures.
 *
 * PARITY PINNING: this file is pinned to oracle/sigops_oracle.py on every edge class and on random inputs
 * (tests/test_oracle.py), and the Python oracle is pinned to the reference's golden vectors, RFC 8032 vectors and
 * OpenSSL (see its header).  Rejecting outcomes are "parity unpinned" with respect to the reference itself (no Rust
 * toolchain in this image; the reference never tests them); they are cross-checked against OpenSSL (Ed25519 verify,
 * ECDSA verify of every recovered key) and libsodium (strict mode) on the whole edge corpus, through this file too.
 *
 * Arithmetic: 4 x 64-bit limbs, generic Montgomery multiplication (unsigned __int128) for all six moduli.
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

typedef struct {
    fe m;    /* modulus */
    u64 n0;  /* -m^-1 mod 2^64 */
    fe r2;   /* R^2 mod m */
    fe one;  /* R mod m */
    fe mm2;  /* m - 2 */
} mctx;

/* ------------------------------------------------------------------------------------------------ bigint */
static int fe_is_zero(const fe* a) { return (a->v[0] | a->v[1] | a->v[2] | a->v[3]) == 0; }
static int fe_eq(const fe* a, const fe* b) {
    return ((a->v[0] ^ b->v[0]) | (a->v[1] ^ b->v[1]) | (a->v[2] ^ b->v[2]) | (a->v[3] ^ b->v[3])) == 0;
}
static int fe_gte(const fe* a, const fe* b) {
    for (int i = 3; i >= 0; i--) {
        if (a->v[i] > b->v[i]) return 1;
        if (a->v[i] < b->v[i]) return 0;
    }
    return 1;
}
static u64 fe_add_raw(fe* r, const fe* a, const fe* b) {
    u128 c = 0;
    for (int i = 0; i < 4; i++) {
        c += (u128)a->v[i] + b->v[i];
        r->v[i] = (u64)c;
        c >>= 64;
    }
    return (u64)c;
}
static u64 fe_sub_raw(fe* r, const fe* a, const fe* b) {
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
static void fe_from_le(fe* r, const uint8_t* b) {
    for (int i = 0; i < 4; i++) {
        u64 w = 0;
        for (int j = 7; j >= 0; j--) w = (w << 8) | b[8 * i + j];
        r->v[i] = w;
    }
}
static void fe_to_le(uint8_t* b, const fe* a) {
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 8; j++) b[8 * i + j] = (uint8_t)(a->v[i] >> (8 * j));
}

/* ------------------------------------------------------------------------------------- modular arithmetic */
static void m_add(const mctx* c, fe* r, const fe* a, const fe* b) {
    fe t, u;
    u64 cy = fe_add_raw(&t, a, b);
    u64 bw = fe_sub_raw(&u, &t, &c->m);
    *r = (cy || !bw) ? u : t;
}
static void m_sub(const mctx* c, fe* r, const fe* a, const fe* b) {
    fe t, u;
    u64 bw = fe_sub_raw(&t, a, b);
    fe_add_raw(&u, &t, &c->m);
    *r = bw ? u : t;
}
static void m_neg(const mctx* c, fe* r, const fe* a) {
    if (fe_is_zero(a))
        *r = *a;
    else
        fe_sub_raw(r, &c->m, a);
}
/* Montgomery product a*b/2^256 mod m (CIOS, rows unrolled); needs a*b < m*2^256; result in [0, m) */
static void m_mul(const mctx* c, fe* r, const fe* a, const fe* b) {
    const u64 a0 = a->v[0], a1 = a->v[1], a2 = a->v[2], a3 = a->v[3];
    const u64 m0 = c->m.v[0], m1 = c->m.v[1], m2 = c->m.v[2], m3 = c->m.v[3];
    const u64 bv[4] = {b->v[0], b->v[1], b->v[2], b->v[3]};
    u64 t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0;
    for (int i = 0; i < 4; i++) {
        const u64 bi = bv[i];
        u128 x;
        u64 cy, t5;
        x = (u128)a0 * bi + t0;
        t0 = (u64)x;
        cy = (u64)(x >> 64);
        x = (u128)a1 * bi + t1 + cy;
        t1 = (u64)x;
        cy = (u64)(x >> 64);
        x = (u128)a2 * bi + t2 + cy;
        t2 = (u64)x;
        cy = (u64)(x >> 64);
        x = (u128)a3 * bi + t3 + cy;
        t3 = (u64)x;
        cy = (u64)(x >> 64);
        x = (u128)t4 + cy;
        t4 = (u64)x;
        t5 = (u64)(x >> 64);
        const u64 q = t0 * c->n0;
        x = (u128)q * m0 + t0;
        cy = (u64)(x >> 64);
        x = (u128)q * m1 + t1 + cy;
        t0 = (u64)x;
        cy = (u64)(x >> 64);
        x = (u128)q * m2 + t2 + cy;
        t1 = (u64)x;
        cy = (u64)(x >> 64);
        x = (u128)q * m3 + t3 + cy;
        t2 = (u64)x;
        cy = (u64)(x >> 64);
        x = (u128)t4 + cy;
        t3 = (u64)x;
        t4 = t5 + (u64)(x >> 64);
    }
    fe s = {{t0, t1, t2, t3}}, u;
    u64 bw = fe_sub_raw(&u, &s, &c->m);
    *r = (t4 || !bw) ? u : s;
}
static void m_sqr(const mctx* c, fe* r, const fe* a) { m_mul(c, r, a, a); }
static void m_to(const mctx* c, fe* r, const fe* a) { m_mul(c, r, a, &c->r2); }
static void m_from(const mctx* c, fe* r, const fe* a) {
    fe one = {{1, 0, 0, 0}};
    m_mul(c, r, a, &one);
}
/* r = a^e (Montgomery domain), fixed 4-bit windows */
static void m_pow(const mctx* c, fe* r, const fe* a, const fe* e) {
    fe tab[16];
    tab[0] = c->one;
    tab[1] = *a;
    for (int i = 2; i < 16; i++) m_mul(c, &tab[i], &tab[i - 1], a);
    fe acc = c->one;
    for (int w = 63; w >= 0; w--) {
        m_sqr(c, &acc, &acc);
        m_sqr(c, &acc, &acc);
        m_sqr(c, &acc, &acc);
        m_sqr(c, &acc, &acc);
        unsigned d = (unsigned)(e->v[w >> 4] >> ((w & 15) * 4)) & 15u;
        if (d) m_mul(c, &acc, &acc, &tab[d]);
    }
    *r = acc;
}
static void m_inv(const mctx* c, fe* r, const fe* a) { m_pow(c, r, a, &c->mm2); }

static void mctx_init(mctx* c, const fe* m) {
    c->m = *m;
    u64 x = 1; /* Newton: x = m^-1 mod 2^64 */
    for (int i = 0; i < 7; i++) x *= 2 - m->v[0] * x;
    c->n0 = (u64)0 - x;
    /* R mod m by 256 modular doublings of 1; R^2 by 256 more */
    fe t = {{1, 0, 0, 0}};
    for (int i = 0; i < 512; i++) {
        fe d, u;
        u64 cy = fe_add_raw(&d, &t, &t);
        u64 bw = fe_sub_raw(&u, &d, m);
        t = (cy || !bw) ? u : d;
        if (i == 255) c->one = t;
    }
    c->r2 = t;
    fe two = {{2, 0, 0, 0}};
    fe_sub_raw(&c->mm2, m, &two);
}

/* ------------------------------------------------------------------------------------------- parameters */
typedef struct {
    fe X, Y, Z; /* Jacobian, Montgomery domain */
    int inf;
} jac;
typedef struct {
    fe x, y;
} aff;

#define GW 7                      /* wNAF width for the generator */
#define GTAB (1 << (GW - 2))      /* odd multiples 1,3,..,2^(GW-1)-1 */
#define PW 5                      /* wNAF width for the per-signature point */
#define PTAB (1 << (PW - 2))

typedef struct {
    mctx fp, fn;
    fe b;          /* Montgomery */
    int a_minus3;  /* a = -3 (P-256) or a = 0 (secp256k1) */
    fe sqrt_exp;   /* (p+1)/4 */
    aff g;
    aff gtab[GTAB];
} swcurve;

static swcurve K1, R1;

typedef struct {
    fe X, Y, Z, T;
} edp;
static struct {
    mctx fp, fl;
    fe d, d2, sqrtm1, p58; /* Montgomery, exponent plain */
    edp B;
    edp btab[GTAB];
    fe L;
} ED;

static pthread_once_t g_once = PTHREAD_ONCE_INIT;

static void fe_from_hex(fe* r, const char* h) {
    uint8_t b[32];
    for (int i = 0; i < 32; i++) {
        unsigned v = 0;
        for (int j = 0; j < 2; j++) {
            char ch = h[2 * i + j];
            v = (v << 4) | (unsigned)(ch <= '9' ? ch - '0' : (ch | 32) - 'a' + 10);
        }
        b[i] = (uint8_t)v;
    }
    fe_from_be(r, b);
}

/* ------------------------------------------------------------------------------- short Weierstrass group */
static void jac_dbl(const swcurve* c, jac* P) {
    const mctx* f = &c->fp;
    if (P->inf) return;
    fe A, B, C, D, E, F, t;
    if (!c->a_minus3) { /* dbl-2009-l */
        m_sqr(f, &A, &P->X);
        m_sqr(f, &B, &P->Y);
        m_sqr(f, &C, &B);
        m_add(f, &t, &P->X, &B);
        m_sqr(f, &t, &t);
        m_sub(f, &t, &t, &A);
        m_sub(f, &t, &t, &C);
        m_add(f, &D, &t, &t);
        m_add(f, &E, &A, &A);
        m_add(f, &E, &E, &A);
        m_sqr(f, &F, &E);
        m_mul(f, &P->Z, &P->Y, &P->Z);
        m_add(f, &P->Z, &P->Z, &P->Z);
        m_add(f, &t, &D, &D);
        m_sub(f, &P->X, &F, &t);
        m_sub(f, &t, &D, &P->X);
        m_mul(f, &t, &E, &t);
        m_add(f, &C, &C, &C);
        m_add(f, &C, &C, &C);
        m_add(f, &C, &C, &C);
        m_sub(f, &P->Y, &t, &C);
    } else { /* dbl-2001-b */
        fe delta, gamma, beta, alpha, u;
        m_sqr(f, &delta, &P->Z);
        m_sqr(f, &gamma, &P->Y);
        m_mul(f, &beta, &P->X, &gamma);
        m_sub(f, &t, &P->X, &delta);
        m_add(f, &u, &P->X, &delta);
        m_mul(f, &alpha, &t, &u);
        m_add(f, &t, &alpha, &alpha);
        m_add(f, &alpha, &alpha, &t);
        m_add(f, &t, &P->Y, &P->Z);
        m_sqr(f, &t, &t);
        m_sub(f, &t, &t, &gamma);
        m_sub(f, &P->Z, &t, &delta);
        m_sqr(f, &t, &alpha);
        m_add(f, &beta, &beta, &beta);
        m_add(f, &beta, &beta, &beta);
        m_add(f, &u, &beta, &beta);
        m_sub(f, &P->X, &t, &u);
        m_sub(f, &t, &beta, &P->X);
        m_mul(f, &t, &alpha, &t);
        m_sqr(f, &gamma, &gamma);
        m_add(f, &gamma, &gamma, &gamma);
        m_add(f, &gamma, &gamma, &gamma);
        m_add(f, &gamma, &gamma, &gamma);
        m_sub(f, &P->Y, &t, &gamma);
    }
}

/* P += Q, both Jacobian; handles infinity, P == Q, P == -Q */
static void jac_add(const swcurve* c, jac* P, const jac* Q) {
    const mctx* f = &c->fp;
    if (Q->inf) return;
    if (P->inf) {
        *P = *Q;
        return;
    }
    fe Z1Z1, Z2Z2, U1, U2, S1, S2, H, r, HH, HHH, V, t;
    m_sqr(f, &Z1Z1, &P->Z);
    m_sqr(f, &Z2Z2, &Q->Z);
    m_mul(f, &U1, &P->X, &Z2Z2);
    m_mul(f, &U2, &Q->X, &Z1Z1);
    m_mul(f, &S1, &Q->Z, &Z2Z2);
    m_mul(f, &S1, &P->Y, &S1);
    m_mul(f, &S2, &P->Z, &Z1Z1);
    m_mul(f, &S2, &Q->Y, &S2);
    m_sub(f, &H, &U2, &U1);
    m_sub(f, &r, &S2, &S1);
    if (fe_is_zero(&H)) {
        if (fe_is_zero(&r))
            jac_dbl(c, P);
        else
            P->inf = 1;
        return;
    }
    m_sqr(f, &HH, &H);
    m_mul(f, &HHH, &H, &HH);
    m_mul(f, &V, &U1, &HH);
    m_sqr(f, &t, &r);
    m_sub(f, &t, &t, &HHH);
    m_sub(f, &t, &t, &V);
    fe X3;
    m_sub(f, &X3, &t, &V);
    m_sub(f, &t, &V, &X3);
    m_mul(f, &t, &r, &t);
    m_mul(f, &HHH, &S1, &HHH);
    m_sub(f, &P->Y, &t, &HHH);
    m_mul(f, &t, &P->Z, &Q->Z);
    m_mul(f, &P->Z, &t, &H);
    P->X = X3;
}
static void jac_from_aff(const swcurve* c, jac* P, const aff* a, int negate) {
    P->X = a->x;
    if (negate)
        m_neg(&c->fp, &P->Y, &a->y);
    else
        P->Y = a->y;
    P->Z = c->fp.one;
    P->inf = 0;
}
static void jac_to_aff(const swcurve* c, aff* a, const jac* P) {
    const mctx* f = &c->fp;
    fe zi, zi2;
    m_inv(f, &zi, &P->Z);
    m_sqr(f, &zi2, &zi);
    m_mul(f, &a->x, &P->X, &zi2);
    m_mul(f, &zi2, &zi2, &zi);
    m_mul(f, &a->y, &P->Y, &zi2);
}

/* width-w non-adjacent form of a 256-bit scalar: digits odd in (-2^(w-1), 2^(w-1)), naf[i] for bit i, 257 entries */
static void wnaf(int8_t* naf, const fe* k, int w) {
    u64 t[5] = {k->v[0], k->v[1], k->v[2], k->v[3], 0};
    memset(naf, 0, 257);
    for (int i = 0; i < 257; i++) {
        if (t[0] & 1) {
            int d = (int)(t[0] & ((1u << w) - 1));
            if (d >= (1 << (w - 1))) d -= 1 << w;
            naf[i] = (int8_t)d;
            /* t -= d */
            if (d >= 0) {
                u64 bw = (u64)d;
                for (int j = 0; j < 5 && bw; j++) {
                    u64 o = t[j];
                    t[j] -= bw;
                    bw = o < bw;
                }
            } else {
                u64 cy = (u64)(-d);
                for (int j = 0; j < 5 && cy; j++) {
                    t[j] += cy;
                    cy = t[j] < cy;
                }
            }
        }
        t[0] = (t[0] >> 1) | (t[1] << 63);
        t[1] = (t[1] >> 1) | (t[2] << 63);
        t[2] = (t[2] >> 1) | (t[3] << 63);
        t[3] = (t[3] >> 1) | (t[4] << 63);
        t[4] >>= 1;
    }
}

/* Q = u1*G + u2*R (R affine; may be NULL with u2 ignored) */
static void sw_double_mul(const swcurve* c, jac* Q, const fe* u1, const fe* u2, const aff* R) {
    int8_t n1[257], n2[257];
    jac ptab[PTAB];
    wnaf(n1, u1, GW);
    if (R) {
        wnaf(n2, u2, PW);
        jac R2;
        jac_from_aff(c, &ptab[0], R, 0);
        R2 = ptab[0];
        jac_dbl(c, &R2);
        for (int i = 1; i < PTAB; i++) {
            ptab[i] = ptab[i - 1];
            jac_add(c, &ptab[i], &R2);
        }
    } else {
        memset(n2, 0, sizeof n2);
    }
    Q->inf = 1;
    memset(&Q->X, 0, sizeof(fe) * 3);
    for (int i = 256; i >= 0; i--) {
        jac_dbl(c, Q);
        if (n1[i]) {
            jac t;
            int d = n1[i];
            jac_from_aff(c, &t, &c->gtab[(d < 0 ? -d : d) >> 1], d < 0);
            jac_add(c, Q, &t);
        }
        if (n2[i]) {
            int d = n2[i];
            jac t = ptab[(d < 0 ? -d : d) >> 1];
            if (d < 0) m_neg(&c->fp, &t.Y, &t.Y);
            jac_add(c, Q, &t);
        }
    }
}

static void sw_init(swcurve* c, const char* p, const char* n, const char* b, const char* gx, const char* gy, int am3) {
    fe t;
    fe_from_hex(&t, p);
    mctx_init(&c->fp, &t);
    fe_from_hex(&t, n);
    mctx_init(&c->fn, &t);
    c->a_minus3 = am3;
    fe_from_hex(&t, b);
    m_to(&c->fp, &c->b, &t);
    /* (p+1)/4 */
    fe one = {{1, 0, 0, 0}}, e;
    fe_add_raw(&e, &c->fp.m, &one); /* p+1 < 2^256 for both primes */
    for (int i = 0; i < 4; i++) c->sqrt_exp.v[i] = (e.v[i] >> 2) | (i < 3 ? e.v[i + 1] << 62 : 0);
    fe_from_hex(&t, gx);
    m_to(&c->fp, &c->g.x, &t);
    fe_from_hex(&t, gy);
    m_to(&c->fp, &c->g.y, &t);
    jac P, G2;
    jac_from_aff(c, &P, &c->g, 0);
    G2 = P;
    jac_dbl(c, &G2);
    for (int i = 0; i < GTAB; i++) {
        jac_to_aff(c, &c->gtab[i], &P);
        jac_add(c, &P, &G2);
    }
}

/* SURVEY.md appendix B.  Returns 0 and writes X||Y, or 1 (invalid) and writes 64 zero bytes. */
static int sw_ecrecover_one(const swcurve* c, const uint8_t* sig, const uint8_t* msg, uint8_t* out) {
    const mctx *fp = &c->fp, *fn = &c->fn;
    memset(out, 0, 64);
    uint8_t sb[32];
    memcpy(sb, sig + 32, 32);
    int parity = sb[0] >> 7;
    sb[0] &= 0x7f;
    fe r, s, z;
    fe_from_be(&r, sig);
    fe_from_be(&s, sb);
    fe_from_be(&z, msg);
    if (fe_is_zero(&r) || fe_is_zero(&s) || fe_gte(&r, &fn->m) || fe_gte(&s, &fn->m)) return 1;
    if (fe_gte(&z, &fn->m)) fe_sub_raw(&z, &z, &fn->m); /* z < 2^256 < 2n */
    /* lift x = r (r < n < p) */
    fe x, t, y, y2;
    m_to(fp, &x, &r);
    m_sqr(fp, &t, &x);
    m_mul(fp, &t, &t, &x);
    if (c->a_minus3) {
        fe x3;
        m_add(fp, &x3, &x, &x);
        m_add(fp, &x3, &x3, &x);
        m_sub(fp, &t, &t, &x3);
    }
    m_add(fp, &t, &t, &c->b);
    m_pow(fp, &y, &t, &c->sqrt_exp);
    m_sqr(fp, &y2, &y);
    if (!fe_eq(&y2, &t)) return 1;
    fe yp;
    m_from(fp, &yp, &y);
    if ((int)(yp.v[0] & 1) != parity) m_neg(fp, &y, &y);
    /* u1 = -z/r, u2 = s/r mod n */
    fe rm, rinv, u1, u2;
    m_to(fn, &rm, &r);
    m_inv(fn, &rinv, &rm);       /* r^-1 * R */
    m_mul(fn, &u2, &rinv, &s);   /* plain */
    m_mul(fn, &u1, &rinv, &z);
    m_neg(fn, &u1, &u1);
    aff R = {x, y};
    jac Q;
    sw_double_mul(c, &Q, &u1, &u2, &R);
    if (Q.inf) return 1;
    aff A;
    jac_to_aff(c, &A, &Q);
    fe ax, ay;
    m_from(fp, &ax, &A.x);
    m_from(fp, &ay, &A.y);
    fe_to_be(out, &ax);
    fe_to_be(out + 32, &ay);
    return 0;
}

/* ------------------------------------------------------------------------------------------------ SHA-512 */
static const u64 SHA512_K[80] = {
    0x428a2f98d728ae22ull, 0x7137449123ef65cdull, 0xb5c0fbcfec4d3b2full, 0xe9b5dba58189dbbcull, 0x3956c25bf348b538ull,
    0x59f111f1b605d019ull, 0x923f82a4af194f9bull, 0xab1c5ed5da6d8118ull, 0xd807aa98a3030242ull, 0x12835b0145706fbeull,
    0x243185be4ee4b28cull, 0x550c7dc3d5ffb4e2ull, 0x72be5d74f27b896full, 0x80deb1fe3b1696b1ull, 0x9bdc06a725c71235ull,
    0xc19bf174cf692694ull, 0xe49b69c19ef14ad2ull, 0xefbe4786384f25e3ull, 0x0fc19dc68b8cd5b5ull, 0x240ca1cc77ac9c65ull,
    0x2de92c6f592b0275ull, 0x4a7484aa6ea6e483ull, 0x5cb0a9dcbd41fbd4ull, 0x76f988da831153b5ull, 0x983e5152ee66dfabull,
    0xa831c66d2db43210ull, 0xb00327c898fb213full, 0xbf597fc7beef0ee4ull, 0xc6e00bf33da88fc2ull, 0xd5a79147930aa725ull,
    0x06ca6351e003826full, 0x142929670a0e6e70ull, 0x27b70a8546d22ffcull, 0x2e1b21385c26c926ull, 0x4d2c6dfc5ac42aedull,
    0x53380d139d95b3dfull, 0x650a73548baf63deull, 0x766a0abb3c77b2a8ull, 0x81c2c92e47edaee6ull, 0x92722c851482353bull,
    0xa2bfe8a14cf10364ull, 0xa81a664bbc423001ull, 0xc24b8b70d0f89791ull, 0xc76c51a30654be30ull, 0xd192e819d6ef5218ull,
    0xd69906245565a910ull, 0xf40e35855771202aull, 0x106aa07032bbd1b8ull, 0x19a4c116b8d2d0c8ull, 0x1e376c085141ab53ull,
    0x2748774cdf8eeb99ull, 0x34b0bcb5e19b48a8ull, 0x391c0cb3c5c95a63ull, 0x4ed8aa4ae3418acbull, 0x5b9cca4f7763e373ull,
    0x682e6ff3d6b2b8a3ull, 0x748f82ee5defb2fcull, 0x78a5636f43172f60ull, 0x84c87814a1f0ab72ull, 0x8cc702081a6439ecull,
    0x90befffa23631e28ull, 0xa4506cebde82bde9ull, 0xbef9a3f7b2c67915ull, 0xc67178f2e372532bull, 0xca273eceea26619cull,
    0xd186b8c721c0c207ull, 0xeada7dd6cde0eb1eull, 0xf57d4f7fee6ed178ull, 0x06f067aa72176fbaull, 0x0a637dc5a2c898a6ull,
    0x113f9804bef90daeull, 0x1b710b35131c471bull, 0x28db77f523047d84ull, 0x32caab7b40c72493ull, 0x3c9ebe0a15c9bebcull,
    0x431d67c49c100d4cull, 0x4cc5d4becb3e42b6ull, 0x597f299cfc657e2aull, 0x5fcb6fab3ad6faecull, 0x6c44198c4a475817ull};

#define ROR(x, n) (((x) >> (n)) | ((x) << (64 - (n))))
static void sha512_block(u64* h, const uint8_t* blk) {
    u64 w[80];
    for (int i = 0; i < 16; i++) {
        u64 x = 0;
        for (int j = 0; j < 8; j++) x = (x << 8) | blk[8 * i + j];
        w[i] = x;
    }
    for (int i = 16; i < 80; i++) {
        u64 s0 = ROR(w[i - 15], 1) ^ ROR(w[i - 15], 8) ^ (w[i - 15] >> 7);
        u64 s1 = ROR(w[i - 2], 19) ^ ROR(w[i - 2], 61) ^ (w[i - 2] >> 6);
        w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    u64 a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
    for (int i = 0; i < 80; i++) {
        u64 S1 = ROR(e, 14) ^ ROR(e, 18) ^ ROR(e, 41);
        u64 ch = (e & f) ^ (~e & g);
        u64 t1 = hh + S1 + ch + SHA512_K[i] + w[i];
        u64 S0 = ROR(a, 28) ^ ROR(a, 34) ^ ROR(a, 39);
        u64 mj = (a & b) ^ (a & c) ^ (b & c);
        u64 t2 = S0 + mj;
        hh = g;
        g = f;
        f = e;
        e = d + t1;
        d = c;
        c = b;
        b = a;
        a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}
/* SHA-512 of three concatenated pieces of any length (streamed block by block) */
static void sha512_3(uint8_t* out, const uint8_t* a, size_t la, const uint8_t* b, size_t lb, const uint8_t* c, size_t lc) {
    u64 h[8] = {0x6a09e667f3bcc908ull, 0xbb67ae8584caa73bull, 0x3c6ef372fe94f82bull, 0xa54ff53a5f1d36f1ull,
                0x510e527fade682d1ull, 0x9b05688c2b3e6c1full, 0x1f83d9abfb41bd6bull, 0x5be0cd19137e2179ull};
    const uint8_t* piece[3] = {a, b, c};
    const size_t plen[3] = {la, lb, lc};
    uint8_t buf[128];
    size_t fill = 0, len = la + lb + lc;
    for (int p = 0; p < 3; p++) {
        for (size_t i = 0; i < plen[p]; i++) {
            buf[fill++] = piece[p][i];
            if (fill == 128) {
                sha512_block(h, buf);
                fill = 0;
            }
        }
    }
    buf[fill++] = 0x80;
    if (fill > 112) {
        memset(buf + fill, 0, 128 - fill);
        sha512_block(h, buf);
        fill = 0;
    }
    memset(buf + fill, 0, 128 - fill);
    u64 bits = (u64)len * 8;
    for (int j = 0; j < 8; j++) buf[127 - j] = (uint8_t)(bits >> (8 * j));
    sha512_block(h, buf);
    for (int i = 0; i < 8; i++)
        for (int j = 0; j < 8; j++) out[8 * i + j] = (uint8_t)(h[i] >> (56 - 8 * j));
}

/* --------------------------------------------------------------------------------------------- ed25519 */
/* 64 little-endian bytes -> integer mod L (plain) */
static void ed_reduce512(fe* k, const uint8_t* h64) {
    const mctx* f = &ED.fl;
    fe lo, hi;
    fe_from_le(&lo, h64);
    fe_from_le(&hi, h64 + 32);
    m_mul(f, &hi, &hi, &f->r2); /* hi * R mod L, canonical */
    m_mul(f, &lo, &lo, &f->r2); /* lo * R */
    m_from(f, &lo, &lo);        /* lo mod L */
    m_add(f, k, &lo, &hi);
}
static void ed_identity(edp* P) {
    memset(P, 0, sizeof *P);
    P->Y = ED.fp.one;
    P->Z = ED.fp.one;
}
/* add-2008-hwcd-3 (a = -1), complete */
static void ed_add(edp* P, const edp* Q, int negq) {
    const mctx* f = &ED.fp;
    fe A, B, C, D, E, F, G, H, t, u, qx = Q->X, qt = Q->T;
    if (negq) {
        m_neg(f, &qx, &qx);
        m_neg(f, &qt, &qt);
    }
    m_sub(f, &t, &P->Y, &P->X);
    m_sub(f, &u, &Q->Y, &qx);
    m_mul(f, &A, &t, &u);
    m_add(f, &t, &P->Y, &P->X);
    m_add(f, &u, &Q->Y, &qx);
    m_mul(f, &B, &t, &u);
    m_mul(f, &C, &P->T, &ED.d2);
    m_mul(f, &C, &C, &qt);
    m_mul(f, &D, &P->Z, &Q->Z);
    m_add(f, &D, &D, &D);
    m_sub(f, &E, &B, &A);
    m_sub(f, &F, &D, &C);
    m_add(f, &G, &D, &C);
    m_add(f, &H, &B, &A);
    m_mul(f, &P->X, &E, &F);
    m_mul(f, &P->Y, &G, &H);
    m_mul(f, &P->Z, &F, &G);
    m_mul(f, &P->T, &E, &H);
}
/* dbl-2008-hwcd */
static void ed_dbl(edp* P) {
    const mctx* f = &ED.fp;
    fe A, B, C, E, G, F, H, t;
    m_sqr(f, &A, &P->X);
    m_sqr(f, &B, &P->Y);
    m_sqr(f, &C, &P->Z);
    m_add(f, &C, &C, &C);
    m_add(f, &t, &P->X, &P->Y);
    m_sqr(f, &t, &t);
    m_sub(f, &t, &t, &A);
    m_sub(f, &E, &t, &B);
    m_sub(f, &G, &B, &A);
    m_sub(f, &F, &G, &C);
    m_add(f, &H, &A, &B);
    m_neg(f, &H, &H);
    m_mul(f, &P->X, &E, &F);
    m_mul(f, &P->Y, &G, &H);
    m_mul(f, &P->Z, &F, &G);
    m_mul(f, &P->T, &E, &H);
}
/* Q = a*B + b*P (P may be NULL), optionally with P negated */
static void ed_double_mul(edp* Q, const fe* a, const fe* b, const edp* P, int negp) {
    int8_t n1[257], n2[257];
    edp ptab[PTAB];
    wnaf(n1, a, GW);
    if (P) {
        wnaf(n2, b, PW);
        ptab[0] = *P;
        edp P2 = *P;
        ed_dbl(&P2);
        for (int i = 1; i < PTAB; i++) {
            ptab[i] = ptab[i - 1];
            ed_add(&ptab[i], &P2, 0);
        }
    } else {
        memset(n2, 0, sizeof n2);
    }
    ed_identity(Q);
    for (int i = 256; i >= 0; i--) {
        ed_dbl(Q);
        if (n1[i]) {
            int d = n1[i];
            ed_add(Q, &ED.btab[(d < 0 ? -d : d) >> 1], d < 0);
        }
        if (n2[i]) {
            int d = n2[i];
            ed_add(Q, &ptab[(d < 0 ? -d : d) >> 1], (d < 0) != (negp != 0));
        }
    }
}
static void ed_compress(uint8_t* out, const edp* P) {
    const mctx* f = &ED.fp;
    fe zi, x, y;
    m_inv(f, &zi, &P->Z);
    m_mul(f, &x, &P->X, &zi);
    m_mul(f, &y, &P->Y, &zi);
    m_from(f, &x, &x);
    m_from(f, &y, &y);
    fe_to_le(out, &y);
    out[31] |= (uint8_t)((x.v[0] & 1) << 7);
}
/* sqrt_ratio_i (curve25519-dalek; src/curve_algos/ed25519_eddsa.rs:160-184); Montgomery in / out */
static int ed_sqrt_ratio_i(fe* r, const fe* u, const fe* v) {
    const mctx* f = &ED.fp;
    fe v3, v7, t, check, nu, nui;
    m_sqr(f, &v3, v);
    m_mul(f, &v3, &v3, v);
    m_sqr(f, &v7, &v3);
    m_mul(f, &v7, &v7, v);
    m_mul(f, &t, u, &v7);
    m_pow(f, &t, &t, &ED.p58);
    m_mul(f, r, u, &v3);
    m_mul(f, r, r, &t);
    m_sqr(f, &check, r);
    m_mul(f, &check, &check, v);
    m_neg(f, &nu, u);
    m_mul(f, &nui, &nu, &ED.sqrtm1);
    int correct = fe_eq(&check, u), flipped = fe_eq(&check, &nu), flipped_i = fe_eq(&check, &nui);
    if (flipped || flipped_i) m_mul(f, r, r, &ED.sqrtm1);
    fe plain;
    m_from(f, &plain, r);
    if (plain.v[0] & 1) m_neg(f, r, r);
    return correct || flipped;
}
/* CompressedEdwardsY::decompress without canonicity check on y; returns 0 when not on the curve */
static int ed_decompress(edp* P, const uint8_t* b) {
    const mctx* f = &ED.fp;
    int sign = b[31] >> 7;
    fe y, yy, u, v, x;
    fe_from_le(&y, b);
    y.v[3] &= 0x7fffffffffffffffull;
    if (fe_gte(&y, &f->m)) fe_sub_raw(&y, &y, &f->m); /* y < 2^255 < 2p */
    m_to(f, &y, &y);
    m_sqr(f, &yy, &y);
    m_sub(f, &u, &yy, &f->one);
    m_mul(f, &v, &yy, &ED.d);
    m_add(f, &v, &v, &f->one);
    if (!ed_sqrt_ratio_i(&x, &u, &v)) return 0;
    if (sign) m_neg(f, &x, &x);
    P->X = x;
    P->Y = y;
    P->Z = f->one;
    m_mul(f, &P->T, &x, &y);
    return 1;
}
/* [8]P == identity (dalek is_small_order) */
static int ed_is_small_order(const edp* P) {
    edp Q = *P;
    ed_dbl(&Q);
    ed_dbl(&Q);
    ed_dbl(&Q);
    return fe_is_zero(&Q.X) && fe_eq(&Q.Y, &Q.Z);
}
/* dalek `verify` (strict == 0) or `verify_strict` (strict != 0) over a message of any length */
static int ed_verify_msg(const uint8_t* sig, const uint8_t* msg, size_t len, const uint8_t* pk, int strict) {
    edp A, Rp;
    if (!ed_decompress(&A, pk)) return 0;
    fe s, k;
    fe_from_le(&s, sig + 32);
    if (fe_gte(&s, &ED.L)) return 0;
    if (strict) {
        edp Rs;
        if (!ed_decompress(&Rs, sig)) return 0;
        if (ed_is_small_order(&Rs) || ed_is_small_order(&A)) return 0;
    }
    uint8_t h[64], enc[32];
    sha512_3(h, sig, 32, pk, 32, msg, len);
    ed_reduce512(&k, h);
    ed_double_mul(&Rp, &s, &k, &A, 1);
    ed_compress(enc, &Rp);
    return memcmp(enc, sig, 32) == 0;
}
static int ed_verify_one(const uint8_t* sig, const uint8_t* msg, const uint8_t* pk) {
    edp A, Rp;
    if (!ed_decompress(&A, pk)) return 0;
    fe s, k;
    fe_from_le(&s, sig + 32);
    if (fe_gte(&s, &ED.L)) return 0;
    uint8_t h[64], enc[32];
    sha512_3(h, sig, 32, pk, 32, msg, 32);
    ed_reduce512(&k, h);
    ed_double_mul(&Rp, &s, &k, &A, 1);
    ed_compress(enc, &Rp);
    return memcmp(enc, sig, 32) == 0;
}

static void ed_init(void) {
    fe t;
    fe_from_hex(&t, "7fffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffed");
    mctx_init(&ED.fp, &t);
    fe_from_hex(&ED.L, "1000000000000000000000000000000014def9dea2f79cd65812631a5cf5d3ed");
    mctx_init(&ED.fl, &ED.L);
    const mctx* f = &ED.fp;
    fe_from_hex(&t, "52036cee2b6ffe738cc740797779e89800700a4d4141d8ab75eb4dca135978a3");
    m_to(f, &ED.d, &t);
    m_add(f, &ED.d2, &ED.d, &ED.d);
    fe_from_hex(&t, "2b8324804fc1df0b2b4d00993dfbd7a72f431806ad2fe478c4ee1b274a0ea0b0");
    m_to(f, &ED.sqrtm1, &t);
    fe_from_hex(&ED.p58, "0ffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffd");
    fe bx, by;
    fe_from_hex(&bx, "216936d3cd6e53fec0a4e231fdd6dc5c692cc7609525a7b2c9562d608f25d51a");
    fe_from_hex(&by, "6666666666666666666666666666666666666666666666666666666666666658");
    m_to(f, &ED.B.X, &bx);
    m_to(f, &ED.B.Y, &by);
    ED.B.Z = f->one;
    m_mul(f, &ED.B.T, &ED.B.X, &ED.B.Y);
    edp P = ED.B, B2 = ED.B;
    ed_dbl(&B2);
    for (int i = 0; i < GTAB; i++) {
        ED.btab[i] = P;
        ed_add(&P, &B2, 0);
    }
}

static void init_all(void) {
    sw_init(&K1, "fffffffffffffffffffffffffffffffffffffffffffffffffffffffefffffc2f",
            "fffffffffffffffffffffffffffffffebaaedce6af48a03bbfd25e8cd0364141",
            "0000000000000000000000000000000000000000000000000000000000000007",
            "79be667ef9dcbbac55a06295ce870b07029bfcdb2dce28d959f2815b16f81798",
            "483ada7726a3c4655da4fbfc0e1108a8fd17b448a68554199c47d08ffb10d4b8", 0);
    sw_init(&R1, "ffffffff00000001000000000000000000000000ffffffffffffffffffffffff",
            "ffffffff00000000ffffffffffffffffbce6faada7179e84f3b9cac2fc632551",
            "5ac635d8aa3a93e7b3ebbd55769886bc651d06b0cc53b0f63bce3c3e27d2604b",
            "6b17d1f2e12c4247f8bce6e563a440f277037d812deb33a0f4a13945d898c296",
            "4fe342e2fe1a7f9b8ee7eb4a7c0f9e162bce33576b315ececbb6406837bf51f5", 1);
    ed_init();
}

/* ---------------------------------------------------------------------------------------- synthetic data */
/* splitmix64 stream keyed by (seed, curve, index, lane): reproducible from any language */
static u64 splitmix(u64* s) {
    u64 z = (*s += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
static void prng32(uint8_t* out, u64 seed, u64 curve, u64 index, u64 lane) {
    u64 s = seed ^ (curve * 0xd1342543de82ef95ull) ^ (index * 0x2545f4914f6cdd1dull) ^ (lane * 0x9e3779b97f4a7c15ull);
    splitmix(&s);
    for (int i = 0; i < 4; i++) {
        u64 w = splitmix(&s);
        memcpy(out + 8 * i, &w, 8);
    }
}
/* scalar in [1, n-1] from 32 PRNG bytes */
static void prng_scalar(const mctx* fn, fe* k, u64 seed, u64 curve, u64 index, u64 lane) {
    uint8_t b[32];
    prng32(b, seed, curve, index, lane);
    fe_from_le(k, b);
    fe one = {{1, 0, 0, 0}}, nm1;
    fe_sub_raw(&nm1, &fn->m, &one);
    /* k mod (n-1) + 1: k < 2^256 < 2(n-1) for both curves' orders?  n > 2^255, so at most one subtraction */
    if (fe_gte(k, &nm1)) fe_sub_raw(k, k, &nm1);
    fe_add_raw(k, k, &one);
}

#define KEY_POOL 4096

/* Montgomery's trick: out[i] = v[i]^-1 for nonzero Montgomery-domain values (one m_inv per call) */
static void batch_inv(const mctx* c, fe* out, const fe* v, size_t n) {
    if (!n) return;
    out[0] = v[0];
    for (size_t i = 1; i < n; i++) m_mul(c, &out[i], &out[i - 1], &v[i]);
    fe inv;
    m_inv(c, &inv, &out[n - 1]);
    for (size_t i = n - 1; i > 0; i--) {
        fe t;
        m_mul(c, &t, &inv, &out[i - 1]);
        m_mul(c, &inv, &inv, &v[i]);
        out[i] = t;
    }
    out[0] = inv;
}

#define GEN_BLOCK 256

/* Signature i: key = pool[i % KEY_POOL], message = PRNG(i), nonce k_i = k_b + (i - b) where b is the first index of i's
 * block of GEN_BLOCK signatures and k_b = PRNG(b): R_i = R_(i-1) + G costs one point addition, and the affine
 * conversions and the k^-1 share one inversion per block (the scheme SURVEY.md 8d proposes for the 1M-row batches).
 * How a nonce was chosen is invisible to a verifier: these are ordinary valid signatures of random keys and messages.
 * s is low-s normalised when low_s (as `Signature::sign` does); the y parity of R goes to bit 7 of byte 32; pks
 * receives the signer's key X||Y. */
static void sw_gen_range(const swcurve* c, int curve_id, u64 seed, size_t lo, size_t hi, int low_s, uint8_t* sigs,
                         uint8_t* msgs, uint8_t* pks, const fe* pool_d, const uint8_t* pool_pk) {
    const mctx *fp = &c->fp, *fn = &c->fn;
    jac G;
    jac_from_aff(c, &G, &c->g, 0);
    fe one_n = {{1, 0, 0, 0}};
    for (size_t base = lo - lo % GEN_BLOCK; base < hi; base += GEN_BLOCK) {
        for (u64 attempt = 0;; attempt++) {
            fe k0;
            prng_scalar(fn, &k0, seed, (u64)curve_id, base, 2 + attempt);
            jac P;
            sw_double_mul(c, &P, &k0, NULL, NULL);
            fe zs[GEN_BLOCK], zinv[GEN_BLOCK], km[GEN_BLOCK], kinv[GEN_BLOCK];
            jac pts[GEN_BLOCK];
            int ok = 1;
            fe k = k0;
            for (int j = 0; j < GEN_BLOCK; j++) {
                if (P.inf || fe_is_zero(&k)) ok = 0;
                pts[j] = P;
                zs[j] = P.Z;
                m_to(fn, &km[j], &k);
                jac_add(c, &P, &G);
                m_add(fn, &k, &k, &one_n);
            }
            if (!ok) continue; /* the block would cross k = 0 mod n: probability ~2^-248 */
            batch_inv(fp, zinv, zs, GEN_BLOCK);
            batch_inv(fn, kinv, km, GEN_BLOCK);
            fe rxs[GEN_BLOCK];
            int par[GEN_BLOCK];
            for (int j = 0; j < GEN_BLOCK; j++) {
                fe zi2, ax, ay, ry;
                m_sqr(fp, &zi2, &zinv[j]);
                m_mul(fp, &ax, &pts[j].X, &zi2);
                m_mul(fp, &zi2, &zi2, &zinv[j]);
                m_mul(fp, &ay, &pts[j].Y, &zi2);
                m_from(fp, &rxs[j], &ax);
                m_from(fp, &ry, &ay);
                par[j] = (int)(ry.v[0] & 1);
                /* x >= n (probability 2^-128) is not encodable in the Fuel format (no "x reduced" recovery id) */
                if (fe_gte(&rxs[j], &fn->m) || fe_is_zero(&rxs[j])) ok = 0;
            }
            if (!ok) continue;
            for (int j = 0; j < GEN_BLOCK; j++) {
                size_t i = base + (size_t)j;
                if (i < lo || i >= hi) continue;
                size_t kidx = i % KEY_POOL;
                uint8_t* sig = sigs + 64 * i;
                uint8_t* msg = msgs + 32 * i;
                fe z, r = rxs[j], s, t, rm, dm, zm;
                for (u64 retry = 0;; retry++) {
                    prng32(msg, seed, (u64)curve_id, i, 1 + 64 * retry);
                    fe_from_be(&z, msg);
                    if (fe_gte(&z, &fn->m)) fe_sub_raw(&z, &z, &fn->m);
                    /* s = k^-1 (z + r d) mod n */
                    m_to(fn, &rm, &r);
                    m_to(fn, &dm, &pool_d[kidx]);
                    m_to(fn, &zm, &z);
                    m_mul(fn, &t, &rm, &dm);
                    m_add(fn, &t, &t, &zm);
                    m_mul(fn, &s, &t, &kinv[j]);
                    m_from(fn, &s, &s);
                    if (!fe_is_zero(&s)) break; /* s == 0 (probability 2^-256): another message */
                }
                int parity = par[j];
                fe ns;
                fe_sub_raw(&ns, &fn->m, &s);
                int high = fe_gte(&s, &ns); /* s > n/2 (n odd so s != n - s) */
                if (low_s && high) {
                    s = ns;
                    parity ^= 1;
                }
                if (s.v[3] >> 63) { /* only s < 2^255 is encodable: flip to the other representative */
                    fe_sub_raw(&s, &fn->m, &s);
                    parity ^= 1;
                }
                fe_to_be(sig, &r);
                fe_to_be(sig + 32, &s);
                sig[32] |= (uint8_t)(parity << 7);
                if (pks) memcpy(pks + 64 * i, pool_pk + 64 * kidx, 64);
            }
            break;
        }
    }
}

/* ed25519: same block scheme with r_i = r_b + (i - b) (any r yields a valid signature R = [r]B, s = r + k a). */
static void ed_gen_range(u64 seed, size_t lo, size_t hi, uint8_t* sigs, uint8_t* msgs, uint8_t* pks, const fe* pool_a,
                         const uint8_t* pool_prefix, const uint8_t* pool_pk) {
    const mctx *fl = &ED.fl, *fp = &ED.fp;
    (void)pool_prefix;
    fe one_l = {{1, 0, 0, 0}};
    for (size_t base = lo - lo % GEN_BLOCK; base < hi; base += GEN_BLOCK) {
        uint8_t rb[64];
        memset(rb, 0, sizeof rb);
        prng32(rb, seed, 2, base, 2);
        prng32(rb + 32, seed, 2, base, 3);
        fe r;
        ed_reduce512(&r, rb);
        edp P;
        ed_double_mul(&P, &r, NULL, NULL, 0);
        fe rs[GEN_BLOCK], zs[GEN_BLOCK], zinv[GEN_BLOCK];
        edp pts[GEN_BLOCK];
        for (int j = 0; j < GEN_BLOCK; j++) {
            pts[j] = P;
            rs[j] = r;
            zs[j] = P.Z; /* never zero: the formulas are complete */
            ed_add(&P, &ED.B, 0);
            m_add(fl, &r, &r, &one_l);
        }
        batch_inv(fp, zinv, zs, GEN_BLOCK);
        for (int j = 0; j < GEN_BLOCK; j++) {
            size_t i = base + (size_t)j;
            if (i < lo || i >= hi) continue;
            size_t kidx = i % KEY_POOL;
            uint8_t* sig = sigs + 64 * i;
            uint8_t* msg = msgs + 32 * i;
            prng32(msg, seed, 2, i, 1);
            fe x, y, k, s, am, km;
            m_mul(fp, &x, &pts[j].X, &zinv[j]);
            m_mul(fp, &y, &pts[j].Y, &zinv[j]);
            m_from(fp, &x, &x);
            m_from(fp, &y, &y);
            fe_to_le(sig, &y);
            sig[31] |= (uint8_t)((x.v[0] & 1) << 7);
            uint8_t h[64];
            sha512_3(h, sig, 32, pool_pk + 32 * kidx, 32, msg, 32);
            ed_reduce512(&k, h);
            /* s = r + k*a mod L */
            m_to(fl, &am, &pool_a[kidx]);
            m_to(fl, &km, &k);
            m_mul(fl, &s, &km, &am);
            m_from(fl, &s, &s);
            m_add(fl, &s, &s, &rs[j]);
            fe_to_le(sig + 32, &s);
            memcpy(pks + 32 * i, pool_pk + 32 * kidx, 32);
        }
    }
}

/* ------------------------------------------------------------------------------------------- threading */
typedef struct {
    int kind; /* 0 ecrecover, 1 ed verify, 2 sw gen, 3 ed gen */
    const swcurve* c;
    int curve_id, low_s;
    u64 seed;
    size_t lo, hi;
    const uint8_t *sigs, *msgs, *pks;
    uint8_t *out, *status, *wsigs, *wmsgs, *wpks;
    const fe* pool_s;
    const uint8_t *pool_b1, *pool_b2;
} job;

static void* worker(void* p) {
    job* j = (job*)p;
    switch (j->kind) {
        case 0:
            for (size_t i = j->lo; i < j->hi; i++) {
                int st = sw_ecrecover_one(j->c, j->sigs + 64 * i, j->msgs + 32 * i, j->out + 64 * i);
                if (j->status) j->status[i] = (uint8_t)st;
            }
            break;
        case 1:
            for (size_t i = j->lo; i < j->hi; i++)
                j->out[i] = (uint8_t)ed_verify_one(j->sigs + 64 * i, j->msgs + 32 * i, j->pks + 32 * i);
            break;
        case 2:
            sw_gen_range(j->c, j->curve_id, j->seed, j->lo, j->hi, j->low_s, j->wsigs, j->wmsgs, j->wpks, j->pool_s,
                         j->pool_b1);
            break;
        case 3:
            ed_gen_range(j->seed, j->lo, j->hi, j->wsigs, j->wmsgs, j->wpks, j->pool_s, j->pool_b1, j->pool_b2);
            break;
    }
    return NULL;
}

static void run_jobs(job* proto, size_t n, int threads) {
    if (threads < 1) threads = 1;
    if ((size_t)threads > n) threads = (int)(n ? n : 1);
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)threads);
    job* js = (job*)malloc(sizeof(job) * (size_t)threads);
    for (int t = 0; t < threads; t++) {
        js[t] = *proto;
        js[t].lo = n * (size_t)t / (size_t)threads;
        js[t].hi = n * (size_t)(t + 1) / (size_t)threads;
        if (t > 0) pthread_create(&th[t], NULL, worker, &js[t]);
    }
    worker(&js[0]);
    for (int t = 1; t < threads; t++) pthread_join(th[t], NULL);
    free(th);
    free(js);
}

/* ------------------------------------------------------------------------------------------- public API */
int oracle_ecrecover(int curve, const uint8_t* sigs, const uint8_t* msgs, size_t n, uint8_t* out, uint8_t* status,
                     int threads) {
    pthread_once(&g_once, init_all);
    if (curve != 0 && curve != 1) return 1;
    job j;
    memset(&j, 0, sizeof j);
    j.kind = 0;
    j.c = curve == 0 ? &K1 : &R1;
    j.sigs = sigs;
    j.msgs = msgs;
    j.out = out;
    j.status = status;
    run_jobs(&j, n, threads);
    return 0;
}

int oracle_ed25519_verify(const uint8_t* sigs, const uint8_t* msgs, const uint8_t* pks, size_t n, uint8_t* valid,
                          int threads) {
    pthread_once(&g_once, init_all);
    job j;
    memset(&j, 0, sizeof j);
    j.kind = 1;
    j.sigs = sigs;
    j.msgs = msgs;
    j.pks = pks;
    j.out = valid;
    run_jobs(&j, n, threads);
    return 0;
}

/* variable-length messages (msg i = msg_bytes[off[i] .. off[i+1])), optionally dalek `verify_strict`; single-threaded */
int oracle_ed25519_verify_msgs(const uint8_t* sigs, const uint8_t* msg_bytes, const uint64_t* off, const uint8_t* pks,
                               size_t n, int strict, uint8_t* valid) {
    pthread_once(&g_once, init_all);
    for (size_t i = 0; i < n; i++)
        valid[i] = (uint8_t)ed_verify_msg(sigs + 64 * i, msg_bytes + off[i], (size_t)(off[i + 1] - off[i]), pks + 32 * i, strict);
    return 0;
}

/* RFC 8032 signature over a message of any length with the key derived from a 32-byte seed: sig 64 B, pk 32 B */
int oracle_ed25519_sign(const uint8_t* seed, const uint8_t* msg, size_t len, uint8_t* sig, uint8_t* pk) {
    pthread_once(&g_once, init_all);
    const mctx* fl = &ED.fl;
    uint8_t h[64], wide[64];
    sha512_3(h, seed, 32, NULL, 0, NULL, 0);
    h[0] &= 248;
    h[31] &= 63;
    h[31] |= 64;
    fe a, amod, r, k, s, am, km;
    fe_from_le(&a, h);
    edp A, Rp;
    ed_double_mul(&A, &a, NULL, NULL, 0);
    ed_compress(pk, &A);
    memset(wide, 0, sizeof wide);
    memcpy(wide, h, 32);
    ed_reduce512(&amod, wide);
    uint8_t hr[64];
    sha512_3(hr, h + 32, 32, msg, len, NULL, 0);
    ed_reduce512(&r, hr);
    ed_double_mul(&Rp, &r, NULL, NULL, 0);
    ed_compress(sig, &Rp);
    sha512_3(hr, sig, 32, pk, 32, msg, len);
    ed_reduce512(&k, hr);
    m_to(fl, &am, &amod);
    m_to(fl, &km, &k);
    m_mul(fl, &s, &km, &am);
    m_from(fl, &s, &s);
    m_add(fl, &s, &s, &r);
    fe_to_le(sig + 32, &s);
    return 0;
}

/* n valid signatures (sigs n*64, msgs n*32) and the signers' public keys (pks n*64, may be NULL) */
int oracle_gen_ecdsa(int curve, uint64_t seed, size_t n, int low_s, uint8_t* sigs, uint8_t* msgs, uint8_t* pks,
                     int threads) {
    pthread_once(&g_once, init_all);
    if (curve != 0 && curve != 1) return 1;
    const swcurve* c = curve == 0 ? &K1 : &R1;
    fe* pool_d = (fe*)malloc(sizeof(fe) * KEY_POOL);
    uint8_t* pool_pk = (uint8_t*)malloc(64 * KEY_POOL);
    size_t np = n < KEY_POOL ? n : KEY_POOL;
    for (size_t i = 0; i < np; i++) {
        prng_scalar(&c->fn, &pool_d[i], seed, (u64)curve, i, 0);
        jac Q;
        sw_double_mul(c, &Q, &pool_d[i], NULL, NULL);
        aff A;
        jac_to_aff(c, &A, &Q);
        fe ax, ay;
        m_from(&c->fp, &ax, &A.x);
        m_from(&c->fp, &ay, &A.y);
        fe_to_be(pool_pk + 64 * i, &ax);
        fe_to_be(pool_pk + 64 * i + 32, &ay);
    }
    job j;
    memset(&j, 0, sizeof j);
    j.kind = 2;
    j.c = c;
    j.curve_id = curve;
    j.low_s = low_s;
    j.seed = seed;
    j.wsigs = sigs;
    j.wmsgs = msgs;
    j.wpks = pks;
    j.pool_s = pool_d;
    j.pool_b1 = pool_pk;
    run_jobs(&j, n, threads);
    free(pool_d);
    free(pool_pk);
    return 0;
}

/* n valid RFC 8032 signatures over 32-byte messages: sigs n*64, msgs n*32, pks n*32 */
int oracle_gen_ed25519(uint64_t seed, size_t n, uint8_t* sigs, uint8_t* msgs, uint8_t* pks, int threads) {
    pthread_once(&g_once, init_all);
    fe* pool_a = (fe*)malloc(sizeof(fe) * KEY_POOL);
    uint8_t* pool_prefix = (uint8_t*)malloc(32 * KEY_POOL);
    uint8_t* pool_pk = (uint8_t*)malloc(32 * KEY_POOL);
    size_t np = n < KEY_POOL ? n : KEY_POOL;
    for (size_t i = 0; i < np; i++) {
        uint8_t sk[32], h[64];
        prng32(sk, seed, 2, i, 0);
        sha512_3(h, sk, 32, NULL, 0, NULL, 0);
        h[0] &= 248;
        h[31] &= 63;
        h[31] |= 64;
        fe a;
        fe_from_le(&a, h);
        memcpy(pool_prefix + 32 * i, h + 32, 32);
        edp A;
        ed_double_mul(&A, &a, NULL, NULL, 0);
        ed_compress(pool_pk + 32 * i, &A);
        /* a mod L for the s computation (a < 2^255 < 16 L): reduce through the 512-bit path */
        uint8_t wide[64];
        memset(wide, 0, sizeof wide);
        memcpy(wide, h, 32);
        ed_reduce512(&pool_a[i], wide);
    }
    job j;
    memset(&j, 0, sizeof j);
    j.kind = 3;
    j.seed = seed;
    j.wsigs = sigs;
    j.wmsgs = msgs;
    j.wpks = pks;
    j.pool_s = pool_a;
    j.pool_b1 = pool_prefix;
    j.pool_b2 = pool_pk;
    run_jobs(&j, n, threads);
    free(pool_a);
    free(pool_prefix);
    free(pool_pk);
    return 0;
}

/* SHA-512 of a short message (<= 256 bytes) -- lets the tests pin this file's hash against hashlib */
int oracle_sha512(const uint8_t* msg, size_t len, uint8_t* out) {
    sha512_3(out, msg, len, NULL, 0, NULL, 0);
    return 0;
}
