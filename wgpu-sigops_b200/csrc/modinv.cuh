// Modular inversion by Bernstein-Yang "safegcd" division steps (the algorithm class libsecp256k1 -- the reference's CPU
// oracle for secp256k1 -- uses in modinv32): 20 batches of 30 branch-free division steps on the low words, each batch
// followed by one 2x2 transition-matrix update of the 9 x 30-bit signed limb vectors (f, g) and (d, e).
//
// Replaces src/wgsl/ff.wgsl:60-120 `ff_inverse` (binary extended GCD: data-dependent branches, one warp lane group per
// path) and the `modpow(z, p-2)` of src/wgsl/secp_curve_utils.wgsl:1-30 / src/wgsl/ed25519_curve.wgsl:204-228.
// Why not Fermat here: an exponentiation costs ~330 modular products (12.6k wide MACs mod p, 33k mod n with generic
// Montgomery) on the INT32 multiply pipe, the kernels' bottleneck; this costs ~1.8k wide MACs plus ~9k shift/mask/add
// operations on the otherwise under-used ALU pipe, and every lane executes the same instruction stream (no divergence).
#pragma once
#include "field.cuh"
#include "consts_gen.cuh"

namespace sigops {

typedef int32_t i32;
typedef int64_t i64;

struct S30 {
    i32 v[9];
};
struct Trans2x2 {
    i32 u, v, q, r;
};

// 30 division steps on the low 30 bits of f and g; returns the new zeta and the transition matrix scaled by 2^30.
// zeta = -(delta + 1/2); all branch-free.
SG_HD i32 divsteps_30(i32 zeta, u32 f0, u32 g0, Trans2x2& t) {
    u32 u = 1, v = 0, q = 0, r = 1, f = f0, g = g0;
#pragma unroll
    for (int i = 0; i < 30; i++) {
        u32 c1 = (u32)(zeta >> 31);  // all ones when zeta < 0
        u32 c2 = 0u - (g & 1u);      // all ones when g is odd
        u32 x = (f ^ c1) - c1, y = (u ^ c1) - c1, z = (v ^ c1) - c1;  // conditionally negated f, u, v
        g += x & c2;
        q += y & c2;
        r += z & c2;
        c1 &= c2;                    // zeta < 0 and g odd: swap roles
        zeta = (zeta ^ (i32)c1) - 1;
        f += g & c1;
        u += q & c1;
        v += r & c1;
        g >>= 1;
        u <<= 1;
        v <<= 1;
    }
    t.u = (i32)u;
    t.v = (i32)v;
    t.q = (i32)q;
    t.r = (i32)r;
    return zeta;
}

// (f, g) <- t * (f, g) / 2^30 (exact)
SG_HD void update_fg_30(S30& f, S30& g, const Trans2x2& t) {
    const i32 M30 = 0x3FFFFFFF;
    const i64 u = t.u, v = t.v, q = t.q, r = t.r;
    i64 cf = u * f.v[0] + v * g.v[0];
    i64 cg = q * f.v[0] + r * g.v[0];
    cf >>= 30;
    cg >>= 30;
#pragma unroll
    for (int i = 1; i < 9; i++) {
        cf += u * f.v[i] + v * g.v[i];
        cg += q * f.v[i] + r * g.v[i];
        f.v[i - 1] = (i32)cf & M30;
        cf >>= 30;
        g.v[i - 1] = (i32)cg & M30;
        cg >>= 30;
    }
    f.v[8] = (i32)cf;
    g.v[8] = (i32)cg;
}

// (d, e) <- t * (d, e) / 2^30 (mod m); keeps d, e in (-2m, m)
SG_HD void update_de_30(S30& d, S30& e, const Trans2x2& t, const S30& m, u32 m_inv30) {
    const i32 M30 = 0x3FFFFFFF;
    const i64 u = t.u, v = t.v, q = t.q, r = t.r;
    const i32 sd = d.v[8] >> 31, se = e.v[8] >> 31;
    i32 md = (t.u & sd) + (t.v & se);
    i32 me = (t.q & sd) + (t.r & se);
    i32 di = d.v[0], ei = e.v[0];
    i64 cd = u * di + v * ei;
    i64 ce = q * di + r * ei;
    md -= (i32)((m_inv30 * (u32)cd + (u32)md) & (u32)M30);
    me -= (i32)((m_inv30 * (u32)ce + (u32)me) & (u32)M30);
    cd += (i64)m.v[0] * md;
    ce += (i64)m.v[0] * me;
    cd >>= 30;
    ce >>= 30;
#pragma unroll
    for (int i = 1; i < 9; i++) {
        di = d.v[i];
        ei = e.v[i];
        cd += u * di + v * ei;
        ce += q * di + r * ei;
        cd += (i64)m.v[i] * md;
        ce += (i64)m.v[i] * me;
        d.v[i - 1] = (i32)cd & M30;
        cd >>= 30;
        e.v[i - 1] = (i32)ce & M30;
        ce >>= 30;
    }
    d.v[8] = (i32)cd;
    e.v[8] = (i32)ce;
}

// r in (-2m, m) (signed limbs), optionally negated by `sign` < 0, brought to [0, m) with canonical 30-bit limbs
SG_HD void normalize_30(S30& r, i32 sign, const S30& m) {
    const i32 M30 = 0x3FFFFFFF;
    i32 cond_add = r.v[8] >> 31;
#pragma unroll
    for (int i = 0; i < 9; i++) r.v[i] += m.v[i] & cond_add;
    i32 cond_neg = sign >> 31;
#pragma unroll
    for (int i = 0; i < 9; i++) r.v[i] = (r.v[i] ^ cond_neg) - cond_neg;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        r.v[i + 1] += r.v[i] >> 30;
        r.v[i] &= M30;
    }
    cond_add = r.v[8] >> 31;
#pragma unroll
    for (int i = 0; i < 9; i++) r.v[i] += m.v[i] & cond_add;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        r.v[i + 1] += r.v[i] >> 30;
        r.v[i] &= M30;
    }
}

SG_HD void s30_from_limbs(S30& r, const u32* a) {
    // limb i = bits [30 i, 30 i + 30) of the 256-bit value
#pragma unroll
    for (int i = 0; i < 9; i++) {
        const int bit = 30 * i, w = bit >> 5, sh = bit & 31;
        u32 lo = a[w] >> sh;
        u32 hi = (sh > 2 && w + 1 < 8) ? (a[w + 1] << (32 - sh)) : 0u;
        r.v[i] = (i32)((lo | hi) & 0x3FFFFFFFu);
    }
}

SG_HD void s30_to_limbs(u32* a, const S30& r) {
    // canonical limbs in [0, 2^30), value < 2^256
#pragma unroll
    for (int w = 0; w < 8; w++) {
        const int bit = 32 * w, i = bit / 30, sh = bit - 30 * i;
        u32 x = (u32)r.v[i] >> sh;
        x |= (u32)r.v[i + 1] << (30 - sh);  // 30 - sh in [2, 30]: always contributes bits
        if (sh > 28 && i + 2 < 9) x |= (u32)r.v[i + 2] << (60 - sh);
        a[w] = x;
    }
}

struct ModInvK1P {
    static SG_HD void mod(S30& m) { const S30 M = {SG_K1_P_S30}; m = M; }
    static constexpr u32 inv30 = SG_K1_P_INV30;
};
struct ModInvK1N {
    static SG_HD void mod(S30& m) { const S30 M = {SG_K1_N_S30}; m = M; }
    static constexpr u32 inv30 = SG_K1_N_INV30;
};
struct ModInvR1P {
    static SG_HD void mod(S30& m) { const S30 M = {SG_R1_P_S30}; m = M; }
    static constexpr u32 inv30 = SG_R1_P_INV30;
};
struct ModInvR1N {
    static SG_HD void mod(S30& m) { const S30 M = {SG_R1_N_S30}; m = M; }
    static constexpr u32 inv30 = SG_R1_N_INV30;
};
struct ModInvEdP {
    static SG_HD void mod(S30& m) { const S30 M = {SG_ED_P_S30}; m = M; }
    static constexpr u32 inv30 = SG_ED_P_INV30;
};

// r = x^-1 mod m for canonical x in [0, m) (plain integers, 8 x 32-bit limbs); x == 0 gives 0.
// One out-of-line copy per modulus (the loop body is ~1.5k instructions).
template <class MI>
struct ModInv {
    static SG_CALL Fe inv_(Fe x) {
        S30 m, d, e, f, g;
        MI::mod(m);
#pragma unroll
        for (int i = 0; i < 9; i++) {
            d.v[i] = 0;
            e.v[i] = 0;
        }
        e.v[0] = 1;
        f = m;
        s30_from_limbs(g, x.v);
        i32 zeta = -1;
#pragma unroll 1
        for (int it = 0; it < 20; it++) {
            Trans2x2 t;
            zeta = divsteps_30(zeta, (u32)f.v[0], (u32)g.v[0], t);
            update_de_30(d, e, t, m, MI::inv30);
            update_fg_30(f, g, t);
        }
        // g == 0 and f == +-gcd == +-1 now (or +-m when x == 0, in which case d == 0)
        normalize_30(d, f.v[8], m);
        Fe r;
        s30_to_limbs(r.v, d);
        return r;
    }
    static SG_HD void inv(u32* r, const u32* x) {
        Fe a;
        copy8(a.v, x);
        Fe b = inv_(a);
        copy8(r, b.v);
    }
};

// Field inversions used by the kernels (to-affine / compress): safegcd on the canonical plain value.
SG_HD void fe_inv(FpK1*, Fe& r, const Fe& a) {
    Fe n;
    FpK1::normalize(n, a);
    r = ModInv<ModInvK1P>::inv_(n);
}
SG_HD void fe_inv(Fp25519*, Fe& r, const Fe& a) {
    Fe n;
    Fp25519::normalize(n, a);
    r = ModInv<ModInvEdP>::inv_(n);
}
// Montgomery domain: a = x R  ->  (x R)^-1 = x^-1 R^-1;  times R^3 (one Montgomery product) = x^-1 R
SG_HD void fe_inv(FpR1*, Fe& r, const Fe& a) {
    const Fe R3 = {SG_R1_P_R3};
    Fe n;
    FpR1::normalize(n, a);
    Fe t = ModInv<ModInvR1P>::inv_(n);
    FpR1::mul(r, t, R3);
}

}  // namespace sigops
