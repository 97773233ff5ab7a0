// Scalar-field arithmetic (mod the group orders n_k1, n_r1 and the ed25519 order L), the secp256k1 GLV
// split and the fixed signed-window recoding used by the interleaved double-scalar multiplication.
//
// Replaces: src/wgsl/ff.wgsl:60-120 `ff_inverse` (binary extended GCD, variable time / divergent) and the
// Barrett `ff_mul` mod n (src/wgsl/ff.wgsl:169-198) used by src/wgsl/secp256k1_ecdsa.wgsl:56-80;
// src/wgsl/ed25519_reduce_fr.wgsl:88-125 (512-bit hash -> mod L); src/wgsl/bigint.wgsl `bigint_to_bits_le`
// (bit decomposition by 256 full-width halvings) -- here scalar digits are read with shifts.
// The GLV lattice constants are the ones the reference carries at src/curve_algos/secp256k1_curve.rs:47-68
// (its split procedure, src/curve_algos/secp256k1_mul.rs:51-85, is test-only there; the WGSL never uses it).
#pragma once
#include "field.cuh"
#include "consts_gen.cuh"
#include "ptab.h"
#include "modinv.cuh"

namespace sigops {

// ---------------------------------------------------------------------------------------------------------
// generic Montgomery arithmetic mod an odd 256-bit modulus, R = 2^256
// ---------------------------------------------------------------------------------------------------------
struct ModK1N {
    typedef ModInvK1N MI;
    static SG_HD void mod(u32* m) {
        const u32 M[8] = SG_K1_N;
        copy8(m, M);
    }
    static SG_HD void r2(u32* m) {
        const u32 M[8] = SG_K1_N_R2;
        copy8(m, M);
    }
    static SG_HD void r3(u32* m) {
        const u32 M[8] = SG_K1_N_R3;
        copy8(m, M);
    }
    static SG_HD void minus2(u32* m) {
        const u32 M[8] = SG_K1_N_MINUS2;
        copy8(m, M);
    }
    static constexpr u32 n0inv = SG_K1_N_N0INV;
};
struct ModR1N {
    typedef ModInvR1N MI;
    static SG_HD void mod(u32* m) {
        const u32 M[8] = SG_R1_N;
        copy8(m, M);
    }
    static SG_HD void r2(u32* m) {
        const u32 M[8] = SG_R1_N_R2;
        copy8(m, M);
    }
    static SG_HD void r3(u32* m) {
        const u32 M[8] = SG_R1_N_R3;
        copy8(m, M);
    }
    static SG_HD void minus2(u32* m) {
        const u32 M[8] = SG_R1_N_MINUS2;
        copy8(m, M);
    }
    static constexpr u32 n0inv = SG_R1_N_N0INV;
};
struct ModEdL {
    static SG_HD void mod(u32* m) {
        const u32 M[8] = SG_ED_L;
        copy8(m, M);
    }
    static SG_HD void r2(u32* m) {
        const u32 M[8] = SG_ED_L_R2;
        copy8(m, M);
    }
    static SG_HD void r3(u32* m) {
        const u32 M[8] = SG_ED_L_R3;
        copy8(m, M);
    }
    static SG_HD void minus2(u32* m) {
        const u32 M[8] = SG_ED_L_MINUS2;
        copy8(m, M);
    }
    static constexpr u32 n0inv = SG_ED_L_N0INV;
};

template <class M>
struct Sc {
    // r = t / 2^256 mod m, for t < 2^256 * m.  Lazy carries: the carry out of each row lands above every limb
    // that still determines a quotient digit, so they are summed once at the end.
    static SG_HD void reduce16(u32* r, const u32* tin) {
        u32 m[8];
        M::mod(m);
        u32 t[16];
#pragma unroll
        for (int i = 0; i < 16; i++) t[i] = tin[i];
        u32 cr[9];
#pragma unroll
        for (int i = 0; i < 9; i++) cr[i] = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            u32 q = t[i] * M::n0inv;
            cr[i] += mad_row4(t + i, m[0], m[2], m[4], m[6], q);
            cr[i + 1] += mad_row4(t + i + 1, m[1], m[3], m[5], m[7], q);
        }
        u32 s[8];
        u32 c = add8(s, t + 8, cr);
        u32 top = cr[8] + c;
        u32 u[8];
        u32 bw = sub8(u, s, m);
        select8(r, (top != 0) || (bw == 0), s, u);
    }
    // Montgomery product a*b/R mod m; requires a < 2^256, b < m (or a*b < 2^256*m)
    static SG_CALL Fe mmul_(Fe a, Fe b) {
        u32 t[16];
        Fe r;
        mul8x8(t, a.v, b.v);
        reduce16(r.v, t);
        return r;
    }
    static SG_CALL Fe msqr_(Fe a) {
        u32 t[16];
        Fe r;
        sqr8(t, a.v);
        reduce16(r.v, t);
        return r;
    }
    // n successive Montgomery squarings in one out-of-line loop (squaring inlined)
    static SG_CALL Fe msqr_n_(Fe a, int n) {
#pragma unroll 1
        for (int i = 0; i < n; i++) {
            u32 t[16];
            sqr8(t, a.v);
            reduce16(a.v, t);
        }
        return a;
    }
    static SG_HD void mmul(u32* r, const u32* a, const u32* b) {
        Fe x, y;
        copy8(x.v, a);
        copy8(y.v, b);
        Fe z = mmul_(x, y);
        copy8(r, z.v);
    }
    static SG_HD void msqr(u32* r, const u32* a) {
        Fe x;
        copy8(x.v, a);
        Fe z = msqr_(x);
        copy8(r, z.v);
    }
    static SG_HD void to_mont(u32* r, const u32* a) {
        u32 r2[8];
        M::r2(r2);
        mmul(r, a, r2);
    }
    static SG_HD void from_mont(u32* r, const u32* a) {
        u32 t[16];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            t[i] = a[i];
            t[i + 8] = 0;
        }
        reduce16(r, t);
    }
    // r = a mod m for a < 2m
    static SG_HD void reduce_once(u32* r, const u32* a) {
        u32 m[8], u[8];
        M::mod(m);
        u32 bw = sub8(u, a, m);
        select8(r, bw == 0, a, u);
    }
    static SG_HD void neg(u32* r, const u32* a) {  // a in [0,m)
        u32 m[8], u[8];
        M::mod(m);
        sub8(u, m, a);
        select8(r, is_zero8(a), u, a);
    }
    static SG_HD bool lt_mod(const u32* a) {
        u32 m[8];
        M::mod(m);
        return !gte8(a, m);
    }
    static SG_HD void r3(u32* r) { M::r3(r); }  // 2^768 mod m
    // plain inverse a^-1 mod m for a in [0, m): safegcd
    static SG_HD void inv_plain(u32* r, const u32* a) { ModInv<typename M::MI>::inv(r, a); }
    // Montgomery-domain inverse by Fermat: a^(m-2), fixed 4-bit windows (256 squarings + 64 + 14 products).
    // Kept as an independent cross-check of inv_plain in the unit tests; the kernels use inv_plain.
    static SG_HD void minv(u32* r, const u32* a) {
        u32 e[8];
        M::minus2(e);
        u32 tab[15][8];  // a^1..a^15
        copy8(tab[0], a);
#pragma unroll 1
        for (int i = 1; i < 15; i++) mmul(tab[i], tab[i - 1], a);
        u32 acc[8];
        {
            u32 d = e[7] >> 28;  // top nibble is nonzero for all three moduli
            copy8(acc, tab[d - 1]);
        }
#pragma unroll 1
        for (int w = 62; w >= 0; w--) {
            {
                Fe x;
                copy8(x.v, acc);
                x = msqr_n_(x, 4);
                copy8(acc, x.v);
            }
            u32 d = (e[w >> 3] >> ((w & 7) * 4)) & 15u;
            if (d) mmul(acc, acc, tab[d - 1]);
        }
        copy8(r, acc);
    }
};

// ---------------------------------------------------------------------------------------------------------
// secp256k1 GLV split: k = k1 + k2*lambda (mod n), |k1|,|k2| < 2^128.  Plain integer arithmetic mod 2^256:
//   c1 = round(k*g1 / 2^384), c2 = round(k*g2 / 2^384)
//   k1 = k - c1*a1 - c2*a2,   k2 = c1*(-b1) - c2*b2        (b2 = a1)
// ---------------------------------------------------------------------------------------------------------
struct GlvSplit {
    u32 k1[5], k2[5];  // magnitudes (fit in 128 bits; limb 4 is zero, kept for the recoding offset)
    bool neg1, neg2;
};

SG_HD void neg256(u32* r, const u32* a) {
    u32 z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    sub8(r, z, a);
}

SG_HD void k1_glv_split(GlvSplit& out, const u32* k) {
    const u32 G1[8] = SG_K1_G1, G2[8] = SG_K1_G2, A1[8] = SG_K1_A1, A2[8] = SG_K1_A2, MB1[8] = SG_K1_MB1;
    u32 t[16], c1[8], c2[8];
    const u32 half[8] = {0, 0, 0, 0x80000000u, 0, 0, 0, 0};  // 2^383 relative to limb 8
    mul8x8(t, k, G1);
    add8(t + 8, t + 8, half);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        c1[i] = t[12 + i];
        c1[4 + i] = 0;
    }
    mul8x8(t, k, G2);
    add8(t + 8, t + 8, half);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        c2[i] = t[12 + i];
        c2[4 + i] = 0;
    }
    u32 p1[16], p2[16], r1[8], r2[8];
    mul8x8(p1, c1, A1);
    mul8x8(p2, c2, A2);
    sub8(r1, k, p1);
    sub8(r1, r1, p2);
    mul8x8(p1, c1, MB1);
    mul8x8(p2, c2, A1);
    sub8(r2, p1, p2);
    out.neg1 = (r1[7] >> 31) != 0;
    out.neg2 = (r2[7] >> 31) != 0;
    if (out.neg1) neg256(r1, r1);
    if (out.neg2) neg256(r2, r2);
#pragma unroll
    for (int i = 0; i < 5; i++) {
        out.k1[i] = r1[i];
        out.k2[i] = r2[i];
    }
}

// ---------------------------------------------------------------------------------------------------------
// fixed signed-window recoding.  For window width W and NW windows, adding the constant
// C = sum_i 2^(W-1) * 2^(W*i) turns the unsigned windows u_i of k' = k + C into signed digits
// d_i = u_i - 2^(W-1) in [-2^(W-1), 2^(W-1)) with k = sum d_i 2^(W*i)   (requires k + C < 2^(W*NW)).
// Every lane adds at the same iterations -> no divergence except on d_i == 0.
// ---------------------------------------------------------------------------------------------------------
// k (NL limbs) += sum_{i < NW} 2^(W-1) * 2^(W*i): the offset that turns unsigned W-bit windows into signed digits.
// The per-limb pattern is a compile-time constant after unrolling.
template <int NL, int W, int NW>
SG_HD void recode_offset(u32* k) {
    u64 c = 0;
#pragma unroll
    for (int l = 0; l < NL; l++) {
        u32 pat = 0;
#pragma unroll
        for (int i = 0; i < NW; i++) {
            const int bit = W * i + W - 1;
            if ((bit >> 5) == l) pat |= 1u << (bit & 31);
        }
        c += (u64)k[l] + pat;
        k[l] = (u32)c;
        c >>= 32;
    }
}

// signed digit of window i (width W) of the offset scalar k'; windows may straddle limbs (the array must extend one
// limb past the last window's low limb)
template <int W>
SG_HD int recode_digit(const u32* kp, int i) {
    const int bit = W * i;
    const int w = bit >> 5, sh = bit & 31;
    u32 v = kp[w] >> sh;
    if (sh + W > 32) v |= kp[w + 1] << (32 - sh);
    return (int)(v & ((1u << W) - 1u)) - (1 << (W - 1));
}

// ---- positional fixed-base tables (ptab.h): signed digits of a 256-bit scalar, lowest window first ----
// kp (9 limbs) = k + pat
SG_HD void ptab_recode(u32* kp, const u32* k, const PTab& t) {
    u64 c = 0;
#pragma unroll
    for (int i = 0; i < 9; i++) {
        c += (u64)(i < 8 ? k[i] : 0u) + t.pat[i];
        kp[i] = (u32)c;
        c >>= 32;
    }
}

// the signed digit of the lowest window of kp, which is then shifted down by one window (1 <= w <= 31)
SG_HD int ptab_pop_digit(u32* kp, u32 w) {
    const int d = (int)(kp[0] & ((1u << w) - 1u)) - (int)(1u << (w - 1));
#pragma unroll
    for (int i = 0; i < 8; i++) kp[i] = (kp[i] >> w) | (kp[i + 1] << (32 - w));
    kp[8] >>= w;
    return d;
}

// Ask L2 for every entry the scalar k will select, well before the additions need them (the lane-group kernels: a lone
// warp cannot hide an HBM miss behind other warps).  No-op in the host simulation.
SG_HD void ptab_prefetch(const PTab& t, const u32* k, int entry_words);

// word offset of entry |d| of window j (d != 0; d == 0 reads entry 1, which the caller discards)
SG_HD size_t ptab_offset(const PTab& t, u32 j, int d, int entry_words) {
    const u32 e = d == 0 ? 0u : (u32)(d < 0 ? -d : d) - 1u;
    return (((size_t)j << (t.w - 1)) + e) * (size_t)entry_words;
}

SG_HD void ptab_prefetch(const PTab& t, const u32* k, int entry_words) {
#if SG_PTX
    u32 kp[9];
    ptab_recode(kp, k, t);
#pragma unroll 1
    for (u32 j = 0; j < t.pos; j++) {
        const int d = ptab_pop_digit(kp, t.w);
        const u32* e = t.base + ptab_offset(t, j, d, entry_words);
        asm volatile("prefetch.global.L2 [%0];" ::"l"(e));
        if (entry_words > 16) asm volatile("prefetch.global.L2 [%0];" ::"l"(e + entry_words - 1));  // 96-byte entries straddle lines
    }
#else
    (void)t;
    (void)k;
    (void)entry_words;
#endif
}

}  // namespace sigops
