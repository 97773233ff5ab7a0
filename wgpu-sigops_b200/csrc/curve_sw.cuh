// Short-Weierstrass group law (Jacobian coordinates) and the fused ECDSA public-key recovery for
// secp256k1 and secp256r1: one signature per thread, everything between the input bytes and the output
// bytes stays in registers / the per-thread table.
//
// Replaces the reference's five-stage pipeline (src/secp256k1_ecdsa.rs:61-213, src/secp256r1_ecdsa.rs:62-214;
// stages src/wgsl/main/secp256k1_ecdsa_main_0..4.wgsl) and its device functions:
//   src/wgsl/secp256k1_ecdsa.wgsl:7-130   `secp256k1_ecrecover_0` / `secp256k1_ecrecover`
//   src/wgsl/secp256k1_curve.wgsl:26-111  projective add-2007-bl / dbl-2007-bl (incomplete: "unsafe")
//   src/wgsl/secp256r1_curve.wgsl:34-146  RCB-2015 complete add / dbl
//   src/wgsl/secp256k1_curve.wgsl:277-294,388-447  `projective_mul` (double-and-add), `projective_fixed_mul`
//   src/wgsl/secp_curve_utils.wgsl:1-30   `projective_to_affine_non_mont`
//   src/wgsl/signature.wgsl:6-21          `decode_signature`
// Differences by design: Jacobian formulas with explicit handling of infinity / P == Q / P == -Q (the reference's
// formulas are wrong there); ONE interleaved Strauss-Shamir pass with fixed signed windows
// (GLV-split into four 128-bit streams on secp256k1) instead of two independent scalar multiplications; and the
// validity checks of the CPU libraries the reference is tested against (r, s range; x not on curve; Q = infinity),
// reported per signature.
#pragma once
#include "scalar.cuh"

namespace sigops {

struct alignas(16) Q4 {
    u32 x, y, z, w;
};

// Per-thread scratch table in 16-byte chunks, interleaved across threads: chunk q of this thread is base[q*stride].
// Adjacent lanes touch adjacent 16-byte slots whatever entry each lane selects (coalesced in global memory,
// conflict-free in shared memory).
struct TabRef {
    Q4* base;
    u32 stride;
};

SG_HD void tab_store_fe(const TabRef& t, int chunk, const Fe& a) {
    Q4 lo = {a.v[0], a.v[1], a.v[2], a.v[3]}, hi = {a.v[4], a.v[5], a.v[6], a.v[7]};
    t.base[(size_t)chunk * t.stride] = lo;
    t.base[(size_t)(chunk + 1) * t.stride] = hi;
}
SG_HD void tab_load_fe(Fe& a, const TabRef& t, int chunk) {
    Q4 lo = t.base[(size_t)chunk * t.stride], hi = t.base[(size_t)(chunk + 1) * t.stride];
    a.v[0] = lo.x;
    a.v[1] = lo.y;
    a.v[2] = lo.z;
    a.v[3] = lo.w;
    a.v[4] = hi.x;
    a.v[5] = hi.y;
    a.v[6] = hi.z;
    a.v[7] = hi.w;
}

struct JacPoint {
    Fe X, Y, Z;
    bool inf;
};

// Curve descriptors are templated on the field flavour: `Cold` = out-of-line field products (FpK1 / FpR1), `Hot` =
// the same field with products inlined at the call site (Inl<>), used only inside the scalar-multiplication loops.
template <class FF>
struct CurveK1T {
    typedef FF F;
    typedef Sc<ModK1N> S;
    typedef CurveK1T<FpK1> Cold;
#if !defined(SG_NO_HOT_INLINE)
    typedef CurveK1T<Inl<FpK1> > Hot;
#else
    typedef CurveK1T<FpK1> Hot;  // out-of-line products everywhere
#endif
    static constexpr bool kGlv = true;
    static constexpr bool kAIsZero = true;
    // t = x^3 + 7
    static SG_HD void rhs(Fe& t, const Fe& x) {
        Fe x2, b;
        F::sqr(x2, x);
        F::mul(t, x2, x);
        F::set_small(b, 7u);
        F::add(t, t, b);
    }
    static SG_HD void mul_beta(Fe& r, const Fe& a) {
        const Fe beta = {SG_K1_BETA};
        F::mul(r, a, beta);
    }
    // r = 3b * a = 21 a: the curve constant of the complete projective formulas (group.cuh)
    static SG_HD void mul_bconst(Fe& r, const Fe& a) { F::mul_u32(r, a, 21u); }
    static SG_HD void mul_bconst3(Fe& r, const Fe& a) { F::mul_u32(r, a, 63u); }  // 9b
    // the generator and the seed 2^-128 G of the positional table (ptab.h), affine x || y in the internal form
#if SG_PTX
    static SG_HD const u32* gen() { return k1_g_dev; }
    static SG_HD const u32* pt0() { return k1_pt0_dev; }
#else
    static SG_HD const u32* gen() { return k1_g_host; }
    static SG_HD const u32* pt0() { return k1_pt0_host; }
#endif
};
typedef CurveK1T<FpK1> CurveK1;

template <class FF>
struct CurveR1T {
    typedef FF F;
    typedef Sc<ModR1N> S;
    typedef CurveR1T<FpR1> Cold;
#if !defined(SG_NO_HOT_INLINE)
    typedef CurveR1T<Inl<FpR1> > Hot;
#else
    typedef CurveR1T<FpR1> Hot;
#endif
    static constexpr bool kGlv = false;
    static constexpr bool kAIsZero = false;  // a = -3
    // t = x^3 - 3x + b   (Montgomery domain)
    static SG_HD void rhs(Fe& t, const Fe& x) {
        const Fe b = {SG_R1_B_MONT};
        Fe x2, x3;
        F::sqr(x2, x);
        F::mul(t, x2, x);
        F::dbl(x3, x);
        F::add(x3, x3, x);
        F::sub(t, t, x3);
        F::add(t, t, b);
    }
    static SG_HD void mul_beta(Fe& r, const Fe& a) { r = a; }
    // r = b * a (Montgomery domain): the curve constant of the complete projective formulas for a = -3 (group.cuh)
    static SG_HD void mul_bconst(Fe& r, const Fe& a) {
        const Fe b = {SG_R1_B_MONT};
        F::mul(r, a, b);
    }
    static SG_HD void mul_bconst3(Fe& r, const Fe& a) {  // unused by the a = -3 formulas; kept for the common interface
        Fe t;
        mul_bconst(t, a);
        F::dbl(r, t);
        F::add(r, r, t);
    }
#if SG_PTX
    static SG_HD const u32* gen() { return r1_g_dev; }
    static SG_HD const u32* pt0() { return r1_pt0_dev; }
#else
    static SG_HD const u32* gen() { return r1_g_host; }
    static SG_HD const u32* pt0() { return r1_pt0_host; }
#endif
};
typedef CurveR1T<FpR1> CurveR1;

// The same curve with the "hot" flavour mapped to the out-of-line products: a kernel instantiated on ColdProducts<C> calls the
// one copy of each field product instead of inlining it at every site.  The lane-group kernels (group.cuh) use it where the
// role programs' instruction fetch costs more than the calls: P-256 always (inlined, its request time grows from 0.78 ms at
// 32 blocks to 1.32 ms at 148; out of line it stays at 0.78-0.79), secp256k1 and ed25519 above ~2,000 signatures
// (profiles/r02_group_cold_products.txt).
template <class C>
struct ColdProducts : C {
    typedef typename C::Cold Hot;
};

// P <- 2P.  a = 0: dbl-2009-l (2M + 5S);  a = -3: dbl-2001-b (3M + 5S).  No point of order 2 on either curve.
template <class C>
SG_HD void jac_dbl(JacPoint& P) {
    typedef typename C::F F;
    if (P.inf) return;
    if (C::kAIsZero) {
        Fe A, B, Cc, D, E, Fq, t;
        F::sqr(A, P.X);
        F::sqr(B, P.Y);
        F::sqr(Cc, B);
        F::add(t, P.X, B);
        F::sqr(t, t);
        F::sub(t, t, A);
        F::sub(t, t, Cc);
        F::dbl(D, t);
        F::dbl(E, A);
        F::add(E, E, A);
        F::sqr(Fq, E);
        F::mul(P.Z, P.Y, P.Z);
        F::dbl(P.Z, P.Z);
        F::dbl(t, D);
        F::sub(P.X, Fq, t);
        F::sub(t, D, P.X);
        F::mul(t, E, t);
        F::template shl<3>(Cc, Cc);  // 8C
        F::sub(P.Y, t, Cc);
    } else {
        Fe delta, gamma, beta, alpha, t, u;
        F::sqr(delta, P.Z);
        F::sqr(gamma, P.Y);
        F::mul(beta, P.X, gamma);
        F::sub(t, P.X, delta);
        F::add(u, P.X, delta);
        F::mul(alpha, t, u);
        F::dbl(t, alpha);
        F::add(alpha, alpha, t);
        F::add(t, P.Y, P.Z);
        F::sqr(t, t);
        F::sub(t, t, gamma);
        F::sub(P.Z, t, delta);
        F::sqr(t, alpha);
        F::template shl<2>(beta, beta);  // 4*beta
        F::dbl(u, beta);                 // 8*beta
        F::sub(P.X, t, u);
        F::sub(t, beta, P.X);
        F::mul(t, alpha, t);
        F::sqr(gamma, gamma);
        F::template shl<3>(gamma, gamma);  // 8*gamma^2
        F::sub(P.Y, t, gamma);
    }
}

// P <- 2P on Y^2 = X^3 + a X + b for an arbitrary a given as a field element (dbl-2007-bl, 2M + 8S): the lane-group
// kernels build the table of multiples of R on a twist whose a depends on the signature (group.cuh).  Not on a hot path.
template <class C>
SG_HD void jac_dbl_a(JacPoint& P, const Fe& a) {
    typedef typename C::F F;
    if (P.inf) return;
    Fe XX, YY, YYYY, ZZ, S, M, t;
    F::sqr(XX, P.X);
    F::sqr(YY, P.Y);
    F::sqr(YYYY, YY);
    F::sqr(ZZ, P.Z);
    F::add(t, P.X, YY);
    F::sqr(t, t);
    F::sub(t, t, XX);
    F::sub(t, t, YYYY);
    F::dbl(S, t);  // S = 2((X + YY)^2 - XX - YYYY) = 4 X YY
    F::add(t, P.Y, P.Z);
    F::sqr(t, t);
    F::sub(t, t, YY);
    F::sub(P.Z, t, ZZ);  // Z3 = 2 Y Z
    F::sqr(ZZ, ZZ);
    F::mul(ZZ, a, ZZ);
    F::dbl(M, XX);
    F::add(M, M, XX);
    F::add(M, M, ZZ);  // M = 3 XX + a ZZ^2
    F::sqr(t, M);
    F::sub(t, t, S);
    F::sub(P.X, t, S);  // X3 = M^2 - 2 S
    F::sub(t, S, P.X);
    F::mul(t, M, t);
    F::template shl<3>(YYYY, YYYY);
    F::sub(P.Y, t, YYYY);  // Y3 = M (S - X3) - 8 YYYY
}

// shared tail of the two additions: given U1,S1 (of P), H = U2-U1, r = S2-S1 and Zm = Z1*Z2 (or Z1), H != 0
template <class F>
SG_HD void jac_add_tail(JacPoint& P, const Fe& U1, const Fe& S1, const Fe& H, const Fe& r, const Fe& Zm) {
    Fe HH, HHH, V, t;
    F::sqr(HH, H);
    F::mul(HHH, H, HH);
    F::mul(V, U1, HH);
    F::sqr(t, r);
    F::sub(t, t, HHH);
    F::sub(t, t, V);
    F::sub(P.X, t, V);
    F::sub(t, V, P.X);
    F::mul(t, r, t);
    F::mul(HHH, S1, HHH);
    F::sub(P.Y, t, HHH);
    F::mul(P.Z, Zm, H);
}

// P <- P + (x2, y2) with an affine second operand that is never infinity (8M + 3S)
template <class C>
SG_HD void jac_madd(JacPoint& P, const Fe& x2, const Fe& y2) {
    typedef typename C::F F;
    if (P.inf) {
        P.X = x2;
        P.Y = y2;
        F::set_one(P.Z);
        P.inf = false;
        return;
    }
    Fe Z1Z1, U2, S2, H, r;
    F::sqr(Z1Z1, P.Z);
    F::mul(U2, x2, Z1Z1);
    F::mul(S2, P.Z, Z1Z1);
    F::mul(S2, y2, S2);
    F::sub(H, U2, P.X);
    F::sub(r, S2, P.Y);
    if (F::is_zero(H)) {
        if (F::is_zero(r))
            jac_dbl<typename C::Cold>(P);  // P == Q: rare, out-of-line products
        else
            P.inf = true;
        return;
    }
    Fe U1 = P.X, S1 = P.Y, Zm = P.Z;
    jac_add_tail<F>(P, U1, S1, H, r, Zm);
}

// P <- P + Q with Q = (X2, Y2, Z2) Jacobian, never infinity (12M + 4S)
template <class C>
SG_HD void jac_add(JacPoint& P, const Fe& X2, const Fe& Y2, const Fe& Z2) {
    typedef typename C::F F;
    if (P.inf) {
        P.X = X2;
        P.Y = Y2;
        P.Z = Z2;
        P.inf = false;
        return;
    }
    Fe Z1Z1, Z2Z2, U1, U2, S1, S2, H, r, Zm;
    F::sqr(Z1Z1, P.Z);
    F::sqr(Z2Z2, Z2);
    F::mul(U1, P.X, Z2Z2);
    F::mul(U2, X2, Z1Z1);
    F::mul(S1, Z2, Z2Z2);
    F::mul(S1, P.Y, S1);
    F::mul(S2, P.Z, Z1Z1);
    F::mul(S2, Y2, S2);
    F::sub(H, U2, U1);
    F::sub(r, S2, S1);
    if (F::is_zero(H)) {
        if (F::is_zero(r))
            jac_dbl<typename C::Cold>(P);  // P == Q: rare, out-of-line products
        else
            P.inf = true;
        return;
    }
    F::mul(Zm, P.Z, Z2);
    jac_add_tail<F>(P, U1, S1, H, r, Zm);
}

// ---------------------------------------------------------------------------------------------------------
// Per-signature table {1..8} * R in AFFINE coordinates (signed 4-bit windows).  Built as Jacobian points, then
// normalised with ONE shared inversion (Montgomery's trick; the inversion itself is the cheap safegcd one), so every
// addition in the main loop is a mixed addition (8M + 3S instead of 12M + 4S): 66 additions x 5 products saved for
// ~70 products of conversion work per signature on secp256k1.
// Scratch layout in 16-byte chunks: entry e (0-based, (e+1)*R): x at 4e, y at 4e+2  (32 chunks);
//   Z_j (j = 2..8) at 32 + 2(j-2);  prefix products c_j = Z_2...Z_j at 46 + 2(j-2)   -> 60 chunks (960 B) per thread.
// ---------------------------------------------------------------------------------------------------------
static constexpr int kSwTabEntries = 8;
static constexpr int kSwTabChunks = 60;
static constexpr int kSwTabZ = 32, kSwTabC = 46;

// Step 1 of the table: the Jacobian multiples 2R..8R (X, Y parked in their final slots, Z in the temp area) and the
// running product of their Z (prefix products stored per entry).  `c` is the running product on entry and exit, so the
// chain can span several signatures' tables (one shared inversion for all of them).
// kTwistA: (x, y) lives on a curve with the coefficient a = a_tw instead of the curve's own (the twist of group.cuh).
template <class C, bool kTwistA>
SG_HD void sw_table_park_a(const TabRef& tab, const Fe& x, const Fe& y, Fe& c, const Fe& a_tw) {
    typedef typename C::F F;
    tab_store_fe(tab, 0, x);
    tab_store_fe(tab, 2, y);
    {
        JacPoint P1, P2, P3, P4, T;
        P1.X = x;
        P1.Y = y;
        F::set_one(P1.Z);
        P1.inf = false;
#define SG_TDBL(P)                \
    do {                          \
        if (kTwistA)              \
            jac_dbl_a<C>(P, a_tw); \
        else                      \
            jac_dbl<C>(P);        \
    } while (0)
        P2 = P1;
        SG_TDBL(P2);
        P3 = P2;
        jac_madd<C>(P3, x, y);
        P4 = P2;
        SG_TDBL(P4);
#define SG_PARK(e, P)                           \
    tab_store_fe(tab, 4 * (e), (P).X);          \
    tab_store_fe(tab, 4 * (e) + 2, (P).Y);      \
    tab_store_fe(tab, kSwTabZ + 2 * ((e)-1), (P).Z)
        SG_PARK(1, P2);
        SG_PARK(2, P3);
        SG_PARK(3, P4);
        T = P4;
        jac_madd<C>(T, x, y);
        SG_PARK(4, T);  // 5R
        T = P3;
        SG_TDBL(T);
        SG_PARK(5, T);  // 6R
        jac_madd<C>(T, x, y);
        SG_PARK(6, T);  // 7R
        T = P4;
        SG_TDBL(T);
        SG_PARK(7, T);  // 8R
#undef SG_PARK
#undef SG_TDBL
    }
    // prefix products (R has prime order n > 8: no multiple is infinity, every Z_j != 0)
    Fe z;
#pragma unroll 1
    for (int j = 2; j <= 8; j++) {
        tab_load_fe(z, tab, kSwTabZ + 2 * (j - 2));
        F::mul(c, c, z);
        tab_store_fe(tab, kSwTabC + 2 * (j - 2), c);
    }
}

template <class C>
SG_HD void sw_table_park(const TabRef& tab, const Fe& x, const Fe& y, Fe& c) {
    sw_table_park_a<C, false>(tab, x, y, c, x);
}

// Step 2: given inv = (running product after this table)^-1 and c_before = the running product before this table, walk
// the table back: zinv_j = inv * c_(j-1), inv <- inv * Z_j; x = X / Z^2, y = Y / Z^3.  On exit inv = c_before^-1.
template <class C>
SG_HD void sw_table_normalize(const TabRef& tab, Fe& inv, const Fe& c_before) {
    typedef typename C::F F;
    Fe z;
#pragma unroll 1
    for (int j = 8; j >= 2; j--) {
        Fe zi, zi2, t;
        if (j > 2)
            tab_load_fe(t, tab, kSwTabC + 2 * (j - 3));
        else
            t = c_before;
        F::mul(zi, inv, t);
        tab_load_fe(z, tab, kSwTabZ + 2 * (j - 2));
        F::mul(inv, inv, z);
        F::sqr(zi2, zi);
        tab_load_fe(t, tab, 4 * (j - 1));
        F::mul(t, t, zi2);
        tab_store_fe(tab, 4 * (j - 1), t);
        F::mul(zi2, zi2, zi);
        tab_load_fe(t, tab, 4 * (j - 1) + 2);
        F::mul(t, t, zi2);
        tab_store_fe(tab, 4 * (j - 1) + 2, t);
    }
}

// one table on its own (unit shims)
template <class C>
SG_HD void sw_build_table(const TabRef& tab, const Fe& x, const Fe& y) {
    typedef typename C::F F;
    Fe one, c, inv;
    F::set_one(one);
    c = one;
    sw_table_park<C>(tab, x, y, c);
    fe_inv((F*)0, inv, c);
    sw_table_normalize<C>(tab, inv, one);
}

// acc += sign(d) * |d| * R (optionally mapped through the endomorphism (x,y) -> (beta*x, y))
template <class C>
SG_HD void sw_add_from_table(JacPoint& acc, const TabRef& tab, int d, bool flip, bool endo) {
    typedef typename C::F F;
    if (d == 0) return;
    int e = (d < 0 ? -d : d) - 1;
    Fe x, y;
    tab_load_fe(x, tab, 4 * e + 0);
    tab_load_fe(y, tab, 4 * e + 2);
    if ((d < 0) != flip) F::neg(y, y);
    if (C::kGlv && endo) C::mul_beta(x, x);
    jac_madd<C>(acc, x, y);
}

// ---- the fixed-base half u1*G from the positional table (ptab.h): `pos` mixed additions, no doublings ----
// entry |d| of window j: affine x || y in the internal form (d == 0 loads entry 1; the caller discards it)
SG_HD void sw_ptab_load(Fe& x, Fe& y, const PTab& t, u32 j, int d) {
    const Q4* q = reinterpret_cast<const Q4*>(t.base + ptab_offset(t, j, d, 16));
    Q4 a = q[0], b = q[1], c = q[2], dd = q[3];
    x.v[0] = a.x; x.v[1] = a.y; x.v[2] = a.z; x.v[3] = a.w;
    x.v[4] = b.x; x.v[5] = b.y; x.v[6] = b.z; x.v[7] = b.w;
    y.v[0] = c.x; y.v[1] = c.y; y.v[2] = c.z; y.v[3] = c.w;
    y.v[4] = dd.x; y.v[5] = dd.y; y.v[6] = dd.z; y.v[7] = dd.w;
}

// acc = sum_j d_j * T[j]  =  2^-D * u1 * G  (infinity when u1 = 0).  The entry of the next window is loaded before the
// current addition so that the gather (HBM / L2, a different line per lane) hides behind ~11 field products.
template <class C>
SG_HD void sw_ptab_sum(JacPoint& acc, const u32* u1, const PTab& t) {
    typedef typename C::F F;
    acc.inf = true;
    F::set_zero(acc.X);
    F::set_zero(acc.Y);
    F::set_zero(acc.Z);
    u32 kp[9];
    ptab_recode(kp, u1, t);
    int d = ptab_pop_digit(kp, t.w);
    Fe x, y;
    sw_ptab_load(x, y, t, 0, d);
#pragma unroll 1
    for (u32 j = 0; j < t.pos; j++) {
        Fe xn = x, yn = y;
        int dn = 0;
        if (j + 1 < t.pos) {
            dn = ptab_pop_digit(kp, t.w);
            sw_ptab_load(xn, yn, t, j + 1, dn);
        }
        if (d != 0) {
            if (d < 0) F::neg(y, y);
            jac_madd<C>(acc, x, y);
        }
        x = xn;
        y = yn;
        d = dn;
    }
}

// Table generation (once per device at init, or at load in the host simulation).
// Window base B_j = 2^(w j) * pt0: `ndbl` = w j doublings of the seed, affine x || y in the internal form.
template <class C>
SG_HD void sw_ptab_base(u32* out16, u32 ndbl) {
    typedef typename C::F F;
    JacPoint P;
    F::from_table(P.X, C::pt0());
    F::from_table(P.Y, C::pt0() + 8);
    F::set_one(P.Z);
    P.inf = false;
#pragma unroll 1
    for (u32 i = 0; i < ndbl; i++) jac_dbl<C>(P);
    Fe zi, zi2, ax, ay;
    fe_inv((F*)0, zi, P.Z);
    F::sqr(zi2, zi);
    F::mul(ax, P.X, zi2);
    F::mul(zi2, zi2, zi);
    F::mul(ay, P.Y, zi2);
    F::normalize(ax, ax);
    F::normalize(ay, ay);
    copy8(out16, ax.v);
    copy8(out16 + 8, ay.v);
}

// Entry m * B (1 <= m <= 2^(w-1)) by double-and-add over the w bits of m
template <class C>
SG_HD void sw_ptab_entry(u32* out16, u32 m, u32 w, const u32* base16) {
    typedef typename C::F F;
    Fe gx, gy;
    F::from_table(gx, base16);
    F::from_table(gy, base16 + 8);
    JacPoint P;
    P.inf = true;
    F::set_zero(P.X);
    F::set_zero(P.Y);
    F::set_zero(P.Z);
#pragma unroll 1
    for (int b = (int)w - 1; b >= 0; b--) {
        jac_dbl<C>(P);
        if ((m >> b) & 1u) jac_madd<C>(P, gx, gy);
    }
    Fe zi, zi2, ax, ay;
    fe_inv((F*)0, zi, P.Z);
    F::sqr(zi2, zi);
    F::mul(ax, P.X, zi2);
    F::mul(zi2, zi2, zi);
    F::mul(ay, P.Y, zi2);
    F::normalize(ax, ax);
    F::normalize(ay, ay);
    copy8(out16, ax.v);
    copy8(out16 + 8, ay.v);
}

// Q = u1*G + u2*R.  The accumulator starts as 2^-D u1 G (sw_ptab_sum); then secp256k1 runs the two GLV halves of u2
// (<= 129 bits each) as 33 signed 4-bit windows sharing D = 128 doublings, secp256r1 one 256-bit stream of 65 windows over
// D = 256 doublings.
// Block-wide rendezvous between the phases of the per-signature program.  The fused kernels run one 512-thread block per
// SM and keep its 16 warps in the same few KB of code at any time: the whole program is ~170 KB of SASS against a
// 32 KB L1.5 / 6 KB L0 instruction cache, and letting warps drift apart cost 12-27% (profiles/r01_variants.md).
// kSync is false for the unit-test shims (non-uniform trip counts) and in the host simulation.
template <bool kSync>
SG_HD void phase_sync() {
#if SG_PTX
    if (kSync) __syncthreads();
#endif
}

// The four doublings of a window: `#pragma unroll 1` keeps the loop body in the instruction cache but pays ~1 register move
// per loop-carried limb and iteration; SG_DBL_UNROLL (1, 2 or 4) trades code size for those moves (profiles/r02_variants.md).
#ifndef SG_DBL_UNROLL
#define SG_DBL_UNROLL 1
#endif
#define SG_PRAGMA_(x) _Pragma(#x)
#define SG_PRAGMA_UNROLL(n) SG_PRAGMA_(unroll n)

template <class C, bool kSync>
SG_HD void sw_double_mul(JacPoint& acc, const u32* u1, const u32* u2, const TabRef& tab, const PTab& gt) {
#if defined(SG_HOT_DBL_ONLY)
    typedef typename C::Cold HA;  // additions with out-of-line products, doublings inlined
#else
    typedef typename C::Hot HA;
#endif
    typedef typename C::Hot H;
    sw_ptab_sum<typename C::Cold>(acc, u1, gt);
    if (C::kGlv) {
        static_assert(!C::kGlv || kPTabShiftK1 == 32 * 4, "the table's scale follows the loop's doublings");
        GlvSplit sr;
        k1_glv_split(sr, u2);
        u32 kp[2][6];
#pragma unroll
        for (int i = 0; i < 5; i++) {
            kp[0][i] = sr.k1[i];
            kp[1][i] = sr.k2[i];
        }
        kp[0][5] = kp[1][5] = 0;
        recode_offset<5, 4, 33>(kp[0]);
        recode_offset<5, 4, 33>(kp[1]);
        const bool flip[2] = {sr.neg1, sr.neg2};
#pragma unroll 1
        for (int i = 32; i >= 0; i--) {
            phase_sync<kSync>();
            if (i != 32) {
                SG_PRAGMA_UNROLL(SG_DBL_UNROLL)
                for (int d = 0; d < 4; d++) {
#if defined(SG_SYNC_DBL)
                    phase_sync<kSync>();
#endif
                    jac_dbl<H>(acc);
                }
            }
#pragma unroll 1
            for (int s = 0; s < 2; s++) sw_add_from_table<HA>(acc, tab, recode_digit<4>(kp[s], i), flip[s], s == 1);
        }
    } else {
        static_assert(C::kGlv || kPTabShiftR1 == 64 * 4, "the table's scale follows the loop's doublings");
        u32 kp[10];
#pragma unroll
        for (int i = 0; i < 8; i++) kp[i] = u2[i];
        kp[8] = kp[9] = 0;
        recode_offset<9, 4, 65>(kp);
#pragma unroll 1
        for (int i = 64; i >= 0; i--) {
            phase_sync<kSync>();
            if (i != 64) {
                SG_PRAGMA_UNROLL(SG_DBL_UNROLL)
                for (int d = 0; d < 4; d++) {
#if defined(SG_SYNC_DBL)
                    phase_sync<kSync>();
#endif
                    jac_dbl<H>(acc);
                }
            }
            sw_add_from_table<HA>(acc, tab, recode_digit<4>(kp, i), false, false);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Batched recovery: one thread walks B <= kSwBatch signatures through the per-signature program phase by phase, so that
// each of the three modular inversions (r^-1 mod n, the table normalisation, Z^-1 of the result) is paid once per B
// signatures by Montgomery's trick (3-4 products per signature instead of a ~25 k-instruction safegcd each: the
// inversions were 16% of all issued instructions, profiles/r01_ncu_instruction_mix.txt).
//
// IO supplies the rows:  io.load(j, sig_w[16], msg_w[8])  and  io.store(j, out_w[16], status).
// sig_w / msg_w are the 64 / 32 input bytes as little-endian-loaded 32-bit words; out_w is X || Y big-endian bytes (again
// as LE-loaded words), all zero when the status is 1 = invalid signature (the CPU libraries' Err(InvalidSignature)).
// Per-thread scratch, in 16-byte chunks: B tables of kSwTabChunks, then per signature the r-chain slot (2), the
// Jacobian result (6) and the Z-chain slot (2).  kSwBatch = 16 (17.9 KB of scratch per thread, 1.4 GB per device): a
// 1,048,576-signature launch is 14 passes = ONE batch; 8 measured 0.9% slower (profiles/r02_variants.md).
// ---------------------------------------------------------------------------------------------------------
#ifndef SG_BATCH
#define SG_BATCH 16
#endif
static constexpr int kSwBatch = SG_BATCH;
static constexpr int kSwBatchChunks = kSwBatch * (kSwTabChunks + 10);

SG_HD TabRef tab_offset(const TabRef& t, int chunks) {
    TabRef r;
    r.base = t.base + (size_t)chunks * t.stride;
    r.stride = t.stride;
    return r;
}

struct SwParsed {
    u32 r[8], s[8], z[8];
    u32 parity;
    bool ok;
};

// decode_signature (y parity = bit 7 of byte 32: src/wgsl/signature.wgsl:6-21, src/tests/mod.rs:151-163) + range checks.
// A rejected signature keeps walking the program on substitute values (r = s = 1) -- no early exit, every thread of the
// block reaches every barrier -- and its outputs are zeroed at the end.
template <class C>
SG_HD void sw_parse(SwParsed& p, const u32* sig_w, const u32* msg_w) {
    typedef typename C::S S;
    be_words_to_limbs(p.r, sig_w);
    u32 sw[8];
#pragma unroll
    for (int i = 0; i < 8; i++) sw[i] = sig_w[8 + i];
    p.parity = (sw[0] >> 7) & 1u;
    sw[0] &= ~0x80u;
    be_words_to_limbs(p.s, sw);
    be_words_to_limbs(p.z, msg_w);
    p.ok = !(is_zero8(p.r) || is_zero8(p.s) || !S::lt_mod(p.r) || !S::lt_mod(p.s));
    if (!p.ok) {
#pragma unroll
        for (int i = 0; i < 8; i++) p.r[i] = p.s[i] = (i == 0) ? 1u : 0u;
    }
    S::reduce_once(p.z, p.z);
}

template <class C, bool kSync, class IO>
SG_HD void sw_ecrecover_batch(int B, IO& io, const TabRef& scratch, const PTab& gt) {
    typedef typename C::F F;
    typedef typename C::S S;
    const int kSA = kSwBatch * kSwTabChunks, kQ = kSA + 2 * kSwBatch, kZP = kQ + 6 * kSwBatch;
    u32 sig_w[16], msg_w[8];
    SwParsed p;
    u32 bad = 0;  // bit j: signature j of the batch is invalid

    // ---- phase A: r_j^-1 mod n for the whole batch (Montgomery domain: values carry a factor 2^256) ----
    {
        u32 c[8], rm[8], inv[8], t[8];
        Fe tmp;
#pragma unroll 1
        for (int j = 0; j < B; j++) {
            io.load(j, sig_w, msg_w);
            sw_parse<C>(p, sig_w, msg_w);
            S::to_mont(rm, p.r);
            if (j == 0)
                copy8(c, rm);
            else
                S::mmul(c, c, rm);
            copy8(tmp.v, c);
            tab_store_fe(scratch, kSA + 2 * j, tmp);
        }
        phase_sync<kSync>();
        S::inv_plain(inv, c);  // (prod r * R)^-1 = (prod r)^-1 R^-1
        S::r3(t);
        S::mmul(inv, inv, t);  // (prod r)^-1 R
#pragma unroll 1
        for (int j = B - 1; j >= 0; j--) {
            io.load(j, sig_w, msg_w);
            sw_parse<C>(p, sig_w, msg_w);
            S::to_mont(rm, p.r);
            if (j > 0) {
                tab_load_fe(tmp, scratch, kSA + 2 * (j - 1));
                S::mmul(t, inv, tmp.v);  // r_j^-1 R
                S::mmul(inv, inv, rm);
            } else {
                copy8(t, inv);
            }
            copy8(tmp.v, t);
            tab_store_fe(scratch, kSA + 2 * j, tmp);
        }
    }
    // ---- phase B1: lift x = r, Jacobian multiples 2R..8R of every signature, one inversion, affine tables ----
    {
        Fe one, c, inv;
        F::set_one(one);
        c = one;
#pragma unroll 1
        for (int j = 0; j < B; j++) {
            phase_sync<kSync>();
            io.load(j, sig_w, msg_w);
            sw_parse<C>(p, sig_w, msg_w);
            if (!p.ok) bad |= 1u << j;
            Fe x, y, t, y2;
            F::from_plain(x, p.r);
            C::rhs(t, x);
            fe_sqrt_candidate((F*)0, y, t);
            F::sqr(y2, y);
            if (!F::eq(y2, t)) {  // x = r is not on the curve: invalid; continue with R = G
                bad |= 1u << j;
                F::from_table(x, C::gen());
                F::from_table(y, C::gen() + 8);
            }
            {
                u32 yp[8];
                F::to_plain(yp, y);
                if ((yp[0] & 1u) != p.parity) F::neg(y, y);
            }
            sw_table_park<C>(tab_offset(scratch, j * kSwTabChunks), x, y, c);
        }
        phase_sync<kSync>();
        fe_inv((F*)0, inv, c);
#pragma unroll 1
        for (int j = B - 1; j >= 0; j--) {
            Fe c_before = one;
            if (j > 0) tab_load_fe(c_before, tab_offset(scratch, (j - 1) * kSwTabChunks), kSwTabC + 2 * 6);
            sw_table_normalize<C>(tab_offset(scratch, j * kSwTabChunks), inv, c_before);
        }
    }
    // ---- phase B2: u1 = -z/r, u2 = s/r, Q = u1*G + u2*R; Jacobian results parked, running product of their Z ----
    {
        Fe one, c, inv;
        F::set_one(one);
        c = one;
#pragma unroll 1
        for (int j = 0; j < B; j++) {
            phase_sync<kSync>();
            io.load(j, sig_w, msg_w);
            sw_parse<C>(p, sig_w, msg_w);
            Fe rinv;
            tab_load_fe(rinv, scratch, kSA + 2 * j);
            u32 u1[8], u2[8];
            S::mmul(u2, rinv.v, p.s);
            S::mmul(u1, rinv.v, p.z);
            S::neg(u1, u1);
            JacPoint Q;
            sw_double_mul<C, kSync>(Q, u1, u2, tab_offset(scratch, j * kSwTabChunks), gt);
            if (Q.inf) {  // Q = infinity: invalid; keep the chain invertible
                bad |= 1u << j;
                F::set_one(Q.Z);
            }
            tab_store_fe(scratch, kQ + 6 * j, Q.X);
            tab_store_fe(scratch, kQ + 6 * j + 2, Q.Y);
            tab_store_fe(scratch, kQ + 6 * j + 4, Q.Z);
            F::mul(c, c, Q.Z);
            tab_store_fe(scratch, kZP + 2 * j, c);
        }
        phase_sync<kSync>();
        fe_inv((F*)0, inv, c);
        // ---- phase C: affine coordinates, big-endian bytes, status ----
#pragma unroll 1
        for (int j = B - 1; j >= 0; j--) {
            Fe X, Y, Z, zi, zi2, ax, ay, cp = one;
            tab_load_fe(X, scratch, kQ + 6 * j);
            tab_load_fe(Y, scratch, kQ + 6 * j + 2);
            tab_load_fe(Z, scratch, kQ + 6 * j + 4);
            if (j > 0) tab_load_fe(cp, scratch, kZP + 2 * (j - 1));
            F::mul(zi, inv, cp);
            F::mul(inv, inv, Z);
            F::sqr(zi2, zi);
            F::mul(ax, X, zi2);
            F::mul(zi2, zi2, zi);
            F::mul(ay, Y, zi2);
            u32 xp[8], yp[8], out_w[16];
            F::to_plain(xp, ax);
            F::to_plain(yp, ay);
            const bool ok = ((bad >> j) & 1u) == 0;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                out_w[i] = ok ? bswap32(xp[7 - i]) : 0u;
                out_w[8 + i] = ok ? bswap32(yp[7 - i]) : 0u;
            }
            io.store(j, out_w, ok ? 0u : 1u);
        }
    }
}

}  // namespace sigops
