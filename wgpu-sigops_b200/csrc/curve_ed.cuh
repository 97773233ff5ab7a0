// ed25519: SHA-512 challenge, mod-L reduction, point decompression, extended twisted-Edwards group law and the
// fused cofactorless verification  R' = [s]B + [k](-A);  valid <=> compress(R') == R bytes.
//
// Replaces the reference's six-stage pipeline (src/ed25519_eddsa.rs:67-257; stages
// src/wgsl/main/ed25519_eddsa_main_0..5.wgsl) and its device functions:
//   src/wgsl/sha512.wgsl:114-194            `sha512_96` (64-bit words emulated as vec2<u32>)
//   src/wgsl/ed25519_reduce_fr.wgsl:88-125  Barrett reduction of the 512-bit hash mod L
//   src/wgsl/ed25519_utils.wgsl:42-118      `sqrt_ratio_i`, `reconstruct_ete_from_y` (dalek port)
//   src/wgsl/ed25519_curve.wgsl:35-101      add-2008-hwcd-3 / dbl-2008-hwcd
//   src/wgsl/ed25519_curve.wgsl:141-292     `ete_mul`, `ete_fixed_mul`;  :19-32,204-228 compress / to-affine
//   src/wgsl/ed25519_eddsa.wgsl:34-63       `compute_neg_a_pt`, `ed25519_verify`
// Semantics follow ed25519-dalek 2.1.1 `VerifyingKey::verify` (what src/tests/ed25519_eddsa.rs:26 asserts):
// s must be canonical (< L); A is decompressed without a canonicity check on y; no cofactor / small-order checks;
// R is never decompressed -- the recomputed point's encoding is compared byte-for-byte.
#pragma once
#include "curve_sw.cuh"

namespace sigops {

// ---------------------------------------------------------------------------------------------------------
// SHA-512 of exactly 96 bytes (R || A || M): one block, padding fixed (src/wgsl/sha512.wgsl:122-123)
// ---------------------------------------------------------------------------------------------------------
#define SG_SHA512_K                                                                                             \
    {                                                                                                           \
        0x428a2f98d728ae22ull, 0x7137449123ef65cdull, 0xb5c0fbcfec4d3b2full, 0xe9b5dba58189dbbcull,             \
            0x3956c25bf348b538ull, 0x59f111f1b605d019ull, 0x923f82a4af194f9bull, 0xab1c5ed5da6d8118ull,         \
            0xd807aa98a3030242ull, 0x12835b0145706fbeull, 0x243185be4ee4b28cull, 0x550c7dc3d5ffb4e2ull,         \
            0x72be5d74f27b896full, 0x80deb1fe3b1696b1ull, 0x9bdc06a725c71235ull, 0xc19bf174cf692694ull,         \
            0xe49b69c19ef14ad2ull, 0xefbe4786384f25e3ull, 0x0fc19dc68b8cd5b5ull, 0x240ca1cc77ac9c65ull,         \
            0x2de92c6f592b0275ull, 0x4a7484aa6ea6e483ull, 0x5cb0a9dcbd41fbd4ull, 0x76f988da831153b5ull,         \
            0x983e5152ee66dfabull, 0xa831c66d2db43210ull, 0xb00327c898fb213full, 0xbf597fc7beef0ee4ull,         \
            0xc6e00bf33da88fc2ull, 0xd5a79147930aa725ull, 0x06ca6351e003826full, 0x142929670a0e6e70ull,         \
            0x27b70a8546d22ffcull, 0x2e1b21385c26c926ull, 0x4d2c6dfc5ac42aedull, 0x53380d139d95b3dfull,         \
            0x650a73548baf63deull, 0x766a0abb3c77b2a8ull, 0x81c2c92e47edaee6ull, 0x92722c851482353bull,         \
            0xa2bfe8a14cf10364ull, 0xa81a664bbc423001ull, 0xc24b8b70d0f89791ull, 0xc76c51a30654be30ull,         \
            0xd192e819d6ef5218ull, 0xd69906245565a910ull, 0xf40e35855771202aull, 0x106aa07032bbd1b8ull,         \
            0x19a4c116b8d2d0c8ull, 0x1e376c085141ab53ull, 0x2748774cdf8eeb99ull, 0x34b0bcb5e19b48a8ull,         \
            0x391c0cb3c5c95a63ull, 0x4ed8aa4ae3418acbull, 0x5b9cca4f7763e373ull, 0x682e6ff3d6b2b8a3ull,         \
            0x748f82ee5defb2fcull, 0x78a5636f43172f60ull, 0x84c87814a1f0ab72ull, 0x8cc702081a6439ecull,         \
            0x90befffa23631e28ull, 0xa4506cebde82bde9ull, 0xbef9a3f7b2c67915ull, 0xc67178f2e372532bull,         \
            0xca273eceea26619cull, 0xd186b8c721c0c207ull, 0xeada7dd6cde0eb1eull, 0xf57d4f7fee6ed178ull,         \
            0x06f067aa72176fbaull, 0x0a637dc5a2c898a6ull, 0x113f9804bef90daeull, 0x1b710b35131c471bull,         \
            0x28db77f523047d84ull, 0x32caab7b40c72493ull, 0x3c9ebe0a15c9bebcull, 0x431d67c49c100d4cull,         \
            0x4cc5d4becb3e42b6ull, 0x597f299cfc657e2aull, 0x5fcb6fab3ad6faecull, 0x6c44198c4a475817ull          \
    }

#if defined(__CUDACC__)
static __device__ const u64 sha512_k_dev[80] = SG_SHA512_K;
#endif
static const u64 sha512_k_host[80] = SG_SHA512_K;

SG_HD u64 sha512_k(int i) {
#if SG_PTX
    return sha512_k_dev[i];
#else
    return sha512_k_host[i];
#endif
}

SG_HD u64 rotr64(u64 x, int n) { return (x >> n) | (x << (64 - n)); }

#define SG_SHA512_IV                                                                                          \
    {                                                                                                         \
        0x6a09e667f3bcc908ull, 0xbb67ae8584caa73bull, 0x3c6ef372fe94f82bull, 0xa54ff53a5f1d36f1ull,           \
            0x510e527fade682d1ull, 0x9b05688c2b3e6c1full, 0x1f83d9abfb41bd6bull, 0x5be0cd19137e2179ull        \
    }

// one compression: st += F(st, w); w is consumed (rolling 16-word schedule)
SG_HD void sha512_compress(u64* st, u64* w) {
    u64 a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
#pragma unroll 1
    for (int t0 = 0; t0 < 80; t0 += 16) {
#pragma unroll
        for (int j = 0; j < 16; j++) {
            if (t0 > 0) {
                u64 w15 = w[(j + 1) & 15], w2 = w[(j + 14) & 15];
                u64 s0 = rotr64(w15, 1) ^ rotr64(w15, 8) ^ (w15 >> 7);
                u64 s1 = rotr64(w2, 19) ^ rotr64(w2, 61) ^ (w2 >> 6);
                w[j] = w[j] + s0 + w[(j + 9) & 15] + s1;
            }
            u64 S1 = rotr64(e, 14) ^ rotr64(e, 18) ^ rotr64(e, 41);
            u64 ch = (e & f) ^ (~e & g);
            u64 t1 = h + S1 + ch + sha512_k(t0 + j) + w[j];
            u64 S0 = rotr64(a, 28) ^ rotr64(a, 34) ^ rotr64(a, 39);
            u64 maj = (a & b) ^ (a & c) ^ (b & c);
            u64 t2 = S0 + maj;
            h = g;
            g = f;
            f = e;
            e = d + t1;
            d = c;
            c = b;
            b = a;
            a = t1 + t2;
        }
    }
    st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
}

SG_HD void sha512_digest_words(u32* out_w, const u64* st) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
        out_w[2 * i] = bswap32((u32)(st[i] >> 32));
        out_w[2 * i + 1] = bswap32((u32)st[i]);
    }
}

// in_w: 24 little-endian-loaded 32-bit words of the 96 message bytes.  out: 64 digest bytes as 16 LE-loaded words.
SG_HD void sha512_96(u32* out_w, const u32* in_w) {
    u64 w[16];
#pragma unroll
    for (int i = 0; i < 12; i++) w[i] = ((u64)bswap32(in_w[2 * i]) << 32) | bswap32(in_w[2 * i + 1]);
    w[12] = 0x8000000000000000ull;
    w[13] = 0;
    w[14] = 0;
    w[15] = 768;  // message length in bits
    u64 st[8] = SG_SHA512_IV;
    sha512_compress(st, w);
    sha512_digest_words(out_w, st);
}

// SHA-512(R || A || M) for a message of any length read bytewise from memory (fuel_crypto::ed25519::verify takes
// arbitrary-length messages; the reference hard-wires 32 bytes, src/wgsl/sha512.wgsl:114-123).
SG_HD void sha512_ram(u32* out_w, const u32* r_w, const u32* a_w, const uint8_t* msg, size_t len) {
    u64 st[8] = SG_SHA512_IV;
    u64 w[16];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        w[i] = ((u64)bswap32(r_w[2 * i]) << 32) | bswap32(r_w[2 * i + 1]);
        w[4 + i] = ((u64)bswap32(a_w[2 * i]) << 32) | bswap32(a_w[2 * i + 1]);
    }
#pragma unroll
    for (int i = 8; i < 16; i++) w[i] = 0;
    const size_t total = 64 + len;
    size_t pos = 64;  // bytes placed so far (stream position)
    size_t mi = 0;
    bool pad_done = false, len_done = false;
    while (!len_done) {
        // fill the current block from byte (pos % 128) on
        int idx = (int)(pos & 127);
        while (idx < 128 && mi < len) {
            w[idx >> 3] |= (u64)msg[mi] << (56 - 8 * (idx & 7));
            idx++;
            mi++;
        }
        pos = (pos & ~(size_t)127) + (size_t)idx;
        if (idx < 128 && !pad_done) {
            w[idx >> 3] |= (u64)0x80 << (56 - 8 * (idx & 7));
            idx++;
            pad_done = true;
        }
        if (pad_done && idx <= 112) {
            w[15] |= (u64)total << 3;  // length in bits (total < 2^61)
            len_done = true;
        }
        sha512_compress(st, w);
#pragma unroll
        for (int i = 0; i < 16; i++) w[i] = 0;
        pos = (pos + 127) & ~(size_t)127;
    }
    sha512_digest_words(out_w, st);
}

// k = (512-bit little-endian integer h) mod L.  h = lo + hi*2^256 = mont(lo, R^2)... two Montgomery products:
// mont(lo, R2) = lo*R, mont(hi, R2) = hi*R, then mont(hi*R, R2) = hi*R^2/R... simpler: lo mod L + (hi*R mod L).
SG_HD void ed_reduce512(u32* k, const u32* h16) {
    typedef Sc<ModEdL> S;
    u32 lo[8], hi[8], r2[8];
    ModEdL::r2(r2);
    // x -> x*R mod L is mmul(x, R2); then from_mont(mmul(lo,R2) + ...) would undo it.  Use:
    //   lo mod L   = mmul(mmul(lo, R2), 1)        (two products, canonical)
    //   hi*R mod L = mmul(hi, R2)                 (one product, canonical)
    S::mmul(hi, h16 + 8, r2);
    S::mmul(lo, h16, r2);
    S::from_mont(lo, lo);
    u32 m[8], s[8], u[8];
    ModEdL::mod(m);
    add8(s, lo, hi);  // < 2L < 2^254: no carry
    u32 bw = sub8(u, s, m);
    select8(k, bw == 0, s, u);
}

// ---------------------------------------------------------------------------------------------------------
// extended twisted Edwards points (a = -1); formulas are complete on the whole curve (d non-square, -1 square)
// ---------------------------------------------------------------------------------------------------------
struct EdPoint {
    Fe X, Y, Z, T;
};

typedef Fp25519 FE;

SG_HD void ed_set_identity(EdPoint& P) {
    FE::set_zero(P.X);
    FE::set_one(P.Y);
    FE::set_one(P.Z);
    FE::set_zero(P.T);
}

// dbl-2008-hwcd: 4S + 3M (+1M for T when the next operation is an addition)
template <class FE>
SG_HD void ed_dbl(EdPoint& P, bool need_t) {
    Fe A, B, C, E, G, F, H, t;
    FE::sqr(A, P.X);
    FE::sqr(B, P.Y);
    FE::sqr(C, P.Z);
    FE::dbl(C, C);
    FE::add(t, P.X, P.Y);
    FE::sqr(t, t);
    FE::sub(t, t, A);
    FE::sub(E, t, B);   // E = 2XY
    FE::sub(G, B, A);   // G = -A + B  (a = -1: D = -A)
    FE::sub(F, G, C);   // F = G - C
    FE::add(H, A, B);
    FE::neg(H, H);      // H = D - B = -A - B
    FE::mul(P.X, E, F);
    FE::mul(P.Y, G, H);
    FE::mul(P.Z, F, G);
    if (need_t) FE::mul(P.T, E, H);
}

// P += (or -=) Q given in cached form (Y2+X2, Y2-X2, Z2, 2d*T2): add-2008-hwcd-3, 8M (+... T optional)
template <class FE>
SG_HD void ed_add_cached(EdPoint& P, const Fe& ypx, const Fe& ymx, const Fe& z2, const Fe& t2d, bool negq, bool need_t) {
    Fe A, B, C, D, E, F, G, H, t;
    FE::sub(t, P.Y, P.X);
    FE::mul(A, t, negq ? ypx : ymx);
    FE::add(t, P.Y, P.X);
    FE::mul(B, t, negq ? ymx : ypx);
    FE::mul(C, P.T, t2d);
    FE::mul(D, P.Z, z2);
    FE::dbl(D, D);
    FE::sub(E, B, A);
    FE::add(H, B, A);
    if (negq) {
        FE::add(F, D, C);
        FE::sub(G, D, C);
    } else {
        FE::sub(F, D, C);
        FE::add(G, D, C);
    }
    FE::mul(P.X, E, F);
    FE::mul(P.Y, G, H);
    FE::mul(P.Z, F, G);
    if (need_t) FE::mul(P.T, E, H);
}

// same with an affine Niels entry (y+x, y-x, 2d*x*y), Z2 = 1: 7M
template <class FE>
SG_HD void ed_add_niels(EdPoint& P, const Fe& ypx, const Fe& ymx, const Fe& xy2d, bool negq, bool need_t) {
    Fe A, B, C, D, E, F, G, H, t;
    FE::sub(t, P.Y, P.X);
    FE::mul(A, t, negq ? ypx : ymx);
    FE::add(t, P.Y, P.X);
    FE::mul(B, t, negq ? ymx : ypx);
    FE::mul(C, P.T, xy2d);
    FE::dbl(D, P.Z);
    FE::sub(E, B, A);
    FE::add(H, B, A);
    if (negq) {
        FE::add(F, D, C);
        FE::sub(G, D, C);
    } else {
        FE::sub(F, D, C);
        FE::add(G, D, C);
    }
    FE::mul(P.X, E, F);
    FE::mul(P.Y, G, H);
    FE::mul(P.Z, F, G);
    if (need_t) FE::mul(P.T, E, H);
}

// table entry e (0-based: (e+1)*(-A)) in cached form occupies chunks [8e, 8e+8): Y+X, Y-X, Z, 2dT
static constexpr int kEdTabEntries = 8;
static constexpr int kEdTabChunks = kEdTabEntries * 8;  // 1 KiB per thread

SG_HD void ed_tab_store(const TabRef& tab, int e, const EdPoint& P) {
    const Fe d2 = {SG_ED_D2};
    Fe t;
    FE::add(t, P.Y, P.X);
    tab_store_fe(tab, 8 * e + 0, t);
    FE::sub(t, P.Y, P.X);
    tab_store_fe(tab, 8 * e + 2, t);
    tab_store_fe(tab, 8 * e + 4, P.Z);
    FE::mul(t, P.T, d2);
    tab_store_fe(tab, 8 * e + 6, t);
}

template <class FE>
SG_HD void ed_add_from_table(EdPoint& acc, const TabRef& tab, int d, bool need_t) {
    if (d == 0) return;
    int e = (d < 0 ? -d : d) - 1;
    Fe ypx, ymx, z2, t2d;
    tab_load_fe(ypx, tab, 8 * e + 0);
    tab_load_fe(ymx, tab, 8 * e + 2);
    tab_load_fe(z2, tab, 8 * e + 4);
    tab_load_fe(t2d, tab, 8 * e + 6);
    ed_add_cached<FE>(acc, ypx, ymx, z2, t2d, d < 0, need_t);
}

SG_HD void ed_load_fe_words(Fe& a, const u32* w) {
    const Q4* q = reinterpret_cast<const Q4*>(w);
    Q4 lo = q[0], hi = q[1];
    a.v[0] = lo.x; a.v[1] = lo.y; a.v[2] = lo.z; a.v[3] = lo.w;
    a.v[4] = hi.x; a.v[5] = hi.y; a.v[6] = hi.z; a.v[7] = hi.w;
}

// ---- the fixed-base half s*B from the positional table (ptab.h): affine Niels triples (y+x, y-x, 2d*x*y), 24 words ----
SG_HD void ed_ptab_load(Fe& ypx, Fe& ymx, Fe& xy2d, const PTab& t, u32 j, int d) {
    const u32* e = t.base + ptab_offset(t, j, d, 24);
    ed_load_fe_words(ypx, e);
    ed_load_fe_words(ymx, e + 8);
    ed_load_fe_words(xy2d, e + 16);
}

// acc = sum_j d_j * T[j] = 2^-252 * s * B in extended coordinates (T valid); the next entry is loaded before the current
// addition (one 96-byte gather per lane)
template <class FE>
SG_HD void ed_ptab_sum(EdPoint& acc, const u32* s8, const PTab& t) {
    ed_set_identity(acc);
    u32 kp[9];
    ptab_recode(kp, s8, t);
    int d = ptab_pop_digit(kp, t.w);
    Fe ypx, ymx, xy2d;
    ed_ptab_load(ypx, ymx, xy2d, t, 0, d);
#pragma unroll 1
    for (u32 j = 0; j < t.pos; j++) {
        Fe a = ypx, b = ymx, c = xy2d;
        int dn = 0;
        if (j + 1 < t.pos) {
            dn = ptab_pop_digit(kp, t.w);
            ed_ptab_load(a, b, c, t, j + 1, dn);
        }
        if (d != 0) ed_add_niels<FE>(acc, ypx, ymx, xy2d, d < 0, true);
        ypx = a;
        ymx = b;
        xy2d = c;
        d = dn;
    }
}

#if SG_PTX
SG_HD const u32* ed_pt0_niels() { return ed_pt0_niels_dev; }
#else
SG_HD const u32* ed_pt0_niels() { return ed_pt0_niels_host; }
#endif

SG_HD void ed_ptab_write(u32* out24, const EdPoint& P) {
    const Fe d2 = {SG_ED_D2};
    Fe zi, x, y, t;
    fe_inv((FE*)0, zi, P.Z);
    FE::mul(x, P.X, zi);
    FE::mul(y, P.Y, zi);
    FE::add(t, y, x);
    FE::normalize(t, t);
    copy8(out24, t.v);
    FE::sub(t, y, x);
    FE::normalize(t, t);
    copy8(out24 + 8, t.v);
    FE::mul(t, x, y);
    FE::mul(t, t, d2);
    FE::normalize(t, t);
    copy8(out24 + 16, t.v);
}

// Window base B_j = 2^(w j) * 2^-252 * B as an affine Niels triple: `ndbl` doublings of the baked seed
SG_HD void ed_ptab_base(u32* out24, u32 ndbl) {
    Fe ypx, ymx, xy2d;
    copy8(ypx.v, ed_pt0_niels());
    copy8(ymx.v, ed_pt0_niels() + 8);
    copy8(xy2d.v, ed_pt0_niels() + 16);
    EdPoint P;
    ed_set_identity(P);
    ed_add_niels<FE>(P, ypx, ymx, xy2d, false, true);
#pragma unroll 1
    for (u32 i = 0; i < ndbl; i++) ed_dbl<FE>(P, false);
    ed_ptab_write(out24, P);
}

// Entry m * B_j (1 <= m <= 2^(w-1)) by double-and-add over the w bits of m.  Runs once per entry at init.
SG_HD void ed_ptab_entry(u32* out24, u32 m, u32 w, const u32* base24) {
    Fe ypx, ymx, xy2d;
    copy8(ypx.v, base24);
    copy8(ymx.v, base24 + 8);
    copy8(xy2d.v, base24 + 16);
    EdPoint P;
    ed_set_identity(P);
#pragma unroll 1
    for (int b = (int)w - 1; b >= 0; b--) {
        ed_dbl<FE>(P, true);
        if ((m >> b) & 1u) ed_add_niels<FE>(P, ypx, ymx, xy2d, false, true);
    }
    ed_ptab_write(out24, P);
}

// sqrt_ratio_i (curve25519-dalek; src/wgsl/ed25519_utils.wgsl:42-88): returns whether u/v is a square and the
// nonnegative root (or of i*u/v when it is not)
SG_HD bool ed_sqrt_ratio_i(Fe& r, const Fe& u, const Fe& v) {
    const Fe sqrtm1 = {SG_ED_SQRTM1};
    Fe v3, v7, t, check, neg_u, neg_u_i;
    FE::sqr(v3, v);
    FE::mul(v3, v3, v);
    FE::sqr(v7, v3);
    FE::mul(v7, v7, v);
    FE::mul(t, u, v7);
    fe_pow_p58(t, t);
    FE::mul(r, u, v3);
    FE::mul(r, r, t);
    FE::sqr(check, r);
    FE::mul(check, check, v);
    FE::neg(neg_u, u);
    FE::mul(neg_u_i, neg_u, sqrtm1);
    bool correct = FE::eq(check, u);
    bool flipped = FE::eq(check, neg_u);
    bool flipped_i = FE::eq(check, neg_u_i);
    if (flipped || flipped_i) FE::mul(r, r, sqrtm1);
    if (FE::is_negative(r)) FE::neg(r, r);
    return correct || flipped;
}

// curve25519-dalek CompressedEdwardsY::decompress: y from the low 255 bits (no canonicity check), x from sqrt_ratio_i
// with the sign bit applied.  Returns false when the encoding is not a curve point (x, y are then meaningless).
SG_HD bool ed_decompress_xy(Fe& x, Fe& y, const u32* enc_w) {
    Fe yy, u, v, one;
    const u32 sign = enc_w[7] >> 31;
    copy8(y.v, enc_w);
    y.v[7] &= 0x7FFFFFFFu;
    const Fe dconst = {SG_ED_D};
    FE::set_one(one);
    FE::sqr(yy, y);
    FE::sub(u, yy, one);
    FE::mul(v, yy, dconst);
    FE::add(v, v, one);
    const bool ok = ed_sqrt_ratio_i(x, u, v);
    if (sign) FE::neg(x, x);
    return ok;
}

// [8]P == identity  (dalek `is_small_order`)
SG_HD bool ed_is_small_order(const Fe& x, const Fe& y) {
    EdPoint P;
    P.X = x;
    P.Y = y;
    FE::set_one(P.Z);
    FE::set_zero(P.T);  // T is not read by the doubling
#pragma unroll 1
    for (int i = 0; i < 3; i++) ed_dbl<FE>(P, false);
    return FE::is_zero(P.X) && FE::eq(P.Y, P.Z);
}

// One signature given the challenge digest dig = SHA-512(R || A || M) (16 LE-loaded words).
// sig_w 16 words (R || s), pk_w 8 words.  kStrict adds dalek's `verify_strict` conditions: R must decompress and
// neither A nor R may have small order.  Returns 1 when the signature verifies, else 0.
// Part 1: everything up to the projective point R' = [s]B + [k](-A).  Returns whether the inputs were acceptable so far.
template <bool kSync, bool kStrict>
SG_HD bool ed_verify_point(EdPoint& acc, const u32* sig_w, const u32* pk_w, const u32* dig, const TabRef& tab, const PTab& btab) {
    typedef Sc<ModEdL> S;
#if !defined(SG_NO_HOT_INLINE)
    typedef Inl<Fp25519> FH;  // products inlined: one doubling, one cached-addition and one Niels-addition site
#else
    typedef Fp25519 FH;
#endif
#if defined(SG_HOT_DBL_ONLY)
    typedef Fp25519 FHA;
#else
    typedef FH FHA;
#endif
    phase_sync<kSync>();
    // No early exit (every thread of the block must reach every phase barrier): a key that does not decompress or a
    // non-canonical s keeps walking the program on whatever values it has and the verdict is forced to 0 at the end.
    Fe x, y;
    bool ok = ed_decompress_xy(x, y, pk_w);
    if (kStrict) {
        Fe rx, ry;
        ok = ed_decompress_xy(rx, ry, sig_w) && ok;
        ok = ok && !ed_is_small_order(rx, ry) && !ed_is_small_order(x, y);
    }
    // -A
    FE::neg(x, x);
    // canonical-scalar check on s
    ok = ok && S::lt_mod(sig_w + 8);
    phase_sync<kSync>();
    u32 k[8];
    ed_reduce512(k, dig);
    // table {1..8} * (-A), cached form
    phase_sync<kSync>();
    {
        EdPoint P1, P2, P3, P4, T;
        P1.X = x;
        P1.Y = y;
        FE::set_one(P1.Z);
        FE::mul(P1.T, x, y);
        ed_tab_store(tab, 0, P1);
        Fe ypx, ymx, t2d;
        const Fe d2 = {SG_ED_D2};
        FE::add(ypx, y, x);
        FE::sub(ymx, y, x);
        FE::mul(t2d, P1.T, d2);
        P2 = P1;
        ed_dbl<FE>(P2, true);
        ed_tab_store(tab, 1, P2);
        P3 = P2;
        ed_add_niels<FE>(P3, ypx, ymx, t2d, false, true);
        ed_tab_store(tab, 2, P3);
        P4 = P2;
        ed_dbl<FE>(P4, true);
        ed_tab_store(tab, 3, P4);
        T = P4;
        ed_add_niels<FE>(T, ypx, ymx, t2d, false, true);
        ed_tab_store(tab, 4, T);
        T = P3;
        ed_dbl<FE>(T, true);
        ed_tab_store(tab, 5, T);
        ed_add_niels<FE>(T, ypx, ymx, t2d, false, true);
        ed_tab_store(tab, 6, T);
        T = P4;
        ed_dbl<FE>(T, true);
        ed_tab_store(tab, 7, T);
    }
    // R' = [s]B + [k](-A): the accumulator starts as 2^-252 [s]B from the positional table, then 64 signed 4-bit windows
    // for k over 252 doublings
    static_assert(kPTabShiftEd == 63 * 4, "the table's scale follows the loop's doublings");
    u32 kp[10];
    copy8(kp, k);
    kp[8] = kp[9] = 0;
    recode_offset<8, 4, 64>(kp);  // k < L < 2^253: k + C < 2^256
    ed_ptab_sum<FE>(acc, sig_w + 8, btab);
#pragma unroll 1
    for (int i = 63; i >= 0; i--) {
        phase_sync<kSync>();
        if (i != 63) {
            SG_PRAGMA_UNROLL(SG_DBL_UNROLL)
            for (int d = 0; d < 4; d++) {
#if defined(SG_SYNC_DBL)
                phase_sync<kSync>();
#endif
                ed_dbl<FH>(acc, d == 3);
            }
        }
        ed_add_from_table<FHA>(acc, tab, recode_digit<4>(kp, i), true);
    }
    // a key that is not a curve point can drive Z to zero: keep the inversion chain invertible (the verdict is 0 anyway)
    if (!ok || FE::is_zero(acc.Z)) {
        ok = false;
        FE::set_one(acc.Z);
    }
    return ok;
}

// Part 2: compress R' with zi = Z^-1 and compare with the signature's R bytes
SG_HD u32 ed_verify_finish(const EdPoint& acc, const Fe& zi, const u32* sig_w, bool ok) {
    Fe ax, ay;
    FE::mul(ax, acc.X, zi);
    FE::mul(ay, acc.Y, zi);
    u32 enc[8];
    FE::to_plain(enc, ay);
    enc[7] |= (FE::is_negative(ax) ? 1u : 0u) << 31;
    return (ok && eq8(enc, sig_w)) ? 1u : 0u;
}

// One signature given the challenge digest (variable-length / strict entry point, unit shims)
template <bool kSync, bool kStrict>
SG_HD u32 ed_verify_core(const u32* sig_w, const u32* pk_w, const u32* dig, const TabRef& tab, const PTab& btab) {
    EdPoint acc;
    const bool ok = ed_verify_point<kSync, kStrict>(acc, sig_w, pk_w, dig, tab, btab);
    phase_sync<kSync>();
    Fe zi;
    fe_inv((FE*)0, zi, acc.Z);
    return ed_verify_finish(acc, zi, sig_w, ok);
}

// Batched verification of the reference's fixed-size case (32-byte messages, non-strict): one thread walks B <= kEdBatch
// signatures and pays the final inversion once (Montgomery's trick), as sw_ecrecover_batch does.
// IO:  io.load(j, sig_w[16], msg_w[8], pk_w[8])  and  io.store(j, verdict).
// Scratch in 16-byte chunks: the table {1..8}(-A) (kEdTabChunks, reused by every signature of the batch), then per
// signature X, Y, Z (6) and the Z-chain slot (2).
static constexpr int kEdBatch = SG_BATCH;
static constexpr int kEdBatchChunks = kEdTabChunks + 8 * kEdBatch;

template <bool kSync, class IO>
SG_HD void ed_verify_batch(int B, IO& io, const TabRef& scratch, const PTab& btab) {
    const int kQ = kEdTabChunks, kZP = kQ + 6 * kEdBatch;
    u32 sig_w[16], msg_w[8], pk_w[8];
    u32 good = 0;
    Fe one, c, inv;
    FE::set_one(one);
    c = one;
#pragma unroll 1
    for (int j = 0; j < B; j++) {
        phase_sync<kSync>();
        io.load(j, sig_w, msg_w, pk_w);
        u32 pre[24], dig[16];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            pre[i] = sig_w[i];
            pre[8 + i] = pk_w[i];
            pre[16 + i] = msg_w[i];
        }
        sha512_96(dig, pre);
        EdPoint acc;
        if (ed_verify_point<kSync, false>(acc, sig_w, pk_w, dig, scratch, btab)) good |= 1u << j;
        tab_store_fe(scratch, kQ + 6 * j, acc.X);
        tab_store_fe(scratch, kQ + 6 * j + 2, acc.Y);
        tab_store_fe(scratch, kQ + 6 * j + 4, acc.Z);
        FE::mul(c, c, acc.Z);
        tab_store_fe(scratch, kZP + 2 * j, c);
    }
    phase_sync<kSync>();
    fe_inv((FE*)0, inv, c);
#pragma unroll 1
    for (int j = B - 1; j >= 0; j--) {
        EdPoint acc;
        Fe zi, cp = one;
        tab_load_fe(acc.X, scratch, kQ + 6 * j);
        tab_load_fe(acc.Y, scratch, kQ + 6 * j + 2);
        tab_load_fe(acc.Z, scratch, kQ + 6 * j + 4);
        if (j > 0) tab_load_fe(cp, scratch, kZP + 2 * (j - 1));
        FE::mul(zi, inv, cp);
        FE::mul(inv, inv, acc.Z);
        io.load(j, sig_w, msg_w, pk_w);
        io.store(j, ed_verify_finish(acc, zi, sig_w, ((good >> j) & 1u) != 0));
    }
}

// The reference's fixed-size case: 32-byte message, non-strict (src/ed25519_eddsa.rs:67-73).
template <bool kSync>
SG_HD u32 ed_verify_one(const u32* sig_w, const u32* msg_w, const u32* pk_w, const TabRef& tab, const PTab& btab) {
    u32 pre[24], dig[16];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        pre[i] = sig_w[i];
        pre[8 + i] = pk_w[i];
        pre[16 + i] = msg_w[i];
    }
    sha512_96(dig, pre);
    return ed_verify_core<kSync, false>(sig_w, pk_w, dig, tab, btab);
}

// Variable-length message, optionally strict (fuel_crypto::ed25519::verify = dalek `verify_strict`).
template <bool kSync>
SG_HD u32 ed_verify_msg(const u32* sig_w, const uint8_t* msg, size_t len, const u32* pk_w, bool strict, const TabRef& tab,
                        const PTab& btab) {
    u32 dig[16];
    sha512_ram(dig, sig_w, pk_w, msg, len);
    return strict ? ed_verify_core<kSync, true>(sig_w, pk_w, dig, tab, btab)
                  : ed_verify_core<kSync, false>(sig_w, pk_w, dig, tab, btab);
}

}  // namespace sigops
