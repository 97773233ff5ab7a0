// The lane-group kernels: several cooperating threads per signature for small batches (latency).
//
// The reference's only strategy is one invocation = one signature (src/wgsl/main/secp256k1_ecdsa_main_0.wgsl:21-30), and so
// is this engine's throughput path (kernels.cuh).  One signature's program is a chain of ~2,900 dependent field operations:
// measured on B200, ONE warp per scheduler needs 536 cycles per 256-bit product, 378 per squaring and ~96 per modular
// addition whatever else the SM has to do (profiles/r02_latency_probe.txt: three independent products interleaved in one
// thread cost the same 536 cycles each -- a lone warp is issue-bound, not latency-bound), which puts a floor of ~0.8 ms
// under any request of up to ~19k signatures.  Splitting ONE product over lanes does not pay on this machine (exchanging and
// re-merging the partial rows costs as much as the multiplies saved); what pays is running the INDEPENDENT products of a
// group-law formula at the same time on DIFFERENT warps -- each warp on its own scheduler with its own multiplier pipe --
// and choosing formulas for depth instead of operation count:
//
//   * short Weierstrass: the complete projective formulas of Renes-Costello-Batina 2015 (the family the reference uses for
//     P-256, src/wgsl/secp256r1_curve.wgsl:34-146).  A doubling is two levels of four (a = 0) or three levels of up to six
//     (a = -3) independent products where the Jacobian dbl-2009-l / dbl-2001-b chains are 5-6 deep; a mixed addition is
//     two (three) levels of five-six.  Being complete they have no P = Q / P = -Q / infinity branches: every thread of the
//     block reaches every barrier whatever its data.
//   * ed25519: the extended twisted-Edwards formulas are four-way parallel as they stand (4 squarings then 4 products).
//
// A signature is served by kRoles threads with the same lane index in kRoles consecutive warps of one block ("roles"; role r
// = warp r).  Roles exchange field elements through a per-block shared-memory mailbox and meet at block barriers; all roles
// hold the same accumulator between group operations.  Role 0 also lifts R (square-root chain) and builds the
// per-signature table while role 1 computes r^-1 mod n, u1, u2 and the GLV split / window recoding.
// Bit-exact with the one-thread-per-signature kernels by construction (same decision procedure, same field arithmetic).
#pragma once
#include <type_traits>
#include "curve_ed.cuh"

namespace sigops {

static constexpr int kGroupRolesSw = 6;  // warps per block of the ecrecover group kernel
static constexpr int kGroupRolesEd = 4;  // and of the ed25519 one
static constexpr int kGroupSigs = 32;    // signatures per block (one per lane)
static constexpr int kMbSlots = 16;      // field elements per signature in the mailbox
static constexpr int kScWords = 32;      // words per signature of the scalar mailbox (recoded digits, flags)
// per-signature work table in shared memory, 16-byte chunks interleaved across the 32 signatures of the block:
//   [0, kSwTabChunks): the affine table build area of sw_build_table;  then 8 entries x (beta*x, -y) = 32 chunks
static constexpr int kGroupSwTabChunks = kSwTabChunks + 32;
static constexpr int kGroupEdTabChunks = kEdTabChunks;
static constexpr size_t kGroupSwSmem = (size_t)kGroupSigs * (kMbSlots * 32 + kScWords * 4 + kGroupSwTabChunks * 16);
static constexpr size_t kGroupEdSmem = (size_t)kGroupSigs * (kMbSlots * 32 + kScWords * 4 + kGroupEdTabChunks * 16);

// Mailbox discipline: the first level of every group operation writes slots 0..7, 14, 15, the second level slots 8..13.  A
// role reads first-level slots between the operation's two barriers and second-level slots after the second one, i.e. before
// the first barrier of the NEXT operation -- so no slot is rewritten while a slower role may still read it.
struct GroupCtx {
    int role;
    Q4* mb;   // this signature's column of the field-element mailbox: 16-byte chunk (slot * 2 + half) * 32
    u32* sc;  // this signature's column of the scalar mailbox: word w * 32
#if !SG_PTX
    void (*host_sync)(void*);  // host simulation: a barrier over the roles of the signature
    void* host_arg;
#endif
    SG_HD void sync() const {
#if SG_PTX
        __syncthreads();
#else
        host_sync(host_arg);
#endif
    }
    // two 16-byte accesses per field element: a lone warp is issue-bound, every instruction counts
    SG_HD void put(int slot, const Fe& a) const {
        const Q4 lo = {a.v[0], a.v[1], a.v[2], a.v[3]}, hi = {a.v[4], a.v[5], a.v[6], a.v[7]};
        mb[(slot * 2) * kGroupSigs] = lo;
        mb[(slot * 2 + 1) * kGroupSigs] = hi;
    }
    SG_HD void get(Fe& a, int slot) const {
        const Q4 lo = mb[(slot * 2) * kGroupSigs], hi = mb[(slot * 2 + 1) * kGroupSigs];
        a.v[0] = lo.x;
        a.v[1] = lo.y;
        a.v[2] = lo.z;
        a.v[3] = lo.w;
        a.v[4] = hi.x;
        a.v[5] = hi.y;
        a.v[6] = hi.z;
        a.v[7] = hi.w;
    }
    SG_HD void put_word(int w, u32 v) const { sc[w * kGroupSigs] = v; }
    SG_HD u32 get_word(int w) const { return sc[w * kGroupSigs]; }
};

// ---------------------------------------------------------------------------------------------------------
// complete projective doubling (RCB 2015 algorithm 9 for a = 0, algorithm 6 for a = -3), (X : Y : Z), identity (0 : 1 : 0)
// ---------------------------------------------------------------------------------------------------------
template <class C>
SG_HD void pj_dbl_g(Fe& X, Fe& Y, Fe& Z, const GroupCtx& g) {
    typedef typename C::Hot HC;  // ColdProducts<C> (curve_sw.cuh) turns this into the out-of-line flavour
    typedef typename HC::F F;
    Fe t, u, v;
    if (C::kAIsZero) {
        // level 1: t0 = Y^2, t1 = Y Z, t2 = b3 Z^2, xy = X Y
        switch (g.role) {
            case 0:
                F::sqr(t, Y);
                g.put(0, t);  // t0
                F::template shl<3>(u, t);
                g.put(1, u);  // 8 t0
                break;
            case 1:
                F::mul(t, Y, Z);
                g.put(2, t);  // t1
                break;
            case 2:
                F::sqr(t, Z);
                HC::mul_bconst(u, t);
                g.put(3, u);  // t2 = 3b Z^2
                HC::mul_bconst3(u, t);
                g.put(4, u);  // 3 t2 = 9b Z^2 (a second small product is shorter than a doubling plus an addition)
                break;
            case 3:
                F::mul(t, X, Y);
                g.put(5, t);  // X Y
                break;
            default:
                break;
        }
        g.sync();
        // level 2: X3a = t2 (8 t0), Z3 = t1 (8 t0), Y3a = (t0 - 3 t2)(t0 + t2), X3 = 2 (t0 - 3 t2) X Y
        switch (g.role) {
            case 0:
                g.get(t, 3);
                g.get(u, 1);
                F::mul(v, t, u);
                g.put(8, v);
                break;
            case 1:
                g.get(t, 2);
                g.get(u, 1);
                F::mul(v, t, u);
                g.put(9, v);
                break;
            case 2:
                g.get(t, 0);
                g.get(u, 4);
                g.get(v, 3);
                F::sub(u, t, u);
                F::add(v, t, v);
                F::mul(v, u, v);
                g.put(10, v);
                break;
            case 3:
                g.get(t, 0);
                g.get(u, 4);
                g.get(v, 5);
                F::sub(u, t, u);
                F::mul(v, u, v);
                F::dbl(v, v);
                g.put(11, v);
                break;
            default:
                break;
        }
        g.sync();
        g.get(X, 11);
        g.get(t, 8);
        g.get(u, 10);
        F::add(Y, t, u);
        g.get(Z, 9);
    } else {
        // level 1 (six products; the two products by the curve constant b ride behind the values they scale)
        switch (g.role) {
            case 0:
                F::sqr(t, X);
                g.put(1, t);  // t0
                F::dbl(u, t);
                F::add(u, u, t);
                g.put(0, u);  // 3 t0
                break;
            case 1:
                F::sqr(t, Y);
                g.put(2, t);  // t1
                break;
            case 2:
                F::sqr(t, Z);
                HC::mul_bconst(u, t);
                g.put(4, u);  // b t2
                F::dbl(u, t);
                F::add(u, u, t);
                g.put(3, u);  // 3 t2
                break;
            case 3:
                F::mul(t, X, Y);
                F::dbl(t, t);
                g.put(5, t);  // 2 X Y
                break;
            case 4:
                F::mul(t, X, Z);
                F::dbl(t, t);
                g.put(6, t);  // 2 X Z
                HC::mul_bconst(u, t);
                g.put(7, u);  // b (2 X Z)
                break;
            default:
                F::mul(t, Y, Z);
                F::dbl(t, t);
                g.put(15, t);  // 2 Y Z  (slots 8..13 belong to the second level: the previous operation may still be read)
                break;
        }
        g.sync();
        // level 2.  Y3m = 3 (b t2 - 2XZ); X3p = t1 - Y3m; Y3p = t1 + Y3m; Z3c = 3 (b 2XZ - 3 t2 - t0); t0p = 3 t0 - 3 t2
        if (g.role <= 1) {
            g.get(t, 4);
            g.get(u, 6);
            F::sub(t, t, u);
            F::dbl(u, t);
            F::add(t, u, t);  // Y3m
            g.get(u, 2);
            F::sub(v, u, t);  // X3p
            if (g.role == 0) {
                F::add(u, u, t);  // Y3p
                F::mul(v, v, u);
                g.put(9, v);  // Y3a = X3p Y3p
            } else {
                g.get(u, 5);
                F::mul(v, v, u);
                g.put(10, v);  // X3a = X3p (2XY)
            }
        } else if (g.role <= 3) {
            g.get(t, 7);
            g.get(u, 3);
            F::sub(t, t, u);
            g.get(v, 1);
            F::sub(t, t, v);
            F::dbl(v, t);
            F::add(t, v, t);  // Z3c
            if (g.role == 2) {
                g.get(v, 0);
                F::sub(v, v, u);  // t0p
                F::mul(v, v, t);
                g.put(11, v);  // t0b = t0p Z3c
            } else {
                g.get(v, 15);
                F::mul(v, v, t);
                g.put(12, v);  // Z3d = (2YZ) Z3c
            }
        } else if (g.role == 4) {
            g.get(t, 15);
            g.get(u, 2);
            F::mul(v, t, u);
            F::template shl<2>(v, v);
            g.put(13, v);  // Z3 = 4 (2YZ) t1
        }
        g.sync();
        g.get(t, 9);
        g.get(u, 11);
        F::add(Y, t, u);
        g.get(t, 10);
        g.get(u, 12);
        F::sub(X, t, u);
        g.get(Z, 13);
    }
}

// complete mixed addition (RCB 2015 algorithm 8 for a = 0, algorithm 5 for a = -3): (X : Y : Z) += (x2, y2) affine, never the
// identity.  `commit` == false (window digit 0): every thread still walks both levels (barriers) and the result is dropped.
template <class C>
SG_HD void pj_madd_g(Fe& X, Fe& Y, Fe& Z, const Fe& x2, const Fe& y2, bool commit, const GroupCtx& g) {
    typedef typename C::Hot HC;  // ColdProducts<C> (curve_sw.cuh) turns this into the out-of-line flavour
    typedef typename HC::F F;
    Fe t, u, v, w;
    if (C::kAIsZero) {
        // Warp w issues on scheduler w % 4: warps 2 and 3 have a scheduler to themselves, (0, 4) and (1, 5) share one.  The
        // heaviest task of each level goes to a lone warp, the shared schedulers get one light and one medium task.
        switch (g.role) {
            case 2:
                F::mul(t, X, x2);
                g.put(0, t);  // t0
                F::dbl(u, t);
                F::add(u, u, t);
                g.put(1, u);  // 3 t0
                break;
            case 0:
                F::mul(t, Y, y2);
                g.put(2, t);  // t1
                break;
            case 3:
                F::add(t, x2, y2);
                F::add(u, X, Y);
                F::mul(t, t, u);
                g.put(3, t);  // (x2 + y2)(X1 + Y1)
                break;
            case 4:
                F::mul(t, y2, Z);
                F::add(t, t, Y);
                g.put(4, t);  // t4 = y2 Z1 + Y1
                break;
            case 1:
                F::mul(t, x2, Z);
                F::add(t, t, X);
                HC::mul_bconst(t, t);
                g.put(5, t);  // Y3b = 3b (x2 Z1 + X1)
                break;
            default:
                HC::mul_bconst(t, Z);
                g.put(6, t);  // t2 = 3b Z1
                break;
        }
        g.sync();
        // level 2.  t3 = t3p - t0 - t1; Z3s = t1 + t2; t1m = t1 - t2
        switch (g.role) {
            case 0:
                g.get(t, 4);
                g.get(u, 5);
                F::mul(v, t, u);
                g.put(8, v);  // X3a = t4 Y3b
                break;
            case 2:
                g.get(t, 3);
                g.get(u, 0);
                F::sub(t, t, u);
                g.get(u, 2);
                F::sub(t, t, u);  // t3
                g.get(v, 6);
                F::sub(u, u, v);  // t1m
                F::mul(v, t, u);
                g.put(9, v);  // t2b = t3 t1m
                break;
            case 1:
                g.get(t, 5);
                g.get(u, 1);
                F::mul(v, t, u);
                g.put(10, v);  // Y3a = Y3b (3 t0)
                break;
            case 4:
                g.get(t, 2);
                g.get(u, 6);
                F::sub(v, t, u);  // t1m
                F::add(w, t, u);  // Z3s
                F::mul(v, v, w);
                g.put(11, v);  // t1b = t1m Z3s
                break;
            case 3:
                g.get(t, 3);
                g.get(u, 0);
                F::sub(t, t, u);
                g.get(u, 2);
                F::sub(t, t, u);  // t3
                g.get(u, 1);
                F::mul(v, t, u);
                g.put(12, v);  // t0b = (3 t0) t3
                break;
            default:
                g.get(t, 2);
                g.get(u, 6);
                F::add(t, t, u);  // Z3s
                g.get(u, 4);
                F::mul(v, t, u);
                g.put(13, v);  // Z3a = Z3s t4
                break;
        }
        g.sync();
        if (commit) {
            g.get(t, 9);
            g.get(u, 8);
            F::sub(X, t, u);
            g.get(t, 11);
            g.get(u, 10);
            F::add(Y, t, u);
            g.get(t, 13);
            g.get(u, 12);
            F::add(Z, t, u);
        }
    } else {
        switch (g.role) {
            case 0:
                F::mul(t, X, x2);
                g.put(0, t);  // t0
                F::dbl(u, t);
                F::add(u, u, t);
                g.put(1, u);  // 3 t0
                break;
            case 1:
                F::mul(t, Y, y2);
                g.put(2, t);  // t1
                break;
            case 2:
                F::add(t, x2, y2);
                F::add(u, X, Y);
                F::mul(t, t, u);
                g.put(3, t);  // (x2 + y2)(X1 + Y1)
                break;
            case 3:
                F::mul(t, y2, Z);
                F::add(t, t, Y);
                g.put(4, t);  // t4
                break;
            case 4:
                F::mul(t, x2, Z);
                F::add(t, t, X);
                g.put(5, t);  // y3 = x2 Z1 + X1
                HC::mul_bconst(u, t);
                g.put(6, u);  // b y3
                break;
            default:
                HC::mul_bconst(t, Z);
                g.put(7, t);  // b Z1
                F::dbl(u, Z);
                F::add(u, u, Z);
                g.put(14, u);  // 3 Z1
                break;
        }
        g.sync();
        // level 2.  t3 = t3p - t0 - t1; X3m = 3 (y3 - b Z1); Z3p = t1 - X3m; X3p = t1 + X3m;
        //           Y3m = 3 (b y3 - 3 Z1 - t0); t0p = 3 t0 - 3 Z1
        if (g.role == 0 || g.role == 1) {
            g.get(t, 6);
            g.get(u, 14);
            F::sub(t, t, u);
            g.get(v, 0);
            F::sub(t, t, v);
            F::dbl(v, t);
            F::add(t, v, t);  // Y3m
            if (g.role == 0) {
                g.get(v, 4);
                F::mul(v, v, t);
                g.put(8, v);  // p1 = t4 Y3m
            } else {
                g.get(v, 1);
                F::sub(v, v, u);  // t0p
                F::mul(v, v, t);
                g.put(9, v);  // p2 = t0p Y3m
            }
        } else if (g.role <= 4) {
            g.get(t, 5);
            g.get(u, 7);
            F::sub(t, t, u);
            F::dbl(u, t);
            F::add(t, u, t);  // X3m
            g.get(u, 2);
            if (g.role == 2) {
                F::sub(v, u, t);  // Z3p
                F::add(u, u, t);  // X3p
                F::mul(v, u, v);
                g.put(10, v);  // p3 = X3p Z3p
            } else if (g.role == 3) {
                F::add(v, u, t);  // X3p
                g.get(t, 3);
                g.get(w, 0);
                F::sub(t, t, w);
                F::sub(t, t, u);  // t3
                F::mul(v, t, v);
                g.put(11, v);  // p4 = t3 X3p
            } else {
                F::sub(v, u, t);  // Z3p
                g.get(t, 4);
                F::mul(v, t, v);
                g.put(12, v);  // p5 = t4 Z3p
            }
        } else {
            g.get(t, 3);
            g.get(u, 0);
            F::sub(t, t, u);
            g.get(u, 2);
            F::sub(t, t, u);  // t3
            g.get(u, 1);
            g.get(v, 14);
            F::sub(u, u, v);  // t0p
            F::mul(v, t, u);
            g.put(13, v);  // p6 = t3 t0p
        }
        g.sync();
        if (commit) {
            g.get(t, 11);
            g.get(u, 8);
            F::sub(X, t, u);
            g.get(t, 10);
            g.get(u, 9);
            F::add(Y, t, u);
            g.get(t, 12);
            g.get(u, 13);
            F::add(Z, t, u);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Q = u1*G + u2*R with the group operations above.  The schedule is sw_double_mul's (curve_sw.cuh): the positional-table
// sum first, then secp256k1's two GLV streams of u2 over 128 doublings, secp256r1's one stream over 256.  Table layout (per signature, shared memory): entry e
// ((e+1) R): x at 4e, y at 4e + 2 (sw_build_table), beta*x at kSwTabChunks + 4e, -y at kSwTabChunks + 4e + 2.
// ---------------------------------------------------------------------------------------------------------
template <class C>
SG_HD void sw_group_add_r(Fe& X, Fe& Y, Fe& Z, const TabRef& tab, int d, bool flip, bool endo, const GroupCtx& g) {
    const int e = d == 0 ? 0 : (d < 0 ? -d : d) - 1;
    Fe x, y;
    tab_load_fe(x, tab, (C::kGlv && endo) ? kSwTabChunks + 4 * e : 4 * e);
    tab_load_fe(y, tab, ((d < 0) != flip) ? kSwTabChunks + 4 * e + 2 : 4 * e + 2);
    pj_madd_g<C>(X, Y, Z, x, y, d != 0, g);
}

// kp / flips: the recoded scalars (sw_group_recode).  secp256k1: the two GLV halves of u2 (2 x 6 words, 2 flip bits), then
// u1 + the positional table's offset (9 words); secp256r1: u2 (10 words), then u1 + offset (9 words).
template <class C>
struct GroupKp {
    static constexpr int kG = C::kGlv ? 12 : 10;  // where the fixed-base scalar starts
    static constexpr int kWords = kG + 9;
};

// The fixed-base half first: X:Y:Z = 2^-D u1 G as `pos` group additions of positional-table entries (ptab.h).  Every role
// loads the entry (the same lines for all the warps of the block); the next one is requested before the current addition.
template <class C>
SG_HD void sw_group_ptab_sum(Fe& X, Fe& Y, Fe& Z, const u32* gk9, const PTab& gt, const GroupCtx& g) {
    typedef typename C::F F;
    F::set_zero(X);
    F::set_one(Y);
    F::set_zero(Z);
    u32 gk[9];
#pragma unroll
    for (int i = 0; i < 9; i++) gk[i] = gk9[i];
    int d = ptab_pop_digit(gk, gt.w);
    Fe x, y;
    sw_ptab_load(x, y, gt, 0, d);
#pragma unroll 1
    for (u32 j = 0; j < gt.pos; j++) {
        Fe xn = x, yn = y;
        int dn = 0;
        if (j + 1 < gt.pos) {
            dn = ptab_pop_digit(gk, gt.w);
            sw_ptab_load(xn, yn, gt, j + 1, dn);
        }
        if (d < 0) F::neg(y, y);
        pj_madd_g<C>(X, Y, Z, x, y, d != 0, g);
        x = xn;
        y = yn;
        d = dn;
    }
}

template <class C>
SG_HD void sw_double_mul_g(Fe& X, Fe& Y, Fe& Z, const u32* kp, u32 flips, const TabRef& tab, const PTab& gt, const GroupCtx& g) {
    sw_group_ptab_sum<C>(X, Y, Z, kp + GroupKp<C>::kG, gt, g);
    if (C::kGlv) {
#pragma unroll 1
        for (int i = 32; i >= 0; i--) {
            if (i != 32) {
#pragma unroll 1
                for (int d = 0; d < 4; d++) pj_dbl_g<C>(X, Y, Z, g);
            }
#pragma unroll 1
            for (int s = 0; s < 2; s++)
                sw_group_add_r<C>(X, Y, Z, tab, recode_digit<4>(kp + 6 * s, i), ((flips >> s) & 1u) != 0, s == 1, g);
        }
    } else {
#pragma unroll 1
        for (int i = 64; i >= 0; i--) {
            if (i != 64) {
#pragma unroll 1
                for (int d = 0; d < 4; d++) pj_dbl_g<C>(X, Y, Z, g);
            }
            sw_group_add_r<C>(X, Y, Z, tab, recode_digit<4>(kp, i), false, false, g);
        }
    }
}

// recoded scalars for sw_double_mul_g (GroupKp<C>::kWords words) and the flip bits
template <class C>
SG_HD u32 sw_group_recode(u32* kp, const u32* u1, const u32* u2, const PTab& gt) {
    u32 flips = 0;
    if (C::kGlv) {
        GlvSplit sr;
        k1_glv_split(sr, u2);
#pragma unroll
        for (int i = 0; i < 5; i++) {
            kp[i] = sr.k1[i];
            kp[6 + i] = sr.k2[i];
        }
        kp[5] = kp[11] = 0;
        recode_offset<5, 4, 33>(kp);
        recode_offset<5, 4, 33>(kp + 6);
        flips = (sr.neg1 ? 1u : 0u) | (sr.neg2 ? 2u : 0u);
    } else {
#pragma unroll
        for (int i = 0; i < 8; i++) kp[i] = u2[i];
        kp[8] = kp[9] = 0;
        recode_offset<9, 4, 65>(kp);
    }
    ptab_recode(kp + GroupKp<C>::kG, u1, gt);
    return flips;
}

// role 0's table: {1..8} R affine (sw_build_table) plus beta*x and -y per entry
template <class C>
SG_HD void sw_group_table(const TabRef& tab, const Fe& x, const Fe& y) {
    typedef typename C::F F;
    sw_build_table<C>(tab, x, y);
#pragma unroll 1
    for (int e = 0; e < kSwTabEntries; e++) {
        Fe a;
        tab_load_fe(a, tab, 4 * e);
        if (C::kGlv) C::mul_beta(a, a);
        tab_store_fe(tab, kSwTabChunks + 4 * e, a);
        tab_load_fe(a, tab, 4 * e + 2);
        F::neg(a, a);
        tab_store_fe(tab, kSwTabChunks + 4 * e + 2, a);
    }
}

// One signature, kGroupRolesSw cooperating threads.  sig_w / msg_w: the input row (every role loads it); out_w / st are
// written by role 0 (returns true there).  tab: the signature's work table (kGroupSwTabChunks chunks, shared by the roles).
template <class C>
SG_HD bool sw_ecrecover_group(const u32* sig_w, const u32* msg_w, u32* out_w, u32* st, const TabRef& tab, const PTab& gtab,
                              const GroupCtx& g) {
    typedef typename C::F F;
    typedef typename C::S S;
    SwParsed p;
    sw_parse<C>(p, sig_w, msg_w);
    constexpr int kKpWords = GroupKp<C>::kWords;
    // The table of multiples of R is built WITHOUT y, in the shadow of the square-root chain.  With w = x^3 + a x + b = y^2,
    // the map (X, Y) -> (X / w, y Y / w^2) sends the twist E'' : Y^2 = X^3 + a w^2 X + b w^3 onto the curve, and P'' = (w x, w^2)
    // on E'' goes to R = (x, y).  The Jacobian formulas never use b, so the multiples e P'' come out of the ordinary table code
    // -- with the a = 0 doubling on secp256k1, with a general-a doubling (a'' = -3 w^2) on P-256; w joins the shared inversion
    // (Montgomery's trick), and when y arrives the y coordinates are scaled by it.
    // Role 0: y (254 squarings + 13 products); role 2: the table; role 1: scalars.
    if (g.role == 0) {
        u32 bad = p.ok ? 0u : 1u;
        Fe x, y, t, y2;
        F::from_plain(x, p.r);
        C::rhs(t, x);
        fe_sqrt_candidate((F*)0, y, t);
        F::sqr(y2, y);
        if (!F::eq(y2, t)) bad = 1u;  // x = r is not on the curve: invalid (the outputs are zeroed; the table is unused)
        {
            u32 yp[8];
            F::to_plain(yp, y);
            if ((yp[0] & 1u) != p.parity) F::neg(y, y);
        }
        g.put(14, y);
        g.put_word(kKpWords + 1, bad);
    } else if (g.role == 2) {
        Fe x, w, xs, ys, c, inv;
        F::from_plain(x, p.r);
        C::rhs(w, x);
        F::mul(xs, w, x);  // P'' = (w x, w^2)
        F::sqr(ys, w);
        c = w;
        if (C::kAIsZero) {
            sw_table_park<C>(tab, xs, ys, c);
        } else {
            Fe a_tw;  // a'' = -3 w^2
            F::dbl(a_tw, ys);
            F::add(a_tw, a_tw, ys);
            F::neg(a_tw, a_tw);
            sw_table_park_a<C, true>(tab, xs, ys, c, a_tw);
        }
        fe_inv((F*)0, inv, c);
        sw_table_normalize<C>(tab, inv, w);  // leaves inv = w^-1
        Fe om2;
        F::sqr(om2, inv);
#pragma unroll 1
        for (int e = 0; e < kSwTabEntries; e++) {
            Fe a;
            tab_load_fe(a, tab, 4 * e);
            F::mul(a, a, inv);  // x_e = X''_e / w
            tab_store_fe(tab, 4 * e, a);
            if (C::kGlv) {
                C::mul_beta(a, a);
                tab_store_fe(tab, kSwTabChunks + 4 * e, a);
            }
            tab_load_fe(a, tab, 4 * e + 2);
            F::mul(a, a, om2);  // y_e / y = Y''_e / w^2
            tab_store_fe(tab, 4 * e + 2, a);
        }
    }
    if (g.role == 1) {
        // r^-1 mod n, u1 = -z/r, u2 = s/r, GLV split, window recoding
        u32 ri[8], rim[8], u1[8], u2[8], kp[kKpWords];
        S::inv_plain(ri, p.r);
        S::to_mont(rim, ri);
        S::mmul(u2, rim, p.s);
        S::mmul(u1, rim, p.z);
        S::neg(u1, u1);
        ptab_prefetch(gtab, u1, 16);  // the fixed-base entries travel to L2 while the other roles finish y and the table
        const u32 flips = sw_group_recode<C>(kp, u1, u2, gtab);
#pragma unroll
        for (int i = 0; i < kKpWords; i++) g.put_word(i, kp[i]);
        g.put_word(kKpWords, flips);
    }
    g.sync();
    {
        // y has arrived: entry e gets y_e = y (Y''_e / w^2) and -y_e, the entries spread over the roles
        Fe y;
        g.get(y, 14);
#pragma unroll 1
        for (int e = g.role; e < kSwTabEntries; e += kGroupRolesSw) {
            Fe a;
            tab_load_fe(a, tab, 4 * e + 2);
            F::mul(a, a, y);
            tab_store_fe(tab, 4 * e + 2, a);
            F::neg(a, a);
            tab_store_fe(tab, kSwTabChunks + 4 * e + 2, a);
        }
        g.sync();
    }
    u32 kp[kKpWords];
#pragma unroll
    for (int i = 0; i < kKpWords; i++) kp[i] = g.get_word(i);
    const u32 flips = g.get_word(kKpWords);
    u32 bad = g.get_word(kKpWords + 1);
    Fe X, Y, Z;
    sw_double_mul_g<C>(X, Y, Z, kp, flips, tab, gtab, g);
    if (g.role != 0) return false;
    if (F::is_zero(Z)) {  // Q = infinity: invalid
        bad = 1u;
        F::set_one(Z);
    }
    Fe zi, ax, ay;
    fe_inv((F*)0, zi, Z);
    F::mul(ax, X, zi);
    F::mul(ay, Y, zi);
    u32 xp[8], yp[8];
    F::to_plain(xp, ax);
    F::to_plain(yp, ay);
    const bool ok = bad == 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        out_w[i] = ok ? bswap32(xp[7 - i]) : 0u;
        out_w[8 + i] = ok ? bswap32(yp[7 - i]) : 0u;
    }
    *st = ok ? 0u : 1u;
    return true;
}

// ---------------------------------------------------------------------------------------------------------
// ed25519: extended twisted Edwards with four roles.  dbl-2008-hwcd = 4 squarings then 4 products, add-2008-hwcd-3 = 4
// products then 4 products (src/wgsl/ed25519_curve.wgsl:35-101 runs the same 7 / 9 products one after the other).
// ---------------------------------------------------------------------------------------------------------
template <class FE>
SG_HD void ed_dbl_g(EdPoint& P, const GroupCtx& g) {
    Fe t, u, v;
    switch (g.role) {
        case 0:
            FE::sqr(t, P.X);
            g.put(0, t);  // A
            break;
        case 1:
            FE::sqr(t, P.Y);
            g.put(1, t);  // B
            break;
        case 2:
            FE::sqr(t, P.Z);
            FE::dbl(t, t);
            g.put(2, t);  // C = 2 Z^2
            break;
        default:
            FE::add(t, P.X, P.Y);
            FE::sqr(t, t);
            g.put(3, t);  // (X + Y)^2
            break;
    }
    g.sync();
    // E = t - A - B, G = B - A, F = G - C, H = -(A + B);  X3 = E F, Y3 = G H, Z3 = F G, T3 = E H
    Fe A, B;
    g.get(A, 0);
    g.get(B, 1);
    switch (g.role) {
        case 0:
            g.get(t, 3);
            FE::sub(t, t, A);
            FE::sub(t, t, B);  // E
            FE::sub(u, B, A);
            g.get(v, 2);
            FE::sub(u, u, v);  // F
            FE::mul(v, t, u);
            g.put(8, v);
            break;
        case 1:
            FE::sub(t, B, A);  // G
            FE::add(u, A, B);
            FE::neg(u, u);  // H
            FE::mul(v, t, u);
            g.put(9, v);
            break;
        case 2:
            FE::sub(t, B, A);  // G
            g.get(v, 2);
            FE::sub(u, t, v);  // F
            FE::mul(v, u, t);
            g.put(10, v);
            break;
        default:
            g.get(t, 3);
            FE::sub(t, t, A);
            FE::sub(t, t, B);  // E
            FE::add(u, A, B);
            FE::neg(u, u);  // H
            FE::mul(v, t, u);
            g.put(11, v);
            break;
    }
    g.sync();
    g.get(P.X, 8);
    g.get(P.Y, 9);
    g.get(P.Z, 10);
    g.get(P.T, 11);
}

// P += (or -=) Q in cached form (Y2+X2, Y2-X2, Z2, 2d*T2); has_z == false: affine Niels (Z2 = 1)
template <class FE>
SG_HD void ed_add_g(EdPoint& P, const Fe& ypx, const Fe& ymx, const Fe& z2, const Fe& t2d, bool has_z, bool negq, bool commit,
                    const GroupCtx& g) {
    Fe t, u, v;
    switch (g.role) {
        case 0:
            FE::sub(t, P.Y, P.X);
            FE::mul(t, t, negq ? ypx : ymx);
            g.put(0, t);  // A
            break;
        case 1:
            FE::add(t, P.Y, P.X);
            FE::mul(t, t, negq ? ymx : ypx);
            g.put(1, t);  // B
            break;
        case 2:
            FE::mul(t, P.T, t2d);
            g.put(2, t);  // C
            break;
        default:
            if (has_z)
                FE::mul(t, P.Z, z2);
            else
                t = P.Z;
            FE::dbl(t, t);
            g.put(3, t);  // D
            break;
    }
    g.sync();
    // E = B - A, H = B + A, F = D -+ C, G = D +- C
    Fe A, B, C, D;
    switch (g.role) {
        case 0:
            g.get(A, 0);
            g.get(B, 1);
            g.get(C, 2);
            g.get(D, 3);
            FE::sub(t, B, A);
            if (negq) FE::add(u, D, C); else FE::sub(u, D, C);
            FE::mul(v, t, u);
            g.put(8, v);  // X3 = E F
            break;
        case 1:
            g.get(A, 0);
            g.get(B, 1);
            g.get(C, 2);
            g.get(D, 3);
            FE::add(t, B, A);
            if (negq) FE::sub(u, D, C); else FE::add(u, D, C);
            FE::mul(v, u, t);
            g.put(9, v);  // Y3 = G H
            break;
        case 2:
            g.get(C, 2);
            g.get(D, 3);
            if (negq) {
                FE::add(t, D, C);
                FE::sub(u, D, C);
            } else {
                FE::sub(t, D, C);
                FE::add(u, D, C);
            }
            FE::mul(v, t, u);
            g.put(10, v);  // Z3 = F G
            break;
        default:
            g.get(A, 0);
            g.get(B, 1);
            FE::sub(t, B, A);
            FE::add(u, B, A);
            FE::mul(v, t, u);
            g.put(11, v);  // T3 = E H
            break;
    }
    g.sync();
    if (commit) {
        g.get(P.X, 8);
        g.get(P.Y, 9);
        g.get(P.Z, 10);
        g.get(P.T, 11);
    }
}

// table entry e = (e+1)(-A) in cached form: every role writes one component (2d*T is a product)
template <class FE>
SG_HD void ed_tab_store_g(const TabRef& tab, int e, const EdPoint& P, const GroupCtx& g) {
    const Fe d2 = {SG_ED_D2};
    Fe t;
    switch (g.role) {
        case 0:
            FE::add(t, P.Y, P.X);
            tab_store_fe(tab, 8 * e + 0, t);
            break;
        case 1:
            FE::sub(t, P.Y, P.X);
            tab_store_fe(tab, 8 * e + 2, t);
            break;
        case 2:
            tab_store_fe(tab, 8 * e + 4, P.Z);
            break;
        default:
            FE::mul(t, P.T, d2);
            tab_store_fe(tab, 8 * e + 6, t);
            break;
    }
}

// One signature, kGroupRolesEd cooperating threads; the verdict is returned by role 0 (the other roles return 0).
// Role 0 decompresses A while role 1 hashes and reduces mod L; the table, the double-scalar loop and nothing else are shared.
// kCold: field products out of line (smaller role programs: faster once every SM runs a block, see ColdProducts in curve_sw.cuh).
template <bool kCold>
SG_HD u32 ed_verify_group(const u32* sig_w, const u32* msg_w, const u32* pk_w, const TabRef& tab, const PTab& btab, const GroupCtx& g) {
    typedef Sc<ModEdL> S;
#if !defined(SG_NO_HOT_INLINE)
    typedef typename std::conditional<kCold, Fp25519, Inl<Fp25519> >::type FH;
#else
    typedef Fp25519 FH;
#endif
    if (g.role == 0) {
        Fe x, y;
        bool ok = ed_decompress_xy(x, y, pk_w);
        FE::neg(x, x);  // -A
        ok = ok && S::lt_mod(sig_w + 8);
        g.put(12, x);
        g.put(13, y);
        g.put_word(20, ok ? 1u : 0u);
    } else if (g.role == 1) {
        ptab_prefetch(btab, sig_w + 8, 24);  // the fixed-base entries travel to L2 during the hash and the decompression
        u32 pre[24], dig[16], k[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            pre[i] = sig_w[i];
            pre[8 + i] = pk_w[i];
            pre[16 + i] = msg_w[i];
        }
        sha512_96(dig, pre);
        ed_reduce512(k, dig);
        u32 kp[10];
        copy8(kp, k);
        kp[8] = kp[9] = 0;
        recode_offset<8, 4, 64>(kp);
#pragma unroll
        for (int i = 0; i < 10; i++) g.put_word(i, kp[i]);
    }
    g.sync();
    u32 kp[10];
#pragma unroll
    for (int i = 0; i < 10; i++) kp[i] = g.get_word(i);
    bool ok = g.get_word(20) != 0;
    // table {1..8} (-A): P1, 2P1, 3P1 = 2P1 + P1, 4P1 = 2 (2P1), 5P1 = 4P1 + P1, 6P1 = 2 (3P1), 7P1 = 6P1 + P1, 8P1 = 2 (4P1)
    {
        EdPoint P1, P2, P3, P4, T;
        g.get(P1.X, 12);
        g.get(P1.Y, 13);
        FE::set_one(P1.Z);
        Fe ypx, ymx, t2d, one;
        FE::set_one(one);
        FE::mul(P1.T, P1.X, P1.Y);
        {
            const Fe d2 = {SG_ED_D2};
            FE::add(ypx, P1.Y, P1.X);
            FE::sub(ymx, P1.Y, P1.X);
            FE::mul(t2d, P1.T, d2);
        }
        ed_tab_store_g<FH>(tab, 0, P1, g);
        P2 = P1;
        ed_dbl_g<FH>(P2, g);
        ed_tab_store_g<FH>(tab, 1, P2, g);
        P3 = P2;
        ed_add_g<FH>(P3, ypx, ymx, one, t2d, false, false, true, g);
        ed_tab_store_g<FH>(tab, 2, P3, g);
        P4 = P2;
        ed_dbl_g<FH>(P4, g);
        ed_tab_store_g<FH>(tab, 3, P4, g);
        T = P4;
        ed_add_g<FH>(T, ypx, ymx, one, t2d, false, false, true, g);
        ed_tab_store_g<FH>(tab, 4, T, g);
        T = P3;
        ed_dbl_g<FH>(T, g);
        ed_tab_store_g<FH>(tab, 5, T, g);
        ed_add_g<FH>(T, ypx, ymx, one, t2d, false, false, true, g);
        ed_tab_store_g<FH>(tab, 6, T, g);
        T = P4;
        ed_dbl_g<FH>(T, g);
        ed_tab_store_g<FH>(tab, 7, T, g);
    }
    g.sync();  // the table is complete before anyone reads it
    // the fixed-base half first: acc = 2^-252 [s]B from the positional table (ptab.h), every role loading the entries
    EdPoint acc;
    ed_set_identity(acc);
    {
        u32 sk[9];
        ptab_recode(sk, sig_w + 8, btab);
        int db = ptab_pop_digit(sk, btab.w);
        Fe bypx, bymx, bxy2d;
        ed_ptab_load(bypx, bymx, bxy2d, btab, 0, db);
#pragma unroll 1
        for (u32 j = 0; j < btab.pos; j++) {
            Fe a = bypx, b = bymx, c = bxy2d;
            int dn = 0;
            if (j + 1 < btab.pos) {
                dn = ptab_pop_digit(sk, btab.w);
                ed_ptab_load(a, b, c, btab, j + 1, dn);
            }
            ed_add_g<FH>(acc, bypx, bymx, bypx, bxy2d, false, db < 0, db != 0, g);
            bypx = a;
            bymx = b;
            bxy2d = c;
            db = dn;
        }
    }
#pragma unroll 1
    for (int i = 63; i >= 0; i--) {
        if (i != 63) {
#pragma unroll 1
            for (int d = 0; d < 4; d++) ed_dbl_g<FH>(acc, g);
        }
        {
            const int d = recode_digit<4>(kp, i);
            const int e = d == 0 ? 0 : (d < 0 ? -d : d) - 1;
            Fe ypx, ymx, z2, t2d;
            tab_load_fe(ypx, tab, 8 * e + 0);
            tab_load_fe(ymx, tab, 8 * e + 2);
            tab_load_fe(z2, tab, 8 * e + 4);
            tab_load_fe(t2d, tab, 8 * e + 6);
            ed_add_g<FH>(acc, ypx, ymx, z2, t2d, true, d < 0, d != 0, g);
        }
    }
    if (g.role != 0) return 0u;
    if (!ok || FE::is_zero(acc.Z)) {
        ok = false;
        FE::set_one(acc.Z);
    }
    Fe zi;
    fe_inv((FE*)0, zi, acc.Z);
    return ed_verify_finish(acc, zi, sig_w, ok);
}

// ---- unit shims (the lane-group twins of K1/R1_DOUBLE_MUL and ED_MULPT): every role calls them; role 0 returns true ----
template <class C>
SG_HD bool unit_double_mul_g(u32* out, const u32* u1, const u32* u2, const u32* xy, const TabRef& tab, const PTab& gtab,
                             const GroupCtx& g) {
    typedef typename C::F F;
    constexpr int kKpWords = GroupKp<C>::kWords;
    if (g.role == 0) {
        Fe x, y;
        F::from_plain(x, xy);
        F::from_plain(y, xy + 8);
        sw_group_table<C>(tab, x, y);
    } else if (g.role == 1) {
        u32 kp[kKpWords];
        const u32 flips = sw_group_recode<C>(kp, u1, u2, gtab);
#pragma unroll
        for (int i = 0; i < kKpWords; i++) g.put_word(i, kp[i]);
        g.put_word(kKpWords, flips);
    }
    g.sync();
    u32 kp[kKpWords];
#pragma unroll
    for (int i = 0; i < kKpWords; i++) kp[i] = g.get_word(i);
    const u32 flips = g.get_word(kKpWords);
    Fe X, Y, Z;
    sw_double_mul_g<C>(X, Y, Z, kp, flips, tab, gtab, g);
    if (g.role != 0) return false;
    for (int i = 0; i < 17; i++) out[i] = 0;
    if (F::is_zero(Z)) {
        out[16] = 1;
        return true;
    }
    Fe zi, ax, ay;
    fe_inv((F*)0, zi, Z);
    F::mul(ax, X, zi);
    F::mul(ay, Y, zi);
    F::to_plain(out, ax);
    F::to_plain(out + 8, ay);
    return true;
}

// k * (x, y) on ed25519 with the group operations (table of cached multiples, signed 4-bit windows): in = k, x, y
SG_HD bool unit_ed_mulpt_g(u32* out, const u32* in, const TabRef& tab, const GroupCtx& g) {
#if !defined(SG_NO_HOT_INLINE)
    typedef Inl<Fp25519> FH;
#else
    typedef Fp25519 FH;
#endif
    EdPoint P1, P2, P3, P4, T, acc;
    Fp25519::from_plain(P1.X, in + 8);
    Fp25519::from_plain(P1.Y, in + 16);
    FE::set_one(P1.Z);
    FE::mul(P1.T, P1.X, P1.Y);
    Fe ypx, ymx, t2d, one;
    FE::set_one(one);
    {
        const Fe d2 = {SG_ED_D2};
        FE::add(ypx, P1.Y, P1.X);
        FE::sub(ymx, P1.Y, P1.X);
        FE::mul(t2d, P1.T, d2);
    }
    ed_tab_store_g<FH>(tab, 0, P1, g);
    P2 = P1;
    ed_dbl_g<FH>(P2, g);
    ed_tab_store_g<FH>(tab, 1, P2, g);
    P3 = P2;
    ed_add_g<FH>(P3, ypx, ymx, one, t2d, false, false, true, g);
    ed_tab_store_g<FH>(tab, 2, P3, g);
    P4 = P2;
    ed_dbl_g<FH>(P4, g);
    ed_tab_store_g<FH>(tab, 3, P4, g);
    T = P4;
    ed_add_g<FH>(T, ypx, ymx, one, t2d, false, false, true, g);
    ed_tab_store_g<FH>(tab, 4, T, g);
    T = P3;
    ed_dbl_g<FH>(T, g);
    ed_tab_store_g<FH>(tab, 5, T, g);
    ed_add_g<FH>(T, ypx, ymx, one, t2d, false, false, true, g);
    ed_tab_store_g<FH>(tab, 6, T, g);
    T = P4;
    ed_dbl_g<FH>(T, g);
    ed_tab_store_g<FH>(tab, 7, T, g);
    g.sync();
    u32 kp[9];
    for (int i = 0; i < 8; i++) kp[i] = in[i];
    kp[8] = 0;
    recode_offset<9, 4, 65>(kp);
    ed_set_identity(acc);
#pragma unroll 1
    for (int i = 64; i >= 0; i--) {
        if (i != 64) {
#pragma unroll 1
            for (int d = 0; d < 4; d++) ed_dbl_g<FH>(acc, g);
        }
        const int d = recode_digit<4>(kp, i);
        const int e = d == 0 ? 0 : (d < 0 ? -d : d) - 1;
        Fe a, b, z2, c;
        tab_load_fe(a, tab, 8 * e + 0);
        tab_load_fe(b, tab, 8 * e + 2);
        tab_load_fe(z2, tab, 8 * e + 4);
        tab_load_fe(c, tab, 8 * e + 6);
        ed_add_g<FH>(acc, a, b, z2, c, true, d < 0, d != 0, g);
    }
    if (g.role != 0) return false;
    Fe zi, ax, ay;
    fe_inv((FE*)0, zi, acc.Z);
    FE::mul(ax, acc.X, zi);
    FE::mul(ay, acc.Y, zi);
    FE::to_plain(out, ax);
    FE::to_plain(out + 8, ay);
    return true;
}

}  // namespace sigops
