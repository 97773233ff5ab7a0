// 256-bit multi-precision primitives on 8 x 32-bit little-endian limbs.
//
// Replaces the reference's 20 x 13-bit limb WGSL bigint layer
// (src/wgsl/bigint.wgsl:13-246, src/wgsl/ff.wgsl:122-138 `bigint_mul`,
// src/wgsl/mont.wgsl:5-70 `mont_mul`) with full 32-bit limbs and PTX carry chains
// (`add.cc/addc.cc`, `mad.lo.cc/madc.hi.cc`).  ptxas fuses each lo/hi pair into one
// IMAD.WIDE.U32(.X) with the carry in a predicate register (full listings: profiles/sass/r02_kern_*.sass.gz).
//
// Every output operand of a multi-instruction asm block is EARLY-CLOBBER ("=&r" / "+&r").  Without it the compiler may let
// an output (in particular the tied input of a read-write accumulator) share a virtual register with another input that
// holds the same value -- e.g. a zero-initialised accumulator limb and a compile-time-zero limb of an operand -- and the
// block then overwrites that input before a later instruction of the same block reads it.  Seen with a squaring of the
// constant 1 (Z of an affine point) inlined into the lane-group kernels: all eight diagonal products read limb 0
// (tools/proto/sqr_one.cu, profiles/r02_variants.md).
//
// Every primitive has two bodies: inline PTX under __CUDA_ARCH__ (the product) and a
// portable uint64_t body used only when the same headers are compiled for the host by
// tests/hostsim (logic tests without a GPU) and by the CPU-only precompute_bases table
// builder.  The product's compute entry points never run the portable bodies.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SG_HD __host__ __device__ __forceinline__
#define SG_D __device__ __forceinline__
// field multiplication / squaring bodies are real functions (one copy per field): the fused kernels call them
// several thousand times per signature from a few hundred sites; inlining every site costs ~40k SASS instructions
// per kernel (I-cache thrash, 15 min ptxas).  Within one translation unit ptxas passes the 8+8 limbs in registers.
#if defined(SG_INLINE_FIELD)
#define SG_CALL __host__ __device__ __forceinline__
#else
#define SG_CALL __host__ __device__ __noinline__
#endif
#else
#define SG_HD inline
#define SG_D inline
#define SG_CALL inline
#endif

#if defined(__CUDA_ARCH__)
#define SG_PTX 1
#else
#define SG_PTX 0
#endif

namespace sigops {

typedef uint32_t u32;
typedef uint64_t u64;

// r = a + b, returns carry out
SG_HD u32 add8(u32* r, const u32* a, const u32* b) {
    u32 c;
#if SG_PTX
    asm("add.cc.u32 %0, %9, %17;\n\t"
        "addc.cc.u32 %1, %10, %18;\n\t"
        "addc.cc.u32 %2, %11, %19;\n\t"
        "addc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21;\n\t"
        "addc.cc.u32 %5, %14, %22;\n\t"
        "addc.cc.u32 %6, %15, %23;\n\t"
        "addc.cc.u32 %7, %16, %24;\n\t"
        "addc.u32 %8, 0, 0;"
        : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(r[7]), "=&r"(c)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
          "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
#else
    u64 t = 0;
    for (int i = 0; i < 8; i++) {
        t += (u64)a[i] + b[i];
        r[i] = (u32)t;
        t >>= 32;
    }
    c = (u32)t;
#endif
    return c;
}

// r = a - b, returns borrow out (1 if a < b)
SG_HD u32 sub8(u32* r, const u32* a, const u32* b) {
    u32 c;
#if SG_PTX
    asm("sub.cc.u32 %0, %9, %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32 %8, 0, 0;"
        : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(r[7]), "=&r"(c)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
          "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
    c &= 1u;  // subc.u32 x,0,0 yields 0 or 0xffffffff
#else
    u64 bw = 0;
    for (int i = 0; i < 8; i++) {
        u64 t = (u64)a[i] - b[i] - bw;
        r[i] = (u32)t;
        bw = (t >> 32) & 1;
    }
    c = (u32)bw;
#endif
    return c;
}

// r += (lo, hi) at limbs 0,1 with carry propagated through all 8 limbs; returns carry out
SG_HD u32 add8_small(u32* r, u32 lo, u32 hi) {
    u32 c;
#if SG_PTX
    asm("add.cc.u32 %0, %0, %9;\n\t"
        "addc.cc.u32 %1, %1, %10;\n\t"
        "addc.cc.u32 %2, %2, 0;\n\t"
        "addc.cc.u32 %3, %3, 0;\n\t"
        "addc.cc.u32 %4, %4, 0;\n\t"
        "addc.cc.u32 %5, %5, 0;\n\t"
        "addc.cc.u32 %6, %6, 0;\n\t"
        "addc.cc.u32 %7, %7, 0;\n\t"
        "addc.u32 %8, 0, 0;"
        : "+&r"(r[0]), "+&r"(r[1]), "+&r"(r[2]), "+&r"(r[3]), "+&r"(r[4]), "+&r"(r[5]), "+&r"(r[6]), "+&r"(r[7]), "=&r"(c)
        : "r"(lo), "r"(hi));
#else
    u64 t = (u64)r[0] + lo;
    r[0] = (u32)t;
    t = (t >> 32) + r[1] + hi;
    r[1] = (u32)t;
    t >>= 32;
    for (int i = 2; i < 8; i++) {
        t += r[i];
        r[i] = (u32)t;
        t >>= 32;
    }
    c = (u32)t;
#endif
    return c;
}

// r -= (lo, hi) at limbs 0,1 with borrow propagated; returns borrow out
SG_HD u32 sub8_small(u32* r, u32 lo, u32 hi) {
    u32 c;
#if SG_PTX
    asm("sub.cc.u32 %0, %0, %9;\n\t"
        "subc.cc.u32 %1, %1, %10;\n\t"
        "subc.cc.u32 %2, %2, 0;\n\t"
        "subc.cc.u32 %3, %3, 0;\n\t"
        "subc.cc.u32 %4, %4, 0;\n\t"
        "subc.cc.u32 %5, %5, 0;\n\t"
        "subc.cc.u32 %6, %6, 0;\n\t"
        "subc.cc.u32 %7, %7, 0;\n\t"
        "subc.u32 %8, 0, 0;"
        : "+&r"(r[0]), "+&r"(r[1]), "+&r"(r[2]), "+&r"(r[3]), "+&r"(r[4]), "+&r"(r[5]), "+&r"(r[6]), "+&r"(r[7]), "=&r"(c)
        : "r"(lo), "r"(hi));
    c &= 1u;
#else
    u64 t = (u64)r[0] - lo;
    r[0] = (u32)t;
    u64 bw = (t >> 32) & 1;
    t = (u64)r[1] - hi - bw;
    r[1] = (u32)t;
    bw = (t >> 32) & 1;
    for (int i = 2; i < 8; i++) {
        t = (u64)r[i] - bw;
        r[i] = (u32)t;
        bw = (t >> 32) & 1;
    }
    c = (u32)bw;
#endif
    return c;
}

// ---------------------------------------------------------------------------------------
// Short carry chains for the modular fix-ups.  A fix-up adds (or subtracts) a one- or two-limb constant; its carry leaves
// the low limbs about once in 2^31 operations, so the hot path stops there and hands the pending carry to a COLD
// block that ptxas cannot if-convert (SG_RARE / SG_COLD below).  (Written as `if (carry) add8_small(...)` the once-in-2^31 path
// became 8 predicated instructions issued by every field operation: 13 % of a secp256k1 recovery's issue slots.)
// ---------------------------------------------------------------------------------------
// Two forms, chosen per field by measurement (profiles/r01_variants.md):
//   CALL: the cold path is a __noinline__ function -- smallest code, but every call site carries the ABI's register
//         constraints (more spills in the P-256 kernel: -2.6 %);
//   LOOP: the cold path is inlined behind a loop whose trip count (0 or 1) the compiler cannot see -- a backward branch
//         is never if-converted and there is no call, but the code grows (ed25519: -4 % against CALL) and cicc takes longer.
#if !defined(__CUDACC__)
#define SG_COLD_CALL inline
#define SG_COLD_LOOP inline
#define SG_RARE_CALL(c) if (c)
#define SG_RARE_LOOP(c) if (c)
#else
#define SG_COLD_CALL __host__ __device__ __noinline__
#define SG_COLD_LOOP __host__ __device__ __forceinline__
#define SG_RARE_CALL(c) if (c)
#define SG_RARE_LOOP(c) for (u32 sg_rare_ = (c); sg_rare_ != 0u; --sg_rare_)
#endif
#if defined(SG_K1_LOOP)
#define SG_COLD_K1 SG_COLD_LOOP
#define SG_RARE_K1 SG_RARE_LOOP
#else
#define SG_COLD_K1 SG_COLD_CALL
#define SG_RARE_K1 SG_RARE_CALL
#endif
#if defined(SG_ED_LOOP)
#define SG_COLD_ED SG_COLD_LOOP
#define SG_RARE_ED SG_RARE_LOOP
#else
#define SG_COLD_ED SG_COLD_CALL
#define SG_RARE_ED SG_RARE_CALL
#endif
#if defined(SG_R1_CALL)
#define SG_COLD_R1 SG_COLD_CALL
#define SG_RARE_R1 SG_RARE_CALL
#else
#define SG_COLD_R1 SG_COLD_LOOP
#define SG_RARE_R1 SG_RARE_LOOP
#endif

// r[0] += v; returns the carry out of limb 0
SG_HD u32 add1_c(u32* r, u32 v) {
    u32 c;
#if SG_PTX
    asm("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, 0, 0;" : "+&r"(r[0]), "=&r"(c) : "r"(v));
#else
    u64 t = (u64)r[0] + v;
    r[0] = (u32)t;
    c = (u32)(t >> 32);
#endif
    return c;
}
SG_HD u32 sub1_b(u32* r, u32 v) {
    u32 c;
#if SG_PTX
    asm("sub.cc.u32 %0, %0, %2;\n\tsubc.u32 %1, 0, 0;" : "+&r"(r[0]), "=&r"(c) : "r"(v));
    c &= 1u;
#else
    u64 t = (u64)r[0] - v;
    r[0] = (u32)t;
    c = (u32)(t >> 32) & 1u;
#endif
    return c;
}
// r[0..2) += (lo, hi); returns the carry out of limb 1
SG_HD u32 add2_c(u32* r, u32 lo, u32 hi) {
    u32 c;
#if SG_PTX
    asm("add.cc.u32 %0, %0, %3;\n\taddc.cc.u32 %1, %1, %4;\n\taddc.u32 %2, 0, 0;"
        : "+&r"(r[0]), "+&r"(r[1]), "=&r"(c)
        : "r"(lo), "r"(hi));
#else
    u64 t = (u64)r[0] + lo;
    r[0] = (u32)t;
    t = (t >> 32) + r[1] + hi;
    r[1] = (u32)t;
    c = (u32)(t >> 32);
#endif
    return c;
}
SG_HD u32 sub2_b(u32* r, u32 lo, u32 hi) {
    u32 c;
#if SG_PTX
    asm("sub.cc.u32 %0, %0, %3;\n\tsubc.cc.u32 %1, %1, %4;\n\tsubc.u32 %2, 0, 0;"
        : "+&r"(r[0]), "+&r"(r[1]), "=&r"(c)
        : "r"(lo), "r"(hi));
    c &= 1u;
#else
    u64 t = (u64)r[0] - lo;
    r[0] = (u32)t;
    u64 bw = (t >> 32) & 1;
    t = (u64)r[1] - hi - bw;
    r[1] = (u32)t;
    c = (u32)(t >> 32) & 1u;
#endif
    return c;
}
// r[0..3) += (l0, l1, l2); returns the carry out of limb 2
SG_HD u32 add3_c(u32* r, u32 l0, u32 l1, u32 l2) {
    u32 c;
#if SG_PTX
    asm("add.cc.u32 %0, %0, %4;\n\taddc.cc.u32 %1, %1, %5;\n\taddc.cc.u32 %2, %2, %6;\n\taddc.u32 %3, 0, 0;"
        : "+&r"(r[0]), "+&r"(r[1]), "+&r"(r[2]), "=&r"(c)
        : "r"(l0), "r"(l1), "r"(l2));
#else
    u64 t = (u64)r[0] + l0;
    r[0] = (u32)t;
    t = (t >> 32) + r[1] + l1;
    r[1] = (u32)t;
    t = (t >> 32) + r[2] + l2;
    r[2] = (u32)t;
    c = (u32)(t >> 32);
#endif
    return c;
}
// cold helpers: r[K..8) += 1 / -= 1, returns the carry / borrow out of limb 7
template <int K>
SG_HD u32 inc_from(u32* r) {
    u32 c = 1;
    for (int i = K; i < 8; i++) {
        r[i] += c;
        c = c & (u32)(r[i] == 0);
    }
    return c;
}
template <int K>
SG_HD u32 dec_from(u32* r) {
    u32 b = 1;
    for (int i = K; i < 8; i++) {
        const u32 was = r[i];
        r[i] -= b;
        b = b & (u32)(was == 0);
    }
    return b;
}

// ---------------------------------------------------------------------------------------
// wide multiply-accumulate rows.  acc[0..2n) += {x0,..,x(n-1)} * b, the n wide products landing on the
// aligned limb pairs (0,1),(2,3),...; one carry chain; returns the carry out of the top limb.
// ---------------------------------------------------------------------------------------
SG_HD u32 mad_row4(u32* acc, u32 x0, u32 x1, u32 x2, u32 x3, u32 b) {
    u32 c;
#if SG_PTX
    asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
        "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
        "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
        "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
        "addc.u32 %8, 0, 0;"
        : "+&r"(acc[0]), "+&r"(acc[1]), "+&r"(acc[2]), "+&r"(acc[3]), "+&r"(acc[4]), "+&r"(acc[5]), "+&r"(acc[6]), "+&r"(acc[7]),
          "=&r"(c)
        : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(b));
#else
    const u32 x[4] = {x0, x1, x2, x3};
    u64 cy = 0;
    for (int j = 0; j < 4; j++) {
        u64 cur = ((u64)acc[2 * j + 1] << 32) | acc[2 * j];
        u64 prod = (u64)x[j] * b;
        u64 s = cur + prod;
        u64 c1 = s < cur;
        u64 s2 = s + cy;
        u64 c2 = s2 < s;
        acc[2 * j] = (u32)s2;
        acc[2 * j + 1] = (u32)(s2 >> 32);
        cy = c1 + c2;
    }
    c = (u32)cy;
#endif
    return c;
}

SG_HD u32 mad_row3(u32* acc, u32 x0, u32 x1, u32 x2, u32 b) {
    u32 c;
#if SG_PTX
    asm("mad.lo.cc.u32 %0, %7, %10, %0;\n\t"
        "madc.hi.cc.u32 %1, %7, %10, %1;\n\t"
        "madc.lo.cc.u32 %2, %8, %10, %2;\n\t"
        "madc.hi.cc.u32 %3, %8, %10, %3;\n\t"
        "madc.lo.cc.u32 %4, %9, %10, %4;\n\t"
        "madc.hi.cc.u32 %5, %9, %10, %5;\n\t"
        "addc.u32 %6, 0, 0;"
        : "+&r"(acc[0]), "+&r"(acc[1]), "+&r"(acc[2]), "+&r"(acc[3]), "+&r"(acc[4]), "+&r"(acc[5]), "=&r"(c)
        : "r"(x0), "r"(x1), "r"(x2), "r"(b));
#else
    const u32 x[3] = {x0, x1, x2};
    u64 cy = 0;
    for (int j = 0; j < 3; j++) {
        u64 cur = ((u64)acc[2 * j + 1] << 32) | acc[2 * j];
        u64 prod = (u64)x[j] * b;
        u64 s = cur + prod;
        u64 c1 = s < cur;
        u64 s2 = s + cy;
        u64 c2 = s2 < s;
        acc[2 * j] = (u32)s2;
        acc[2 * j + 1] = (u32)(s2 >> 32);
        cy = c1 + c2;
    }
    c = (u32)cy;
#endif
    return c;
}

SG_HD u32 mad_row2(u32* acc, u32 x0, u32 x1, u32 b) {
    u32 c;
#if SG_PTX
    asm("mad.lo.cc.u32 %0, %5, %7, %0;\n\t"
        "madc.hi.cc.u32 %1, %5, %7, %1;\n\t"
        "madc.lo.cc.u32 %2, %6, %7, %2;\n\t"
        "madc.hi.cc.u32 %3, %6, %7, %3;\n\t"
        "addc.u32 %4, 0, 0;"
        : "+&r"(acc[0]), "+&r"(acc[1]), "+&r"(acc[2]), "+&r"(acc[3]), "=&r"(c)
        : "r"(x0), "r"(x1), "r"(b));
#else
    const u32 x[2] = {x0, x1};
    u64 cy = 0;
    for (int j = 0; j < 2; j++) {
        u64 cur = ((u64)acc[2 * j + 1] << 32) | acc[2 * j];
        u64 prod = (u64)x[j] * b;
        u64 s = cur + prod;
        u64 c1 = s < cur;
        u64 s2 = s + cy;
        u64 c2 = s2 < s;
        acc[2 * j] = (u32)s2;
        acc[2 * j + 1] = (u32)(s2 >> 32);
        cy = c1 + c2;
    }
    c = (u32)cy;
#endif
    return c;
}

SG_HD u32 mad_row1(u32* acc, u32 x0, u32 b) {
    u32 c;
#if SG_PTX
    asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\t"
        "madc.hi.cc.u32 %1, %3, %4, %1;\n\t"
        "addc.u32 %2, 0, 0;"
        : "+&r"(acc[0]), "+&r"(acc[1]), "=&r"(c)
        : "r"(x0), "r"(b));
#else
    u64 cur = ((u64)acc[1] << 32) | acc[0];
    u64 s = cur + (u64)x0 * b;
    c = (u32)(s < cur);
    acc[0] = (u32)s;
    acc[1] = (u32)(s >> 32);
#endif
    return c;
}

// acc[0..8) += {x0..x3} * b on aligned pairs, and the chain continues into a FRESH pair: (acc[8], acc[9]) = x4 * b2 + carry.
// Appending the next row's top product to this row's chain means the carry out of the chain never has to be
// materialised in a register pair of its own (a SEL plus a zeroing IMAD.MOV on the multiplier pipe per row).
SG_HD void mad_row5(u32* acc, u32 x0, u32 x1, u32 x2, u32 x3, u32 b, u32 x4, u32 b2) {
#if SG_PTX
    asm("mad.lo.cc.u32 %0, %10, %14, %0;\n\t"
        "madc.hi.cc.u32 %1, %10, %14, %1;\n\t"
        "madc.lo.cc.u32 %2, %11, %14, %2;\n\t"
        "madc.hi.cc.u32 %3, %11, %14, %3;\n\t"
        "madc.lo.cc.u32 %4, %12, %14, %4;\n\t"
        "madc.hi.cc.u32 %5, %12, %14, %5;\n\t"
        "madc.lo.cc.u32 %6, %13, %14, %6;\n\t"
        "madc.hi.cc.u32 %7, %13, %14, %7;\n\t"
        "madc.lo.cc.u32 %8, %15, %16, 0;\n\t"
        "madc.hi.u32 %9, %15, %16, 0;"
        : "+&r"(acc[0]), "+&r"(acc[1]), "+&r"(acc[2]), "+&r"(acc[3]), "+&r"(acc[4]), "+&r"(acc[5]), "+&r"(acc[6]), "+&r"(acc[7]),
          "=&r"(acc[8]), "=&r"(acc[9])
        : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(b), "r"(x4), "r"(b2));
#else
    u32 c = mad_row4(acc, x0, x1, x2, x3, b);
    u64 p = (u64)x4 * b2 + c;
    acc[8] = (u32)p;
    acc[9] = (u32)(p >> 32);
#endif
}

// acc[0..6) += {x0,x1,x2} * b on aligned pairs; the carry is propagated into the live pair (acc[6], acc[7])
SG_HD void mad_row3c(u32* acc, u32 x0, u32 x1, u32 x2, u32 b) {
#if SG_PTX
    asm("mad.lo.cc.u32 %0, %8, %11, %0;\n\t"
        "madc.hi.cc.u32 %1, %8, %11, %1;\n\t"
        "madc.lo.cc.u32 %2, %9, %11, %2;\n\t"
        "madc.hi.cc.u32 %3, %9, %11, %3;\n\t"
        "madc.lo.cc.u32 %4, %10, %11, %4;\n\t"
        "madc.hi.cc.u32 %5, %10, %11, %5;\n\t"
        "addc.cc.u32 %6, %6, 0;\n\t"
        "addc.u32 %7, %7, 0;"
        : "+&r"(acc[0]), "+&r"(acc[1]), "+&r"(acc[2]), "+&r"(acc[3]), "+&r"(acc[4]), "+&r"(acc[5]), "+&r"(acc[6]), "+&r"(acc[7])
        : "r"(x0), "r"(x1), "r"(x2), "r"(b));
#else
    u32 c = mad_row3(acc, x0, x1, x2, b);
    u64 t = (((u64)acc[7] << 32) | acc[6]) + c;
    acc[6] = (u32)t;
    acc[7] = (u32)(t >> 32);
#endif
}

// (lo,hi) = x*b written to acc[0],acc[1] (no accumulate)
SG_HD void mul_wide(u32* acc, u32 x, u32 b) {
#if SG_PTX
    asm("mul.lo.u32 %0, %2, %3;\n\tmul.hi.u32 %1, %2, %3;" : "=&r"(acc[0]), "=&r"(acc[1]) : "r"(x), "r"(b));
#else
    u64 p = (u64)x * b;
    acc[0] = (u32)p;
    acc[1] = (u32)(p >> 32);
#endif
}

// r[0..16) = e[0..16) + (o[0..15) << 32): merge of the even/odd column accumulators
SG_HD void merge_even_odd(u32* r, const u32* e, const u32* o) {
#if SG_PTX
    r[0] = e[0];
    asm("add.cc.u32 %0, %15, %30;\n\t"
        "addc.cc.u32 %1, %16, %31;\n\t"
        "addc.cc.u32 %2, %17, %32;\n\t"
        "addc.cc.u32 %3, %18, %33;\n\t"
        "addc.cc.u32 %4, %19, %34;\n\t"
        "addc.cc.u32 %5, %20, %35;\n\t"
        "addc.cc.u32 %6, %21, %36;\n\t"
        "addc.cc.u32 %7, %22, %37;\n\t"
        "addc.cc.u32 %8, %23, %38;\n\t"
        "addc.cc.u32 %9, %24, %39;\n\t"
        "addc.cc.u32 %10, %25, %40;\n\t"
        "addc.cc.u32 %11, %26, %41;\n\t"
        "addc.cc.u32 %12, %27, %42;\n\t"
        "addc.cc.u32 %13, %28, %43;\n\t"
        "addc.u32 %14, %29, %44;"
        : "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(r[7]), "=&r"(r[8]), "=&r"(r[9]),
          "=&r"(r[10]), "=&r"(r[11]), "=&r"(r[12]), "=&r"(r[13]), "=&r"(r[14]), "=&r"(r[15])
        : "r"(e[1]), "r"(e[2]), "r"(e[3]), "r"(e[4]), "r"(e[5]), "r"(e[6]), "r"(e[7]), "r"(e[8]), "r"(e[9]), "r"(e[10]),
          "r"(e[11]), "r"(e[12]), "r"(e[13]), "r"(e[14]), "r"(e[15]), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]),
          "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]), "r"(o[8]), "r"(o[9]), "r"(o[10]), "r"(o[11]), "r"(o[12]),
          "r"(o[13]), "r"(o[14]));
#else
    r[0] = e[0];
    u64 t = 0;
    for (int k = 1; k < 16; k++) {
        t += (u64)e[k] + o[k - 1];
        r[k] = (u32)t;
        t >>= 32;
    }
#endif
}

// r[0..16) = a[0..8) * b[0..8): 64 wide MACs on even/odd column accumulators (e: products a_i*b_j with i+j even at limb
// i+j; o: i+j odd at limb i+j-1; result = e + (o << 32)).  Row j adds a*b_j; a row whose chain ends on a live limb takes
// the NEXT row's top product a7*b_(j+1) -- which lands on the fresh pair just above -- as a fifth element, and that next
// row then runs three products plus a two-limb carry propagation (mad_row5 / mad_row3c): no carry limb is ever
// materialised except the final o[14].
SG_HD void mul8x8_school(u32* r, const u32* a, const u32* b) {
    u32 e[16], o[16];
#pragma unroll
    for (int i = 0; i < 16; i++) {
        e[i] = 0;
        o[i] = 0;
    }
    // row 0: plain products
    mul_wide(e + 0, a[0], b[0]);
    mul_wide(e + 2, a[2], b[0]);
    mul_wide(e + 4, a[4], b[0]);
    mul_wide(e + 6, a[6], b[0]);
    mul_wide(o + 0, a[1], b[0]);
    mul_wide(o + 2, a[3], b[0]);
    mul_wide(o + 4, a[5], b[0]);
    mul_wide(o + 6, a[7], b[0]);
#if defined(SG_MUL_CARRY_LIMBS)
#pragma unroll
    for (int i = 1; i < 8; i += 2) {
        o[i + 7] = mad_row4(o + i - 1, a[0], a[2], a[4], a[6], b[i]);
        mad_row4(e + i + 1, a[1], a[3], a[5], a[7], b[i]);
        if (i + 1 < 8) {
            e[i + 9] = mad_row4(e + i + 1, a[0], a[2], a[4], a[6], b[i + 1]);
            mad_row4(o + i + 1, a[1], a[3], a[5], a[7], b[i + 1]);
        }
    }
#else
    // row 1: o[0..8) += a_even*b1, then (o8,o9) = a7*b2 + carry;  e[2..10) += a_odd*b1 (top pair fresh: no carry out)
    mad_row5(o + 0, a[0], a[2], a[4], a[6], b[1], a[7], b[2]);
    mad_row4(e + 2, a[1], a[3], a[5], a[7], b[1]);
#pragma unroll
    for (int j = 2; j < 8; j += 2) {
        // even row j: e[j..j+8) += a_even*b_j, then (e[j+8], e[j+9]) = a7*b_(j+1) + carry;
        //             o[j..j+6) += a1,a3,a5 * b_j, carry into the live pair (o[j+6], o[j+7]) = a7*b_j (+ earlier carry)
        mad_row5(e + j, a[0], a[2], a[4], a[6], b[j], a[7], b[j + 1]);
        mad_row3c(o + j, a[1], a[3], a[5], b[j]);
        // odd row j+1: e[j+2..j+8) += a1,a3,a5 * b_(j+1), carry into the live pair (e[j+8], e[j+9]);
        //              o[j..j+8) += a_even*b_(j+1), then (o[j+8], o[j+9]) = a7*b_(j+2) + carry (last row: carry -> o[14])
        mad_row3c(e + j + 2, a[1], a[3], a[5], b[j + 1]);
        if (j + 2 < 8)
            mad_row5(o + j, a[0], a[2], a[4], a[6], b[j + 1], a[7], b[j + 2]);
        else
            o[14] = mad_row4(o + j, a[0], a[2], a[4], a[6], b[j + 1]);
    }
#endif
    merge_even_odd(r, e, o);
}

// ---- one-level Karatsuba variant (experiment, -DSG_KARATSUBA): three 4x4 products (48 wide MACs instead of 64) and
// ~65 more add / logic instructions.  Trades multiplier-pipe time for ALU-pipe time; see profiles/r01_variants.md.
// r[0..8) = a[0..4) * b[0..4): 16 wide MACs on even/odd column accumulators, carries in explicit limbs
SG_HD void mul4x4(u32* r, const u32* a, const u32* b) {
    u32 e[8], o[8];
    mul_wide(e + 0, a[0], b[0]);
    mul_wide(e + 2, a[2], b[0]);
    mul_wide(o + 0, a[1], b[0]);
    mul_wide(o + 2, a[3], b[0]);
    e[4] = e[5] = e[6] = e[7] = 0;
    o[4] = o[5] = o[6] = o[7] = 0;
    o[4] = mad_row2(o + 0, a[0], a[2], b[1]);
    mad_row2(e + 2, a[1], a[3], b[1]);
    e[6] = mad_row2(e + 2, a[0], a[2], b[2]);
    mad_row2(o + 2, a[1], a[3], b[2]);
    o[6] = mad_row2(o + 2, a[0], a[2], b[3]);
    mad_row2(e + 4, a[1], a[3], b[3]);
    r[0] = e[0];
#if SG_PTX
    asm("add.cc.u32 %0, %7, %14;\n\t"
        "addc.cc.u32 %1, %8, %15;\n\t"
        "addc.cc.u32 %2, %9, %16;\n\t"
        "addc.cc.u32 %3, %10, %17;\n\t"
        "addc.cc.u32 %4, %11, %18;\n\t"
        "addc.cc.u32 %5, %12, %19;\n\t"
        "addc.u32 %6, %13, %20;"
        : "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(r[7])
        : "r"(e[1]), "r"(e[2]), "r"(e[3]), "r"(e[4]), "r"(e[5]), "r"(e[6]), "r"(e[7]), "r"(o[0]), "r"(o[1]), "r"(o[2]),
          "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]));
#else
    u64 t = 0;
    for (int k = 1; k < 8; k++) {
        t += (u64)e[k] + o[k - 1];
        r[k] = (u32)t;
        t >>= 32;
    }
#endif
}

// s[0..4) = x[0..4) + x[4..8), returns the carry (0 or 1)
SG_HD u32 fold_halves(u32* s, const u32* x) {
    u32 c;
#if SG_PTX
    asm("add.cc.u32 %0, %5, %9;\n\t"
        "addc.cc.u32 %1, %6, %10;\n\t"
        "addc.cc.u32 %2, %7, %11;\n\t"
        "addc.cc.u32 %3, %8, %12;\n\t"
        "addc.u32 %4, 0, 0;"
        : "=&r"(s[0]), "=&r"(s[1]), "=&r"(s[2]), "=&r"(s[3]), "=&r"(c)
        : "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(x[4]), "r"(x[5]), "r"(x[6]), "r"(x[7]));
#else
    u64 t = 0;
    for (int i = 0; i < 4; i++) {
        t += (u64)x[i] + x[i + 4];
        s[i] = (u32)t;
        t >>= 32;
    }
    c = (u32)t;
#endif
    return c;
}

SG_HD void mul8x8_kara(u32* r, const u32* a, const u32* b) {
    u32 z0[8], z2[8], zm[9], sa[4], sb[4];
    mul4x4(z0, a, b);
    mul4x4(z2, a + 4, b + 4);
    const u32 ca = fold_halves(sa, a), cb = fold_halves(sb, b);
    mul4x4(zm, sa, sb);
    // (sa + ca 2^128)(sb + cb 2^128) = zm + (ca ? sb : 0) 2^128 + (cb ? sa : 0) 2^128 + (ca & cb) 2^256
    const u32 ma = 0u - ca, mb = 0u - cb;
    u32 ta[4], tb[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        ta[i] = sb[i] & ma;
        tb[i] = sa[i] & mb;
    }
    zm[8] = ca & cb;
#if SG_PTX
    asm("add.cc.u32 %0, %0, %5;\n\t"
        "addc.cc.u32 %1, %1, %6;\n\t"
        "addc.cc.u32 %2, %2, %7;\n\t"
        "addc.cc.u32 %3, %3, %8;\n\t"
        "addc.u32 %4, %4, 0;\n\t"
        "add.cc.u32 %0, %0, %9;\n\t"
        "addc.cc.u32 %1, %1, %10;\n\t"
        "addc.cc.u32 %2, %2, %11;\n\t"
        "addc.cc.u32 %3, %3, %12;\n\t"
        "addc.u32 %4, %4, 0;"
        : "+&r"(zm[4]), "+&r"(zm[5]), "+&r"(zm[6]), "+&r"(zm[7]), "+&r"(zm[8])
        : "r"(ta[0]), "r"(ta[1]), "r"(ta[2]), "r"(ta[3]), "r"(tb[0]), "r"(tb[1]), "r"(tb[2]), "r"(tb[3]));
    // zm -= z0; zm -= z2   (the middle term a_lo b_hi + a_hi b_lo, below 2^257)
    asm("sub.cc.u32 %0, %0, %9;\n\t"
        "subc.cc.u32 %1, %1, %10;\n\t"
        "subc.cc.u32 %2, %2, %11;\n\t"
        "subc.cc.u32 %3, %3, %12;\n\t"
        "subc.cc.u32 %4, %4, %13;\n\t"
        "subc.cc.u32 %5, %5, %14;\n\t"
        "subc.cc.u32 %6, %6, %15;\n\t"
        "subc.cc.u32 %7, %7, %16;\n\t"
        "subc.u32 %8, %8, 0;\n\t"
        "sub.cc.u32 %0, %0, %17;\n\t"
        "subc.cc.u32 %1, %1, %18;\n\t"
        "subc.cc.u32 %2, %2, %19;\n\t"
        "subc.cc.u32 %3, %3, %20;\n\t"
        "subc.cc.u32 %4, %4, %21;\n\t"
        "subc.cc.u32 %5, %5, %22;\n\t"
        "subc.cc.u32 %6, %6, %23;\n\t"
        "subc.cc.u32 %7, %7, %24;\n\t"
        "subc.u32 %8, %8, 0;"
        : "+&r"(zm[0]), "+&r"(zm[1]), "+&r"(zm[2]), "+&r"(zm[3]), "+&r"(zm[4]), "+&r"(zm[5]), "+&r"(zm[6]), "+&r"(zm[7]), "+&r"(zm[8])
        : "r"(z0[0]), "r"(z0[1]), "r"(z0[2]), "r"(z0[3]), "r"(z0[4]), "r"(z0[5]), "r"(z0[6]), "r"(z0[7]), "r"(z2[0]),
          "r"(z2[1]), "r"(z2[2]), "r"(z2[3]), "r"(z2[4]), "r"(z2[5]), "r"(z2[6]), "r"(z2[7]));
    // r = z0 + zm 2^128 + z2 2^256
    r[0] = z0[0];
    r[1] = z0[1];
    r[2] = z0[2];
    r[3] = z0[3];
    asm("add.cc.u32 %0, %12, %20;\n\t"
        "addc.cc.u32 %1, %13, %21;\n\t"
        "addc.cc.u32 %2, %14, %22;\n\t"
        "addc.cc.u32 %3, %15, %23;\n\t"
        "addc.cc.u32 %4, %16, %24;\n\t"
        "addc.cc.u32 %5, %17, %25;\n\t"
        "addc.cc.u32 %6, %18, %26;\n\t"
        "addc.cc.u32 %7, %19, %27;\n\t"
        "addc.cc.u32 %8, %29, %28;\n\t"
        "addc.cc.u32 %9, %30, 0;\n\t"
        "addc.cc.u32 %10, %31, 0;\n\t"
        "addc.u32 %11, %32, 0;"
        : "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(r[7]), "=&r"(r[8]), "=&r"(r[9]), "=&r"(r[10]), "=&r"(r[11]), "=&r"(r[12]),
          "=&r"(r[13]), "=&r"(r[14]), "=&r"(r[15])
        : "r"(z0[4]), "r"(z0[5]), "r"(z0[6]), "r"(z0[7]), "r"(z2[0]), "r"(z2[1]), "r"(z2[2]), "r"(z2[3]), "r"(zm[0]),
          "r"(zm[1]), "r"(zm[2]), "r"(zm[3]), "r"(zm[4]), "r"(zm[5]), "r"(zm[6]), "r"(zm[7]), "r"(zm[8]), "r"(z2[4]),
          "r"(z2[5]), "r"(z2[6]), "r"(z2[7]));
#else
    u64 t = 0;
    for (int i = 0; i < 4; i++) {
        t += (u64)zm[4 + i] + ta[i] + tb[i];
        zm[4 + i] = (u32)t;
        t >>= 32;
    }
    zm[8] += (u32)t;
    for (int pass = 0; pass < 2; pass++) {
        const u32* z = pass ? z2 : z0;
        u64 bw = 0;
        for (int i = 0; i < 9; i++) {
            u64 d = (u64)zm[i] - (i < 8 ? z[i] : 0u) - bw;
            zm[i] = (u32)d;
            bw = (d >> 32) & 1;
        }
    }
    for (int i = 0; i < 4; i++) r[i] = z0[i];
    t = 0;
    for (int i = 4; i < 16; i++) {
        t += (u64)(i < 8 ? z0[i] : z2[i - 8]) + (i < 13 ? zm[i - 4] : 0);
        r[i] = (u32)t;
        t >>= 32;
    }
#endif
}

SG_HD void mul8x8(u32* r, const u32* a, const u32* b) {
#if defined(SG_KARATSUBA)
    mul8x8_kara(r, a, b);
#else
    mul8x8_school(r, a, b);
#endif
}

// r[0..16) += sum a_i^2 * 2^(64 i)
SG_HD void mad_diag8(u32* r, const u32* a) {
#if SG_PTX
    asm("mad.lo.cc.u32 %0, %16, %16, %0;\n\t"
        "madc.hi.cc.u32 %1, %16, %16, %1;\n\t"
        "madc.lo.cc.u32 %2, %17, %17, %2;\n\t"
        "madc.hi.cc.u32 %3, %17, %17, %3;\n\t"
        "madc.lo.cc.u32 %4, %18, %18, %4;\n\t"
        "madc.hi.cc.u32 %5, %18, %18, %5;\n\t"
        "madc.lo.cc.u32 %6, %19, %19, %6;\n\t"
        "madc.hi.cc.u32 %7, %19, %19, %7;\n\t"
        "madc.lo.cc.u32 %8, %20, %20, %8;\n\t"
        "madc.hi.cc.u32 %9, %20, %20, %9;\n\t"
        "madc.lo.cc.u32 %10, %21, %21, %10;\n\t"
        "madc.hi.cc.u32 %11, %21, %21, %11;\n\t"
        "madc.lo.cc.u32 %12, %22, %22, %12;\n\t"
        "madc.hi.cc.u32 %13, %22, %22, %13;\n\t"
        "madc.lo.cc.u32 %14, %23, %23, %14;\n\t"
        "madc.hi.u32 %15, %23, %23, %15;"
        : "+&r"(r[0]), "+&r"(r[1]), "+&r"(r[2]), "+&r"(r[3]), "+&r"(r[4]), "+&r"(r[5]), "+&r"(r[6]), "+&r"(r[7]), "+&r"(r[8]),
          "+&r"(r[9]), "+&r"(r[10]), "+&r"(r[11]), "+&r"(r[12]), "+&r"(r[13]), "+&r"(r[14]), "+&r"(r[15])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]));
#else
    u64 cy = 0;
    for (int j = 0; j < 8; j++) {
        u64 cur = ((u64)r[2 * j + 1] << 32) | r[2 * j];
        u64 prod = (u64)a[j] * a[j];
        u64 s = cur + prod;
        u64 c1 = s < cur;
        u64 s2 = s + cy;
        u64 c2 = s2 < s;
        r[2 * j] = (u32)s2;
        r[2 * j + 1] = (u32)(s2 >> 32);
        cy = c1 + c2;
    }
#endif
}

// r[0..16) = a^2: 28 cross products (doubled by a funnel shift) + 8 diagonal squares accumulated on top.
SG_HD void sqr8(u32* r, const u32* a) {
    u32 e[16], o[16];
#pragma unroll
    for (int i = 0; i < 16; i++) {
        e[i] = 0;
        o[i] = 0;
    }
    // row 0 (a0 * a1..a7): odd columns 1,3,5,7 -> o[0..8); even columns 2,4,6 -> e[2..8)
    mul_wide(o + 0, a[1], a[0]);
    mul_wide(o + 2, a[3], a[0]);
    mul_wide(o + 4, a[5], a[0]);
    mul_wide(o + 6, a[7], a[0]);
    mul_wide(e + 2, a[2], a[0]);
    mul_wide(e + 4, a[4], a[0]);
    mul_wide(e + 6, a[6], a[0]);
    // row 1 (a1 * a2..a7): even columns 4,6,8 -> e[4..10); odd columns 3,5,7 -> o[2..8) carry -> o[8]
    mad_row3(e + 4, a[3], a[5], a[7], a[1]);
    o[8] = mad_row3(o + 2, a[2], a[4], a[6], a[1]);
    // row 2 (a2 * a3..a7): odd columns 5,7,9 -> o[4..10); even columns 6,8 -> e[6..10) carry -> e[10]
    mad_row3(o + 4, a[3], a[5], a[7], a[2]);
    e[10] = mad_row2(e + 6, a[4], a[6], a[2]);
    // row 3 (a3 * a4..a7): even columns 8,10 -> e[8..12); odd columns 7,9 -> o[6..10) carry -> o[10]
    mad_row2(e + 8, a[5], a[7], a[3]);
    o[10] = mad_row2(o + 6, a[4], a[6], a[3]);
    // row 4 (a4 * a5..a7): odd columns 9,11 -> o[8..12); even column 10 -> e[10..12) carry -> e[12]
    mad_row2(o + 8, a[5], a[7], a[4]);
    e[12] = mad_row1(e + 10, a[6], a[4]);
    // row 5 (a5 * a6,a7): even column 12 -> e[12..14); odd column 11 -> o[10..12) carry -> o[12]
    mad_row1(e + 12, a[7], a[5]);
    o[12] = mad_row1(o + 10, a[6], a[5]);
    // row 6 (a6 * a7): odd column 13 -> o[12..14)
    mad_row1(o + 12, a[7], a[6]);
    u32 t[16];
    merge_even_odd(t, e, o);
    // double the cross terms
#pragma unroll
    for (int k = 15; k > 0; k--) r[k] = (t[k] << 1) | (t[k - 1] >> 31);
    r[0] = t[0] << 1;
    // add the diagonal a_i^2 at limb pairs (2i, 2i+1): one 16-instruction chain (no carry out: a^2 < 2^512)
    mad_diag8(r, a);
}

}  // namespace sigops
