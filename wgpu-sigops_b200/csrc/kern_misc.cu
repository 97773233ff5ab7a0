// Small kernels: SHA-256 on either side of recovery, the fixed-base table generator, the unit-test shim (the role of
// src/wgsl/tests/*.wgsl) and the integer-pipe micro-benchmark; plus their launchers.
#include <cuda_runtime.h>

#include "kernels.cuh"
#include "launch.h"

using namespace sigops;

namespace sigops {

// SHA-256 of n variable-length messages (one per thread): the `Message::new` prehash done on the device.
__global__ void __launch_bounds__(256) sha256_msgs_kernel(const uint8_t* __restrict__ bytes,
                                                          const unsigned long long* __restrict__ off, size_t n,
                                                          u32* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u32 d[8];
    sha256_ram(d, bytes + off[i], (size_t)(off[i + 1] - off[i]));
#pragma unroll
    for (int j = 0; j < 8; j++) out[i * 8 + j] = d[j];
}

// Fuel address of each recovered key: SHA-256(X || Y); 32 zero bytes where the recovery was rejected.
__global__ void __launch_bounds__(256) sha256_pubkeys_kernel(const u32* __restrict__ pubkeys, const uint8_t* __restrict__ status,
                                                             size_t n, u32* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u32 in[16], d[8];
#pragma unroll
    for (int j = 0; j < 16; j++) in[j] = pubkeys[i * 16 + j];
    sha256_64(d, in);
    const bool bad = status && status[i] != 0;
#pragma unroll
    for (int j = 0; j < 8; j++) out[i * 8 + j] = bad ? 0u : d[j];
}

// Positional fixed-base tables (ptab.h), generated once per device at init from the baked seeds 2^-D G (consts_gen.cuh).
// Replaces the 16-entry tables of src/precompute.rs:14-69 that the reference uploads on every call
// (src/secp256k1_ecdsa.rs:108).  Two launches per curve: the window bases B_j = 2^(w j) * seed (one thread each), then one
// thread per entry m * B_j (w doublings, ~w/2 additions and one inversion: about a fifth of a signature).
template <int CURVE>
__global__ void __launch_bounds__(64) gen_ptab_bases_kernel(u32* bases, u32 w, u32 pos) {
    const u32 j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= pos) return;
    u32 e[24];
    if (CURVE == 0) sw_ptab_base<CurveK1>(e, w * j);
    if (CURVE == 1) sw_ptab_base<CurveR1>(e, w * j);
    if (CURVE == 2) ed_ptab_base(e, w * j);
    const int words = CURVE == 2 ? 24 : 16;
    for (int i = 0; i < words; i++) bases[(size_t)j * words + i] = e[i];
}

template <int CURVE>
__global__ void __launch_bounds__(128) gen_ptab_kernel(u32* tab, const u32* __restrict__ bases, u32 w, u32 pos) {
    const size_t per = (size_t)1 << (w - 1);
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= per * pos) return;
    const u32 j = (u32)(t >> (w - 1)), m = (u32)(t & (per - 1)) + 1u;
    const int words = CURVE == 2 ? 24 : 16;
    u32 e[24];
    if (CURVE == 0) sw_ptab_entry<CurveK1>(e, m, w, bases + (size_t)j * words);
    if (CURVE == 1) sw_ptab_entry<CurveR1>(e, m, w, bases + (size_t)j * words);
    if (CURVE == 2) ed_ptab_entry(e, m, w, bases + (size_t)j * words);
    Q4* dst = reinterpret_cast<Q4*>(tab + t * words);
    for (int q = 0; q < words / 4; q++) {
        Q4 v = {e[4 * q], e[4 * q + 1], e[4 * q + 2], e[4 * q + 3]};
        dst[q] = v;
    }
}

__global__ void __launch_bounds__(kBlock) unit_kernel(int op, const u32* __restrict__ in, size_t n, u32* __restrict__ out,
                                                      Q4* __restrict__ scratch, const __grid_constant__ PTab k1g,
                                                      const __grid_constant__ PTab r1g, const __grid_constant__ PTab edb) {
    int in_w, out_w;
    unit_shape(op, in_w, out_w);
    const size_t nthreads = (size_t)gridDim.x * blockDim.x;
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    TabRef tab;
    tab.base = scratch + gid;
    tab.stride = (u32)nthreads;
    for (size_t i = gid; i < n; i += nthreads) {
        u32 a[32], r[17];
        for (int j = 0; j < 32; j++) a[j] = j < in_w ? in[i * in_w + j] : 0u;
        for (int j = 0; j < 17; j++) r[j] = 0;
        unit_dispatch(op, r, a, tab, k1g, r1g, edb);
        for (int j = 0; j < out_w; j++) out[i * out_w + j] = r[j];
    }
}

// ---- integer-pipe micro-benchmark: 8 independent accumulator chains per thread, fully unrolled inner block ----
template <int KIND>
__global__ void __launch_bounds__(256) imad_peak_kernel(u32* sink, int iters, u32 seed) {
    u32 a0 = seed + threadIdx.x, a1 = a0 * 3 + 1, a2 = a0 * 5 + 2, a3 = a0 * 7 + 3;
    u32 a4 = a0 * 9 + 4, a5 = a0 * 11 + 5, a6 = a0 * 13 + 6, a7 = a0 * 15 + 7;
    u32 b0 = a7, b1 = a6, b2 = a5, b3 = a4, b4 = a3, b5 = a2, b6 = a1, b7 = a0;
    u32 x = seed | 1u, y = (seed >> 3) | 5u;
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            if (KIND == 0) {
                asm volatile(
                    "mad.lo.u32 %0, %0, %8, %9;\n\tmad.lo.u32 %1, %1, %8, %9;\n\tmad.lo.u32 %2, %2, %8, %9;\n\t"
                    "mad.lo.u32 %3, %3, %8, %9;\n\tmad.lo.u32 %4, %4, %8, %9;\n\tmad.lo.u32 %5, %5, %8, %9;\n\t"
                    "mad.lo.u32 %6, %6, %8, %9;\n\tmad.lo.u32 %7, %7, %8, %9;"
                    : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3), "+r"(a4), "+r"(a5), "+r"(a6), "+r"(a7)
                    : "r"(x), "r"(y));
            } else if (KIND == 1) {
                // 8 independent 64-bit accumulators (aK,bK) += x*y : IMAD.WIDE.U32 without carry chain
                asm volatile(
                    "{.reg .u64 t0,t1,t2,t3,t4,t5,t6,t7;\n\t"
                    "mov.b64 t0,{%0,%8}; mov.b64 t1,{%1,%9}; mov.b64 t2,{%2,%10}; mov.b64 t3,{%3,%11};\n\t"
                    "mov.b64 t4,{%4,%12}; mov.b64 t5,{%5,%13}; mov.b64 t6,{%6,%14}; mov.b64 t7,{%7,%15};\n\t"
                    "mad.wide.u32 t0,%0,%16,t0; mad.wide.u32 t1,%1,%16,t1; mad.wide.u32 t2,%2,%16,t2; mad.wide.u32 t3,%3,%16,t3;\n\t"
                    "mad.wide.u32 t4,%4,%16,t4; mad.wide.u32 t5,%5,%16,t5; mad.wide.u32 t6,%6,%16,t6; mad.wide.u32 t7,%7,%16,t7;\n\t"
                    "mov.b64 {%0,%8},t0; mov.b64 {%1,%9},t1; mov.b64 {%2,%10},t2; mov.b64 {%3,%11},t3;\n\t"
                    "mov.b64 {%4,%12},t4; mov.b64 {%5,%13},t5; mov.b64 {%6,%14},t6; mov.b64 {%7,%15},t7;}"
                    : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3), "+r"(a4), "+r"(a5), "+r"(a6), "+r"(a7), "+r"(b0),
                      "+r"(b1), "+r"(b2), "+r"(b3), "+r"(b4), "+r"(b5), "+r"(b6), "+r"(b7)
                    : "r"(x));
            } else if (KIND == 2) {
                // two 4-product carry chains (the shape of one row of the field multiplication)
                asm volatile(
                    "mad.lo.cc.u32 %0,%16,%17,%0; madc.hi.cc.u32 %1,%16,%17,%1; madc.lo.cc.u32 %2,%17,%16,%2; madc.hi.cc.u32 %3,%17,%16,%3;\n\t"
                    "madc.lo.cc.u32 %4,%16,%16,%4; madc.hi.cc.u32 %5,%16,%16,%5; madc.lo.cc.u32 %6,%17,%17,%6; madc.hi.u32 %7,%17,%17,%7;\n\t"
                    "mad.lo.cc.u32 %8,%16,%17,%8; madc.hi.cc.u32 %9,%16,%17,%9; madc.lo.cc.u32 %10,%17,%16,%10; madc.hi.cc.u32 %11,%17,%16,%11;\n\t"
                    "madc.lo.cc.u32 %12,%16,%16,%12; madc.hi.cc.u32 %13,%16,%16,%13; madc.lo.cc.u32 %14,%17,%17,%14; madc.hi.u32 %15,%17,%17,%15;"
                    : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3), "+r"(a4), "+r"(a5), "+r"(a6), "+r"(a7), "+r"(b0),
                      "+r"(b1), "+r"(b2), "+r"(b3), "+r"(b4), "+r"(b5), "+r"(b6), "+r"(b7)
                    : "r"(x), "r"(y));
            } else if (KIND == 5) {
                // FP64 pipe: 8 independent DFMA chains (the a/b registers are reinterpreted pairwise as doubles)
                asm volatile(
                    "{.reg .f64 d0,d1,d2,d3,d4,d5,d6,d7,m;\n\t"
                    "mov.b64 d0,{%0,%8}; mov.b64 d1,{%1,%9}; mov.b64 d2,{%2,%10}; mov.b64 d3,{%3,%11};\n\t"
                    "mov.b64 d4,{%4,%12}; mov.b64 d5,{%5,%13}; mov.b64 d6,{%6,%14}; mov.b64 d7,{%7,%15};\n\t"
                    "mov.b64 m,{%16,%17};\n\t"
                    "fma.rn.f64 d0,d0,m,m; fma.rn.f64 d1,d1,m,m; fma.rn.f64 d2,d2,m,m; fma.rn.f64 d3,d3,m,m;\n\t"
                    "fma.rn.f64 d4,d4,m,m; fma.rn.f64 d5,d5,m,m; fma.rn.f64 d6,d6,m,m; fma.rn.f64 d7,d7,m,m;\n\t"
                    "mov.b64 {%0,%8},d0; mov.b64 {%1,%9},d1; mov.b64 {%2,%10},d2; mov.b64 {%3,%11},d3;\n\t"
                    "mov.b64 {%4,%12},d4; mov.b64 {%5,%13},d5; mov.b64 {%6,%14},d6; mov.b64 {%7,%15},d7;}"
                    : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3), "+r"(a4), "+r"(a5), "+r"(a6), "+r"(a7), "+r"(b0),
                      "+r"(b1), "+r"(b2), "+r"(b3), "+r"(b4), "+r"(b5), "+r"(b6), "+r"(b7)
                    : "r"(x), "r"(y));
            } else if (KIND == 6) {
                asm volatile(
                    "mad.hi.u32 %0, %0, %8, %9;\n\tmad.hi.u32 %1, %1, %8, %9;\n\tmad.hi.u32 %2, %2, %8, %9;\n\t"
                    "mad.hi.u32 %3, %3, %8, %9;\n\tmad.hi.u32 %4, %4, %8, %9;\n\tmad.hi.u32 %5, %5, %8, %9;\n\t"
                    "mad.hi.u32 %6, %6, %8, %9;\n\tmad.hi.u32 %7, %7, %8, %9;"
                    : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3), "+r"(a4), "+r"(a5), "+r"(a6), "+r"(a7)
                    : "r"(x), "r"(y));
            } else if (KIND == 3) {
                asm volatile(
                    "add.u32 %0, %0, %8;\n\tadd.u32 %1, %1, %9;\n\tadd.u32 %2, %2, %8;\n\tadd.u32 %3, %3, %9;\n\t"
                    "add.u32 %4, %4, %8;\n\tadd.u32 %5, %5, %9;\n\tadd.u32 %6, %6, %8;\n\tadd.u32 %7, %7, %9;"
                    : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3), "+r"(a4), "+r"(a5), "+r"(a6), "+r"(a7)
                    : "r"(x), "r"(y));
            } else {
                asm volatile(
                    "{.reg .u64 t0,t1,t2,t3;\n\t"
                    "mov.b64 t0,{%0,%4}; mov.b64 t1,{%1,%5}; mov.b64 t2,{%2,%6}; mov.b64 t3,{%3,%7};\n\t"
                    "mad.wide.u32 t0,%0,%16,t0; add.u32 %8,%8,%17; mad.wide.u32 t1,%1,%16,t1; add.u32 %9,%9,%17;\n\t"
                    "mad.wide.u32 t2,%2,%16,t2; add.u32 %10,%10,%17; mad.wide.u32 t3,%3,%16,t3; add.u32 %11,%11,%17;\n\t"
                    "mad.wide.u32 t0,%1,%16,t0; add.u32 %12,%12,%17; mad.wide.u32 t1,%2,%16,t1; add.u32 %13,%13,%17;\n\t"
                    "mad.wide.u32 t2,%3,%16,t2; add.u32 %14,%14,%17; mad.wide.u32 t3,%0,%16,t3; add.u32 %15,%15,%17;\n\t"
                    "mov.b64 {%0,%4},t0; mov.b64 {%1,%5},t1; mov.b64 {%2,%6},t2; mov.b64 {%3,%7},t3;}"
                    : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3), "+r"(a4), "+r"(a5), "+r"(a6), "+r"(a7), "+r"(b0),
                      "+r"(b1), "+r"(b2), "+r"(b3), "+r"(b4), "+r"(b5), "+r"(b6), "+r"(b7)
                    : "r"(x), "r"(y));
            }
        }
    }
    u32 r = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7 ^ b0 ^ b1 ^ b2 ^ b3 ^ b4 ^ b5 ^ b6 ^ b7;
    if (r == 0x12345678u) sink[blockIdx.x * blockDim.x + threadIdx.x] = r;
}


int kl_sha256_msgs(cudaStream_t st, const uint8_t* bytes, const unsigned long long* off, size_t n, u32* out) {
    sha256_msgs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(bytes, off, n, out);
    return (int)cudaGetLastError();
}
int kl_sha256_pubkeys(cudaStream_t st, const u32* pubkeys, const uint8_t* status, size_t n, u32* out) {
    sha256_pubkeys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(pubkeys, status, n, out);
    return (int)cudaGetLastError();
}
int kl_gen_ptab(cudaStream_t st, int curve, u32 w, u32* bases, u32* tab) {
    const u32 pos = ptab_positions(w);
    const size_t entries = ptab_entries(w);
    const unsigned gb = (pos + 63) / 64, ge = (unsigned)((entries + 127) / 128);
    switch (curve) {
        case 0:
            gen_ptab_bases_kernel<0><<<gb, 64, 0, st>>>(bases, w, pos);
            gen_ptab_kernel<0><<<ge, 128, 0, st>>>(tab, bases, w, pos);
            break;
        case 1:
            gen_ptab_bases_kernel<1><<<gb, 64, 0, st>>>(bases, w, pos);
            gen_ptab_kernel<1><<<ge, 128, 0, st>>>(tab, bases, w, pos);
            break;
        case 2:
            gen_ptab_bases_kernel<2><<<gb, 64, 0, st>>>(bases, w, pos);
            gen_ptab_kernel<2><<<ge, 128, 0, st>>>(tab, bases, w, pos);
            break;
        default:
            return -1;
    }
    return (int)cudaGetLastError();
}
int kl_unit(const KLaunch& l, int op, const u32* in, size_t n, u32* out, void* scratch, const PTab& k1g, const PTab& r1g, const PTab& edb) {
    unit_kernel<<<l.grid, l.tpb, 0, l.stream>>>(op, in, n, out, (Q4*)scratch, k1g, r1g, edb);
    return (int)cudaGetLastError();
}
int kl_unit_setup(int* max_blocks_per_sm) {
    return (int)cudaOccupancyMaxActiveBlocksPerMultiprocessor(max_blocks_per_sm, unit_kernel, kBlock, 0);
}
int kl_imad_peak(int kind, int grid, int block, cudaStream_t st, u32* sink, int iters, u32 seed) {
    switch (kind) {
        case 0: imad_peak_kernel<0><<<grid, block, 0, st>>>(sink, iters, seed); break;
        case 1: imad_peak_kernel<1><<<grid, block, 0, st>>>(sink, iters, seed); break;
        case 2: imad_peak_kernel<2><<<grid, block, 0, st>>>(sink, iters, seed); break;
        case 3: imad_peak_kernel<3><<<grid, block, 0, st>>>(sink, iters, seed); break;
        case 4: imad_peak_kernel<4><<<grid, block, 0, st>>>(sink, iters, seed); break;
        case 5: imad_peak_kernel<5><<<grid, block, 0, st>>>(sink, iters, seed); break;
        case 6: imad_peak_kernel<6><<<grid, block, 0, st>>>(sink, iters, seed); break;
        default: return -1;
    }
    return (int)cudaGetLastError();
}

}  // namespace sigops
