// __global__ entry points: one fused kernel per curve (replaces the 5 / 6 / 7 dispatches of
// src/secp256k1_ecdsa.rs:61-213, src/secp256r1_ecdsa.rs:62-214 and src/ed25519_eddsa.rs:67-257 and the
// intermediate storage buffers between them), the unit-test shims (the role of src/wgsl/tests/*.wgsl) and
// the integer-pipe micro-benchmark.
//
// Launch geometry: 1-D grid of one 512-thread block per SM (smaller blocks spread over all SMs for small batches); thread t
// owns rows t, t + T, t + 2T, ... of the shard -- no power-of-two padding, no 3-D workgroup lookup table
// (src/benchmarks/mod.rs:10-53 `compute_num_workgroups`, src/secp256k1_ecdsa.rs:24-47).
#pragma once
#include "../../include/sigops.h"
#include "curve_ed.cuh"
#include "sha256.cuh"

namespace sigops {

#ifndef SG_BLOCK
#define SG_BLOCK 512
#endif
static constexpr int kBlock = SG_BLOCK;
// Launch bounds: 512 threads x 1 block per SM => at most 128 registers per thread, 16 warps per SM.  Measured on B200
// (profiles/r01_variants.md): 14-16 warps/SM beat the unconstrained build (188-250 registers, 8 warps/SM) by 10-14%
// despite ~1 KB of spills per thread, and ONE block per SM with a barrier per phase of the per-signature program beats
// seven independent 64-thread blocks by another 12-27% (instruction-cache locality).
#ifndef SG_MINB_SW
#define SG_MINB_SW 1
#endif
#ifndef SG_MINB_ED
#define SG_MINB_ED 1
#endif

// Barriers inside the per-signature program (between its phases and once per window of the main loop) on top of the
// one per signature.  Measured on B200 (profiles/r01_variants.md): with out-of-line products they cost 2-4%; together
// with the loop's products inlined (SG_HOT_INLINE, a 70-86 KB loop body that all 16 warps then stream in lockstep) they
// win 3-5%.  Both are on by default; -DSG_NO_HOT_INLINE / -DSG_NO_INNER_SYNC switch them off.
#if defined(SG_NO_INNER_SYNC)
static constexpr bool kInnerSync = false;
#else
static constexpr bool kInnerSync = true;
#endif

#if defined(__CUDACC__)

// Fixed-base table staged in shared memory: with one block per SM the whole 227 KB is free, and a 2048-entry table
// (128 KB for the secp curves' j*G, 192 KB for ed25519's Niels triples) fits.  `smem_words` == 0 (small batches, where
// the copy would cost more than the gathers) leaves the table in global memory / L2.
extern __shared__ __align__(16) u32 sg_smem_table[];
__device__ __forceinline__ const u32* stage_table(const u32* __restrict__ gtab, u32 smem_words) {
    if (smem_words == 0) return gtab;
    const Q4* src = reinterpret_cast<const Q4*>(gtab);
    Q4* dst = reinterpret_cast<Q4*>(sg_smem_table);
    for (u32 i = threadIdx.x; i < smem_words / 4; i += blockDim.x) dst[i] = src[i];
    __syncthreads();
    return sg_smem_table;
}

// Row access of one thread's batch: item j of the batch is row first + j * stride of the shard.  Rows past the end are
// clamped on load (the thread redoes the last signature so that it reaches every barrier) and dropped on store.
struct SwDeviceIO {
    const Q4* sigs;
    const Q4* msgs;
    Q4* out;
    uint8_t* status;
    size_t n, first, stride;
    __device__ __forceinline__ void load(int j, u32* sig_w, u32* msg_w) const {
        size_t i = first + (size_t)j * stride;
        if (i >= n) i = n - 1;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            Q4 v = sigs[4 * i + q];
            sig_w[4 * q + 0] = v.x;
            sig_w[4 * q + 1] = v.y;
            sig_w[4 * q + 2] = v.z;
            sig_w[4 * q + 3] = v.w;
        }
#pragma unroll
        for (int q = 0; q < 2; q++) {
            Q4 v = msgs[2 * i + q];
            msg_w[4 * q + 0] = v.x;
            msg_w[4 * q + 1] = v.y;
            msg_w[4 * q + 2] = v.z;
            msg_w[4 * q + 3] = v.w;
        }
    }
    __device__ __forceinline__ void store(int j, const u32* out_w, u32 st) const {
        const size_t i = first + (size_t)j * stride;
        if (i >= n) return;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            Q4 v = {out_w[4 * q + 0], out_w[4 * q + 1], out_w[4 * q + 2], out_w[4 * q + 3]};
            out[4 * i + q] = v;
        }
        if (status) status[i] = (uint8_t)st;
    }
};

template <class C>
__global__ void __launch_bounds__(kBlock, SG_MINB_SW) ecrecover_kernel(const Q4* __restrict__ sigs, const Q4* __restrict__ msgs,
                                                                       size_t n, Q4* __restrict__ out, uint8_t* __restrict__ status,
                                                                       Q4* __restrict__ scratch, const u32* __restrict__ gtab_g,
                                                                       u32 smem_words) {
    // secp256k1: the j*G half of the table is staged (the lambda*j*G half stays in L2); secp256r1: the whole table
    const u32* gtab_s = stage_table(gtab_g, smem_words);
    const size_t nthreads = (size_t)gridDim.x * blockDim.x;
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    TabRef tab;
    tab.base = scratch + gid;
    tab.stride = (u32)nthreads;
    SwDeviceIO io = {sigs, msgs, out, status, n, 0, nthreads};
    // Each thread takes the rows gid, gid + T, gid + 2T, ... (T = threads in the grid) in batches of up to kSwBatch; the
    // batch size is uniform over the grid, so every thread of a block reaches every barrier.
    const size_t passes = (n + nthreads - 1) / nthreads;
    for (size_t pass = 0; pass < passes; pass += kSwBatch) {
        const int B = (int)((passes - pass) < (size_t)kSwBatch ? (passes - pass) : (size_t)kSwBatch);
        phase_sync<true>();
        io.first = pass * nthreads + gid;
        sw_ecrecover_batch<C, kInnerSync>(B, io, tab, gtab_s, gtab_g);
    }
}

struct EdDeviceIO {
    const Q4* sigs;
    const Q4* msgs;
    const Q4* pks;
    uint8_t* valid;
    size_t n, first, stride;
    __device__ __forceinline__ void load(int j, u32* sig_w, u32* msg_w, u32* pk_w) const {
        size_t i = first + (size_t)j * stride;
        if (i >= n) i = n - 1;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            Q4 v = sigs[4 * i + q];
            sig_w[4 * q + 0] = v.x;
            sig_w[4 * q + 1] = v.y;
            sig_w[4 * q + 2] = v.z;
            sig_w[4 * q + 3] = v.w;
        }
#pragma unroll
        for (int q = 0; q < 2; q++) {
            Q4 v = msgs[2 * i + q];
            msg_w[4 * q + 0] = v.x;
            msg_w[4 * q + 1] = v.y;
            msg_w[4 * q + 2] = v.z;
            msg_w[4 * q + 3] = v.w;
            Q4 p = pks[2 * i + q];
            pk_w[4 * q + 0] = p.x;
            pk_w[4 * q + 1] = p.y;
            pk_w[4 * q + 2] = p.z;
            pk_w[4 * q + 3] = p.w;
        }
    }
    __device__ __forceinline__ void store(int j, u32 v) const {
        const size_t i = first + (size_t)j * stride;
        if (i < n) valid[i] = (uint8_t)v;
    }
};

__global__ void __launch_bounds__(kBlock, SG_MINB_ED) ed25519_verify_kernel(const Q4* __restrict__ sigs, const Q4* __restrict__ msgs,
                                                                            const Q4* __restrict__ pks, size_t n,
                                                                            uint8_t* __restrict__ valid, Q4* __restrict__ scratch,
                                                                            const u32* __restrict__ btab_g, u32 smem_words) {
    const u32* btab = stage_table(btab_g, smem_words);
    const size_t nthreads = (size_t)gridDim.x * blockDim.x;
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    TabRef tab;
    tab.base = scratch + gid;
    tab.stride = (u32)nthreads;
    EdDeviceIO io = {sigs, msgs, pks, valid, n, 0, nthreads};
    const size_t passes = (n + nthreads - 1) / nthreads;
    for (size_t pass = 0; pass < passes; pass += kEdBatch) {
        const int B = (int)((passes - pass) < (size_t)kEdBatch ? (passes - pass) : (size_t)kEdBatch);
        phase_sync<true>();
        io.first = pass * nthreads + gid;
        ed_verify_batch<kInnerSync>(B, io, tab, btab);
    }
}

// ed25519 with variable-length messages and optional strict semantics (what
// fuel_crypto::ed25519::verify needs; the reference hard-wires 32-byte messages, src/wgsl/sha512.wgsl:114-123).
// msg_bytes: all messages back to back; msg_off[i] .. msg_off[i+1] delimit message i (n + 1 offsets).
__global__ void __launch_bounds__(kBlock, SG_MINB_ED) ed25519_verify_msgs_kernel(
    const Q4* __restrict__ sigs, const uint8_t* __restrict__ msg_bytes, const unsigned long long* __restrict__ msg_off,
    const Q4* __restrict__ pks, size_t n, int strict, uint8_t* __restrict__ valid, Q4* __restrict__ scratch,
    const u32* __restrict__ btab) {
    const size_t nthreads = (size_t)gridDim.x * blockDim.x;
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    TabRef tab;
    tab.base = scratch + gid;
    tab.stride = (u32)nthreads;
    for (size_t base = (size_t)blockIdx.x * blockDim.x; base < n; base += nthreads) {
        phase_sync<true>();
        size_t i = base + threadIdx.x;
        const bool live = i < n;
        if (!live) i = n - 1;
        u32 sig_w[16], pk_w[8];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            Q4 v = sigs[4 * i + q];
            sig_w[4 * q + 0] = v.x;
            sig_w[4 * q + 1] = v.y;
            sig_w[4 * q + 2] = v.z;
            sig_w[4 * q + 3] = v.w;
        }
#pragma unroll
        for (int q = 0; q < 2; q++) {
            Q4 p = pks[2 * i + q];
            pk_w[4 * q + 0] = p.x;
            pk_w[4 * q + 1] = p.y;
            pk_w[4 * q + 2] = p.z;
            pk_w[4 * q + 3] = p.w;
        }
        const unsigned long long lo = msg_off[i], hi = msg_off[i + 1];
        const u32 v = ed_verify_msg<kInnerSync>(sig_w, msg_bytes + lo, (size_t)(hi - lo), pk_w, strict != 0, tab, btab);
        if (live) valid[i] = (uint8_t)v;
    }
}

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

// Fixed-base tables, generated once per device at init: thread j writes entry j (the (j+1)-th multiple) of
//   k1tab  [2][kGTabEntries][16]  j*G and lambda*j*G        r1tab [kGTabEntries][16]  j*G (Montgomery form)
//   edtab  [kGTabEntries][24]     j*B as affine Niels triples
// from the baked single generators (consts_gen.cuh).  Replaces the 16-entry tables of src/precompute.rs:14-69 that the
// reference uploads on every call (src/secp256k1_ecdsa.rs:108).
__global__ void __launch_bounds__(64) gen_tables_kernel(u32* k1tab, u32* r1tab, u32* edtab) {
    const u32 j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= (u32)kGTabEntries) return;
    u32 e[24];
    sw_gtab_entry<CurveK1>(e, j + 1, false, k1_g_dev);
    for (int i = 0; i < 16; i++) k1tab[(size_t)j * 16 + i] = e[i];
    sw_gtab_entry<CurveK1>(e, j + 1, true, k1_g_dev);
    for (int i = 0; i < 16; i++) k1tab[((size_t)kGTabEntries + j) * 16 + i] = e[i];
    sw_gtab_entry<CurveR1>(e, j + 1, false, r1_g_dev);
    for (int i = 0; i < 16; i++) r1tab[(size_t)j * 16 + i] = e[i];
    ed_btab_entry(e, j + 1, ed_b_niels_dev);
    for (int i = 0; i < 24; i++) edtab[(size_t)j * 24 + i] = e[i];
}

#endif  // __CUDACC__

// ---------------------------------------------------------------------------------------------------------
// unit-test shims: one item per thread, words in / words out
// ---------------------------------------------------------------------------------------------------------
SG_HD void unit_shape(int op, int& in_w, int& out_w) {
    in_w = 8;
    out_w = 8;
    switch (op) {
        case SIGOPS_UNIT_K1_MUL: case SIGOPS_UNIT_K1_ADD: case SIGOPS_UNIT_K1_SUB:
        case SIGOPS_UNIT_R1_MUL: case SIGOPS_UNIT_R1_ADD: case SIGOPS_UNIT_R1_SUB:
        case SIGOPS_UNIT_ED_MUL: case SIGOPS_UNIT_ED_ADD: case SIGOPS_UNIT_ED_SUB:
        case SIGOPS_UNIT_K1N_MUL: case SIGOPS_UNIT_R1N_MUL:
            in_w = 16;
            break;
        case SIGOPS_UNIT_EDL_REDUCE512:
            in_w = 16;
            break;
        case SIGOPS_UNIT_SHA512_96:
            in_w = 24;
            out_w = 16;
            break;
        case SIGOPS_UNIT_SHA256_64:
            in_w = 16;
            out_w = 8;
            break;
        case SIGOPS_UNIT_K1_GLV:
            out_w = 12;
            break;
        case SIGOPS_UNIT_MUL8X8:
            in_w = 16;
            out_w = 16;
            break;
        case SIGOPS_UNIT_SQR8:
            out_w = 16;
            break;
        case SIGOPS_UNIT_K1_MULPT: case SIGOPS_UNIT_R1_MULPT:
            in_w = 24;
            out_w = 17;
            break;
        case SIGOPS_UNIT_ED_MULPT:
            in_w = 24;
            out_w = 16;
            break;
        case SIGOPS_UNIT_K1_DOUBLE_MUL: case SIGOPS_UNIT_R1_DOUBLE_MUL:
            in_w = 32;
            out_w = 17;
            break;
        case SIGOPS_UNIT_RAW_ADDSUB:
            in_w = 17;
            out_w = 16;
            break;
        case SIGOPS_UNIT_RAW_REDUCE16:
            in_w = 17;
            out_w = 8;
            break;
        default:
            break;
    }
}

template <class F>
SG_HD void unit_raw_addsub(u32* out, const u32* in) {
    Fe a, b, s, d;
    copy8(a.v, in);
    copy8(b.v, in + 8);
    F::add(s, a, b);
    F::sub(d, a, b);
    copy8(out, s.v);
    copy8(out + 8, d.v);
}

template <class F>
SG_HD void unit_field(int which, u32* out, const u32* in) {
    // which: 0 mul, 1 sqr, 2 add, 3 sub, 4 inv, 5 sqrt-candidate
    Fe a, b, r;
    F::from_plain(a, in);
    if (which == 0 || which == 2 || which == 3) F::from_plain(b, in + 8);
    switch (which) {
        case 0: F::mul(r, a, b); break;
        case 1: F::sqr(r, a); break;
        case 2: F::add(r, a, b); break;
        case 3: F::sub(r, a, b); break;
        default: break;
    }
    F::to_plain(out, r);
}

template <class C>
SG_HD void unit_double_mul(u32* out, const u32* u1, const u32* u2, const u32* xy, const TabRef& tab, const u32* gtab) {
    typedef typename C::F F;
    Fe x, y;
    F::from_plain(x, xy);
    F::from_plain(y, xy + 8);
    sw_build_table<C>(tab, x, y);
    JacPoint Q;
    sw_double_mul<C, false>(Q, u1, u2, tab, gtab, gtab);
    for (int i = 0; i < 17; i++) out[i] = 0;
    if (Q.inf) {
        out[16] = 1;
        return;
    }
    Fe zi, zi2, ax, ay;
    fe_inv((F*)0, zi, Q.Z);
    F::sqr(zi2, zi);
    F::mul(ax, Q.X, zi2);
    F::mul(zi2, zi2, zi);
    F::mul(ay, Q.Y, zi2);
    F::to_plain(out, ax);
    F::to_plain(out + 8, ay);
}

// dispatcher shared by the device shim kernel and the host simulation
SG_HD void unit_dispatch(int op, u32* out, const u32* in, const TabRef& tab, const u32* k1g, const u32* r1g,
                         const u32* edb) {
    switch (op) {
        case SIGOPS_UNIT_K1_MUL: case SIGOPS_UNIT_K1_SQR: case SIGOPS_UNIT_K1_ADD: case SIGOPS_UNIT_K1_SUB:
            unit_field<FpK1>(op - SIGOPS_UNIT_K1_MUL, out, in);
            break;
        case SIGOPS_UNIT_R1_MUL: case SIGOPS_UNIT_R1_SQR: case SIGOPS_UNIT_R1_ADD: case SIGOPS_UNIT_R1_SUB:
            unit_field<FpR1>(op - SIGOPS_UNIT_R1_MUL, out, in);
            break;
        case SIGOPS_UNIT_ED_MUL: case SIGOPS_UNIT_ED_SQR: case SIGOPS_UNIT_ED_ADD: case SIGOPS_UNIT_ED_SUB:
            unit_field<Fp25519>(op - SIGOPS_UNIT_ED_MUL, out, in);
            break;
        case SIGOPS_UNIT_K1_INV: case SIGOPS_UNIT_K1_SQRT: {
            Fe a, r;
            FpK1::from_plain(a, in);
            if (op == SIGOPS_UNIT_K1_INV) fe_inv((FpK1*)0, r, a); else fe_sqrt_candidate((FpK1*)0, r, a);
            FpK1::to_plain(out, r);
            break;
        }
        case SIGOPS_UNIT_R1_INV: case SIGOPS_UNIT_R1_SQRT: {
            Fe a, r;
            FpR1::from_plain(a, in);
            if (op == SIGOPS_UNIT_R1_INV) fe_inv((FpR1*)0, r, a); else fe_sqrt_candidate((FpR1*)0, r, a);
            FpR1::to_plain(out, r);
            break;
        }
        case SIGOPS_UNIT_ED_INV: case SIGOPS_UNIT_ED_POW_P58: {
            Fe a, r;
            Fp25519::from_plain(a, in);
            if (op == SIGOPS_UNIT_ED_INV) fe_inv((Fp25519*)0, r, a); else fe_pow_p58(r, a);
            Fp25519::to_plain(out, r);
            break;
        }
        case SIGOPS_UNIT_K1N_MUL: {
            u32 am[8];
            Sc<ModK1N>::to_mont(am, in);
            Sc<ModK1N>::mmul(out, am, in + 8);
            break;
        }
        case SIGOPS_UNIT_R1N_MUL: {
            u32 am[8];
            Sc<ModR1N>::to_mont(am, in);
            Sc<ModR1N>::mmul(out, am, in + 8);
            break;
        }
        case SIGOPS_UNIT_K1N_INV:
            Sc<ModK1N>::inv_plain(out, in);
            break;
        case SIGOPS_UNIT_R1N_INV:
            Sc<ModR1N>::inv_plain(out, in);
            break;
        case SIGOPS_UNIT_K1N_INV_FERMAT: {
            u32 am[8], im[8];
            Sc<ModK1N>::to_mont(am, in);
            Sc<ModK1N>::minv(im, am);
            Sc<ModK1N>::from_mont(out, im);
            break;
        }
        case SIGOPS_UNIT_K1_INV_FERMAT: {
            Fe a, r;
            FpK1::from_plain(a, in);
            fe_inv_fermat((FpK1*)0, r, a);
            FpK1::to_plain(out, r);
            break;
        }
        case SIGOPS_UNIT_R1_INV_FERMAT: {
            Fe a, r;
            FpR1::from_plain(a, in);
            fe_inv_fermat((FpR1*)0, r, a);
            FpR1::to_plain(out, r);
            break;
        }
        case SIGOPS_UNIT_ED_INV_FERMAT: {
            Fe a, r;
            Fp25519::from_plain(a, in);
            fe_inv_fermat((Fp25519*)0, r, a);
            Fp25519::to_plain(out, r);
            break;
        }
        case SIGOPS_UNIT_EDL_REDUCE512:
            ed_reduce512(out, in);
            break;
        case SIGOPS_UNIT_SHA512_96:
            sha512_96(out, in);
            break;
        case SIGOPS_UNIT_SHA256_64:
            sha256_64(out, in);
            break;
        case SIGOPS_UNIT_K1_GLV: {
            GlvSplit s;
            k1_glv_split(s, in);
            for (int i = 0; i < 5; i++) {
                out[i] = s.k1[i];
                out[5 + i] = s.k2[i];
            }
            out[10] = s.neg1;
            out[11] = s.neg2;
            break;
        }
        case SIGOPS_UNIT_RAW_ADDSUB:
            if (in[0] == 0) unit_raw_addsub<FpK1>(out, in + 1);
            else if (in[0] == 1) unit_raw_addsub<FpR1>(out, in + 1);
            else unit_raw_addsub<Fp25519>(out, in + 1);
            break;
        case SIGOPS_UNIT_RAW_REDUCE16:
            if (in[0] == 0) FpK1::reduce16(out, in + 1);
            else Fp25519::reduce16(out, in + 1);
            break;
        case SIGOPS_UNIT_MUL8X8:
            mul8x8(out, in, in + 8);
            break;
        case SIGOPS_UNIT_SQR8:
            sqr8(out, in);
            break;
        case SIGOPS_UNIT_K1_MULPT: {
            const u32 zero[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            unit_double_mul<CurveK1>(out, zero, in, in + 8, tab, k1g);
            break;
        }
        case SIGOPS_UNIT_R1_MULPT: {
            const u32 zero[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            unit_double_mul<CurveR1>(out, zero, in, in + 8, tab, r1g);
            break;
        }
        case SIGOPS_UNIT_K1_DOUBLE_MUL:
            unit_double_mul<CurveK1>(out, in, in + 8, in + 16, tab, k1g);
            break;
        case SIGOPS_UNIT_R1_DOUBLE_MUL:
            unit_double_mul<CurveR1>(out, in, in + 8, in + 16, tab, r1g);
            break;
        case SIGOPS_UNIT_ED_MULPT: {
            // k*(x,y) by the same windowed ladder the verifier uses (table of multiples, signed 4-bit windows)
            Fe x, y;
            Fp25519::from_plain(x, in + 8);
            Fp25519::from_plain(y, in + 16);
            EdPoint P1, P2, P3, P4, T, acc;
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
            P2 = P1; ed_dbl<FE>(P2, true); ed_tab_store(tab, 1, P2);
            P3 = P2; ed_add_niels<FE>(P3, ypx, ymx, t2d, false, true); ed_tab_store(tab, 2, P3);
            P4 = P2; ed_dbl<FE>(P4, true); ed_tab_store(tab, 3, P4);
            T = P4; ed_add_niels<FE>(T, ypx, ymx, t2d, false, true); ed_tab_store(tab, 4, T);
            T = P3; ed_dbl<FE>(T, true); ed_tab_store(tab, 5, T);
            ed_add_niels<FE>(T, ypx, ymx, t2d, false, true); ed_tab_store(tab, 6, T);
            T = P4; ed_dbl<FE>(T, true); ed_tab_store(tab, 7, T);
            u32 kp[9];
            for (int i = 0; i < 8; i++) kp[i] = in[i];
            kp[8] = 0;
            recode_offset<9, 4, 65>(kp);
            ed_set_identity(acc);
            for (int i = 64; i >= 0; i--) {
                if (i != 64) {
                    for (int d = 0; d < 4; d++) ed_dbl<FE>(acc, d == 3);
                }
                ed_add_from_table<FE>(acc, tab, recode_digit<4>(kp, i), true);
            }
            Fe zi, ax, ay;
            fe_inv((FE*)0, zi, acc.Z);
            FE::mul(ax, acc.X, zi);
            FE::mul(ay, acc.Y, zi);
            FE::to_plain(out, ax);
            FE::to_plain(out + 8, ay);
            (void)edb;
            break;
        }
        default:
            break;
    }
}

#if defined(__CUDACC__)
__global__ void __launch_bounds__(kBlock) unit_kernel(int op, const u32* __restrict__ in, size_t n, u32* __restrict__ out,
                                                      Q4* __restrict__ scratch, const u32* k1g, const u32* r1g,
                                                      const u32* edb) {
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
#endif  // __CUDACC__

}  // namespace sigops
