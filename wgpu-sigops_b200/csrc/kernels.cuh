// Shared pieces of the __global__ entry points (kern_*.cu): one fused kernel per curve (replaces the 5 / 6 / 7 dispatches of
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
        case SIGOPS_UNIT_K1_GROUP_DOUBLE_MUL: case SIGOPS_UNIT_R1_GROUP_DOUBLE_MUL:
            in_w = 32;
            out_w = 17;
            break;
        case SIGOPS_UNIT_ED_GROUP_MULPT:
            in_w = 24;
            out_w = 16;
            break;
        case SIGOPS_UNIT_ED_FIXED_MUL:
            out_w = 16;
            break;
        case SIGOPS_UNIT_RAW_ADDSUB:
            in_w = 17;
            out_w = 16;
            break;
        case SIGOPS_UNIT_RAW_REDUCE16:
            in_w = 17;
            out_w = 8;
            break;
        case SIGOPS_UNIT_RAW_SHL:
            in_w = 10;
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
SG_HD void unit_raw_shl(u32* out, const u32* in) {
    Fe a, r;
    copy8(a.v, in + 1);
    if (in[0] == 2)
        F::template shl<2>(r, a);
    else
        F::template shl<3>(r, a);
    copy8(out, r.v);
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
SG_HD void unit_double_mul(u32* out, const u32* u1, const u32* u2, const u32* xy, const TabRef& tab, const PTab& gtab) {
    typedef typename C::F F;
    Fe x, y;
    F::from_plain(x, xy);
    F::from_plain(y, xy + 8);
    sw_build_table<C>(tab, x, y);
    JacPoint Q;
    sw_double_mul<C, false>(Q, u1, u2, tab, gtab);
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
SG_HD void unit_dispatch(int op, u32* out, const u32* in, const TabRef& tab, const PTab& k1g, const PTab& r1g,
                         const PTab& edb) {
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
        case SIGOPS_UNIT_RAW_SHL:
            if (in[0] == 0) unit_raw_shl<FpK1>(out, in + 1);
            else if (in[0] == 1) unit_raw_shl<FpR1>(out, in + 1);
            else unit_raw_shl<Fp25519>(out, in + 1);
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
            break;
        }
        case SIGOPS_UNIT_ED_FIXED_MUL: {
            // s * B through the positional table alone: the table's sum is 2^-252 s B, 252 doublings undo the scale
            EdPoint acc;
            ed_ptab_sum<FE>(acc, in, edb);
            for (int i = 0; i < kPTabShiftEd; i++) ed_dbl<FE>(acc, false);
            Fe zi, ax, ay;
            fe_inv((FE*)0, zi, acc.Z);
            FE::mul(ax, acc.X, zi);
            FE::mul(ay, acc.Y, zi);
            FE::to_plain(out, ax);
            FE::to_plain(out + 8, ay);
            break;
        }
        default:
            break;
    }
}


}  // namespace sigops

