// ed25519 ecverify kernels (fixed 32-byte messages; variable-length / strict) and their launchers
// (replaces src/wgsl/main/ed25519_eddsa_main*.wgsl).
#include <cuda_runtime.h>

#include "kernels.cuh"
#include "launch.h"

using namespace sigops;

namespace sigops {

__global__ void __launch_bounds__(kBlock, SG_MINB_ED) ed25519_verify_kernel(const Q4* __restrict__ sigs, const Q4* __restrict__ msgs,
                                                                            const Q4* __restrict__ pks, size_t n,
                                                                            uint8_t* __restrict__ valid, Q4* __restrict__ scratch,
                                                                            const __grid_constant__ PTab btab) {
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
    const __grid_constant__ PTab btab) {
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

int kl_ed_verify(const KLaunch& l, const void* sigs, const void* msgs, const void* pks, size_t n, uint8_t* valid, void* scratch,
                 const PTab& btab) {
    ed25519_verify_kernel<<<l.grid, l.tpb, 0, l.stream>>>((const Q4*)sigs, (const Q4*)msgs, (const Q4*)pks, n, valid, (Q4*)scratch, btab);
    return (int)cudaGetLastError();
}
int kl_ed_verify_msgs(const KLaunch& l, const void* sigs, const uint8_t* msg_bytes, const unsigned long long* msg_off, const void* pks,
                      size_t n, int strict, uint8_t* valid, void* scratch, const PTab& btab) {
    ed25519_verify_msgs_kernel<<<l.grid, l.tpb, 0, l.stream>>>((const Q4*)sigs, msg_bytes, msg_off, (const Q4*)pks, n, strict, valid,
                                                              (Q4*)scratch, btab);
    return (int)cudaGetLastError();
}
int kl_ed_setup(int* max_blocks_per_sm, int* max_blocks_per_sm_msgs) {
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(max_blocks_per_sm, ed25519_verify_kernel, kBlock, 0);
    if (e != cudaSuccess) return (int)e;
    return (int)cudaOccupancyMaxActiveBlocksPerMultiprocessor(max_blocks_per_sm_msgs, ed25519_verify_msgs_kernel, kBlock, 0);
}

}  // namespace sigops
