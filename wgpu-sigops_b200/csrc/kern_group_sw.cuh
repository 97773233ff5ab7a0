// The lane-group ecrecover kernel (group.cuh): kGroupRolesSw cooperating warps per 32 signatures.  Instantiated by
// kern_k1g.cu / kern_r1g.cu.
#pragma once
#include <cuda_runtime.h>

#include "group.cuh"
#include "launch.h"

using namespace sigops;

namespace sigops {

template <class C>
__global__ void __launch_bounds__(kGroupRolesSw * 32) ecrecover_group_kernel(const Q4* __restrict__ sigs, const Q4* __restrict__ msgs,
                                                                             size_t n, Q4* __restrict__ out,
                                                                             uint8_t* __restrict__ status,
                                                                             const __grid_constant__ PTab gtab) {
    extern __shared__ __align__(16) u32 sg_group_smem[];
    const int lane = threadIdx.x & 31, role = threadIdx.x >> 5;
    Q4* mb = reinterpret_cast<Q4*>(sg_group_smem);
    u32* sc = sg_group_smem + kMbSlots * 8 * kGroupSigs;
    Q4* tabq = reinterpret_cast<Q4*>(sc + kScWords * kGroupSigs);
    GroupCtx g;
    g.role = role;
    g.mb = mb + lane;
    g.sc = sc + lane;
    TabRef tab;
    tab.base = tabq + lane;
    tab.stride = kGroupSigs;
    // every block walks groups of 32 signatures; rows past the end are clamped (the lane redoes the last signature so that
    // it reaches every barrier) and dropped on store
    for (size_t base = (size_t)blockIdx.x * kGroupSigs; base < n; base += (size_t)gridDim.x * kGroupSigs) {
        size_t i = base + lane;
        const bool live = i < n;
        if (!live) i = n - 1;
        u32 sig_w[16], msg_w[8], out_w[16], st = 0;
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
        const bool writer = sw_ecrecover_group<C>(sig_w, msg_w, out_w, &st, tab, gtab, g);
        if (writer && live) {
#pragma unroll
            for (int q = 0; q < 4; q++) {
                Q4 v = {out_w[4 * q + 0], out_w[4 * q + 1], out_w[4 * q + 2], out_w[4 * q + 3]};
                out[4 * i + q] = v;
            }
            if (status) status[i] = (uint8_t)st;
        }
        __syncthreads();  // the mailbox and the table are reused by the next group
    }
}

template <class C>
int launch_ecrecover_group(const KLaunch& l, const void* sigs, const void* msgs, size_t n, void* out, uint8_t* status,
                           const PTab& gtab) {
    ecrecover_group_kernel<C><<<l.grid, kGroupRolesSw * 32, kGroupSwSmem, l.stream>>>((const Q4*)sigs, (const Q4*)msgs, n, (Q4*)out,
                                                                                     status, gtab);
    return (int)cudaGetLastError();
}

template <class C>
int setup_ecrecover_group(int* max_blocks_per_sm) {
    cudaError_t e = cudaFuncSetAttribute(ecrecover_group_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGroupSwSmem);
    if (e != cudaSuccess) return (int)e;
    return (int)cudaOccupancyMaxActiveBlocksPerMultiprocessor(max_blocks_per_sm, ecrecover_group_kernel<C>, kGroupRolesSw * 32,
                                                              kGroupSwSmem);
}

}  // namespace sigops
