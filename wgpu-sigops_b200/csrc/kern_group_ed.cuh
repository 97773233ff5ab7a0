// The ed25519 lane-group kernel (group.cuh): kGroupRolesEd cooperating warps per 32 signatures.  Instantiated by kern_edg.cu
// (field products inlined) and kern_edgc.cu (out of line).
#pragma once
#include <cuda_runtime.h>

#include "group.cuh"
#include "launch.h"

namespace sigops {

template <bool kCold>
__global__ void __launch_bounds__(kGroupRolesEd * 32) ed25519_verify_group_kernel(const Q4* __restrict__ sigs, const Q4* __restrict__ msgs,
                                                                                  const Q4* __restrict__ pks, size_t n,
                                                                                  uint8_t* __restrict__ valid,
                                                                                  const __grid_constant__ PTab btab) {
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
    for (size_t base = (size_t)blockIdx.x * kGroupSigs; base < n; base += (size_t)gridDim.x * kGroupSigs) {
        size_t i = base + lane;
        const bool live = i < n;
        if (!live) i = n - 1;
        u32 sig_w[16], msg_w[8], pk_w[8];
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
        const u32 v = ed_verify_group<kCold>(sig_w, msg_w, pk_w, tab, btab, g);
        if (role == 0 && live) valid[i] = (uint8_t)v;
        __syncthreads();
    }
}

template <bool kCold>
int launch_ed_group(const KLaunch& l, const void* sigs, const void* msgs, const void* pks, size_t n, uint8_t* valid, const PTab& btab) {
    ed25519_verify_group_kernel<kCold><<<l.grid, kGroupRolesEd * 32, kGroupEdSmem, l.stream>>>((const Q4*)sigs, (const Q4*)msgs,
                                                                                               (const Q4*)pks, n, valid, btab);
    return (int)cudaGetLastError();
}

template <bool kCold>
int setup_ed_group(int* max_blocks_per_sm) {
    cudaError_t e = cudaFuncSetAttribute(ed25519_verify_group_kernel<kCold>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGroupEdSmem);
    if (e != cudaSuccess) return (int)e;
    return (int)cudaOccupancyMaxActiveBlocksPerMultiprocessor(max_blocks_per_sm, ed25519_verify_group_kernel<kCold>, kGroupRolesEd * 32,
                                                              kGroupEdSmem);
}

}  // namespace sigops
