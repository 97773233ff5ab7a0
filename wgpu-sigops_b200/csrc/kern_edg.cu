// ed25519 ecverify, lane-group kernel (small batches, field products inlined), and the lane-group unit-test shim: kernels + launchers.
#include <cuda_runtime.h>

#include "kern_group_ed.cuh"
#include "../../include/sigops.h"

using namespace sigops;

namespace sigops {

// unit shim: one item per lane, the roles of a block cooperate on 32 items (ops SIGOPS_UNIT_*_GROUP_*)
__global__ void __launch_bounds__(kGroupRolesSw * 32) unit_group_kernel(int op, const u32* __restrict__ in, size_t n, u32* __restrict__ out,
                                                                        const __grid_constant__ PTab k1g,
                                                                        const __grid_constant__ PTab r1g) {
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
    const int in_w = op == SIGOPS_UNIT_ED_GROUP_MULPT ? 24 : 32, out_w = op == SIGOPS_UNIT_ED_GROUP_MULPT ? 16 : 17;
    for (size_t base = (size_t)blockIdx.x * kGroupSigs; base < n; base += (size_t)gridDim.x * kGroupSigs) {
        size_t i = base + lane;
        const bool live = i < n;
        if (!live) i = n - 1;
        u32 a[32], r[17];
        for (int j = 0; j < 32; j++) a[j] = j < in_w ? in[i * in_w + j] : 0u;
        for (int j = 0; j < 17; j++) r[j] = 0;
        bool writer = false;
        if (op == SIGOPS_UNIT_K1_GROUP_DOUBLE_MUL)
            writer = unit_double_mul_g<CurveK1>(r, a, a + 8, a + 16, tab, k1g, g);
        else if (op == SIGOPS_UNIT_R1_GROUP_DOUBLE_MUL)
            writer = unit_double_mul_g<CurveR1>(r, a, a + 8, a + 16, tab, r1g, g);
        else  // launched with kGroupRolesEd warps per block
            writer = unit_ed_mulpt_g(r, a, tab, g);
        if (writer && live)
            for (int j = 0; j < out_w; j++) out[i * out_w + j] = r[j];
        __syncthreads();
    }
}

int kl_ed_group(const KLaunch& l, const void* sigs, const void* msgs, const void* pks, size_t n, uint8_t* valid, const PTab& btab) {
    return launch_ed_group<false>(l, sigs, msgs, pks, n, valid, btab);
}
int kl_ed_group_setup(int* max_blocks_per_sm) {
    cudaError_t e = cudaFuncSetAttribute(unit_group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGroupSwSmem);
    if (e != cudaSuccess) return (int)e;
    return setup_ed_group<false>(max_blocks_per_sm);
}
int kl_unit_group(const KLaunch& l, int op, const u32* in, size_t n, u32* out, const PTab& k1g, const PTab& r1g) {
    const int roles = op == SIGOPS_UNIT_ED_GROUP_MULPT ? kGroupRolesEd : kGroupRolesSw;
    unit_group_kernel<<<l.grid, roles * 32, kGroupSwSmem, l.stream>>>(op, in, n, out, k1g, r1g);
    return (int)cudaGetLastError();
}

}  // namespace sigops
