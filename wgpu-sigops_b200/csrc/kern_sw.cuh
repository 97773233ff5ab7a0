// The fused ecrecover kernel (one per short-Weierstrass curve) and its launcher; instantiated by kern_k1.cu / kern_r1.cu,
// one translation unit per curve so that the library builds in parallel.
#pragma once
#include <cuda_runtime.h>

#include "kernels.cuh"
#include "launch.h"

using namespace sigops;

namespace sigops {

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

template <class C>
int launch_ecrecover(const KLaunch& l, const void* sigs, const void* msgs, size_t n, void* out, uint8_t* status, void* scratch,
                     const u32* gtab, u32 smem_words) {
    ecrecover_kernel<C><<<l.grid, l.tpb, (size_t)smem_words * 4, l.stream>>>((const Q4*)sigs, (const Q4*)msgs, n, (Q4*)out, status,
                                                                            (Q4*)scratch, gtab, smem_words);
    return (int)cudaGetLastError();
}

template <class C>
int setup_ecrecover(int* max_blocks_per_sm) {
    cudaError_t e = cudaFuncSetAttribute(ecrecover_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGTabEntries * 16 * 4);
    if (e != cudaSuccess) return (int)e;
    return (int)cudaOccupancyMaxActiveBlocksPerMultiprocessor(max_blocks_per_sm, ecrecover_kernel<C>, kBlock, 0);
}

}  // namespace sigops
