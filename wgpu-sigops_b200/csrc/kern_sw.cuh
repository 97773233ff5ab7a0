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
                                                                       Q4* __restrict__ scratch, const __grid_constant__ PTab gtab) {
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
        sw_ecrecover_batch<C, kInnerSync>(B, io, tab, gtab);
    }
}

template <class C>
int launch_ecrecover(const KLaunch& l, const void* sigs, const void* msgs, size_t n, void* out, uint8_t* status, void* scratch,
                     const PTab& gtab) {
    ecrecover_kernel<C><<<l.grid, l.tpb, 0, l.stream>>>((const Q4*)sigs, (const Q4*)msgs, n, (Q4*)out, status, (Q4*)scratch, gtab);
    return (int)cudaGetLastError();
}

template <class C>
int setup_ecrecover(int* max_blocks_per_sm) {
    return (int)cudaOccupancyMaxActiveBlocksPerMultiprocessor(max_blocks_per_sm, ecrecover_kernel<C>, kBlock, 0);
}

}  // namespace sigops
