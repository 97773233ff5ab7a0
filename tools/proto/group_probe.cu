// Cycles per lane-group operation (development tool): one block of `roles` warps on one SM, N dependent group operations.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o group_probe tools/proto/group_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../../wgpu-sigops_b200/csrc/group.cuh"
using namespace sigops;

template <int MODE>
__global__ void probe(u32* out, long long* cyc, int iters) {
    extern __shared__ __align__(16) u32 smem[];
    const int lane = threadIdx.x & 31, role = threadIdx.x >> 5;
    GroupCtx g;
    g.role = role;
    g.mb = reinterpret_cast<Q4*>(smem) + lane;
    g.sc = smem + kMbSlots * 8 * kGroupSigs + lane;
    Fe X, Y, Z, x2, y2;
    for (int i = 0; i < 8; i++) {
        X.v[i] = 0x9e3779b9u * (lane + 1 + i);
        Y.v[i] = 0x85ebca6bu * (lane + 7 + 3 * i);
        Z.v[i] = X.v[i] ^ 0x1234567u;
        x2.v[i] = X.v[i] + 0x7654321u;
        y2.v[i] = X.v[i] * 3u;
    }
    EdPoint P;
    P.X = X; P.Y = Y; P.Z = Z; P.T = x2;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) pj_dbl_g<CurveK1>(X, Y, Z, g);
        if (MODE == 1) pj_madd_g<CurveK1>(X, Y, Z, x2, y2, true, g);
        if (MODE == 2) pj_dbl_g<CurveR1>(X, Y, Z, g);
        if (MODE == 3) pj_madd_g<CurveR1>(X, Y, Z, x2, y2, true, g);
        if (MODE == 4) ed_dbl_g<Inl<Fp25519> >(P, g);
        if (MODE == 5) ed_add_g<Inl<Fp25519> >(P, x2, y2, X, Y, true, false, true, g);
        if (MODE == 6) { g.put(0, X); g.sync(); g.get(X, 0); }            // one exchange: put + barrier + get
        if (MODE == 7) g.sync();                                           // bare barrier
        if (MODE == 8) { g.put(0, X); g.put(1, Y); g.get(X, 1); g.get(Y, 0); }  // 2 puts + 2 gets, no barrier
    }
    long long t1 = clock64();
    u32 acc = 0;
    for (int i = 0; i < 8; i++) acc ^= X.v[i] ^ Y.v[i] ^ Z.v[i] ^ P.X.v[i] ^ P.Y.v[i] ^ P.Z.v[i] ^ P.T.v[i];
    out[threadIdx.x] = acc;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(const char* name, int roles) {
    u32* out;
    long long* cyc;
    cudaMalloc(&out, 4 * 32 * 8);
    cudaMalloc(&cyc, 8);
    const int iters = 1000;
    cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    for (int rep = 0; rep < 2; rep++) probe<MODE><<<1, 32 * roles, 65536>>>(out, cyc, iters);
    long long h = 0;
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    cudaError_t e = cudaDeviceSynchronize();
    printf("%-40s roles %d: %8.1f cycles per op %s\n", name, roles, (double)h / iters, e ? cudaGetErrorString(e) : "");
}

int main() {
    run<0>("k1 pj_dbl_g", 6);
    run<0>("k1 pj_dbl_g", 4);
    run<1>("k1 pj_madd_g", 6);
    run<2>("r1 pj_dbl_g", 6);
    run<3>("r1 pj_madd_g", 6);
    run<4>("ed ed_dbl_g", 4);
    run<5>("ed ed_add_g (cached)", 4);
    run<6>("put + barrier + get", 4);
    run<6>("put + barrier + get", 6);
    run<7>("bare barrier", 4);
    run<7>("bare barrier", 6);
    run<8>("2 puts + 2 gets", 4);
    return 0;
}
