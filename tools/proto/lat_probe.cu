// Latency probe (development tool, not part of the library): cycles per field operation / point doubling for ONE warp per
// SM sub-partition, with dependent and with interleaved independent chains.  Drives the design of the small-batch kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 [-maxrregcount=N] -o lat_probe tools/proto/lat_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../../wgpu-sigops_b200/csrc/kernels.cuh"
using namespace sigops;

template <int MODE>
__global__ void probe(u32* out, long long* cyc, int iters) {
    Fe x, y, x1, x2, x3;
    for (int i = 0; i < 8; i++) {
        x.v[i] = 0x9e3779b9u * (threadIdx.x + 1 + i);
        y.v[i] = 0x85ebca6bu * (threadIdx.x + 7 + 3 * i);
        x1.v[i] = x.v[i] ^ 0x1234567u;
        x2.v[i] = x.v[i] + 0x7654321u;
        x3.v[i] = x.v[i] * 3u;
    }
    JacPoint P, Q;
    P.X = x; P.Y = y; P.Z = x1; P.inf = false;
    Q.X = x2; Q.Y = x3; Q.Z = y; Q.inf = false;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) x = FpK1::mul_body(x, y);                      // dependent, inlined
        if (MODE == 1) x = FpK1::mul_(x, y);                          // dependent, out of line
        if (MODE == 2) { x = FpK1::mul_body(x, y); x1 = FpK1::mul_body(x1, y); x2 = FpK1::mul_body(x2, y); }  // 3 independent
        if (MODE == 3) x = FpK1::sqr_body(x);
        if (MODE == 4) { x = FpK1::sqr_body(x); x1 = FpK1::sqr_body(x1); x2 = FpK1::sqr_body(x2); }
        if (MODE == 5) jac_dbl<CurveK1::Hot>(P);                      // inlined products
        if (MODE == 6) jac_dbl<CurveK1>(P);                           // out-of-line products
        if (MODE == 7) { jac_dbl<CurveK1::Hot>(P); jac_dbl<CurveK1::Hot>(Q); }  // two independent points
        if (MODE == 8) jac_madd<CurveK1::Hot>(P, x2, x3);
        if (MODE == 9) { FpK1::add(x, x, y); FpK1::sub(x, x, x1); }   // two dependent add/sub
        if (MODE == 10) x = FpR1::mul_body(x, y);
        if (MODE == 11) x = Fp25519::mul_body(x, y);
        if (MODE == 12) { x = FpK1::mul_body(x, y); x1 = FpK1::mul_body(x1, y); }
    }
    long long t1 = clock64();
    u32 acc = 0;
    for (int i = 0; i < 8; i++) acc ^= x.v[i] ^ x1.v[i] ^ x2.v[i] ^ P.X.v[i] ^ P.Y.v[i] ^ P.Z.v[i] ^ Q.X.v[i] ^ Q.Y.v[i] ^ Q.Z.v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(const char* name, int per_iter, int warps) {
    u32* out;
    long long* cyc;
    cudaMalloc(&out, 4 * 32 * 64);
    cudaMalloc(&cyc, 8);
    const int iters = 2000;
    for (int rep = 0; rep < 2; rep++) probe<MODE><<<1, 32 * warps>>>(out, cyc, iters);
    long long h = 0;
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    cudaError_t e = cudaDeviceSynchronize();
    printf("%-44s warps/SM %2d: %8.1f cycles per op%s\n", name, warps, (double)h / iters / per_iter, e ? cudaGetErrorString(e) : "");
    cudaFree(out);
    cudaFree(cyc);
}

int main() {
    for (int w : {1, 4, 8, 16}) {
        // w warps in ONE block on one SM: w/4 per scheduler (w = 1: a single warp)
        run<0>("k1 mul, dependent, inlined", 1, w);
        run<1>("k1 mul, dependent, out of line", 1, w);
        run<12>("k1 mul x2 independent (per product)", 2, w);
        run<2>("k1 mul x3 independent (per product)", 3, w);
        run<3>("k1 sqr, dependent, inlined", 1, w);
        run<4>("k1 sqr x3 independent (per square)", 3, w);
        run<5>("k1 jac_dbl, inlined products", 1, w);
        run<6>("k1 jac_dbl, out-of-line products", 1, w);
        run<7>("k1 jac_dbl x2 independent (per doubling)", 2, w);
        run<8>("k1 jac_madd, inlined products", 1, w);
        run<9>("k1 add+sub dependent (per pair)", 1, w);
        run<10>("r1 mul, dependent, inlined", 1, w);
        run<11>("25519 mul, dependent, inlined", 1, w);
    }
    return 0;
}
