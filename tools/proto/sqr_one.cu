#include <cstdio>
#include <cuda_runtime.h>
#include "../../wgpu-sigops_b200/csrc/kernels.cuh"
using namespace sigops;
__global__ void k(const u32* in, u32* out) {
    u32 a[8] = {1, 0, 0, 0, 0, 0, 0, 0}, r[16], d[16];
    sqr8(r, a);                                  // constant input
    for (int i = 0; i < 16; i++) out[i] = r[i];
    for (int i = 0; i < 16; i++) d[i] = 0;
    mad_diag8(d, a);
    for (int i = 0; i < 16; i++) out[16 + i] = d[i];
    u32 b[8] = {3, 0, 0, 0, 5, 0, 0, 0};
    for (int i = 0; i < 16; i++) d[i] = 0;
    mad_diag8(d, b);
    for (int i = 0; i < 16; i++) out[32 + i] = d[i];
    u32 e[16], o[16];
    for (int i = 0; i < 16; i++) { e[i] = 0; o[i] = 0; }
    e[0] = 7; o[3] = 9;
    merge_even_odd(d, e, o);
    for (int i = 0; i < 16; i++) out[48 + i] = d[i];
}
int main() {
    u32 *d, *o, ho[64];
    cudaMalloc(&d, 256); cudaMalloc(&o, sizeof ho);
    k<<<1, 32>>>(d, o); cudaMemcpy(ho, o, sizeof ho, cudaMemcpyDeviceToHost);
    for (int j = 0; j < 4; j++) { for (int i = 0; i < 16; i++) printf("%x ", ho[16 * j + i]); printf("\n"); }
    return 0;
}
