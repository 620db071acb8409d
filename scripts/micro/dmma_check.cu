#include <cstdio>
#include "../../ciaoalgorithms.jl_b200/csrc/common.cuh"
__global__ void k(double *out) {
    double v = (double)(threadIdx.x * threadIdx.x + 1) * 0.125;
    out[threadIdx.x] = warp_sum_mma(v, threadIdx.x);
}
int main() {
    double *o, h[32]; cudaMalloc(&o, 256); k<<<1, 32>>>(o); cudaMemcpy(h, o, 256, cudaMemcpyDeviceToHost);
    double want = 0; for (int i = 0; i < 32; ++i) want += (i * i + 1) * 0.125;
    int same = 1; for (int i = 0; i < 32; ++i) same &= (h[i] == h[0]);
    printf("dmma warp sum: got %.6f want %.6f all-lanes-equal %d (%s)\n", h[0], want, same, cudaGetErrorString(cudaGetLastError()));
    return !(h[0] == want && same);
}
