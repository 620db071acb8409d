// Latency of a dependent chain of fp64 tensor-core MMAs (mma.sync.m8n8k4.f64) vs a shuffle-add round, single warp.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b, double c0, double c1) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
                 : "=d"(d0), "=d"(d1) : "d"(a), "d"(b), "d"(c0), "d"(c1));
}
__global__ void k_dmma(double *out, int iters, long long *cyc) {
    double a = threadIdx.x * 1e-3 + 1.0, d0 = 0, d1 = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) { dmma(d0, d1, a, 1.0, d0, d1); a = d0 * 1e-9 + 1.0; }  // A depends on the previous D
    long long t1 = clock64();
    out[threadIdx.x] = d0 + d1; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_dmma_acc(double *out, int iters, long long *cyc) {
    double a = threadIdx.x * 1e-3 + 1.0, d0 = 0, d1 = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) dmma(d0, d1, a, 1.0, d0, d1);   // accumulator chain only
    long long t1 = clock64();
    out[threadIdx.x] = d0 + d1; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_shfl(double *out, int iters, long long *cyc) {
    double v = threadIdx.x * 1e-3 + 1.0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) v += __shfl_xor_sync(0xffffffffu, v, 1 << (i % 5));
    long long t1 = clock64();
    out[threadIdx.x] = v; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_check(double *out) {   // total of the 32 lane values via DMMA, SHFL, DMMA, DADD
    double v = (double)(threadIdx.x + 1) * 0.125;
    double s0, s1; dmma(s0, s1, v, 1.0, 0.0, 0.0);              // s0 = s1 = sum over the 4 lanes of my group (lane>>2)
    int L = threadIdx.x, src = 4 * ((L & 3) + 4 * ((L >> 2) & 1));
    double b = __shfl_sync(0xffffffffu, s0, src);
    double t0, t1; dmma(t0, t1, 1.0, b, 0.0, 0.0);
    out[threadIdx.x] = t0 + t1;
}
int main() {
    double *out; long long *cyc, h; cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 8);
    int it = 20000;
    k_dmma<<<1, 32>>>(out, it, cyc); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("DMMA + dependent DFMA chain: %.1f cycles per iteration\n", (double)h / it);
    k_dmma_acc<<<1, 32>>>(out, it, cyc); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("DMMA accumulator chain: %.1f cycles per DMMA\n", (double)h / it);
    k_shfl<<<1, 32>>>(out, it, cyc); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("shfl_xor(double) + DADD round: %.1f cycles\n", (double)h / it);
    k_check<<<1, 32>>>(out); double ho[32]; cudaMemcpy(ho, out, 256, cudaMemcpyDeviceToHost);
    printf("DMMA tree total: lane0 %.4f lane17 %.4f lane31 %.4f (expected %.4f) err %s\n", ho[0], ho[17], ho[31], 0.125 * 32 * 33 / 2, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
