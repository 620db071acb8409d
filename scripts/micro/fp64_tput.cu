// Throughput of the fp64 pipe per SM sub-partition: independent streams of DFMA, DADD, DMMA (m8n8k4) and double shuffles,
// 1 / 2 / 4 warps per sub-partition (one CTA on one SM).  Prints cycles per warp instruction per sub-partition.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int OP>
__global__ void k(double *out, int iters, long long *cyc) {
    double x[8], a = threadIdx.x * 1e-3 + 1.0, b = 1.0 + 1e-9 * threadIdx.x;
    double y[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = i + a; y[i] = i - a; }
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) x[i] = fma(x[i], a, b);
            if (OP == 1) x[i] = x[i] + b;
            if (OP == 2) dmma(x[i], y[i], a, b);
            if (OP == 3) x[i] += __shfl_xor_sync(0xffffffffu, x[i], 1 + (it & 15));
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i] + y[i];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
    double *out; long long *cyc, h; cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 8);
    const int it = 4000;
    const char *names[4] = {"DFMA", "DADD", "DMMA.8x8x4", "SHFL.f64 + DADD"};
    for (int op = 0; op < 4; ++op)
        for (int wps = 1; wps <= 4; wps *= 2) {
            const int T = 128 * wps;
            if (op == 0) k<0><<<1, T>>>(out, it, cyc);
            if (op == 1) k<1><<<1, T>>>(out, it, cyc);
            if (op == 2) k<2><<<1, T>>>(out, it, cyc);
            if (op == 3) k<3><<<1, T>>>(out, it, cyc);
            cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            printf("%-16s %d warps per sub-partition: %.2f cycles per warp instruction per sub-partition (%s)\n", names[op], wps,
                   (double)h / ((double)it * 8 * wps), cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
