// fp64 pipe microbenchmark: dependent-chain latency and per-SM throughput of DFMA/DMUL/DADD on this GPU.
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void dfma_kernel(double *out, double a, double b, int iters, long long *cycles) {
    double x[ILP];
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
    }
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
template <int ILP>
void run(int threads, int iters) {
    double *out; long long *cyc, h;
    cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
    dfma_kernel<ILP><<<148, threads>>>(out, 1.0000001, 1e-9, iters, cyc);
    cudaDeviceSynchronize();
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    dfma_kernel<ILP><<<148, threads>>>(out, 1.0000001, 1e-9, iters, cyc);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double per = (double)h / ((double)iters * ILP);
    printf("threads/SM %4d ILP %2d: %.2f cycles per DFMA per warp (chain cycles/iter %.1f), chip %.2f TFLOP/s, per SM %.2f DFMA/clk\n", threads, ILP, per,
           (double)h / iters, 2.0 * 148 * threads * (double)iters * ILP / (ms * 1e-3) / 1e12, (double)threads * ILP * iters / (double)h);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<1>(32, 20000); run<2>(32, 20000); run<4>(32, 20000); run<8>(32, 20000); run<16>(32, 10000);
    run<1>(128, 20000); run<4>(128, 20000); run<8>(128, 20000);
    run<4>(256, 20000); run<8>(512, 10000); run<8>(1024, 10000);
    return 0;
}
