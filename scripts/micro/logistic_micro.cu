// Latency (dependent chain, one warp) and accuracy of the logistic coefficient c(u) = −μ y / (1 + exp(y u)):
//   ref  : exp() + __ddiv_rn  (CUDA math library: Horner polynomial + division subroutine)
//   fast : common.cuh fast_exp (Estrin) + fast_div (MUFU seed, cubic step, one exact-remainder correction)
// Accuracy is reported in ulps of c against long-double host arithmetic.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CIAO_MICRO 1
#include "../../ciaoalgorithms.jl_b200/csrc/fastmath.cuh"

__device__ __forceinline__ double coef_ref(double u, double y, double mu) {
    double e = exp(__dmul_rn(y, u));
    return __ddiv_rn(__dmul_rn(-mu, y), __dadd_rn(1.0, e));
}
__device__ __forceinline__ double coef_fast(double u, double y, double mu) { return logistic_coef_fast(u, y, mu); }

template <int WHICH>
__global__ void k_lat(double *out, int iters, long long *cyc) {
    double u = 0.3 + threadIdx.x * 0.0, c = 0.0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        c = WHICH == 0 ? coef_ref(u, 1.0, 1.0) : coef_fast(u, 1.0, 1.0);
        u = fma(c, 1e-3, u);  // next input depends on the previous output
    }
    long long t1 = clock64();
    out[threadIdx.x] = c + u;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
template <int WHICH>
__global__ void k_exp_lat(double *out, int iters, long long *cyc) {
    double u = 0.3, c = 0.0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        c = WHICH == 0 ? exp(u) : fast_exp(u);
        u = fma(c, 1e-9, 0.3);
    }
    long long t1 = clock64();
    out[threadIdx.x] = c + u;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_eval(const double *u, const double *y, double *cref, double *cfast, double *efast, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        cref[i] = coef_ref(u[i], y[i], 1.0);
        cfast[i] = coef_fast(u[i], y[i], 1.0);
        efast[i] = fast_exp(u[i]);
    }
}
static double ulps(double got, long double want) {
    if (want == 0) return got == 0 ? 0 : 1e9;
    int e;
    frexpl(want, &e);
    long double ulp = ldexpl(1.0L, e - 53);
    return (double)fabsl(((long double)got - want) / ulp);
}
int main() {
    double *out; long long *cyc, h;
    cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 8);
    int it = 20000;
    k_lat<0><<<1, 32>>>(out, it, cyc); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("coef exp()+__ddiv_rn : %.1f cycles\n", (double)h / it - 8);
    k_lat<1><<<1, 32>>>(out, it, cyc); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("coef fast            : %.1f cycles\n", (double)h / it - 8);
    k_exp_lat<0><<<1, 32>>>(out, it, cyc); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("exp()                : %.1f cycles\n", (double)h / it - 8);
    k_exp_lat<1><<<1, 32>>>(out, it, cyc); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("fast_exp             : %.1f cycles\n", (double)h / it - 8);

    const int n = 1 << 20;
    double *hu = (double *)malloc(n * 8), *hy = (double *)malloc(n * 8), *hr = (double *)malloc(n * 8), *hf = (double *)malloc(n * 8),
           *he = (double *)malloc(n * 8);
    srand(1);
    for (int i = 0; i < n; ++i) {
        double r = rand() / (double)RAND_MAX;
        double span = (i % 4 == 0) ? 745.0 : (i % 4 == 1 ? 40.0 : (i % 4 == 2 ? 2.0 : 1e-3));
        hu[i] = (2 * r - 1) * span;
        hy[i] = (rand() & 1) ? 1.0 : -1.0;
    }
    hu[0] = 0.0; hu[1] = 1000.0; hu[2] = -1000.0; hu[3] = 709.9; hu[4] = -709.9; hu[5] = 1e-300; hy[1] = hy[2] = hy[3] = hy[4] = 1.0;
    double *du, *dy, *dr, *df, *de;
    cudaMalloc(&du, n * 8); cudaMalloc(&dy, n * 8); cudaMalloc(&dr, n * 8); cudaMalloc(&df, n * 8); cudaMalloc(&de, n * 8);
    cudaMemcpy(du, hu, n * 8, cudaMemcpyHostToDevice); cudaMemcpy(dy, hy, n * 8, cudaMemcpyHostToDevice);
    k_eval<<<n / 256, 256>>>(du, dy, dr, df, de, n);
    cudaMemcpy(hr, dr, n * 8, cudaMemcpyDeviceToHost); cudaMemcpy(hf, df, n * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(he, de, n * 8, cudaMemcpyDeviceToHost);
    double mr = 0, mf = 0, me = 0, mabs = 0; int nbad = 0;
    for (int i = 0; i < n; ++i) {
        long double t = (long double)hy[i] * hu[i];
        long double want = -(long double)hy[i] / (1.0L + expl(t));
        double a = ulps(hr[i], want), b = ulps(hf[i], want);
        if (fabsl(want) > 1e-290L) {  // below that the clamped exponent changes bits that no tolerance can see
            if (a > mr) mr = a;
            if (b > mf) { mf = b; }
            if (b > 4 && nbad++ < 5) printf("  bad: u=%.17g y=%g ref=%.17g fast=%.17g want=%.17Lg\n", hu[i], hy[i], hr[i], hf[i], want);
        } else {
            double ad = fabs(hf[i] - (double)want);
            if (ad > mabs) mabs = ad;
        }
        if (fabs(hu[i]) < 700) { double c = ulps(he[i], expl((long double)hu[i])); if (c > me) me = c; }
    }
    printf("max ulp error of c: library %.2f, fast %.2f; fast_exp %.2f ulp; |c| < 1e-290: max abs diff %.3g\n", mr, mf, me, mabs);
    printf("edge: u=0 %.17g, u=1000 %.3g, u=-1000 %.17g, u=709.9 %.3g, u=-709.9 %.17g (%s)\n", hf[0], hf[1], hf[2], hf[3], hf[4],
           cudaGetErrorString(cudaGetLastError()));
    return 0;
}
