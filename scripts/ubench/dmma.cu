// Micro-benchmark: issue / pipe cost of DFMA vs DMMA (mma.sync.m8n8k4.f64) vs a mix, one CTA of 896 threads per SM.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b, double c0, double c1) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
                 : "=d"(d0), "=d"(d1) : "d"(a), "d"(b), "d"(c0), "d"(c1));
}
template <int MODE>   // 0: 8 DFMA per iteration; 1: 2 DMMA per iteration; 2: 4 DFMA + 1 DMMA; 3: 8 DFMA + 16 IADD; 4: 2 DMMA + 16 IADD
__global__ void __launch_bounds__(896, 1) k(double *out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = 1, x2 = 2, x3 = 3, x4 = 4, x5 = 5, x6 = 6, x7 = 7;
    double m0 = 0, m1 = 0, m2 = 0, m3 = 0;
    int i0 = threadIdx.x, i1 = 1, i2 = 2, i3 = 3;
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0 || MODE == 3) {
            x0 = fma(a, x0, b); x1 = fma(a, x1, b); x2 = fma(a, x2, b); x3 = fma(a, x3, b);
            x4 = fma(a, x4, b); x5 = fma(a, x5, b); x6 = fma(a, x6, b); x7 = fma(a, x7, b);
        }
        if (MODE == 1 || MODE == 4) { dmma(m0, m1, a, b, m0, m1); dmma(m2, m3, a, b, m2, m3); }
        if (MODE == 2) {
            x0 = fma(a, x0, b); x1 = fma(a, x1, b); x2 = fma(a, x2, b); x3 = fma(a, x3, b);
            dmma(m0, m1, a, b, m0, m1);
        }
        if (MODE == 3 || MODE == 4) {
#pragma unroll
            for (int r = 0; r < 4; ++r) { i0 = i0 * 3 + i1; i1 = i1 * 5 + i2; i2 = i2 * 7 + i3; i3 = i3 * 9 + i0; }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7 + m0 + m1 + m2 + m3 + i0 + i1 + i2 + i3;
}
template <int MODE> void run(const char *name, double *d, int fma_per_thread_iter) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 200000;
    k<MODE><<<148, 896>>>(d, 1000, 0.999, 0.001);
    cudaEventRecord(e0);
    k<MODE><<<148, 896>>>(d, iters, 0.999, 0.001);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double cyc = ms * 1e-3 * 1.965e9;                 // SM cycles
    printf("%-28s %8.3f ms  %7.2f cycles per iteration and scheduler (7 warps each)  = %.2f cycles per warp-iteration; %.1f fp64 FMA/clk/SM\n",
           name, ms, cyc / iters, cyc / iters / 7.0, (double)fma_per_thread_iter * 896 * iters / cyc);
}
int main() {
    double *d; cudaMalloc(&d, 148 * 896 * 8);
    run<0>("8 DFMA", d, 8);
    run<1>("2 DMMA m8n8k4", d, 16);          // 2 x 256 FMA per warp = 16 per thread
    run<2>("4 DFMA + 1 DMMA", d, 12);
    run<3>("8 DFMA + 16 IMAD", d, 8);
    run<4>("2 DMMA + 16 IMAD", d, 16);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
