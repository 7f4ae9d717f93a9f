// FP64 dependent-issue latency and throughput vs (chains per thread, warps per SM sub-partition).
#include <cstdio>
#include <cuda_runtime.h>
template <int C>
__global__ void k(double *out, long long *cyc, int iters, double seed) {
    double a[C];
#pragma unroll
    for (int c = 0; c < C; ++c) a[c] = seed + c + threadIdx.x;
    const double m = 0.999999, b = 1e-9;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < C; ++c) a[c] = fma(a[c], m, b);
    }
    long long t1 = clock64();
    double s = 0; 
#pragma unroll
    for (int c = 0; c < C; ++c) s += a[c];
    if (s == 12345.678) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int C> void run(int warps_per_smsp) {
    double *d; long long *dc; cudaMalloc(&d, 8); cudaMalloc(&dc, 8);
    const int iters = 4096; int threads = 32 * 4 * warps_per_smsp;   // one CTA on one SM
    k<C><<<1, threads>>>(d, dc, iters, 1.0);
    k<C><<<1, threads>>>(d, dc, iters, 2.0);
    long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
    double per_iter = (double)c / iters;
    printf("chains=%d warps/smsp=%d cycles/iter=%.2f  cycles per DFMA per warp=%.2f  smsp DFMA/cycle=%.3f\n", C, warps_per_smsp,
           per_iter, per_iter / C, (double)C * warps_per_smsp / per_iter);
    cudaFree(d); cudaFree(dc);
}
int main() {
    run<1>(1); run<2>(1); run<4>(1); run<8>(1);
    run<1>(2); run<1>(4); run<1>(8); run<2>(4); run<2>(6); run<2>(8); run<4>(4); run<4>(8);
    return 0;
}
