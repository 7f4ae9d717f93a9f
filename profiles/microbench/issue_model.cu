// Does an FP64 warp instruction occupy the issue port for one cycle or two on sm_100a?
// Kernel k<R> runs 8 independent DFMA chains per thread with R independent integer ops
// interleaved per DFMA.  If non-FP64 instructions issue in the FP64 pipe's shadow, time(R=1)
// == time(R=0); if an FP64 instruction holds the port for two cycles, time grows with R.
#include <cstdio>
#include <cuda_runtime.h>

template <int R>
__global__ void __launch_bounds__(256) k(double *out, int *iout, int iters, double seed) {
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    int i0 = threadIdx.x, i1 = i0 + 1, i2 = i0 + 2, i3 = i0 + 3, i4 = i0 + 4, i5 = i0 + 5, i6 = i0 + 6, i7 = i0 + 7;
    const double m = 0.999999, b = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, b); if (R > 0) i0 = (i0 ^ i) + 0x9e3779b9; if (R > 1) i0 = (i0 >> 3) ^ i1; if (R > 2) i0 = i0 * 3 + i2;
        a1 = fma(a1, m, b); if (R > 0) i1 = (i1 ^ i) + 0x7f4a7c15; if (R > 1) i1 = (i1 >> 5) ^ i2; if (R > 2) i1 = i1 * 5 + i3;
        a2 = fma(a2, m, b); if (R > 0) i2 = (i2 ^ i) + 0x94d049bb; if (R > 1) i2 = (i2 >> 7) ^ i3; if (R > 2) i2 = i2 * 7 + i4;
        a3 = fma(a3, m, b); if (R > 0) i3 = (i3 ^ i) + 0xbf58476d; if (R > 1) i3 = (i3 >> 9) ^ i4; if (R > 2) i3 = i3 * 9 + i5;
        a4 = fma(a4, m, b); if (R > 0) i4 = (i4 ^ i) + 0x1ce4e5b9; if (R > 1) i4 = (i4 >> 11) ^ i5; if (R > 2) i4 = i4 * 11 + i6;
        a5 = fma(a5, m, b); if (R > 0) i5 = (i5 ^ i) + 0x133111eb; if (R > 1) i5 = (i5 >> 13) ^ i6; if (R > 2) i5 = i5 * 13 + i7;
        a6 = fma(a6, m, b); if (R > 0) i6 = (i6 ^ i) + 0x632be59b; if (R > 1) i6 = (i6 >> 15) ^ i7; if (R > 2) i6 = i6 * 15 + i0;
        a7 = fma(a7, m, b); if (R > 0) i7 = (i7 ^ i) + 0xd9b4e019; if (R > 1) i7 = (i7 >> 17) ^ i0; if (R > 2) i7 = i7 * 17 + i1;
    }
    const double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    const int t = i0 ^ i1 ^ i2 ^ i3 ^ i4 ^ i5 ^ i6 ^ i7;
    if (s == 12345.678) out[0] = s;
    if (t == 0x12345678) iout[0] = t;
}

template <int R> float run(double *d, int *di, int iters, int grid) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<R><<<grid, 256>>>(d, di, iters, 1.0);
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0); k<R><<<grid, 256>>>(d, di, iters, 2.0 + r); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    double *d; int *di; cudaMalloc(&d, 8); cudaMalloc(&di, 4);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int iters = 1 << 15, grid = sms * 8;
    float t0 = run<0>(d, di, iters, grid), t1 = run<1>(d, di, iters, grid), t2 = run<2>(d, di, iters, grid), t3 = run<3>(d, di, iters, grid);
    double dfma = 8.0 * iters * grid * 256;
    printf("{\"dfma_tflops_R0\": %.2f, \"ms\": {\"R0\": %.3f, \"R1\": %.3f, \"R2\": %.3f, \"R3\": %.3f}, "
           "\"ratio_vs_R0\": {\"R1\": %.3f, \"R2\": %.3f, \"R3\": %.3f}}\n",
           2 * dfma / (t0 * 1e-3) / 1e12, t0, t1, t2, t3, t1 / t0, t2 / t0, t3 / t0);
    return 0;
}
