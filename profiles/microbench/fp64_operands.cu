// FP64 throughput vs number of distinct register operands (register-file bandwidth).
// 8 independent chains per thread, 8 warps per sub-partition: the pipe is saturated; what varies
// is how many distinct 64-bit registers each instruction reads.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(1024) k(double *out, const double *in, long long *cyc, int iters) {
    double a[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) a[c] = in[c] + threadIdx.x;
    double x = in[8], y = in[9], z = in[10], w = in[11];   // runtime values: stay in registers
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            if (MODE == 0) a[c] = fma(a[c], 0.999999, 1e-9);          // 1 register operand (+2 immediates)
            if (MODE == 1) a[c] = fma(a[c], x, 1e-9);                 // 2 register operands
            if (MODE == 2) a[c] = fma(a[c], x, y);                    // 3 register operands, 2 shared by all
            if (MODE == 3) a[c] = fma(a[c], (c & 1) ? x : z, (c & 2) ? y : w);   // 3 operands, some variety
            if (MODE == 4) a[c] = fma(a[(c + 1) & 7], a[(c + 2) & 7], a[c]);      // 3 operands, all distinct & changing
            if (MODE == 5) a[c] = a[c] * x;                           // DMUL 2 regs
            if (MODE == 6) a[c] = a[c] + x;                           // DADD 2 regs
            if (MODE == 7) a[c] = a[(c + 1) & 7] * a[(c + 2) & 7];   // DMUL 2 distinct changing regs
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) s += a[c];
    if (s == 12345.678) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int MODE> void run(const char *name, double *d, double *in, long long *dc) {
    const int iters = 2048;
    k<MODE><<<1, 1024>>>(d, in, dc, iters);
    k<MODE><<<1, 1024>>>(d, in, dc, iters);
    long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
    printf("%-58s cycles per FP64 instr per sub-partition = %.2f\n", name, (double)c / iters / 8 / 8);
}
int main() {
    double *d, *in; long long *dc; cudaMalloc(&d, 8); cudaMalloc(&in, 128); cudaMalloc(&dc, 8);
    double h[16] = {1, 2, 3, 4, 5, 6, 7, 8, 0.999999, 1e-9, 0.999998, 2e-9};
    cudaMemcpy(in, h, 128, cudaMemcpyHostToDevice);
    run<0>("DFMA a = a*imm + imm (1 register operand)", d, in, dc);
    run<1>("DFMA a = a*x + imm (2 register operands)", d, in, dc);
    run<2>("DFMA a = a*x + y (3 register operands, x y shared)", d, in, dc);
    run<3>("DFMA a = a*(x|z) + (y|w) (3 register operands)", d, in, dc);
    run<4>("DFMA a = b*c + a (3 distinct changing registers)", d, in, dc);
    run<5>("DMUL a = a*x", d, in, dc);
    run<6>("DADD a = a+x", d, in, dc);
    run<7>("DMUL a = b*c (2 distinct changing registers)", d, in, dc);
    return 0;
}
