// Floor of "one kernel launch + one synchronise" on this box, for reading the dff_ latency:
// an empty kernel, launched and waited for, 20000 times.
#include <cstdio>
#include <time.h>
__global__ void empty_kernel(int *p) { if (p) *p = 1; }
static double now_us() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3; }
int main() {
    cudaStream_t st; cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    int *flag; cudaHostAlloc(&flag, 64, cudaHostAllocMapped);
    for (int i = 0; i < 200; ++i) { empty_kernel<<<1, 32, 0, st>>>(nullptr); cudaStreamSynchronize(st); }
    const int n = 20000;
    double t0 = now_us();
    for (int i = 0; i < n; ++i) { empty_kernel<<<1, 640, 0, st>>>(nullptr); cudaStreamSynchronize(st); }
    double a = (now_us() - t0) / n;
    t0 = now_us();
    for (int i = 0; i < n; ++i) {          // completion through a flag in mapped memory, host spins
        *(volatile int *)flag = 0;
        empty_kernel<<<1, 640, 0, st>>>(flag);
        while (*(volatile int *)flag == 0) {}
    }
    double b = (now_us() - t0) / n;
    cudaStreamSynchronize(st);
    printf("{\"empty_kernel_launch_plus_stream_sync_us\": %.2f, \"empty_kernel_launch_plus_flag_spin_us\": %.2f}\n", a, b);
    return 0;
}
