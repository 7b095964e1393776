// micro-benchmark: dependent-issue latencies of the sweep's critical chain on one warp
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, int n, double c)
{
    double a = out[threadIdx.x];
    long long t0, t1;
    // DADD chain
    t0 = clock64();
    for (int i = 0; i < n; i++) a = a - c;
    t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    // DMUL+DADD chain
    t0 = clock64();
    for (int i = 0; i < n; i++) a = a - c * a;
    t1 = clock64();
    if (threadIdx.x == 0) cyc[1] = t1 - t0;
    // SHFL(64-bit) + DMUL + DADD chain
    t0 = clock64();
    for (int i = 0; i < n; i++) { double v = __shfl_sync(0xffffffffu, a, (threadIdx.x + 31) & 31); a = a - c * v; }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[2] = t1 - t0;
    // SHFL + DMUL + 3x DADD (the W = 3 step)
    t0 = clock64();
    for (int i = 0; i < n; i++) { double v = __shfl_sync(0xffffffffu, a, (threadIdx.x + 31) & 31); double p = c * v; a = ((a - p) - p) - p; }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[3] = t1 - t0;
    // + st.cg each step
    t0 = clock64();
    for (int i = 0; i < n; i++) { double v = __shfl_sync(0xffffffffu, a, (threadIdx.x + 31) & 31); double p = c * v; a = ((a - p) - p) - p;
        asm volatile("st.global.cg.f64 [%0], %1;" ::"l"(out + 32 * (i & 1023) + threadIdx.x), "d"(a)); }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[4] = t1 - t0;
    out[threadIdx.x] = a;
}
int main()
{
    double* d; long long* c; cudaMalloc(&d, 32 * 1024 * 8 + 256); cudaMalloc(&c, 64); cudaMemset(d, 0, 32 * 1024 * 8);
    const int n = 4096;
    for (int r = 0; r < 2; r++) k<<<1, 32>>>(d, c, n, 1e-9);
    long long h[5]; cudaMemcpy(h, c, 40, cudaMemcpyDeviceToHost);
    const char* names[] = {"DADD", "DMUL+DADD", "SHFL64+DMUL+DADD", "SHFL64+DMUL+3xDADD", "SHFL64+DMUL+3xDADD+ST.CG"};
    for (int i = 0; i < 5; i++) printf("%-28s %.1f cycles/iter\n", names[i], (double)h[i] / n);
    return 0;
}
