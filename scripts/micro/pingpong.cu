// micro-benchmark: latency of passing a value between two SMs through global memory (st + polling ld),
// the mechanism of the sweeps' cross-group dependencies.  Block 0 and block `peer` ping-pong n times.
#include <cstdio>
#include <cuda_runtime.h>
template <int LD>
__device__ __forceinline__ double ld(const double* p)
{
    double v;
    if (LD == 0) asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p));
    if (LD == 1) asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p));
    if (LD == 2) asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
template <int LD>
__global__ void k(double* buf, long long* cyc, int n, int peer, int slot)
{
    if (blockIdx.x != 0 && blockIdx.x != peer) return;
    const int me = blockIdx.x == 0 ? 0 : 1;
    double* mine = buf + me * 64 + threadIdx.x;      // I write here
    double* theirs = buf + (1 - me) * 64 + threadIdx.x; // I poll here
    long long t0 = clock64();
    for (int i = 1; i <= n; i++)
    {
        if (me == 0)
        {
            asm volatile("st.global.cg.f64 [%0], %1;" ::"l"(mine), "d"((double)i));
            while (ld<LD>(theirs) != (double)i) {}
        }
        else
        {
            while (ld<LD>(theirs) != (double)i) {}
            asm volatile("st.global.cg.f64 [%0], %1;" ::"l"(mine), "d"((double)i));
        }
    }
    long long t1 = clock64();
    if (me == 0 && threadIdx.x == 0) cyc[slot] = t1 - t0;
}
int main()
{
    double* d; long long* c;
    cudaMalloc(&d, 4096); cudaMalloc(&c, 256);
    const int n = 2000;
    int peers[] = {1, 2, 8, 37, 74, 100, 147};
    long long h[32];
    for (int pi = 0; pi < 7; pi++)
    {
        cudaMemset(d, 0, 4096);
        k<0><<<148, 32>>>(d, c, n, peers[pi], pi);
        cudaDeviceSynchronize();
    }
    cudaMemset(d, 0, 4096); k<1><<<148, 32>>>(d, c, n, 74, 7); cudaDeviceSynchronize();
    cudaMemset(d, 0, 4096); k<2><<<148, 32>>>(d, c, n, 74, 8); cudaDeviceSynchronize();
    cudaMemcpy(h, c, 72, cudaMemcpyDeviceToHost);
    for (int pi = 0; pi < 7; pi++) printf("ld.cg  block 0 <-> block %3d : one-way %.0f cycles\n", peers[pi], (double)h[pi] / n / 2);
    printf("ld.relaxed.gpu  0 <-> 74 : one-way %.0f cycles\n", (double)h[7] / n / 2);
    printf("ld.volatile     0 <-> 74 : one-way %.0f cycles\n", (double)h[8] / n / 2);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
