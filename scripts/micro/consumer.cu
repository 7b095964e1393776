// micro-benchmark: the sweep consumer's block loop in isolation (one warp, operands in shared memory)
// STORE: 0 none, 1 st.global.cg, 2 plain st.global, 3 st.relaxed.gpu, 4 STS + TMA bulk store per block, 5 STS only
#include <cstdio>
#include <cuda_runtime.h>
template <int SKEW, int STORE, bool POLL>
__global__ void k(double* out, long long* cyc, int nBlocks, int slot)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* sm = reinterpret_cast<double*>(smem);
    for (int i = threadIdx.x; i < 4 * 8 * 3 * 32; i += blockDim.x) sm[i] = 1e-9 * (i % 7);
    volatile unsigned* cnt = reinterpret_cast<volatile unsigned*>(smem + 4 * 8 * 3 * 256);
    double* obuf = reinterpret_cast<double*>(smem + 4 * 8 * 3 * 256 + 128); // 4 x 2 KB
    if (threadIdx.x < 4) cnt[threadIdx.x] = 8;
    __syncthreads();
    if (warp != 0) return;
    double h[SKEW];
    for (int k2 = 0; k2 < SKEW; k2++) h[k2] = 1.0 + lane;
    double* outPtr = out + lane;
    const int src = (lane + 31) & 31;
    long long t0 = clock64();
    int st = 0;
    for (int blk = 0; blk < nBlocks; blk++)
    {
        if (POLL)
        {
            unsigned c;
            do c = cnt[st];
            while ((c & 0xff) != 8);
        }
        const double* base = sm + st * 8 * 3 * 32 + lane;
        double a0[8], c0[8], c1[8];
#pragma unroll
        for (int q = 0; q < 8; q++)
        {
            a0[q] = base[q * 96];
            c0[q] = base[q * 96 + 32];
            c1[q] = base[q * 96 + 64];
        }
        if (STORE == 4 && blk >= 4) asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
#pragma unroll
        for (int q = 0; q < 8; q++)
        {
            const double sh = __shfl_sync(0xffffffffu, h[SKEW - 1], src);
            const double pre = a0[q] - c0[q] * sh;
            const double acc = pre - c1[q] * h[0];
            if (STORE == 1) asm volatile("st.global.cg.f64 [%0], %1;" ::"l"(outPtr + q * 32), "d"(acc));
            if (STORE == 2) outPtr[q * 32] = acc;
            if (STORE == 3) asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(outPtr + q * 32), "d"(acc));
            if (STORE >= 4) obuf[st * 256 + q * 32 + lane] = acc;
#pragma unroll
            for (int k2 = SKEW - 1; k2 > 0; k2--) h[k2] = h[k2 - 1];
            h[0] = acc;
        }
        if (STORE == 4)
        {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0)
            {
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 2048;" ::"l"(outPtr - lane),
                             "r"((unsigned)__cvta_generic_to_shared(obuf + st * 256))
                             : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
        outPtr += 8 * 32;
        if (POLL && lane == 0) cnt[st] = 8;
        if (++st == 4) st = 0;
    }
    if (STORE == 4) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    long long t1 = clock64();
    if (lane == 0) cyc[slot] = t1 - t0;
    out[lane] = h[0];
}
int main()
{
    const int nBlocks = 512;
    double* d; long long* c;
    cudaMalloc(&d, (size_t)nBlocks * 8 * 32 * 8 + 1024); cudaMalloc(&c, 128);
    const int smem = 4 * 8 * 3 * 256 + 128 + 4 * 2048;
    for (int r = 0; r < 2; r++)
    {
        k<2, 0, true><<<1, 32, smem>>>(d, c, nBlocks, 0);
        k<2, 1, true><<<1, 32, smem>>>(d, c, nBlocks, 1);
        k<2, 2, true><<<1, 32, smem>>>(d, c, nBlocks, 2);
        k<2, 3, true><<<1, 32, smem>>>(d, c, nBlocks, 3);
        k<2, 4, true><<<1, 32, smem>>>(d, c, nBlocks, 4);
        k<2, 5, true><<<1, 32, smem>>>(d, c, nBlocks, 5);
        k<3, 5, true><<<1, 32, smem>>>(d, c, nBlocks, 6);
        k<3, 4, true><<<1, 32, smem>>>(d, c, nBlocks, 7);
    }
    long long h[8]; cudaMemcpy(h, c, 64, cudaMemcpyDeviceToHost);
    const char* names[] = {"skew2 nostore", "skew2 st.cg", "skew2 st plain", "skew2 st.relaxed.gpu", "skew2 STS+TMA store", "skew2 STS only", "skew3 STS only", "skew3 STS+TMA store"};
    for (int i = 0; i < 8; i++) printf("%-24s %.1f cycles/step\n", names[i], (double)h[i] / (nBlocks * 8));
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
