// How long does nanosleep really sleep?  (B200, one warp, and 32 warps per SM)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void probe(unsigned ns, long long *out, int reps)
{
    long long t0 = clock64();
    for (int i = 0; i < reps; i++) __nanosleep(ns);
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = (t1 - t0) / reps;
}
int main()
{
    long long *d; cudaMalloc(&d, 8 * 1024);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    for (unsigned ns : {100u, 200u, 500u, 1000u, 2000u, 4000u, 8000u, 16000u, 64000u, 1000000u}) {
        for (int blocks : {1, 148 * 4}) {
            probe<<<blocks, 256>>>(ns, d, 200);
            cudaDeviceSynchronize();
            long long h[1024]; cudaMemcpy(h, d, 8 * blocks, cudaMemcpyDeviceToHost);
            double avg = 0; for (int i = 0; i < blocks; i++) avg += h[i]; avg /= blocks;
            printf("nanosleep(%u) blocks %d: %.0f cycles = %.0f ns\n", ns, blocks, avg, avg * 1e6 / clk);
        }
    }
    return 0;
}
