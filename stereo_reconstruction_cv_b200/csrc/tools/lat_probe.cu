// lat_probe.cu -- standalone latency probes (one warp, clock64) for design decisions of the sweep
// kernels: sub-warp min reduction by shuffles vs redux.sync with per-group member masks.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat_probe lat_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int LPC>
__device__ __forceinline__ uint32_t gmin_shfl(uint32_t t)
{
#pragma unroll
    for (int off = LPC / 2; off >= 1; off >>= 1) t = min(t, __shfl_xor_sync(0xFFFFFFFFu, t, off, LPC));
    return t;
}
template <int LPC>
__device__ __forceinline__ uint32_t gmin_redux(uint32_t t, uint32_t gmask)
{
    return __reduce_min_sync(gmask, t);
}

template <int LPC, int MODE>
__global__ void k_probe(uint32_t *out, long long *cyc, uint32_t seed)
{
    const int lane = threadIdx.x & 31;
    const uint32_t gmask = (LPC == 32 ? 0xFFFFFFFFu : ((1u << LPC) - 1u)) << ((lane / LPC) * LPC);
    uint32_t v = (lane * 2654435761u + seed) >> 4;
    uint32_t acc = 0;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 1024; it++) {
        uint32_t r = MODE == 0 ? gmin_shfl<LPC>(v) : gmin_redux<LPC>(v, gmask);
        acc += r;
        v = v * 1664525u + r + 1013904223u;      // dependent chain
    }
    long long t1 = clock64();
    out[threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

template <int LPC>
static void run()
{
    uint32_t *o0, *o1; long long *c;
    cudaMalloc(&o0, 128); cudaMalloc(&o1, 128); cudaMalloc(&c, 16);
    uint32_t h0[32], h1[32]; long long c0, c1;
    k_probe<LPC, 0><<<1, 32>>>(o0, c, 12345u); cudaMemcpy(&c0, c, 8, cudaMemcpyDeviceToHost);
    k_probe<LPC, 1><<<1, 32>>>(o1, c, 12345u); cudaMemcpy(&c1, c, 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(h0, o0, 128, cudaMemcpyDeviceToHost); cudaMemcpy(h1, o1, 128, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int i = 0; i < 32; i++) bad += h0[i] != h1[i];
    printf("LPC=%2d  shfl chain %.1f cyc/iter   redux %.1f cyc/iter   mismatches %d  (%s)\n", LPC, c0 / 1024.0, c1 / 1024.0, bad,
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(o0); cudaFree(o1); cudaFree(c);
}

int main()
{
    run<2>(); run<4>(); run<8>(); run<16>(); run<32>();
    return 0;
}
