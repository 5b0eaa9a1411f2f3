// sgbm_microbench.cu -- issue-rate microbenchmark of the packed 16-bit integer instructions the
// path kernels are built from (VIMNMX3.U16x2, VIADDMNMX.U16x2, IADD3, PRMT, SHFL).  The result is
// the denominator of the ALU roofline in bench.py (MEASURED_PEAKS.json has no integer peak).
#include "sgbm_common.cuh"

#define MB_CHAINS 8
#define MB_ITERS 4096

template <int WHICH>
__global__ void __launch_bounds__(256) k_microbench(uint32_t *out, uint32_t a, uint32_t b)
{
    uint32_t x[MB_CHAINS];
#pragma unroll
    for (int i = 0; i < MB_CHAINS; i++) x[i] = threadIdx.x * 0x10003u + i * 0x70005u + blockIdx.x;
    for (int it = 0; it < MB_ITERS; it++) {
#pragma unroll
        for (int i = 0; i < MB_CHAINS; i++) {
            if (WHICH == 0) x[i] = __vimin3_u16x2(x[i], a, b) + 0u;
            else if (WHICH == 1) x[i] = __viaddmin_u16x2(x[i], a, b);
            else if (WHICH == 2) x[i] = x[i] + a - b;
            else if (WHICH == 3) x[i] = __byte_perm(x[i], a, 0x5432);
            else if (WHICH == 4) x[i] = __shfl_xor_sync(0xFFFFFFFFu, x[i], 1);
            else if (WHICH == 5) x[i] = __vminu2(x[i], a);
            else if (WHICH == 7) x[i] = __funnelshift_r(x[i], a, 16);
            else if (WHICH == 8) x[i] = (x[i] >> 16) | (a << 16);
            else if (WHICH == 9) x[i] = __vmaxu2(x[i], a) ^ b;
            else {   // 6: the path-step mix per packed register: PRMT, VIMNMX3, VIADDMNMX, IADD3, VIADDMNMX(S), VIMNMX
                uint32_t s = __byte_perm(x[i], a, 0x5432);
                uint32_t m3 = __vimin3_u16x2(s, x[(i + 1) % MB_CHAINS], b);
                uint32_t bb = __viaddmin_u16x2(m3, a, x[i]);
                uint32_t ln = bb + a - b;
                x[i] = __vminu2(__viaddmin_u16x2(x[i], ln, 0x7FFF7FFFu), s);
            }
        }
        a += 0x10001u;   // keep the compiler from hoisting
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < MB_CHAINS; i++) r ^= x[i];
    if (r == 0x12345678u) out[0] = r;
}

template <int WHICH>
static int run_one(double opsPerIter, double *out)
{
    int dev = 0, sms = 0;
    SGBM_CUDA_CHECK(cudaGetDevice(&dev));
    SGBM_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    uint32_t *d = nullptr;
    SGBM_CUDA_CHECK(cudaMalloc(&d, 64));
    cudaEvent_t e0, e1;
    SGBM_CUDA_CHECK(cudaEventCreate(&e0));
    SGBM_CUDA_CHECK(cudaEventCreate(&e1));
    const int grid = sms * 8;
    k_microbench<WHICH><<<grid, 256>>>(d, 3u, 5u);
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        SGBM_CUDA_CHECK(cudaEventRecord(e0));
        k_microbench<WHICH><<<grid, 256>>>(d, 3u, 5u);
        SGBM_CUDA_CHECK(cudaEventRecord(e1));
        SGBM_CUDA_CHECK(cudaEventSynchronize(e1));
        float ms = 0;
        SGBM_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    SGBM_CUDA_CHECK(cudaGetLastError());
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    double laneOps = (double)grid * 256 * MB_ITERS * MB_CHAINS * opsPerIter;
    *out = laneOps / (best * 1e-3) / 1e9;
    return 0;
}

int sgbm_run_microbench(int which, double *out)
{
    switch (which) {
    case 0: return run_one<0>(1, out);
    case 1: return run_one<1>(1, out);
    case 2: return run_one<2>(1, out);
    case 3: return run_one<3>(1, out);
    case 4: return run_one<4>(1, out);
    case 5: return run_one<5>(1, out);
    case 6: return run_one<6>(6, out);
    case 7: return run_one<7>(1, out);
    case 8: return run_one<8>(1, out);
    case 9: return run_one<9>(2, out);
    }
    return sgbm_fail(-1, "unknown microbenchmark %d", which);
}
