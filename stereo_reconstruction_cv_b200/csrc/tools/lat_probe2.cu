// lat_probe2.cu -- single-warp latency of the building blocks of the sweep kernel (clock64):
// path_step (dependent row-to-row chain), mbarrier test_wait / try_wait / arrive, LDS.128 vector loads.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I.. -o lat_probe2 lat_probe2.cu
#include <cstdio>
#include "../sgbm_common.cuh"
int sgbm_fail_cuda(cudaError_t, const char *, const char *, int) { return -4; }
void sgbm_count_launch(int) {}
int sgbm_fail(int c, const char *, ...) { return c; }

template <int NREG, int LPC, int MODE>
__global__ void k_path(uint32_t *out, long long *cyc, uint32_t seed, int iters)
{
    __shared__ __align__(16) uint16_t sm[32 * 64 * 4];
    __shared__ uint64_t bar[4];
    const int lane = threadIdx.x & 31, lg = lane % LPC;
    for (int i = lane; i < 32 * 64 * 2; i += 32) reinterpret_cast<uint32_t *>(sm)[i] = (i * 2654435761u + seed) & 0x03FF03FFu;
    if (lane == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_fence_init(); }
    __syncwarp();
    if (lane == 0) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar[0])) : "memory"); }
    __syncwarp();
    uint32_t L[NREG], m = 0;
#pragma unroll
    for (int j = 0; j < NREG; j++) L[j] = 0;
    const uint32_t P1p = 200u * 0x10001u, P2mP1p = 600u * 0x10001u;
    uint32_t acc = 0;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
        if (MODE == 0 || MODE == 1) {                 // path step with C from shared memory (MODE 1: + in place)
            uint32_t Cc[NREG];
            load_vec<NREG, LPC>(Cc, sm + ((it & 3) * (32 / LPC) + (lane / LPC)) * (2 * NREG * LPC), lg);
            m = path_step<NREG, LPC>(L, L, m, Cc, P1p, P2mP1p, lg, LPC - 1);
        } else if (MODE == 2) {                       // test_wait on a completed phase, result consumed at once
            acc += mbar_test_wait(&bar[0], 0) ? 1u : 0u;
        } else if (MODE == 3) {                       // try_wait on a completed phase
            acc += mbar_try_wait(&bar[0], 0) ? 1u : 0u;
        } else if (MODE == 4) {                       // arrive (count 1: every arrive completes a phase)
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar[1])) : "memory");
        } else if (MODE == 5) {                       // local_min + group_min only (dependent through m)
            L[0] = m + it;
            uint32_t t = local_min<NREG>(L);
            m = group_min<LPC>(t);
        } else if (MODE == 6) {                       // group_min only
            m = group_min<LPC>(m + it);
        } else if (MODE == 7) {                       // LDS.128 x NREG/4 dependent
            uint32_t Cc[NREG];
            load_vec<NREG, LPC>(Cc, sm + ((m & 3) * (32 / LPC) + (lane / LPC)) * (2 * NREG * LPC), lg);
            m = Cc[0] & 3;
        } else if (MODE == 8) {                       // __syncwarp + lane0 arrive (hand-off pattern)
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar[1])) : "memory");
        }
    }
    long long t1 = clock64();
#pragma unroll
    for (int j = 0; j < NREG; j++) acc ^= L[j];
    out[threadIdx.x] = acc ^ m;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

template <int NREG, int LPC, int MODE>
static void run(const char *what)
{
    uint32_t *o; long long *c;
    cudaMalloc(&o, 128); cudaMalloc(&c, 16);
    long long h = 0;
    const int iters = 2048;
    k_path<NREG, LPC, MODE><<<1, 32>>>(o, c, 1u, iters);
    k_path<NREG, LPC, MODE><<<1, 32>>>(o, c, 1u, iters);
    cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("NREG=%2d LPC=%2d %-44s %7.1f cyc/iter (%s)\n", NREG, LPC, what, (double)h / iters, cudaGetErrorString(cudaGetLastError()));
    cudaFree(o); cudaFree(c);
}

int main()
{
    run<16, 8, 0>("LDS C + path_step (row-to-row chain)");
    run<8, 8, 0>("LDS C + path_step (row-to-row chain)");
    run<12, 8, 0>("LDS C + path_step (row-to-row chain)");
    run<4, 32, 0>("LDS C + path_step (row-to-row chain)");
    run<8, 16, 0>("LDS C + path_step (row-to-row chain)");
    run<16, 8, 5>("local_min + group_min");
    run<8, 8, 5>("local_min + group_min");
    run<8, 8, 6>("group_min");
    run<8, 32, 6>("group_min");
    run<8, 8, 7>("dependent vector load (LDS.128 x NREG/4)");
    run<16, 8, 7>("dependent vector load (LDS.128 x NREG/4)");
    run<8, 8, 2>("mbarrier.test_wait (complete), consumed");
    run<8, 8, 3>("mbarrier.try_wait (complete), consumed");
    run<8, 8, 4>("mbarrier.arrive");
    run<8, 8, 8>("__syncwarp + lane-0 mbarrier.arrive");
    return 0;
}
