// sgbm_common.cuh -- shared definitions of the sm_100a dense-stereo kernels.
//
// Arithmetic contract: SURVEY.md Appendix A (the validated restatement of what
// cv2.StereoSGBM.compute does at the reference call site main.ipynb:655-668).
//
// Data layout in HBM (all internal volumes are uint16, values are non-negative int16 costs):
//   volume[y][x1][pos]   x1 = x - minX1 in [0,W1),  pos in [0,Dp)
// A column's disparity vector is owned by a group of LPC lanes of one warp, each lane holding
// 2*NREG consecutive disparities as NREG packed u16x2 registers (lane l: d = l*2*NREG ...).
// In memory the vector is stored chunk-interleaved so that the 16-byte chunk k of lane l sits at
// byte offset 16*(LPC*k + l): every 128-bit load/store of a group is one contiguous segment.
//   pos(d) = 8*(LPC*k + l) + e   with l = d / (2*NREG), k = (d % (2*NREG)) / 8, e = d % 8
// Dp = 2*NREG*LPC >= D; lanes >= lanesUsed = D/(2*NREG) are whole padding lanes.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <mutex>

#define SGBM_MAX_S 0x7FFF7FFFu      // packed saturation value 32767 (A.4)
#define SGBM_INF2  0xFFFFFFFFu      // packed +inf for out-of-range disparity neighbours

struct Geo {
    int W, H, cn;
    int minD, D, maxD, r, P1, P2, UR, DMD, ftzero, INV, minX1, maxX1, W1, mode;
    int nreg, lpc, Dp, lanesUsed;
    int lpcShift;                    // log2(lpc)
    long long rowStride;             // W1 * Dp  (u16 elements)
};

__host__ __device__ __forceinline__ int sgbm_pos(int d, int nreg, int lpc)
{
    int l = d / (2 * nreg), i = d % (2 * nreg);
    return 8 * (lpc * (i >> 3) + l) + (i & 7);
}

// ---- packed u16x2 helpers (map 1:1 onto VIMNMX / VIMNMX3 / VIADDMNMX .U16x2 on sm_100a) -------
__device__ __forceinline__ uint32_t pmin(uint32_t a, uint32_t b) { return __vminu2(a, b); }
__device__ __forceinline__ uint32_t pmin3(uint32_t a, uint32_t b, uint32_t c) { return __vimin3_u16x2(a, b, c); }
__device__ __forceinline__ uint32_t paddmin(uint32_t a, uint32_t b, uint32_t c) { return __viaddmin_u16x2(a, b, c); }
__device__ __forceinline__ uint32_t pswap(uint32_t a) { return __byte_perm(a, 0, 0x1032); }
__device__ __forceinline__ uint32_t pbcast(uint32_t v) { return (v & 0xFFFFu) * 0x10001u; }

// Group-wide (LPC lanes, aligned) min of a packed value whose halves are already equal.
template <int LPC>
__device__ __forceinline__ uint32_t group_min(uint32_t t)
{
#pragma unroll
    for (int off = LPC / 2; off >= 1; off >>= 1) t = pmin(t, __shfl_xor_sync(0xFFFFFFFFu, t, off, LPC));
    return t;
}

// Local min over NREG packed registers, both halves combined and broadcast to both halves.
template <int NREG>
__device__ __forceinline__ uint32_t local_min(const uint32_t (&v)[NREG])
{
    uint32_t t = v[0];
#pragma unroll
    for (int j = 1; j + 1 < NREG; j += 2) t = pmin3(t, v[j], v[j + 1]);
    if ((NREG & 1) == 0) t = pmin(t, v[NREG - 1]);
    return pmin(t, pswap(t));
}

// Vector load/store of one lane's NREG registers (NREG % 4 == 0) from a column base pointer.
template <int NREG, int LPC>
__device__ __forceinline__ void load_vec(uint32_t (&v)[NREG], const uint16_t *col, int lg)
{
    const uint4 *p = reinterpret_cast<const uint4 *>(col) + lg;
#pragma unroll
    for (int k = 0; k < NREG / 4; k++) {
        uint4 q = p[LPC * k];
        v[4 * k + 0] = q.x; v[4 * k + 1] = q.y; v[4 * k + 2] = q.z; v[4 * k + 3] = q.w;
    }
}
template <int NREG, int LPC>
__device__ __forceinline__ void load_vec_nc(uint32_t (&v)[NREG], const uint16_t *col, int lg)
{
    const uint4 *p = reinterpret_cast<const uint4 *>(col) + lg;
#pragma unroll
    for (int k = 0; k < NREG / 4; k++) {
        uint4 q = __ldg(p + LPC * k);
        v[4 * k + 0] = q.x; v[4 * k + 1] = q.y; v[4 * k + 2] = q.z; v[4 * k + 3] = q.w;
    }
}
template <int NREG, int LPC>
__device__ __forceinline__ void store_vec(const uint32_t (&v)[NREG], uint16_t *col, int lg)
{
    uint4 *p = reinterpret_cast<uint4 *>(col) + lg;
#pragma unroll
    for (int k = 0; k < NREG / 4; k++)
        p[LPC * k] = make_uint4(v[4 * k + 0], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
}


// ---- host-side knobs and per-device set-up --------------------------------------------------------
// Tuning / test knobs.  They are read from the environment ONCE, in sgbm_create, and live in the handle;
// the launchers see the knobs of the handle whose call is running on this thread (sgbm_knobs()).  Nothing
// reads the environment per frame.  The two hooks that make results INVALID on purpose (dbgNoSync,
// dbgStall) exist only in builds with -DSGBM_DEBUG_HOOKS (libsgbm_b200_dbg.so, used by one test).
struct SgbmKnobs {
    int nreg = 0;                        // SGBM_NREG: force the registers-per-lane of the lane mapping (0 = auto)
    int vr = 0;                          // SGBM_VR: rows per super-step of the sweeps (0 = default)
    int sweepK = 0, sweepNSC = 0, sweepNSI = 0, sweepNWW = 0;   // SGBM_SWEEP_K / _NSC / _NSI / _NWW ring depths, WTA warps per row
    int sweepPF = -1;                    // SGBM_SWEEP_PF: rows of L2 prefetch ahead of the spilling sweep's bulk copies (-1 = default 8, 0 = off)
    int sweepRPS = 0;                    // SGBM_SWEEP_RPS: rows per ring stage of the sweeps (0 = auto, 1, 2)
    int sweepWRG = 0;                    // SGBM_SWEEP_WRG: cap on the row groups of the WTA warps (0 = as many as fit)
    int sweepW = 1;                      // SGBM_SWEEP_W=0: winner-take-all on role C instead of the W role
    int sweep = 1;                       // SGBM_SWEEP=0: lock-step k_vertical instead of the role-specialised sweep
    int rowstep = 0;                     // SGBM_ROWSTEP=1: row-at-a-time fallback
    int cost2 = 1, cost3 = 1;            // SGBM_COST2=0 / SGBM_COST3=0: older cost-kernel generations
    int cost3NXG = 0, cost3RB = 0;       // SGBM_COST3_NXG / _RB
    int cost3Pad = 0;                    // SGBM_COST3_PAD: extra dynamic shared memory (bytes) requested by the cost kernel (occupancy experiments)
    int nstg = 0;                        // SGBM_NSTG: staging depth of k_vertical
    int sweepSat = 0;                    // SGBM_SWEEP_SAT=1: force the saturating S accumulation
    int hhSplit = 1;                     // SGBM_HH_SPLIT=0: MODE_HH feeds L_hB into the forward sweep instead of the backward one
    int verbose = 0;                     // SGBM_VERBOSE: print launch geometries to stderr
    int dbgNoSync = 0, dbgStall = 0;     // SGBM_DBG_NOSYNC / SGBM_DBG_STALL (debug-hook builds only)
    char tracePath[256] = "";            // SGBM_SWEEP_TRACE (tracing builds only)
    // facts about the device the handle lives on
    int device = 0, numSMs = 0, maxSmemOptin = 0;
};
const SgbmKnobs &sgbm_knobs();           // sgbm_api.cu: knobs of the call running on this thread

#ifdef SGBM_DEBUG_HOOKS
#define SGBM_DBG_HOOK(x) (x)
#else
#define SGBM_DBG_HOOK(x) 0
#endif

// Per-device one-time set-up of a kernel (cudaFuncSetAttribute is per device; one process may hold handles
// on several devices and drive them from several threads).  Usage:
//   static unsigned long long done = 0;
//   { SgbmDeviceOnce once(done); if (once.first) { SGBM_CUDA_CHECK(cudaFuncSetAttribute(...)); once.done(); } }
// The lock is held until the set-up has finished, so no thread launches before the attribute is set.
std::mutex &sgbm_setup_mutex();          // sgbm_api.cu
struct SgbmDeviceOnce {
    std::unique_lock<std::mutex> lk;
    unsigned long long &mask;
    unsigned long long bit = 1;
    bool first = true;
    explicit SgbmDeviceOnce(unsigned long long &m) : lk(sgbm_setup_mutex()), mask(m)
    {
        int dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess) bit = 1ull << (dev & 63);
        first = !(mask & bit);
    }
    void done() { mask |= bit; }
};

// ---- shared memory by 32-bit address --------------------------------------------------------------
// The sweeps touch several shared-memory rings per row.  With generic pointers the compiler re-derives
// every address from the kernel arguments, the thread id and the shared window base at each use (the role
// warps run under a tight register cap and rematerialise instead of keeping the pointer): eight integer
// instructions in front of every group of LDS / STS.  A 32-bit shared address made opaque once (sm_keep)
// stays in one register, and [reg + immediate] addressing covers a lane's chunks.
__device__ __forceinline__ uint32_t sm_keep(uint32_t v)
{
    uint32_t r;
    asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v));
    return r;
}
template <int OFF>
__device__ __forceinline__ uint4 lds128(uint32_t addr)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4+%5];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr), "n"(OFF));
    return v;
}
template <int OFF>
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    asm volatile("st.shared.v4.u32 [%0+%1], {%2, %3, %4, %5};" ::"r"(addr), "n"(OFF), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint32_t lds16(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts16(uint32_t addr, uint32_t v)
{
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
// One lane's NREG registers of a column whose shared address (lane chunk included) is `addr`.
template <int NREG, int LPC, int K = 0>
__device__ __forceinline__ void lds_vec(uint32_t (&v)[NREG], uint32_t addr)
{
    if constexpr (K < NREG / 4) {
        const uint4 q = lds128<16 * LPC * K>(addr);
        v[4 * K + 0] = q.x; v[4 * K + 1] = q.y; v[4 * K + 2] = q.z; v[4 * K + 3] = q.w;
        lds_vec<NREG, LPC, K + 1>(v, addr);
    }
}
template <int NREG, int LPC, int K = 0>
__device__ __forceinline__ void sts_vec(const uint32_t (&v)[NREG], uint32_t addr)
{
    if constexpr (K < NREG / 4) {
        sts128<16 * LPC * K>(addr, v[4 * K + 0], v[4 * K + 1], v[4 * K + 2], v[4 * K + 3]);
        sts_vec<NREG, LPC, K + 1>(v, addr);
    }
}

// ---- TMA bulk copy (cp.async.bulk, SASS UBLKCP) + mbarrier helpers ------------------------------
// One elected lane arms an mbarrier with the byte count and issues global->shared bulk copies; the
// consumers wait on the barrier's phase parity.  Waits are bounded: a protocol error traps instead
// of hanging the GPU.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// Non-blocking probe: issued early, consumed later, so its latency overlaps independent work.
__device__ __forceinline__ bool mbar_test_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// The same operations on a barrier named by its 32-bit shared-memory address: kernels that touch many
// barriers per row (sgbm_sweep.cu) keep those instead of generic pointers -- no generic->shared conversion
// at every use, one register per barrier array instead of two.
struct SmemBar {
    uint32_t addr;
    __device__ __forceinline__ SmemBar operator[](int i) const { return SmemBar{addr + 8u * (uint32_t)i}; }
};
__device__ __forceinline__ void mbar_init(SmemBar bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar.addr), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(SmemBar bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar.addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, SmemBar bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(bar.addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(SmemBar bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar.addr), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_test_wait(SmemBar bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar.addr), "r"(parity) : "memory");
    return ok != 0;
}
// Blocking wait.  The retry passes a suspend-time hint so that a waiting warp sleeps in hardware
// instead of burning issue slots in a poll loop; ~80 s without progress traps (protocol error; the
// bound is generous because profilers that patch the kernel slow it down by orders of magnitude).
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    if (mbar_try_wait(bar, parity)) return;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .u32 n;\n\t"
        "mov.u32 n, 0;\n"
        "SGBM_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 0x989680;\n\t"
        "@p bra SGBM_DONE;\n\t"
        "add.u32 n, n, 1;\n\t"
        "setp.lt.u32 p, n, 8192;\n\t"
        "@p bra SGBM_WAIT;\n\t"
        "trap;\n"
        "SGBM_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// The same on a barrier named by its shared address.
__device__ __forceinline__ void mbar_wait(SmemBar bar, uint32_t parity)
{
    if (mbar_try_wait(bar, parity)) return;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .u32 n;\n\t"
        "mov.u32 n, 0;\n"
        "SGBM_WAITA:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 0x989680;\n\t"
        "@p bra SGBM_DONEA;\n\t"
        "add.u32 n, n, 1;\n\t"
        "setp.lt.u32 p, n, 8192;\n\t"
        "@p bra SGBM_WAITA;\n\t"
        "trap;\n"
        "SGBM_DONEA:\n\t"
        "}" ::"r"(bar.addr), "r"(parity) : "memory");
}

// One step of the SGM recurrence (A.4) for one column, distributed over a group of LPC lanes:
//   Ln(d) = C(d) + min(Lp(d), Lp(d-1)+P1, Lp(d+1)+P1, m+P2) - m,   m = min_k Lp(k)
// Lp / mp : predecessor's path costs and their (packed, broadcast) minimum.
// Returns the new packed-broadcast minimum of Ln.  lg = lane in group, lastLane = lanesUsed-1.
// Unsigned 16-bit arithmetic: every intermediate is <= m + P2 <= 65534 or <= C + P2.
template <int NREG, int LPC>
__device__ __forceinline__ uint32_t path_step(uint32_t (&Ln)[NREG], const uint32_t (&Lp)[NREG],
                                              uint32_t mp, const uint32_t (&C)[NREG], uint32_t P1p,
                                              uint32_t P2mP1p, int lg, int lastLane)
{
    uint32_t up = __shfl_up_sync(0xFFFFFFFFu, Lp[NREG - 1], 1, LPC);
    uint32_t dn = __shfl_down_sync(0xFFFFFFFFu, Lp[0], 1, LPC);
    if (lg == 0) up = SGBM_INF2;                 // L(-1) = +inf
    if (lg >= lastLane) dn = SGBM_INF2;          // L(D)  = +inf
    const uint32_t k1 = mp + P2mP1p;             // (m + P2 - P1) in both halves, no carry (<= 65535)
    uint32_t sPrev = __byte_perm(up, Lp[0], 0x5432);   // (L[2j-1], L[2j]) for j = 0
#pragma unroll
    for (int j = 0; j < NREG; j++) {
        uint32_t nxt = (j + 1 < NREG) ? Lp[j + 1] : dn;
        uint32_t sNext = __byte_perm(Lp[j], nxt, 0x5432);   // (L[2j+1], L[2j+2])
        uint32_t a = pmin3(sPrev, sNext, k1);               // min(L(d-1), L(d+1), m+P2-P1)
        uint32_t b = paddmin(a, P1p, Lp[j]);                // min(a + P1, L(d))
        Ln[j] = b + C[j] - mp;                              // halves stay in [0, 65535]: plain add
        sPrev = sNext;
    }
    uint32_t t = local_min<NREG>(Ln);
    if (lg > lastLane) t = SGBM_INF2;
    return group_min<LPC>(t);
}

// OR-masks of a lane inside its group: all ones where the disparity neighbour below / above the lane's
// range is outside [0, D) (L(-1) = L(D) = +inf, A.4) and where the whole lane is padding.  Held in registers
// (made opaque once) so that the row loop does not re-derive them from the thread id.
struct LaneMasks { uint32_t up, dn, pad; };
__device__ __forceinline__ LaneMasks lane_masks(int lg, int lastLane)
{
    LaneMasks m;
    m.up = sm_keep(lg == 0 ? 0xFFFFFFFFu : 0u);
    m.dn = sm_keep(lg >= lastLane ? 0xFFFFFFFFu : 0u);
    m.pad = sm_keep(lg > lastLane ? 0xFFFFFFFFu : 0u);
    return m;
}

// path_step above, in place, with the lane masks instead of per-step comparisons (the row loops of the sweeps and the
// column loop of the horizontal kernel).
template <int NREG, int LPC>
__device__ __forceinline__ uint32_t path_step_m(uint32_t (&L)[NREG], uint32_t mp, const uint32_t (&C)[NREG], uint32_t P1p,
                                                uint32_t P2mP1p, const LaneMasks &lm)
{
    const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, L[NREG - 1], 1, LPC) | lm.up;
    const uint32_t dn = __shfl_down_sync(0xFFFFFFFFu, L[0], 1, LPC) | lm.dn;
    uint32_t sPrev = __byte_perm(up, L[0], 0x5432);        // (L[2j-1], L[2j]) for j = 0
    const uint32_t k1 = mp + P2mP1p;                       // (m + P2 - P1) in both halves, no carry (<= 65535)
#pragma unroll
    for (int j = 0; j < NREG; j++) {
        const uint32_t nxt = (j + 1 < NREG) ? L[j + 1] : dn;
        const uint32_t sNext = __byte_perm(L[j], nxt, 0x5432);
        const uint32_t b = paddmin(pmin3(sPrev, sNext, k1), P1p, L[j]);
        L[j] = b + C[j] - mp;
        sPrev = sNext;
    }
    return group_min<LPC>(local_min<NREG>(L) | lm.pad);
}

// numDisparities that are not a multiple of 8 (cv2 accepts them): the volumes are as wide as the next multiple of 8, the
// cost volume carries a large constant at the padding disparities d in [D, Dc) (sgbm_api.cu: k_pad_cost), which makes them
// inert in the path step (never the minimum, never the better neighbour of d = D - 1), and the winner-take-all hides
// them: in the last used lane the registers from local index padFrom on are forced to all ones.
template <int NREG>
__device__ __forceinline__ void mask_pad_regs(uint32_t (&S)[NREG], bool lastUsedLane, int padFrom)
{
    if (lastUsedLane) {
#pragma unroll
        for (int j = 0; j < NREG; j++) {
            if (2 * j >= padFrom) S[j] = 0xFFFFFFFFu;
            else if (2 * j + 1 >= padFrom) S[j] |= 0xFFFF0000u;
        }
    }
}
// local index (inside the last used lane) of the first padding disparity; 2 * nreg when there is none
__host__ __device__ __forceinline__ int sgbm_pad_from(const Geo &g) { return g.D - (g.lanesUsed - 1) * 2 * g.nreg; }

// A multiply-add by an opaque +-1 (a kernel argument the compiler cannot fold) is an IMAD and issues on the
// FMA pipe; the packed min / max / permute instructions issue on the ALU pipe only, at half rate.  Used
// by sgbm_cost3.cu, where the ALU pipe is the nearest bound.  (Tried in the path step as well, b + (C - m)
// as two IMADs: one more instruction per word, sweeps 2 % SLOWER -- they are latency-, not pipe-bound.)
__device__ __forceinline__ uint32_t fma_mad(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// Path start (predecessor outside the image): L = C, m = min C.
template <int NREG, int LPC>
__device__ __forceinline__ uint32_t path_start(uint32_t (&Ln)[NREG], const uint32_t (&C)[NREG], int lg,
                                               int lastLane)
{
#pragma unroll
    for (int j = 0; j < NREG; j++) Ln[j] = C[j];
    uint32_t t = local_min<NREG>(Ln);
    if (lg > lastLane) t = SGBM_INF2;
    return group_min<LPC>(t);
}

// Arguments of the vertical sweep kernel (sgbm_paths.cu), filled by sgbm_api.cu.
struct VertArgs {
    Geo g;
    const uint16_t *C;        // cost volume
    const uint16_t *Calt;     // 3WAY: re-clamped first r rows of stripes 1..3 ([3][r][W1][Dp]) or null
    const uint16_t *inA;      // volumes added into S before the paths of this sweep (may alias sout)
    const uint16_t *inB;      // second input volume or null
    uint16_t *sout;           // non-null: store S (no WTA)
    uint16_t *sdbg;           // test hook: also store the final S of WTA rows here (or null)
    int16_t *raw;             // WTA output, dense H x W int16 (pre-filled with INV)
    unsigned int *d2key;      // WTA disp2 splat keys, dense H x W, pre-filled with 0xFFFFFFFF
    int SW, nstrips;          // max columns per strip (slots), number of strips
    int R;                    // rows per super-step between halo exchanges (NDIR = 3)
    int nslots, nstg, nAB;    // exchange slots, TMA staging depth, staged input volumes (1 or 2)
    unsigned int ssmOff, stgCOff, stgABOff, barOff;   // shared-memory layout (bytes)
    int backward;             // 0: rows 0..H-1, 1: rows H-1..0
    int threeway;             // MODE_SGBM_3WAY rules (stripes via blockIdx.y, tie-break, uniqueness)
    int ss, ov;               // 3WAY stripe height and overlap
    uint16_t *haloA;          // [nstrips][2][R][Dp + 8]  (x-1)-path state of each strip's last R columns
    uint16_t *haloC;          // [nstrips][2][R][Dp + 8]  (x+1)-path state of each strip's first R columns
    unsigned int *flagA;      // [nstrips] super-steps published
    unsigned int *flagC;      // unused
    int dbgNoSync;            // experiment only: skip neighbour-strip waits (results invalid)
    uint16_t *rowState;       // [2][3][W1][Dp + 8] path state of the row-at-a-time fallback (sgbm_sweep.cu)
    unsigned int *watchDev;   // [8] device words of the sweep's hand-off watchdog (sticky until reported)
    unsigned int *watchHost;  // pinned host copy, refreshed after every sweep (or null)
};

#define SGBM_CUDA_CHECK(call)                                                              \
    do {                                                                                   \
        cudaError_t _e = (call);                                                           \
        if (_e != cudaSuccess) return sgbm_fail_cuda(_e, #call, __FILE__, __LINE__);       \
    } while (0)

int sgbm_fail_cuda(cudaError_t e, const char *what, const char *file, int line);
void sgbm_count_launch(int n);          // bump the process-wide kernel launch counter
int sgbm_fail(int code, const char *fmt, ...);
