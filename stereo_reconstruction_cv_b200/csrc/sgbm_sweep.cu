// sgbm_sweep.cu -- the 3-direction sweep of StereoSGBM.compute (main.ipynb:668) for sm_100a:
// one top-down (or bottom-up) pass that advances the vertical path (0,-1) and the two diagonal
// paths (-1,-1), (+1,-1) for every column, accumulates S and either spills it once (MODE_HH
// forward sweep) or runs the fused winner-take-all.
//
// Design (replaces the lock-step k_vertical<.,.,3>, which spent its time in a per-row block
// barrier): one persistent CTA per column strip, warp-specialised into three path roles that each
// keep their path state in REGISTERS and never exchange it inside the CTA:
//   role V : one lane group per column, fixed column              -> vertical path
//   role A : lane groups that move one column to the right per row -> path with predecessor (x-1)
//   role C : lane groups that move one column to the left per row  -> path with predecessor (x+1)
// A lane group of role A/C follows "its" diagonal through the strip, so the predecessor state is
// simply what the group computed on the previous row.  What flows between the roles is the partial
// sum S, through a ring of K shared-memory row slots:  V writes L_h + L_v, A adds its path,
// C adds its path and finishes the pixel (spill or WTA).  Cost rows (and the L_h / S_fwd input
// rows) are staged by a producer warp with TMA bulk copies into multi-stage rings.  All hand-offs
// are mbarrier full/empty pairs, so the roles drift apart by up to K rows and no warp ever waits at
// a block-wide barrier.
//
// In the last sweep of MODE_SGBM / MODE_HH a fourth role W (WROLE kernels) takes the winner-take-all off
// role C: C writes the finished S back into the slot, W reads it, and setmaxnreg moves registers from
// the W / producer warpgroups to the path warpgroups.  The W warps form row groups that take alternate rows
// (one pass of the WTA is a ~1500-cycle dependent chain: with narrow strips its latency, not its throughput,
// would set the row period).  k_rowstep at the end of this file is the row-at-a-time fallback for geometries
// no persistent kernel holds.
//
// A ring stage holds one or two image rows (template parameter RPS): with two, a role waits, advances and arrives once per
// two rows ("wait, (path step, S update) x 2, arrive"), super-steps are multiples of the stage height and the last stage of
// an odd-height image holds one row.  It pays only when the cost ring stays deep (>= 4 stages): see sweep_plan.
//
// The row loops are written for instruction count (round 2): every ring is addressed through a RingPos carried in
// registers (32-bit shared addresses, ld/st.shared with immediate offsets, advanced by additions), the barriers of a
// stage sit side by side, a diagonal role knows its column as one number in staged-window coordinates, and
// super-steps are an outer loop.  profiles/r02_sweep_roles_instruction_counts.md has the per-role counts.
//
// Strips exchange diagonal state every R rows (a "super-step"): at its end, role A publishes the
// state of the strip's last R columns, role C that of its first R columns (global memory, a ring of
// 64 / R super-step slots with one release flag per ring entry).  At the next super-step R chains per role restart from the
// neighbour's published columns, R-1 of them in halo columns outside the strip (the redundant
// triangle that pays for syncing every R rows instead of every row).  Group (b, i) -- batch b,
// index i -- restarts at super-steps n == b (mod NB) at halo position i, so at every row each
// column of the strip is covered by exactly one chain.
#include "sgbm_common.cuh"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

struct SweepArgs {
    Geo g;
    const uint16_t *C, *inA, *inB;
    uint16_t *sout, *sdbg;
    int16_t *raw;
    unsigned int *d2key;
    int SW, nstrips, R, NB;          // max columns per strip, strips, rows per super-step, batches
    int nwA, nwV;                    // warps of each diagonal role / of the vertical role
    int aA, aV, nwW, wPass;          // WROLE: warps allocated per role (multiples of 4), WTA warps, their passes per row
    int wRG, wPR;                    // WROLE: row groups of the WTA warps (group j takes rows t = j mod wRG), warps per row
    int NSC, NSI, K, nAB;            // ring depths (cost rows, input rows, S slots), input volumes
    unsigned int stgCOff, stgIOff, pOff, ssmOff, barOff;
    int backward;
    uint16_t *haloA, *haloC;         // [nstrips][64 / R][R][Dp + 8]
    unsigned int *flagA, *flagC;     // [nstrips][64 / R][R] super-step (+1) each halo ring entry was last published for
    int dbgNoSync;
    int dbgStall;                    // test hook (SGBM_DBG_STALL=1): role V withholds one hand-off so that the watchdog trips
    unsigned int urMagic;            // floor(2^32 / (100 - uniquenessRatio)) + 1
    unsigned int P1p, P2mP1p;        // P1 and P2 - P1 in both 16-bit halves
    // ring geometry in bytes (host-computed so that the row loops add constants instead of multiplying)
    unsigned int colB, cStrideB, cSpanB, iBB, iStrideB, iSpanB, pStrideB, pSpanB, cBarSpan, iBarSpan, pBarSpan;
    unsigned int *dbg;               // [8] hand-off watchdog: {tripped, strip, warp, wait id, row, ...}, zeroed per launch
    unsigned long long *trace;       // debug: clock64 time stamps of one strip [row][role 4][8] (or null)
    int traceStrip;
    int pfDist;                      // rows by which the producer's L2 prefetch runs ahead of its bulk copies (0 = off)
    // Rows per ring stage (1 or 2).  With two rows per stage every hand-off (wait, arrive, ring advance) is paid once per
    // two image rows; cRowB / iRowB / pRowB are the bytes of one row inside a stage of the cost / input / S ring.
    // New fields go LAST: the register allocation of the WTA kernels is sensitive to the layout of the fields above.
    int rps;
    unsigned int cRowB, iRowB, pRowB;
    int wtaSlow;                     // numDisparities % 8 != 0: sweep_wta hides the padding disparities of the last used lane
};

// clock64 time stamps of one strip: compiled in only with -DSGBM_SWEEP_TRACING (make TRACE=1); the
// predicated-off stamps cost ~12 instructions per warp and row otherwise
#ifndef SGBM_SWEEP_TRACING
#define SWEEP_TR(ROLE, IDX, COND) do { } while (0)
#else
#define SWEEP_TR(ROLE, IDX, COND)                                                                          \
    do {                                                                                                   \
        if (a.trace && (COND) && (threadIdx.x & 31) == 0 && (int)blockIdx.x == a.traceStrip)                \
            a.trace[((size_t)t * 4 + (ROLE)) * 8 + (IDX)] = (unsigned long long)clock64();                 \
    } while (0)
#endif


__device__ __forceinline__ void mbar_arrive(SmemBar bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar.addr) : "memory");
}
// Bounded waits of the sweep.  A hand-off that makes no progress for ~2 s (a protocol error: it never
// happens in the validated configurations) does not trap -- that would poison the CUDA context -- but
// records where it happened, raises the abort word and returns; every other wait then returns at once,
// the kernel drains with garbage, and the host reports SGBM_E_CUDA for the frame.
#define SWEEP_WAIT_LIMIT 4000000000ll
#ifdef SGBM_SWEEP_TRACING
__device__ unsigned int g_sweep_prog[1024 * 8];          // tracing builds: current row of every role of every strip
#define SWEEP_PROG(ROLE, COND) do { if ((COND) && (threadIdx.x & 31) == 0) g_sweep_prog[blockIdx.x * 8 + (ROLE)] = (unsigned)t; } while (0)
#else
#define SWEEP_PROG(ROLE, COND) do { } while (0)
#endif
__device__ __noinline__ void sweep_timeout(unsigned int *dbg, int id, int t)
{
    if (atomicExch(&dbg[0], 1u) == 0u) {
        dbg[1] = blockIdx.x; dbg[2] = threadIdx.x >> 5; dbg[3] = (unsigned)id; dbg[4] = (unsigned)t;
        __threadfence();
#ifdef SGBM_SWEEP_TRACING
        for (int s = (int)blockIdx.x - 3; s <= (int)blockIdx.x + 3; s++)
            if (s >= 0 && s < (int)gridDim.x)
                printf("strip %d: V %u  A0 %u  Alast %u  C0 %u  Clast %u  W %u  P %u\n", s, g_sweep_prog[s * 8 + 0], g_sweep_prog[s * 8 + 1],
                       g_sweep_prog[s * 8 + 2], g_sweep_prog[s * 8 + 3], g_sweep_prog[s * 8 + 4], g_sweep_prog[s * 8 + 5], g_sweep_prog[s * 8 + 6]);
#endif
    }
}
__device__ __forceinline__ void sweep_wait(const SweepArgs &a, SmemBar bar, uint32_t parity, int id, int t)
{
    if (mbar_try_wait(bar, parity)) return;
    // the retry loop stays inline PTX (no call, two scratch registers): waiting here is the normal case
    uint32_t to;
    asm volatile(
        "{\n\t"
        ".reg .pred p, q;\n\t"
        ".reg .u32 f;\n\t"
        ".reg .u64 c0, c1;\n\t"
        "mov.u32 %0, 0;\n\t"
        "mov.u64 c0, %%clock64;\n"
        "SGBM_SW_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 0x186A0;\n\t"
        "@p bra SGBM_SW_DONE;\n\t"
        "ld.volatile.global.u32 f, [%3];\n\t"
        "setp.ne.u32 q, f, 0;\n\t"
        "@q bra SGBM_SW_DONE;\n\t"
        "mov.u64 c1, %%clock64;\n\t"
        "sub.u64 c1, c1, c0;\n\t"
        "setp.lt.u64 q, c1, 4000000000;\n\t"
        "@q bra SGBM_SW_WAIT;\n\t"
        "mov.u32 %0, 1;\n"
        "SGBM_SW_DONE:\n\t"
        "}" : "=r"(to) : "r"(bar.addr), "r"(parity), "l"(a.dbg) : "memory");
    if (to) sweep_timeout(a.dbg, id, t);
}
__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int *p)
{
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(unsigned int *p, unsigned int v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
template <int NREG, int LPC>
__device__ __forceinline__ void load_vec_l2(uint32_t (&v)[NREG], const uint16_t *col, int lg)
{
    const uint4 *p = reinterpret_cast<const uint4 *>(col) + lg;
#pragma unroll
    for (int k = 0; k < NREG / 4; k++) {
        uint4 q = __ldcg(p + LPC * k);
        v[4 * k + 0] = q.x; v[4 * k + 1] = q.y; v[4 * k + 2] = q.z; v[4 * k + 3] = q.w;
    }
}

// Barriers.  The stages of a ring keep their barriers side by side, so that a role carries ONE barrier
// address per ring and reaches the others through immediate offsets:
//   cost ring   stage q: barC + 16 q   { +0 fullC, +8 emptyC }
//   input ring  stage q: barI + 16 q   { +0 fullI, +8 emptyI }
//   S ring      slot  q: barP + 32 q   { +0 fullV, +8 fullM, +16 fullW, +24 freeP }
#define BAR_FULL 0u
#define BAR_EMPTY 8u
#define BAR_FULLV 0u
#define BAR_FULLM 8u
#define BAR_FULLW 16u
#define BAR_FREEP 24u
struct SweepSmem {
    uint16_t *ssm;                   // [groups of role C][Dp]  WTA scratch of the kernels without the W role
    uint32_t aC, aI, aP;             // 32-bit shared addresses of the rings: stgC [NSC][SW + 2(R-1)][Dp], stgI [NSI][nAB][SW][Dp], P [K][SW][Dp]
    uint32_t barC, barI, barP;
};
__device__ __forceinline__ SweepSmem sweep_carve(const SweepArgs &a, uint8_t *smem)
{
    SweepSmem s;
    s.ssm = reinterpret_cast<uint16_t *>(smem + a.ssmOff);
    const uint32_t base = smem_u32(smem);
    s.aC = base + a.stgCOff; s.aI = base + a.stgIOff; s.aP = base + a.pOff;
    s.barC = base + a.barOff;
    s.barI = s.barC + 16u * (uint32_t)a.NSC;
    s.barP = s.barI + 16u * (uint32_t)a.NSI;
    return s;
}

// A role's position in one ring: everything the row loop needs of it lives in these registers and is
// advanced by additions -- no per-row multiplications, no re-derivation from the kernel arguments.
struct RingPos {
    uint32_t data;                   // shared address of this thread's data in the current stage
    uint32_t bar;                    // shared address of the current stage's barrier group
    uint32_t par;                    // phase parity of the current pass over the ring
    int left;                        // stages until the ring wraps
};
__device__ __forceinline__ RingPos ring_start(uint32_t data, uint32_t bar, int depth)
{
    RingPos r;
    r.data = sm_keep(data); r.bar = sm_keep(bar); r.par = 0u; r.left = depth;
    return r;
}
// stride / span = bytes per stage / per ring of the data, barStep / barSpan the same for the barriers
__device__ __forceinline__ void ring_advance(RingPos &r, uint32_t stride, uint32_t span, uint32_t barStep, uint32_t barSpan, int depth)
{
    r.data += stride; r.bar += barStep;
    if (--r.left == 0) { r.left = depth; r.data -= span; r.bar -= barSpan; r.par ^= 1u; }
}
// the same when the row loop has already moved r.data over the rows of the stage (full stages only: the last, partial stage
// of an image is never followed by another one)
__device__ __forceinline__ void ring_advance_moved(RingPos &r, uint32_t span, uint32_t barStep, uint32_t barSpan, int depth)
{
    r.bar += barStep;
    if (--r.left == 0) { r.left = depth; r.data -= span; r.bar -= barSpan; r.par ^= 1u; }
}
// the same for a role that visits every n-th stage (n <= depth)
__device__ __forceinline__ void ring_advance_n(RingPos &r, int n, uint32_t stride, uint32_t span, uint32_t barStep, uint32_t barSpan, int depth)
{
    r.data += (uint32_t)n * stride; r.bar += (uint32_t)n * barStep; r.left -= n;
    if (r.left <= 0) { r.left += depth; r.data -= span; r.bar -= barSpan; r.par ^= 1u; }
}
__device__ __forceinline__ void bar_arrive(uint32_t addr) { mbar_arrive(SmemBar{addr}); }
// (measured in round 2: without the early probes -- every hand-off one blocking wait -- all sweeps are within 0.02 ms)
__device__ __forceinline__ bool bar_test(uint32_t addr, uint32_t parity) { return mbar_test_wait(SmemBar{addr}, parity); }
// L2 prefetch of a contiguous global range (no destination, no completion): the producer runs it several rows
// ahead of the bulk copies, so that the copies are served from L2 -- a prefetch depth that does not cost
// shared memory the way a deeper staging ring does.
__device__ __forceinline__ void bulk_prefetch_l2(const void *src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s_a(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// Rows per ring stage (template parameter RPS of everything below): 1 or 2.  The 16-register lane mappings
// (numDisparities >= 256 and a few odd sizes) never have the shared memory for two rows per stage at the widths the
// persistent sweep holds, so only RPS = 1 is instantiated for them.
constexpr int SWEEP_RPS_MAX_NREG = 12;

// ---- producer: TMA bulk copies of the cost rows and the input rows of every ring stage ---------------
// PF: with an L2 prefetch a.pfDist rows ahead -- the producer of the spilling (forward) sweep of MODE_HH, whose speed
// follows the depth of its cost ring.  Compiled apart on purpose: in the winner-take-all kernels the producer lives in
// the 40-register warpgroup, and any growth of its code showed up as spills there (cfg3 WTA sweep 3.20 -> 3.37 ms).
template <int RPS, bool PF>
__device__ __forceinline__ void sweep_producer_t(const SweepArgs &a, const SweepSmem &s, int xs, int xe, int yBegin,
                                                 int yStep, int nRows)
{
    const Geo &g = a.g;
    const int Dp = g.Dp, HG = a.R - 1;
    const int scol0 = xs - HG;
    const int clo = max(scol0, 0), chi = min(xe + HG, g.W1);
    const uint32_t colB = a.colB;
    const uint32_t bytesC = (uint32_t)(chi - clo) * colB, bytesI = (uint32_t)(xe - xs) * colB;
    RingPos rc = ring_start(s.aC + (uint32_t)(clo - scol0) * colB, s.barC, a.NSC);
    RingPos ri = ring_start(s.aI, s.barI, a.NSI);
    const uint16_t *srcC = a.C + (size_t)yBegin * g.rowStride + (size_t)clo * Dp;
    const uint16_t *srcA = a.inA + (size_t)yBegin * g.rowStride + (size_t)xs * Dp;
    const uint16_t *srcB = a.nAB > 1 ? a.inB + (size_t)yBegin * g.rowStride + (size_t)xs * Dp : nullptr;
    const long long rowStep = (long long)yStep * g.rowStride;
    constexpr int rps = RPS;
    const int pf = PF ? a.pfDist : 0;
    const long long pfOff = (long long)pf * rowStep;
    // (a freshly initialised barrier passes a wait on the opposite parity at once: the first pass over a ring needs no test)
    for (int t = 0; t < nRows; t += rps) {
        const int rows = min(rps, nRows - t);
        if (PF && pf > 0 && t + pf + rps <= nRows) {
            for (int r = 0; r < rows; r++) {
                bulk_prefetch_l2(srcC + pfOff + r * rowStep, bytesC);
                bulk_prefetch_l2(srcA + pfOff + r * rowStep, bytesI);
                if (srcB) bulk_prefetch_l2(srcB + pfOff + r * rowStep, bytesI);
            }
        }
        SWEEP_TR(3, 0, true);
        sweep_wait(a, SmemBar{rc.bar + BAR_EMPTY}, rc.par ^ 1u, 1, t);
        SWEEP_TR(3, 1, true);
        mbar_expect_tx(SmemBar{rc.bar + BAR_FULL}, bytesC * (uint32_t)rows);
        for (int r = 0; r < rows; r++) bulk_g2s_a(rc.data + (uint32_t)r * a.cRowB, srcC + r * rowStep, bytesC, rc.bar + BAR_FULL);
        sweep_wait(a, SmemBar{ri.bar + BAR_EMPTY}, ri.par ^ 1u, 2, t);
        mbar_expect_tx(SmemBar{ri.bar + BAR_FULL}, bytesI * (uint32_t)(a.nAB * rows));
        for (int r = 0; r < rows; r++) {
            bulk_g2s_a(ri.data + (uint32_t)r * a.iRowB, srcA + r * rowStep, bytesI, ri.bar + BAR_FULL);
            if (srcB) bulk_g2s_a(ri.data + (uint32_t)r * a.iRowB + a.iBB, srcB + r * rowStep, bytesI, ri.bar + BAR_FULL);
        }
        SWEEP_TR(3, 2, true);
        ring_advance(rc, a.cStrideB, a.cSpanB, 16u, a.cBarSpan, a.NSC);
        ring_advance(ri, a.iStrideB, a.iSpanB, 16u, a.iBarSpan, a.NSI);
        srcC += rows * rowStep; srcA += rows * rowStep;
        if (srcB) srcB += rows * rowStep;
    }
}

// S accumulation: saturating (A.4) unless the host proved that the sum of all paths fits 16 bits, in
// which case plain adds are exact and the clamp to 32767 is applied once when the pixel is finished.
template <bool SAT>
__device__ __forceinline__ uint32_t sacc(uint32_t s, uint32_t x) { return SAT ? paddmin(s, x, SGBM_MAX_S) : s + x; }

// ---- role V: vertical path, starts the S slot of every row ------------------------------------------
template <int NREG, int LPC, bool SAT, int RPS>
__device__ __forceinline__ void sweep_role_v(const SweepArgs &a, const SweepSmem &s, int rwarp, int SW, int nRows)
{
    constexpr int GPW = 32 / LPC;
    const Geo &g = a.g;
    const int lane = threadIdx.x & 31, lg = lane % LPC;
    const int HG = a.R - 1;
    const int gi = rwarp * GPW + lane / LPC;
    const int ci = gi < SW ? gi : SW - 1;
    const uint32_t own = sm_keep(gi < SW ? 1u : 0u), lane0 = sm_keep(lane == 0 ? 1u : 0u);
    const LaneMasks lm = lane_masks(lg, g.lanesUsed - 1);
    const bool hasB = a.nAB > 1;
    const int NSC = a.NSC, NSI = a.NSI, K = a.K;
    // this lane's chunk of its column inside stage 0 of each ring, and the ring geometry in bytes
    const uint32_t colB = a.colB, laneB = 16u * (uint32_t)lg;
    RingPos rc = ring_start(s.aC + (uint32_t)(HG + ci) * colB + laneB, s.barC, NSC);
    RingPos ri = ring_start(s.aI + (uint32_t)ci * colB + laneB, s.barI, NSI);
    RingPos rp = ring_start(s.aP + (uint32_t)ci * colB + laneB, s.barP, K);
    uint32_t LB[NREG], mB = 0;
#pragma unroll
    for (int j = 0; j < NREG; j++) LB[j] = 0;
    constexpr int rps = RPS;
    bool okC = false;                                     // early probe of the stage's cost rows
    for (int t = 0; t < nRows; t += rps) {
        const int rows = min(rps, nRows - t);
        SWEEP_PROG(0, rwarp == 0);
        SWEEP_TR(0, 0, rwarp == 0);
        if (!okC) sweep_wait(a, SmemBar{rc.bar + BAR_FULL}, rc.par, 3, t);
        SWEEP_TR(0, 1, rwarp == 0);
        // probes whose latency hides behind the first path step
        bool okI = bar_test(ri.bar + BAR_FULL, ri.par);
        bool okP = bar_test(rp.bar + BAR_FREEP, rp.par ^ 1u);   // (a fresh barrier passes the opposite parity at once)
        for (int r = rows; r > 0; r--) {
            uint32_t S[NREG];
            {
                uint32_t Cc[NREG];
                lds_vec<NREG, LPC>(Cc, rc.data);
                mB = path_step_m<NREG, LPC>(LB, mB, Cc, a.P1p, a.P2mP1p, lm);
            }
            if constexpr (RPS == 1) {
                // one row per stage: the cost row is in registers, hand the stage back now and look at the next one
                __syncwarp();
                if (lane0) bar_arrive(rc.bar + BAR_EMPTY);
                ring_advance(rc, a.cStrideB, a.cSpanB, 16u, a.cBarSpan, NSC);
                okC = bar_test(rc.bar + BAR_FULL, rc.par);
            } else {
                rc.data += a.cRowB;
            }
            SWEEP_TR(0, 2, rwarp == 0);
            if (!okI) { sweep_wait(a, SmemBar{ri.bar + BAR_FULL}, ri.par, 4, t); okI = true; }
            lds_vec<NREG, LPC>(S, ri.data);
#pragma unroll
            for (int j = 0; j < NREG; j++) S[j] = sacc<SAT>(S[j], LB[j]);
            if (hasB) {
                uint32_t Bv[NREG];
                lds_vec<NREG, LPC>(Bv, ri.data + a.iBB);
#pragma unroll
                for (int j = 0; j < NREG; j++) S[j] = sacc<SAT>(S[j], Bv[j]);
            }
            SWEEP_TR(0, 3, rwarp == 0);
            if (!okP) { sweep_wait(a, SmemBar{rp.bar + BAR_FREEP}, rp.par ^ 1u, 5, t); okP = true; }
            SWEEP_TR(0, 4, rwarp == 0);
            if (own) sts_vec<NREG, LPC>(S, rp.data);
            ri.data += a.iRowB; rp.data += a.pRowB;
        }
        // the stage's rows are consumed: hand the cost and input stages back, publish the S slot, look at the next stage
        __syncwarp();
        if (lane0) {
            if constexpr (RPS > 1) bar_arrive(rc.bar + BAR_EMPTY);
            if (!(SGBM_DBG_HOOK(a.dbgStall) && t <= 5 && 5 < t + rps && blockIdx.x == 0)) bar_arrive(rp.bar + BAR_FULLV);
            bar_arrive(ri.bar + BAR_EMPTY);
        }
        SWEEP_TR(0, 5, rwarp == 0);
        ring_advance_moved(ri, a.iSpanB, 16u, a.iBarSpan, NSI);
        ring_advance_moved(rp, a.pSpanB, 32u, a.pBarSpan, K);
        if constexpr (RPS > 1) {
            ring_advance_moved(rc, a.cSpanB, 16u, a.cBarSpan, NSC);
            okC = bar_test(rc.bar + BAR_FULL, rc.par);
        }
    }
}

// ---- winner-take-all of one pixel (A.5), S distributed over the lane group --------------------------
// Minimum and first arg-minimum come out of ONE reduction over 32-bit keys (S << 16 | d): the key
// order is the order "(cost, disparity)", so the smallest key is the first minimum.  The
// uniqueness scan re-reads S from shared scratch with the window around the winner masked and
// combines the lanes' verdicts with one ballot.
template <int NREG, int LPC>
__device__ __forceinline__ void sweep_wta(const SweepArgs &a, uint32_t (&S)[NREG], uint16_t *ssm, int lg, bool own,
                                          int x1, int y)
{
    const Geo &g = a.g;
    const int lastLane = g.lanesUsed - 1;
    if (a.wtaSlow) mask_pad_regs<NREG>(S, lg == lastLane, sgbm_pad_from(g));    // (a no-op unless numDisparities % 8 != 0)
    __syncwarp();
    store_vec<NREG, LPC>(S, ssm, lg);                     // scratch for the sub-pixel neighbours / masked re-scan
    uint32_t km[NREG];
#pragma unroll
    for (int j = 0; j < NREG; j++) {
        const uint32_t idx2 = (uint32_t)(2 * j) | ((uint32_t)(2 * j + 1) << 16);
        const uint32_t klo = __byte_perm(S[j], idx2, 0x1054);           // S(2j)   << 16 | 2j
        const uint32_t khi = __byte_perm(S[j], idx2, 0x3276);           // S(2j+1) << 16 | 2j+1
        km[j] = min(klo, khi);
    }
    uint32_t key = km[0];
#pragma unroll
    for (int j = 1; j + 1 < NREG; j += 2) key = __vimin3_u32(key, km[j], km[j + 1]);
    if ((NREG & 1) == 0) key = min(key, km[NREG - 1]);
    key += (uint32_t)(lg * 2 * NREG);                                   // lane-local index -> disparity
    if (lg > lastLane) key = 0xFFFFFFFFu;
#pragma unroll
    for (int off = LPC / 2; off >= 1; off >>= 1) key = min(key, __shfl_xor_sync(0xFFFFFFFFu, key, off, LPC));
    const int minS = (int)(key >> 16);
    const int best = (minS == 32767) ? -1 : (int)(key & 0xFFFFu);       // first minimum (A.5)
    __syncwarp();
    int Sm = 0, Sp = 0;
    const bool interior = best > 0 && best < g.D - 1;
    if (lg == 0 && interior) {
        Sm = ssm[sgbm_pos(best - 1, NREG, LPC)];
        Sp = ssm[sgbm_pos(best + 1, NREG, LPC)];
    }
    bool reject = false;
    if (g.UR > 0) {
        const int av = 100 - g.UR;                                       // S(d)*(100-UR) < minS*100  <=>  S(d) < T
        // ceil(100*minS / av) by multiply-high with M = floor(2^32/av)+1: exact for numerators < 2^22
        const unsigned num = (unsigned)(100 * minS + av - 1);
        const int T = av > 0 ? min((int)(av == 1 ? num : __umulhi(num, a.urMagic)), 32768) : (minS > 0 ? 32768 : 0);
        if (lg == 0) {
#pragma unroll
            for (int dd = -1; dd <= 1; dd++) {
                const int d = best + dd;
                if (d >= 0 && d < g.D) ssm[sgbm_pos(d, NREG, LPC)] = 0xFFFFu;
            }
        }
        __syncwarp();
        uint32_t S2[NREG];
        load_vec<NREG, LPC>(S2, ssm, lg);
        const uint32_t t2 = local_min<NREG>(S2);
        const bool viol = lg <= lastLane && (int)(t2 & 0xFFFFu) < T;
        const unsigned ball = __ballot_sync(0xFFFFFFFFu, viol);
        const int lane = threadIdx.x & 31;
        const unsigned gmask = (LPC == 32 ? 0xFFFFFFFFu : ((1u << LPC) - 1u)) << (lane & ~(LPC - 1));
        reject = (ball & gmask) != 0u;
    }
    if (lg == 0 && own) {
        const int x = x1 + g.minX1;
        int out = g.INV;
        if (!reject) {
            const int x2 = x - best - g.minD;
            if (minS < 32767 && x2 >= 0 && x2 < g.W)
                atomicMin(a.d2key + (size_t)y * g.W + x2, ((unsigned)minS << 16) | (0xFFFFu - (unsigned)x1));
            int dq = best * 16;
            if (interior) {
                const int den = max(Sm + Sp - 2 * minS, 1);
                // C division (toward zero) through one IEEE float division: |num| < 2^20 and 2*den < 2^18 are
                // exact floats and the rounded quotient cannot reach the next integer (error < 1/(2*den))
                dq += (int)((float)((Sm - Sp) * 16 + den) / (float)(2 * den));
            }
            out = dq + g.minD * 16;
        }
        a.raw[(size_t)y * g.W + x] = (int16_t)out;
    }
}

// ---- roles A (DIR = +1) and C (DIR = -1, finishes the pixel) -----------------------------------------
template <int NREG, int LPC, int DIR, bool SAT, bool WROLE, int RPS>
__device__ __forceinline__ void sweep_role_diag(const SweepArgs &a, const SweepSmem &s, int rwarp, int strip, int xs,
                                                int xe, int yBegin, int yStep, int nRows)
{
    constexpr int GPW = 32 / LPC;
    // Fixed order V -> A -> C, role C finishes the pixel.  Two alternatives were built and measured
    // (4K, D=256, WTA sweep 4.1 ms) and are slower: letting A and C take turns in finishing rows chains
    // the two roles' rows into one serial sequence (5.5 ms); moving the winner-take-all tail onto roles
    // V / A (even / odd rows) couples V and A row by row through the slot release and the SM's ALU pipe
    // is already 55 % busy, so the freed time of role C cannot be used (7.0 ms).
    constexpr bool FINAL = DIR < 0;
    const Geo &g = a.g;
    const int lane = threadIdx.x & 31, lg = lane % LPC;
    const int Dp = g.Dp, lastLane = g.lanesUsed - 1, R = a.R, HG = R - 1, NB = a.NB;
    const int gg = rwarp * GPW + lane / LPC;
    const int b = gg / R, i = gg - b * R;
    const bool exists = b < NB;
    int p = i + ((NB - b) % NB) * R;                      // position at row 0 (chains "started" before the image)
    const bool hasNbr = DIR > 0 ? strip > 0 : strip < a.nstrips - 1;
    const bool canPub = DIR > 0 ? strip < a.nstrips - 1 : strip > 0;
    const size_t haloStride = (size_t)Dp + 8;
    // Halo ring: 64 column entries per strip = HS = 64 / R super-step slots.  A strip's publishing role can
    // run ahead of the neighbour's consuming role by the S-ring depth plus one super-step, i.e. by up to
    // (K + R) / R + 1 slots, and nothing but the depth of this ring holds it back.
    const int HS = 64 / R;
    uint16_t *haloOut = (DIR > 0 ? a.haloA : a.haloC) + (size_t)strip * 64 * haloStride;
    // One flag per halo ring ENTRY (slot, column), not per column: the warps of a role drift apart by up to
    // the S-ring depth, so the same column can be published for super-step n+1 by one warp before another
    // warp has published it for super-step n.  A per-column "latest super-step" flag then lets the
    // neighbour read a slot that is not written yet, and a late smaller value can even make it wait
    // forever (seen as rare mismatches / watchdog trips with R <= 3 rows per super-step).
    unsigned int *flagOut = (DIR > 0 ? a.flagA : a.flagC) + (size_t)strip * 64;
    const int nbr = DIR > 0 ? strip - 1 : strip + 1;
    const int hidx = DIR > 0 ? i : R - 1 - i;             // which published column chain i continues
    const uint16_t *haloIn = (DIR > 0 ? a.haloA : a.haloC) + ((size_t)nbr * 64 + hidx) * haloStride;
    const unsigned int *flagIn = (DIR > 0 ? a.flagA : a.flagC) + (size_t)nbr * 64 + hidx;
    uint16_t *ssm = s.ssm + (size_t)gg * Dp;          // (kernels without the W role only)

    uint32_t L[NREG], m = 0;
#pragma unroll
    for (int j = 0; j < NREG; j++) L[j] = 0;              // "predecessor outside" == L = 0, m = 0 (A.4)
    const int NSC = a.NSC, K = a.K;
    const LaneMasks lm = lane_masks(lg, lastLane);
    const uint32_t lane0 = sm_keep(lane == 0 ? 1u : 0u);
    // this lane's chunk of column 0 of stage 0 of the cost ring / of slot 0 of the S ring
    const uint32_t colB = a.colB;
    RingPos rc = ring_start(s.aC + 16u * (uint32_t)lg, s.barC, NSC);
    RingPos rp = ring_start(s.aP + 16u * (uint32_t)lg, s.barP, K);
    // WROLE: the finished S goes back into the slot and the winner-take-all warps release it
    constexpr uint32_t WAIT_BAR = FINAL ? BAR_FULLM : BAR_FULLV;
    constexpr uint32_t DONE_BAR = FINAL ? (WROLE ? BAR_FULLW : BAR_FREEP) : BAR_FULLM;
    // Column bookkeeping in window coordinates: q = column - (xs - HG) is the group's place in the staged cost row
    // (q = p for the path that moves right, q = WW - 1 - p for the one that moves left).  Everything the row loop asks
    // about the column is an unsigned range test on q against per-thread constants made opaque once:
    //   active  <=>  q - aLo < aLen      (the chain is inside the strip + its incoming halo, and inside the image)
    //   own     <=>  active and q - HG < SWs
    const int SWs = xe - xs, WW = SWs + 2 * HG;
    int aLoI, aHiI;
    if (DIR > 0) { aLoI = max(0, HG - xs); aHiI = HG + SWs; }            // col >= 0, col < xe
    else { aLoI = HG; aHiI = min(WW, g.W1 - xs + HG); }                  // col >= xs, col < W1
    const uint32_t aLo = sm_keep((uint32_t)aLoI), aLen = sm_keep(exists && aHiI > aLoI ? (uint32_t)(aHiI - aLoI) : 0u);
    const uint32_t hgU = (uint32_t)HG, swsU = (uint32_t)SWs;
    // strips at the image border: the chain enters the image when col <= 0 (col >= W1 - 1): predecessor outside
    const uint32_t edge = sm_keep(hasNbr ? 0u : 1u);
    const int edgeQ = DIR > 0 ? HG - xs : g.W1 - 1 - xs + HG;            // q of column 0 / column W1 - 1
    const uint32_t pubOn = sm_keep(canPub ? 1u : 0u);
    const int pubQ0 = DIR > 0 ? HG + SWs - R : HG;                      // q of the first published column
    int hs = 0, hsPrev = 0;                               // n % HS and (n - 1) % HS: halo ring slots of this / the previous super-step
    int nm = 0, tq = 0;                                   // tq: first row of the current super-step
    // spilling sweep: byte pointer to this lane's chunk of column xs - HG of the current output row (advanced per row,
    // so that the store address is one multiply-add instead of two 64-bit products)
    char *soutRow = (FINAL && !WROLE && a.sout)
                        ? reinterpret_cast<char *>(a.sout) + ((long long)yBegin * g.rowStride + (long long)(xs - HG) * Dp) * 2 + 16 * lg : nullptr;
    const long long soutStep = (long long)yStep * g.rowStride * 2;
    constexpr int rps = RPS;
    bool okC = false;                                     // early probe of the stage's cost rows
    for (int n = 0; tq < nRows; n++) {
        // ---- super-step start: batch nm restarts from the neighbour's published columns -------------
        if (n > 0) {
            const bool restart = exists && nm == b;
            if (restart) p = i;
            if (hasNbr) {
                if (restart && lg == 0 && !SGBM_DBG_HOOK(a.dbgNoSync)) {
                    const long long t0 = clock64();
                    const unsigned int *fl = flagIn + hsPrev * R;
                    while (ld_acquire_u32(fl) < (unsigned)n) {
                        __nanosleep(32);                  // (spinning instead: cfg3 +0.12 ms, small frames -0.01 ms)
                        if (*reinterpret_cast<volatile unsigned int *>(a.dbg) != 0u) break;
                        if (clock64() - t0 > SWEEP_WAIT_LIMIT) { sweep_timeout(a.dbg, DIR > 0 ? 6 : 7, tq); break; }
                    }
                }
                __syncwarp();
                if (restart) {
                    const uint16_t *h = haloIn + (size_t)hsPrev * R * haloStride;
                    load_vec_l2<NREG, LPC>(L, h, lg);
                    m = __ldcg(reinterpret_cast<const unsigned int *>(h + Dp));
                }
            } else if (restart) {
#pragma unroll
                for (int j = 0; j < NREG; j++) L[j] = 0;
                m = 0;
            }
        }
        const int rowsHere = min(R, nRows - tq);
        const bool pubStep = pubOn && tq + R < nRows;     // a full super-step that is not the image's last rows
        int q = DIR > 0 ? p : WW - 1 - p;
        // (rl counts the rows left in the super-step; the row index t = tq + rowsHere - rl is only needed off the hot path.
        //  R is a multiple of the rows per ring stage, so a stage never straddles two super-steps.)
        for (int rl = rowsHere; rl > 0; rl -= rps) {
            const int rows = min(rps, rl);
#if defined(SGBM_SWEEP_TRACING)
            const int t = tq + rowsHere - rl;
#endif
            SWEEP_PROG(DIR > 0 ? 1 : 3, rwarp == 0);
            SWEEP_PROG(DIR > 0 ? 2 : 4, rwarp == a.nwA - 1);
            SWEEP_TR(DIR > 0 ? 1 : 2, 0, rwarp == a.nwA / 2);
            if (!okC) sweep_wait(a, SmemBar{rc.bar + BAR_FULL}, rc.par, DIR > 0 ? 8 : 9, tq);
            SWEEP_TR(DIR > 0 ? 1 : 2, 1, rwarp == a.nwA / 2);
            bool okS = bar_test(rp.bar + WAIT_BAR, rp.par);   // latency hides behind the first path step
            for (int r = rows; r > 0; r--) {
                const bool active = (uint32_t)q - aLo < aLen;
                const bool own = active && (uint32_t)q - hgU < swsU;
                if (edge && (DIR > 0 ? q <= edgeQ : q >= edgeQ)) {   // chain enters the image: predecessor outside
#pragma unroll
                    for (int j = 0; j < NREG; j++) L[j] = 0;
                    m = 0;
                }
                if (__any_sync(0xFFFFFFFFu, active)) {       // inactive groups compute garbage that is never used
                    uint32_t Cc[NREG];
                    lds_vec<NREG, LPC>(Cc, rc.data + (active ? (uint32_t)q : hgU) * colB);
                    m = path_step_m<NREG, LPC>(L, m, Cc, a.P1p, a.P2mP1p, lm);
                }
                if constexpr (RPS == 1) {
                    // one row per stage: the cost row is in registers, hand the stage back now and look at the next one
                    __syncwarp();
                    if (lane0) bar_arrive(rc.bar + BAR_EMPTY);
                    ring_advance(rc, a.cStrideB, a.cSpanB, 16u, a.cBarSpan, NSC);
                    okC = bar_test(rc.bar + BAR_FULL, rc.par);
                } else {
                    rc.data += a.cRowB;
                }
                // ---- super-step end: publish the columns the neighbour continues ------------------------
                if (pubStep && rl + r == rows + 1) {              // last row of the super-step
                    const int pi = q - pubQ0;
                    const bool pub = own && pi >= 0 && pi < R;
                    if (pub) {
                        uint16_t *h = haloOut + ((size_t)hs * R + pi) * haloStride;
                        store_vec<NREG, LPC>(L, h, lg);
                        if (lg == 0) *reinterpret_cast<unsigned int *>(h + Dp) = m;
                    }
                    __syncwarp();
                    if (pub && lg == 0) {
                        // st.release.gpu is itself a release fence: cumulative over the group's stores, which the
                        // __syncwarp above has ordered before this lane (a separate __threadfence() doubled the
                        // MEMBAR / CCTL.IVALL cost of every publication)
                        st_release_u32(flagOut + hs * R + pi, (unsigned)(n + 1));
                    }
                }
                // ---- S slot of this row -----------------------------------------------------------------
                SWEEP_TR(DIR > 0 ? 1 : 2, 2, rwarp == a.nwA / 2);
                if (!okS) { sweep_wait(a, SmemBar{rp.bar + WAIT_BAR}, rp.par, DIR > 0 ? 10 : 11, tq); okS = true; }
                SWEEP_TR(DIR > 0 ? 1 : 2, 3, rwarp == a.nwA / 2);
                uint32_t S[NREG];
                if (own) {
                    const uint32_t ps = rp.data + ((uint32_t)q - hgU) * colB;
                    lds_vec<NREG, LPC>(S, ps);
#pragma unroll
                    for (int j = 0; j < NREG; j++) S[j] = sacc<SAT>(S[j], L[j]);
                    if (!FINAL || WROLE) sts_vec<NREG, LPC>(S, ps);
                } else if (FINAL && !WROLE) {
#pragma unroll
                    for (int j = 0; j < NREG; j++) S[j] = SGBM_MAX_S;
                }
                if (FINAL && !WROLE && __any_sync(0xFFFFFFFFu, own)) {
                    const int y = yBegin + (tq + rowsHere - rl + rows - r) * yStep;
                    const int x1 = own ? xs - HG + q : xs;
                    if (a.sout) {
                        if (own) {
                            uint4 *po = reinterpret_cast<uint4 *>(soutRow + (size_t)((uint32_t)q * colB));
#pragma unroll
                            for (int k4 = 0; k4 < NREG / 4; k4++) po[LPC * k4] = make_uint4(S[4 * k4 + 0], S[4 * k4 + 1], S[4 * k4 + 2], S[4 * k4 + 3]);
                        }
                    } else {
                        if (!SAT) {
#pragma unroll
                            for (int j = 0; j < NREG; j++) S[j] = pmin(S[j], SGBM_MAX_S);
                        }
                        if (a.sdbg && own) store_vec<NREG, LPC>(S, a.sdbg + (size_t)y * g.rowStride + (size_t)x1 * Dp, lg);
                        sweep_wta<NREG, LPC>(a, S, ssm, lg, own, x1, y);
                    }
                }
                q += DIR;
                if (FINAL && !WROLE) soutRow += soutStep;
                rp.data += a.pRowB;
            }
            // the stage's rows are consumed: hand the cost stage back, pass the S slot on, look at the next stage
            __syncwarp();
            if (lane0) {
                if constexpr (RPS > 1) bar_arrive(rc.bar + BAR_EMPTY);
                bar_arrive(rp.bar + DONE_BAR);
            }
            SWEEP_TR(DIR > 0 ? 1 : 2, 4, rwarp == a.nwA / 2);
            ring_advance_moved(rp, a.pSpanB, 32u, a.pBarSpan, K);
            if constexpr (RPS > 1) {
                ring_advance_moved(rc, a.cSpanB, 16u, a.cBarSpan, NSC);
                okC = bar_test(rc.bar + BAR_FULL, rc.par);
            }
            SWEEP_TR(DIR > 0 ? 1 : 2, 5, rwarp == a.nwA / 2);
        }
        p += rowsHere; tq += rowsHere;
        if (++nm == NB) nm = 0;
        hsPrev = hs;
        if (++hs == HS) hs = 0;
    }
}

// ---- role W (WROLE kernels): winner-take-all of the finished rows --------------------------------------
// Reads the final S of a pixel from the ring slot role C left it in, uses the slot itself as the scratch
// of the masked uniqueness re-scan, and releases the slot.  S in the slot is the plain (unclamped) sum
// when !SAT; the clamp to 32767 (A.4) is applied to what is read.
template <int NREG, int LPC, bool SAT>
__device__ __forceinline__ void sweep_wta_slot(const SweepArgs &a, uint32_t colA, int lg, uint32_t padMask, unsigned gmask,
                                               bool own, int x1, int y)
{
    // colA: shared address of the column's vector in the ring slot (lane chunk NOT included)
    const Geo &g = a.g;
    const uint32_t laneA = colA + 16u * (uint32_t)lg;
    uint32_t key = 0xFFFFFFFFu;
    {
        uint32_t S[NREG];
        lds_vec<NREG, LPC>(S, laneA);
        // The clamp to 32767 is monotone, so it commutes with the min reduction; a saturated minimum makes
        // the pixel invalid whichever disparity carries it, so only the debug dump needs clamped vectors.
        if (!SAT && a.sdbg) {
#pragma unroll
            for (int j = 0; j < NREG; j++) S[j] = pmin(S[j], SGBM_MAX_S);
        }
        if (a.sdbg && own) store_vec<NREG, LPC>(S, a.sdbg + (size_t)y * g.rowStride + (size_t)x1 * g.Dp, lg);
#pragma unroll
        for (int j = 0; j < NREG; j += 2) {
            const uint32_t i0 = (uint32_t)(2 * j) | ((uint32_t)(2 * j + 1) << 16), i1 = i0 + 0x00020002u;
            const uint32_t k0 = min(__byte_perm(S[j], i0, 0x1054), __byte_perm(S[j], i0, 0x3276));
            const uint32_t k1 = min(__byte_perm(S[j + 1], i1, 0x1054), __byte_perm(S[j + 1], i1, 0x3276));
            key = __vimin3_u32(key, k0, k1);
        }
    }
    key = (key + (uint32_t)(lg * 2 * NREG)) | padMask;                  // lane-local index -> disparity; padding lanes drop out
#pragma unroll
    for (int off = LPC / 2; off >= 1; off >>= 1) key = min(key, __shfl_xor_sync(0xFFFFFFFFu, key, off, LPC));
    const int minS = min((int)(key >> 16), 32767);
    const int best = (minS == 32767) ? -1 : (int)(key & 0xFFFFu);       // first minimum (A.5)
    int Sm = 0, Sp = 0;
    const bool interior = best > 0 && best < g.D - 1;
    if (lg == 0 && interior) {
        Sm = min((int)lds16(colA + 2u * (uint32_t)sgbm_pos(best - 1, NREG, LPC)), 32767);
        Sp = min((int)lds16(colA + 2u * (uint32_t)sgbm_pos(best + 1, NREG, LPC)), 32767);
    }
    bool reject = false;
    if (g.UR > 0) {
        const int av = 100 - g.UR;                                       // S(d)*(100-UR) < minS*100  <=>  S(d) < T
        const unsigned num = (unsigned)(100 * minS + av - 1);
        const int T = av > 0 ? min((int)(av == 1 ? num : __umulhi(num, a.urMagic)), 32768) : (minS > 0 ? 32768 : 0);
        __syncwarp();
        if (lg == 0 && own) {
#pragma unroll
            for (int dd = -1; dd <= 1; dd++) {
                const int d = best + dd;
                if (d >= 0 && d < g.D) sts16(colA + 2u * (uint32_t)sgbm_pos(d, NREG, LPC), 0xFFFFu);
            }
        }
        __syncwarp();
        uint32_t S2[NREG];
        lds_vec<NREG, LPC>(S2, laneA);
        const uint32_t t2 = local_min<NREG>(S2) | padMask;
        const bool viol = padMask == 0u && min((int)(t2 & 0xFFFFu), 32767) < T;
        reject = (__ballot_sync(0xFFFFFFFFu, viol) & gmask) != 0u;
    }
    if (lg == 0 && own) {
        const int x = x1 + g.minX1;
        int out = g.INV;
        if (!reject) {
            const int x2 = x - best - g.minD;
            if (minS < 32767 && x2 >= 0 && x2 < g.W)
                atomicMin(a.d2key + (size_t)y * g.W + x2, ((unsigned)minS << 16) | (0xFFFFu - (unsigned)x1));
            int dq = best * 16;
            if (interior) {
                const int den = max(Sm + Sp - 2 * minS, 1);
                dq += (int)((float)((Sm - Sp) * 16 + den) / (float)(2 * den));   // exactness: see sweep_wta
            }
            out = dq + g.minD * 16;
        }
        a.raw[(size_t)y * g.W + x] = (int16_t)out;
    }
}

template <int NREG, int LPC, bool SAT, int RPS>
__device__ __forceinline__ void sweep_role_w(const SweepArgs &a, const SweepSmem &s, int wwarp, int xs, int SW, int yBegin,
                                             int yStep, int nRows)
{
    constexpr int GPW = 32 / LPC;
    const int lane = threadIdx.x & 31, lg = lane % LPC;
    const int K = a.K;
    const uint32_t colB = a.colB;
    const uint32_t padMask = sm_keep(lg > a.g.lanesUsed - 1 ? 0xFFFFFFFFu : 0u);
    const unsigned gmask = sm_keep((LPC == 32 ? 0xFFFFFFFFu : ((1u << LPC) - 1u)) << (lane & ~(LPC - 1)));
    const uint32_t lane0 = sm_keep(lane == 0 ? 1u : 0u);
    // The winner-take-all of a pixel is one long dependent chain (two reductions, a masked re-scan, a division):
    // ~1500 cycles per pass whatever the strip width.  With narrow strips that latency, not throughput, would set
    // the row period of the whole kernel, so the W warps are split into wRG row groups that take alternate rows.
    const int wRG = a.wRG, wPR = a.wPR;
    const int rgrp = wwarp / wPR, wq = wwarp - rgrp * wPR;
    constexpr int rps = RPS;
    RingPos rp = ring_start(s.aP, s.barP, K);
    if (rgrp) ring_advance_n(rp, rgrp, a.pStrideB, a.pSpanB, 32u, a.pBarSpan, K);
    for (int t = rgrp * rps; t < nRows; t += wRG * rps) {
        const int rows = min(rps, nRows - t);
        SWEEP_PROG(5, wwarp == 0);
        SWEEP_TR(3, 4, wwarp == 0);
        sweep_wait(a, SmemBar{rp.bar + BAR_FULLW}, rp.par, 12, t);
        SWEEP_TR(3, 5, wwarp == 0);
        uint32_t dp = rp.data;
        for (int r = 0; r < rows; r++) {
            const int y = yBegin + (t + r) * yStep;
            for (int it = 0; it < a.wPass; it++) {
                const int gi = (it * wPR + wq) * GPW + lane / LPC;
                const bool own = gi < SW;
                if (!__any_sync(0xFFFFFFFFu, own)) continue;
                const int ci = own ? gi : SW - 1;
                // (groups without a column read column SW-1 along with the warp; all their writes are guarded by own)
                sweep_wta_slot<NREG, LPC, SAT>(a, dp + (uint32_t)ci * colB, lg, padMask, gmask, own, xs + ci, y);
            }
            dp += a.pRowB;
        }
        __syncwarp();
        SWEEP_TR(3, 6, wwarp == 0);
        if (lane0) bar_arrive(rp.bar + BAR_FREEP);
        ring_advance_n(rp, wRG, a.pStrideB, a.pSpanB, 32u, a.pBarSpan, K);
    }
}

template <int NREG, bool WROLE> struct SweepMaxThreads { static const int value = (NREG >= 12 && !WROLE) ? 768 : 1024; };

template <int NREG, int LPC, bool SAT, bool WROLE, int RPS>
__global__ void __launch_bounds__((SweepMaxThreads<NREG, WROLE>::value), 1) k_sweep(SweepArgs a)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const Geo &g = a.g;
    const int strip = blockIdx.x;
    const int xs = (int)((long long)strip * g.W1 / a.nstrips);
    const int xe = (int)((long long)(strip + 1) * g.W1 / a.nstrips);
    const int nRows = g.H, yBegin = a.backward ? g.H - 1 : 0, yStep = a.backward ? -1 : 1;
    const SweepSmem s = sweep_carve(a, smem);
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        const int nCons = a.nwV + 2 * a.nwA;
        for (int q = 0; q < a.NSC; q++) { mbar_init(SmemBar{s.barC + 16u * q + BAR_FULL}, 1); mbar_init(SmemBar{s.barC + 16u * q + BAR_EMPTY}, nCons); }
        for (int q = 0; q < a.NSI; q++) { mbar_init(SmemBar{s.barI + 16u * q + BAR_FULL}, 1); mbar_init(SmemBar{s.barI + 16u * q + BAR_EMPTY}, a.nwV); }
        for (int q = 0; q < a.K; q++) {
            mbar_init(SmemBar{s.barP + 32u * q + BAR_FULLV}, a.nwV); mbar_init(SmemBar{s.barP + 32u * q + BAR_FULLM}, a.nwA);
            mbar_init(SmemBar{s.barP + 32u * q + BAR_FREEP}, WROLE ? a.wPR : a.nwA); mbar_init(SmemBar{s.barP + 32u * q + BAR_FULLW}, a.nwA);
        }
        mbar_fence_init();
    }
    __syncthreads();
    if (WROLE) {
        // Warp layout [V: aV][A: aA][C: aA][W: nwW, producer, idle: 8 in all], every role a whole number of
        // warpgroups so that registers can move between them: the three path roles need ~80 registers at
        // NREG >= 12, the winner-take-all and producer warps give theirs up (1024 threads x 64 at launch).
        constexpr bool SPLIT = NREG >= 12;
        if (warp < a.aV + 2 * a.aA) {
            if (SPLIT) asm volatile("setmaxnreg.inc.sync.aligned.u32 72;");
            if (warp < a.aV) {
                if (warp < a.nwV) sweep_role_v<NREG, LPC, SAT, RPS>(a, s, warp, xe - xs, nRows);
            } else if (warp < a.aV + a.aA) {
                if (warp - a.aV < a.nwA) sweep_role_diag<NREG, LPC, +1, SAT, true, RPS>(a, s, warp - a.aV, strip, xs, xe, yBegin, yStep, nRows);
            } else {
                if (warp - a.aV - a.aA < a.nwA) sweep_role_diag<NREG, LPC, -1, SAT, true, RPS>(a, s, warp - a.aV - a.aA, strip, xs, xe, yBegin, yStep, nRows);
            }
        } else {
            if (SPLIT) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
            const int ww = warp - a.aV - 2 * a.aA;
            if (ww < a.nwW) sweep_role_w<NREG, LPC, SAT, RPS>(a, s, ww, xs, xe - xs, yBegin, yStep, nRows);
            else if (ww == a.nwW && (threadIdx.x & 31) == 0) sweep_producer_t<RPS, false>(a, s, xs, xe, yBegin, yStep, nRows);
        }
        return;
    }
    if (warp < a.nwV) {
        sweep_role_v<NREG, LPC, SAT, RPS>(a, s, warp, xe - xs, nRows);
    } else if (warp < a.nwV + a.nwA) {
        sweep_role_diag<NREG, LPC, +1, SAT, false, RPS>(a, s, warp - a.nwV, strip, xs, xe, yBegin, yStep, nRows);
    } else if (warp < a.nwV + 2 * a.nwA) {
        sweep_role_diag<NREG, LPC, -1, SAT, false, RPS>(a, s, warp - a.nwV - a.nwA, strip, xs, xe, yBegin, yStep, nRows);
    } else if ((threadIdx.x & 31) == 0) {
        sweep_producer_t<RPS, true>(a, s, xs, xe, yBegin, yStep, nRows);
    }
}

// =================================================================================================
// Host side
// =================================================================================================
static size_t sweep_layout(SweepArgs &a, int groupsC, bool scratch)
{
    const Geo &g = a.g;
    const size_t col = (size_t)g.Dp * 2;
    const size_t rps = (size_t)a.rps;
    size_t off = 0;
    a.stgCOff = (unsigned)off; off += (size_t)a.NSC * rps * (a.SW + 2 * (a.R - 1)) * col;
    a.stgIOff = (unsigned)off; off += (size_t)a.NSI * rps * a.nAB * a.SW * col;
    a.pOff = (unsigned)off; off += (size_t)a.K * rps * a.SW * col;
    a.ssmOff = (unsigned)off; if (scratch) off += (size_t)groupsC * col;
    off = (off + 15) & ~(size_t)15;
    a.barOff = (unsigned)off; off += (size_t)(2 * a.NSC + 2 * a.NSI + 4 * a.K) * 8;
    a.colB = (unsigned)col;
    a.cRowB = (unsigned)((a.SW + 2 * (a.R - 1)) * col);
    a.cStrideB = a.cRowB * (unsigned)rps; a.cSpanB = a.cStrideB * (unsigned)a.NSC; a.cBarSpan = 16u * (unsigned)a.NSC;
    a.iBB = (unsigned)(a.SW * col); a.iRowB = a.iBB * (unsigned)a.nAB;
    a.iStrideB = a.iRowB * (unsigned)rps; a.iSpanB = a.iStrideB * (unsigned)a.NSI; a.iBarSpan = 16u * (unsigned)a.NSI;
    a.pRowB = a.iBB;
    a.pStrideB = a.pRowB * (unsigned)rps; a.pSpanB = a.pStrideB * (unsigned)a.K; a.pBarSpan = 32u * (unsigned)a.K;
    return off;
}

// Strip / warp / ring geometry of one launch (a.nAB must be set).  False when the geometry does not fit.
static bool sweep_plan(const Geo &g, int numSMs, bool wrole, bool wta, int maxThreads, int maxSmem, SweepArgs &a, int *threadsOut,
                       size_t *smemOut)
{
    const int GPW = 32 / g.lpc;
    // Rows per super-step (1..8 are covered by tests/test_gpu_parity.py::test_sweep_rows_per_superstep and a repeatability
    // run, 12 / 16 by test_sweep_schedule_knobs).  8 is the fastest with 4 or more lanes per column; with two lanes per
    // column (numDisparities <= 32) a warp holds 16 column groups, the halo chains of 16 rows per super-step cost no extra
    // warps, and half as many strip hand-offs win (4K D=16: 1.37 -> 1.18 ms).
    const int Rmin = 1;
    const SgbmKnobs &kn = sgbm_knobs();
    int R = g.lpc <= 2 ? 16 : 8;
    if (kn.vr > 0) R = kn.vr >= Rmin ? kn.vr : Rmin;
    if (R > 16) R = 16;
    int Kwant = 5, NSCwant = 5, NSIwant = 3;
    if (kn.sweepK >= 1 && kn.sweepK <= 16) Kwant = kn.sweepK;
    if (kn.sweepNSC >= 2 && kn.sweepNSC <= 16) NSCwant = kn.sweepNSC;
    if (kn.sweepNSI >= 2 && kn.sweepNSI <= 16) NSIwant = kn.sweepNSI;
    for (; R >= Rmin; R--) {
        int nstrips = numSMs;
        const int minCols = R > 2 ? R : 2;                // every strip owns >= R (and >= 2) columns
        if (nstrips > g.W1 / minCols) nstrips = g.W1 / minCols;
        if (nstrips < 1) nstrips = 1;
        if (nstrips == 1 && R > 1) continue;              // a single strip has no halos
        const int SWmax = (g.W1 + nstrips - 1) / nstrips;
        const int NB = (SWmax + R - 1 + R - 1) / R;       // ceil((SW + HG) / R)
        const int groupsA = NB * R;
        a.nwA = (groupsA + GPW - 1) / GPW;
        a.nwV = (SWmax + GPW - 1) / GPW;
        int threads;
        if (wrole) {
            a.aA = (a.nwA + 3) & ~3; a.aV = (a.nwV + 3) & ~3;
            // 8 warps in the last warpgroup: WTA warps, the producer, idle.  wPR warps share a row, the rest of
            // the seven form further row groups (fixed below, once the S-ring depth is known)
            a.wPR = a.nwV < 7 ? a.nwV : 7;
            if (kn.sweepNWW >= 1 && kn.sweepNWW <= 7) a.wPR = kn.sweepNWW;
            a.wPass = (SWmax + a.wPR * GPW - 1) / (a.wPR * GPW);
            a.wRG = 1; a.nwW = a.wPR;
            threads = (a.aV + 2 * a.aA + 8) * 32;
        } else {
            threads = (a.nwV + 2 * a.nwA + 1) * 32;
        }
        if (threads > maxThreads) continue;
        a.SW = SWmax; a.nstrips = nstrips; a.R = R; a.NB = NB;
        int wRGwant = wrole ? 7 / a.wPR : 1;              // row groups of the WTA warps: each holds one S slot while it works
        if (kn.sweepWRG >= 1 && kn.sweepWRG < wRGwant) wRGwant = kn.sweepWRG;
        // Rows per ring stage.  Two rows per stage halve the hand-off work per image row (waits, arrives, ring advances:
        // about half of a role's instructions when a lane holds few registers) at the price of coarser rings; taken when the
        // rings keep their depth in ROWS (three cost stages, three S stages for the path roles, two input stages) --
        // otherwise (4K at numDisparities = 256: shared memory is full with one row per stage) one row per stage.
        const int rpsFirst = (kn.sweepRPS == 1 || (R & 1) || g.nreg > SWEEP_RPS_MAX_NREG) ? 1 : 2;
        for (int rps = rpsFirst; rps >= 1; rps--) {
            a.rps = rps;
            if (rps == 2) {
                // Measured (720p D=128, 1080p D=192, 4K D=16; sweep time against one row per stage): what matters is the depth
                // of the cost ring -- 3 stages: no gain, 4: -8 %, 5: -10 %, 8: -11 .. -12 %; two input stages are enough,
                // and the S ring wants two or three slots beside those the WTA row groups hold.
                const int K2 = kn.sweepK ? Kwant : 4 + wRGwant;
                static const int tries2[][2] = {{0, 8}, {0, 6}, {0, 5}, {-1, 5}, {-1, 4}, {-2, 4}};
                bool ok = false;
                for (const auto &tr : tries2) {
                    a.K = K2 + (kn.sweepK ? 0 : tr[0]); a.NSC = kn.sweepNSC ? NSCwant : tr[1]; a.NSI = kn.sweepNSI ? NSIwant : 2;
                    if (a.K < (wrole ? 3 : 2)) a.K = wrole ? 3 : 2;
                    if (wrole) { a.wRG = wRGwant < a.K - 2 ? wRGwant : (a.K - 2 > 1 ? a.K - 2 : 1); a.nwW = a.wPR * a.wRG; }
                    const size_t smem = sweep_layout(a, a.nwA * GPW, wta && !wrole);
                    if (smem <= (size_t)maxSmem) { *threadsOut = threads; *smemOut = smem; ok = true; break; }
                }
                if (ok) return true;
                if (kn.sweepRPS == 2) {                   // forced (tests): shrink the rings as far as the protocol allows
                    a.K = wrole ? 3 : 2; a.NSC = 2; a.NSI = 2;
                    if (wrole) { a.wRG = 1; a.nwW = a.wPR; }
                    const size_t smem2 = sweep_layout(a, a.nwA * GPW, wta && !wrole);
                    if (smem2 <= (size_t)maxSmem) { *threadsOut = threads; *smemOut = smem2; return true; }
                }
                continue;
            }
            // ring depths: shrink until the layout fits
            static const int tries[][3] = {{0, 0, 0}, {0, 0, -1}, {0, -1, -1}, {-1, -1, -1}, {-1, -2, -1}, {-2, -2, -1}, {-2, -3, -1}};
            for (const auto &tr : tries) {
                a.K = Kwant + (kn.sweepK ? 0 : wRGwant - 1) + tr[0]; a.NSC = NSCwant + tr[1]; a.NSI = NSIwant + tr[2];
                if (a.K < 1) a.K = 1;
                if (a.NSC < 2) a.NSC = 2;
                if (a.NSI < 2) a.NSI = 2;
                if (wrole) {                                  // the path roles keep at least two slots to themselves
                    a.wRG = wRGwant < a.K - 2 ? wRGwant : (a.K - 2 > 1 ? a.K - 2 : 1);
                    a.nwW = a.wPR * a.wRG;
                }
                const size_t smem = sweep_layout(a, a.nwA * GPW, wta && !wrole);
                if (smem <= (size_t)maxSmem) { *threadsOut = threads; *smemOut = smem; return true; }
            }
        }
    }
    return false;
}

// Test hook (sgbm_debug_sweep_plan): the schedule the planner picks for a geometry, without a device.
// out[16] = { found, nstrips, SW, R, NB, rps, K, NSC, NSI, nwV, nwA, nwW, wRG, wPR, threads, smem bytes }
bool sgbm_sweep_plan_debug(const Geo &g, int numSMs, int maxSmem, int wrole, int nAB, int *out)
{
    SweepArgs a;
    memset(&a, 0, sizeof(a));
    a.g = g; a.nAB = nAB;
    int threads = 0;
    size_t smem = 0;
    const int maxThreads = wrole ? 1024 : (g.nreg >= 12 ? 768 : 1024);
    const bool found = sweep_plan(g, numSMs, wrole != 0, true, maxThreads, maxSmem, a, &threads, &smem);
    const int v[16] = {found ? 1 : 0, a.nstrips, a.SW, a.R, a.NB, a.rps, a.K, a.NSC, a.NSI, a.nwV, a.nwA, a.nwW, a.wRG, a.wPR, threads, (int)smem};
    for (int i = 0; i < 16; i++) out[i] = v[i];
    return found;
}

// Whether the persistent sweeps of mode SGBM (0) / HH (1) hold this geometry on `numSMs` SMs: what the
// batch entry points ask before they run two frames side by side, each on half of the GPU.
bool sgbm_sweep_fits(const Geo &g, int numSMs, int mode)
{
    const int maxSmem = sgbm_knobs().maxSmemOptin;
    SweepArgs a;
    int threads = 0;
    size_t smem = 0;
    memset(&a, 0, sizeof(a));
    a.g = g;
    a.nAB = mode == 1 ? 1 : 2;                            // last sweep: S_fwd (HH) or L_hA + L_hB (SGBM)
    if (!sweep_plan(g, numSMs, true, true, 1024, maxSmem, a, &threads, &smem) || a.nstrips > numSMs) return false;
    if (mode == 1) {                                      // forward sweep of MODE_HH
        memset(&a, 0, sizeof(a));
        a.g = g; a.nAB = 2;
        if (!sweep_plan(g, numSMs, false, false, g.nreg >= 12 ? 768 : 1024, maxSmem, a, &threads, &smem)) return false;
    }
    return true;
}

// Launch of one planned sweep with the kernel instantiation for its rows per ring stage.
template <int NREG, int LPC, bool SAT, bool WROLE, int RPS>
static int launch_sweep_k(SweepArgs &a, const VertArgs &va, int numSMs, int threads, size_t smem, bool wta, cudaStream_t st)
{
    const Geo &g = va.g;
    auto kern = k_sweep<NREG, LPC, SAT, WROLE, RPS>;
    const SgbmKnobs &kn = sgbm_knobs();
    static unsigned long long attrDone = 0;   // one bit per device: function attributes are per device
    {
        SgbmDeviceOnce once(attrDone);
        if (once.first) {
            SGBM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kn.maxSmemOptin));
            once.done();
        }
    }
    int occ = 0;
    SGBM_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
    if (occ * numSMs < a.nstrips) return 1;
    SGBM_CUDA_CHECK(cudaMemsetAsync(a.flagA, 0, sizeof(unsigned int) * 2 * (size_t)a.nstrips * 64, st));
    a.flagC = a.flagA + (size_t)a.nstrips * 64;
    a.dbg = va.watchDev;                                  // sticky: zeroed with the workspace and after a report
    const char *tracePath = kn.tracePath[0] ? kn.tracePath : nullptr;   // debug: dump one strip's time stamps to a file
#ifndef SGBM_SWEEP_TRACING
    if (tracePath) { fprintf(stderr, "SGBM_SWEEP_TRACE needs a build with make TRACE=1\n"); tracePath = nullptr; }
#endif
    const size_t traceBytes = (size_t)g.H * 4 * 8 * sizeof(unsigned long long);
    if (tracePath) {
        SGBM_CUDA_CHECK(cudaMalloc(&a.trace, traceBytes));
        SGBM_CUDA_CHECK(cudaMemsetAsync(a.trace, 0, traceBytes, st));
        a.traceStrip = a.nstrips / 2;
    }
    if (kn.verbose)
        fprintf(stderr, "sweep: wrole=%d wta=%d strips=%d SW=%d R=%d NB=%d nwV=%d nwA=%d nwW=%d (x%d row groups) wPass=%d rows/stage=%d K=%d NSC=%d NSI=%d threads=%d smem=%zu\n",
                (int)WROLE, (int)wta, a.nstrips, a.SW, a.R, a.NB, a.nwV, a.nwA, a.nwW, a.wRG, a.wPass, a.rps, a.K, a.NSC, a.NSI, threads, smem);
    void *args[] = {&a};
    SGBM_CUDA_CHECK(cudaLaunchCooperativeKernel((void *)kern, dim3(a.nstrips), dim3(threads), args, smem, st));
    sgbm_count_launch(1);
    if (va.watchHost)                                         // the caller checks it at its next synchronisation point
        SGBM_CUDA_CHECK(cudaMemcpyAsync(va.watchHost, va.watchDev, 8 * sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
    if (tracePath) {
        SGBM_CUDA_CHECK(cudaStreamSynchronize(st));
        void *hbuf = malloc(traceBytes);
        SGBM_CUDA_CHECK(cudaMemcpy(hbuf, a.trace, traceBytes, cudaMemcpyDeviceToHost));
        if (FILE *f = fopen(tracePath, "wb")) { fwrite(hbuf, 1, traceBytes, f); fclose(f); }
        free(hbuf);
        cudaFree(a.trace);
        fprintf(stderr, "sweep trace: H=%d strips=%d R=%d NB=%d nwV=%d nwA=%d K=%d NSC=%d NSI=%d rows/stage=%d threads=%d smem=%zu\n", g.H,
                a.nstrips, a.R, a.NB, a.nwV, a.nwA, a.K, a.NSC, a.NSI, a.rps, threads, smem);
    }
    return 0;
}

// Returns 0 on success, 1 if this geometry does not fit the role-specialised sweep (the caller falls
// back to k_vertical), negative on error.  WROLE: winner-take-all sweep with the dedicated WTA role.
template <int NREG, int LPC, bool SAT, bool WROLE>
static int launch_sweep_t(const VertArgs &va, int numSMs, cudaStream_t st)
{
    const Geo &g = va.g;
    const SgbmKnobs &kn = sgbm_knobs();
    SweepArgs a;
    memset(&a, 0, sizeof(a));
    a.g = g; a.C = va.C; a.inA = va.inA; a.inB = va.inB; a.sout = va.sout; a.sdbg = va.sdbg; a.raw = va.raw;
    a.d2key = va.d2key; a.backward = va.backward; a.haloA = va.haloA; a.haloC = va.haloC; a.flagA = va.flagA;
    a.flagC = va.flagC; a.dbgNoSync = va.dbgNoSync;
    a.dbgStall = SGBM_DBG_HOOK(kn.dbgStall);
    a.nAB = va.inB ? 2 : 1;
    a.urMagic = g.UR < 99 ? 0xFFFFFFFFu / (unsigned)(100 - g.UR) + 1u : 0u;
    a.pfDist = kn.sweepPF >= 0 ? kn.sweepPF : 8;
    a.P1p = (unsigned)g.P1 * 0x10001u; a.P2mP1p = (unsigned)(g.P2 - g.P1) * 0x10001u;
    a.wtaSlow = sgbm_pad_from(g) < 2 * g.nreg ? 1 : 0;
    const bool wta = va.sout == nullptr;
    if (WROLE && !wta) return 1;
    const int maxThreads = SweepMaxThreads<NREG, WROLE>::value;
    int threads = 0;
    size_t smem = 0;
    const bool found = sweep_plan(g, numSMs, WROLE, wta, maxThreads, kn.maxSmemOptin, a, &threads, &smem);
    if (!found) return 1;
    if constexpr (NREG <= SWEEP_RPS_MAX_NREG) {
        if (a.rps == 2) return launch_sweep_k<NREG, LPC, SAT, WROLE, 2>(a, va, numSMs, threads, smem, wta, st);
    }
    return launch_sweep_k<NREG, LPC, SAT, WROLE, 1>(a, va, numSMs, threads, smem, wta, st);
}

// the winner-take-all sweep first tries the kernel with the dedicated WTA role (SGBM_SWEEP_W=0 disables it)
template <int NREG, int LPC, bool SAT>
static int launch_sweep_any(const VertArgs &va, int numSMs, cudaStream_t st)
{
    if (va.sout == nullptr) {
        // numDisparities % 8 != 0 stays on the kernels whose winner-take-all masks the padding disparities in registers
        // (sweep_wta): the W warps of the 16-register mappings run under a 40-register cap, and ANY code added to their pixel
        // loop -- even an out-of-line call behind a uniform branch -- showed up as spills (cfg3 WTA sweep 3.17 -> 3.37 ms)
        if (sgbm_knobs().sweepW && sgbm_pad_from(va.g) >= 2 * va.g.nreg) {
            const int rc = launch_sweep_t<NREG, LPC, SAT, true>(va, numSMs, st);
            if (rc <= 0) return rc;
        }
    }
    return launch_sweep_t<NREG, LPC, SAT, false>(va, numSMs, st);
}

// =================================================================================================
// Row-at-a-time fallback: any width.  One launch per image row, one lane group per column; the state of
// the three paths of the previous row lives in global memory ([2][3][W1][Dp + 8], ping-pong, L2 resident).
// No device-side synchronisation at all (the launches order the rows), ~4 us per row: the last resort for
// images whose strips do not fit the persistent kernels (e.g. 7680 wide with numDisparities = 256), and
// an independent implementation for the tests (SGBM_ROWSTEP=1).
// =================================================================================================
struct RowArgs {
    SweepArgs sw;            // g, C, inA, inB, sout, sdbg, raw, d2key, urMagic are used
    uint16_t *state;
    int y, first, cur;
};

template <int NREG, int LPC>
__global__ void __launch_bounds__(128) k_rowstep(RowArgs a)
{
    extern __shared__ __align__(16) uint8_t smem[];
    constexpr int GPW = 32 / LPC;
    const SweepArgs &sa = a.sw;
    const Geo &g = sa.g;
    const int lane = threadIdx.x & 31, lg = lane % LPC;
    const int gIn = (threadIdx.x >> 5) * GPW + lane / LPC;            // group within the CTA
    const int xRaw = blockIdx.x * (blockDim.x / LPC) + gIn;
    const bool own = xRaw < g.W1;
    const int x = own ? xRaw : g.W1 - 1;
    const int Dp = g.Dp, lastLane = g.lanesUsed - 1;
    const size_t sst = (size_t)Dp + 8, plane = (size_t)g.W1 * sst;
    const uint16_t *prev = a.state + (size_t)(a.cur ^ 1) * 3 * plane;
    uint16_t *cur = a.state + (size_t)a.cur * 3 * plane;
    const uint32_t P1p = (uint32_t)g.P1 * 0x10001u, P2mP1p = (uint32_t)(g.P2 - g.P1) * 0x10001u;
    const size_t off = (size_t)a.y * g.rowStride + (size_t)x * Dp;
    uint32_t Cc[NREG], S[NREG];
    load_vec_nc<NREG, LPC>(Cc, sa.C + off, lg);
    load_vec_nc<NREG, LPC>(S, sa.inA + off, lg);
    if (sa.inB) {
        uint32_t Bv[NREG];
        load_vec_nc<NREG, LPC>(Bv, sa.inB + off, lg);
#pragma unroll
        for (int j = 0; j < NREG; j++) S[j] = paddmin(S[j], Bv[j], SGBM_MAX_S);
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {                         // k = 0: (0,-1), 1: (-1,-1), 2: (+1,-1)
        const int xp = x + (k == 1 ? -1 : (k == 2 ? 1 : 0));
        uint32_t L[NREG], m = 0;
#pragma unroll
        for (int j = 0; j < NREG; j++) L[j] = 0;          // "predecessor outside" == L = 0, m = 0 (A.4)
        if (!a.first && xp >= 0 && xp < g.W1) {
            const uint16_t *p = prev + (size_t)k * plane + (size_t)xp * sst;
            load_vec_l2<NREG, LPC>(L, p, lg);
            m = __ldcg(reinterpret_cast<const unsigned int *>(p + Dp));
        }
        m = path_step<NREG, LPC>(L, L, m, Cc, P1p, P2mP1p, lg, lastLane);
        if (own) {
            uint16_t *q = cur + (size_t)k * plane + (size_t)x * sst;
            store_vec<NREG, LPC>(L, q, lg);
            if (lg == 0) *reinterpret_cast<unsigned int *>(q + Dp) = m;
        }
#pragma unroll
        for (int j = 0; j < NREG; j++) S[j] = paddmin(S[j], L[j], SGBM_MAX_S);
    }
    if (sa.sout) {
        if (own) store_vec<NREG, LPC>(S, sa.sout + off, lg);
    } else {
        if (sa.sdbg && own) store_vec<NREG, LPC>(S, sa.sdbg + off, lg);
        sweep_wta<NREG, LPC>(sa, S, reinterpret_cast<uint16_t *>(smem) + (size_t)gIn * Dp, lg, own, x, a.y);
    }
}

template <int NREG, int LPC>
static int launch_rowstep_t(const VertArgs &va, cudaStream_t st)
{
    const Geo &g = va.g;
    RowArgs a;
    memset(&a, 0, sizeof(a));
    a.sw.g = g; a.sw.C = va.C; a.sw.inA = va.inA; a.sw.inB = va.inB; a.sw.sout = va.sout; a.sw.sdbg = va.sdbg;
    a.sw.raw = va.raw; a.sw.d2key = va.d2key;
    a.sw.urMagic = g.UR < 99 ? 0xFFFFFFFFu / (unsigned)(100 - g.UR) + 1u : 0u;
    a.sw.wtaSlow = sgbm_pad_from(g) < 2 * g.nreg ? 1 : 0;
    a.state = va.rowState;
    if (!a.state) return sgbm_fail(-3, "row-step fallback has no state buffer");
    const int threads = 128, groups = threads / LPC;
    const size_t smem = (size_t)groups * g.Dp * 2;
    const int grid = (g.W1 + groups - 1) / groups;
    for (int t = 0; t < g.H; t++) {
        a.y = va.backward ? g.H - 1 - t : t;
        a.first = t == 0;
        a.cur = t & 1;
        k_rowstep<NREG, LPC><<<grid, threads, smem, st>>>(a);
    }
    sgbm_count_launch(g.H);
    SGBM_CUDA_CHECK(cudaGetLastError());
    return 0;
}

#define ROWSTEP_DISPATCH(NREG_, LPC_) if (g.nreg == NREG_ && g.lpc == LPC_) return launch_rowstep_t<NREG_, LPC_>(a, st);
int sgbm_launch_rowstep(const VertArgs &a, cudaStream_t st)
{
    const Geo &g = a.g;
#ifdef SGBM_FAST_BUILD      // development builds: only the lane mappings of the BASELINE configurations (make FAST=1)
    ROWSTEP_DISPATCH(4, 2) ROWSTEP_DISPATCH(8, 8) ROWSTEP_DISPATCH(12, 8) ROWSTEP_DISPATCH(16, 8)
#else
    ROWSTEP_DISPATCH(4, 2) ROWSTEP_DISPATCH(4, 4) ROWSTEP_DISPATCH(4, 8) ROWSTEP_DISPATCH(4, 16) ROWSTEP_DISPATCH(4, 32)
    ROWSTEP_DISPATCH(8, 2) ROWSTEP_DISPATCH(8, 4) ROWSTEP_DISPATCH(8, 8) ROWSTEP_DISPATCH(8, 16) ROWSTEP_DISPATCH(8, 32)
    ROWSTEP_DISPATCH(12, 2) ROWSTEP_DISPATCH(12, 4) ROWSTEP_DISPATCH(12, 8) ROWSTEP_DISPATCH(12, 16) ROWSTEP_DISPATCH(12, 32)
    ROWSTEP_DISPATCH(16, 2) ROWSTEP_DISPATCH(16, 4) ROWSTEP_DISPATCH(16, 8) ROWSTEP_DISPATCH(16, 16) ROWSTEP_DISPATCH(16, 32)
#endif
    return sgbm_fail(-3, "no kernel for lane mapping nreg=%d lpc=%d", g.nreg, g.lpc);
}

#define SWEEP_DISPATCH(NREG_, LPC_)                                                               \
    if (g.nreg == NREG_ && g.lpc == LPC_)                                                         \
        return sat ? launch_sweep_any<NREG_, LPC_, true>(a, numSMs, st) : launch_sweep_any<NREG_, LPC_, false>(a, numSMs, st);

int sgbm_launch_sweep(const VertArgs &a, int numSMs, cudaStream_t st)
{
    const Geo &g = a.g;
    // Largest value the sum over all paths of the mode can reach (A.2-A.4): every path cost is
    // <= C + P2 and C <= cn * (2r+1)^2 * (max BT cost of the gradient plane + (255 >> 2)).
    const long long pixMax = (long long)(2 * g.ftzero < 255 ? 2 * g.ftzero : 255) + 63;
    const long long cMax = (long long)g.cn * (2 * g.r + 1) * (2 * g.r + 1) * pixMax;
    const int npaths = g.mode == 1 ? 8 : 5;
    bool sat = npaths * (cMax + g.P2) > 65535;
    sat = sat || sgbm_knobs().sweepSat != 0;
#ifdef SGBM_FAST_BUILD
    SWEEP_DISPATCH(4, 2) SWEEP_DISPATCH(8, 8) SWEEP_DISPATCH(12, 8) SWEEP_DISPATCH(16, 8)
#else
    SWEEP_DISPATCH(4, 2) SWEEP_DISPATCH(4, 4) SWEEP_DISPATCH(4, 8) SWEEP_DISPATCH(4, 16) SWEEP_DISPATCH(4, 32)
    SWEEP_DISPATCH(8, 2) SWEEP_DISPATCH(8, 4) SWEEP_DISPATCH(8, 8) SWEEP_DISPATCH(8, 16) SWEEP_DISPATCH(8, 32)
    SWEEP_DISPATCH(12, 2) SWEEP_DISPATCH(12, 4) SWEEP_DISPATCH(12, 8) SWEEP_DISPATCH(12, 16) SWEEP_DISPATCH(12, 32)
    SWEEP_DISPATCH(16, 2) SWEEP_DISPATCH(16, 4) SWEEP_DISPATCH(16, 8) SWEEP_DISPATCH(16, 16) SWEEP_DISPATCH(16, 32)
#endif
    return 1;
}
