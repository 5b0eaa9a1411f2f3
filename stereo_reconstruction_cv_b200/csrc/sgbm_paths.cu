// sgbm_paths.cu -- stages 3+4 of StereoSGBM.compute (main.ipynb:668) for sm_100a:
//   k_horizontal : the two horizontal SGM paths (-1,0) and (+1,0); one lane group per image row
//                  walks the row, state in registers, min over disparities by warp shuffles.
//   k_vertical   : the vertical + two diagonal paths of one sweep (top-down or bottom-up), all
//                  columns in parallel, one persistent CTA per column strip; diagonal state is
//                  exchanged through shared memory inside a strip and through L2-resident halo
//                  buffers + release/acquire flags between neighbouring strips.  The sweep either
//                  spills S once (MODE_HH forward sweep) or runs the fused winner-take-all:
//                  uniqueness, sub-pixel fit, disp2 splat (A.5 / A.6).
// Recurrence and saturation rules: SURVEY.md Appendix A.4; packed u16x2 DPX arithmetic
// (VIMNMX3 / VIADDMNMX) as described in sgbm_common.cuh.
#include "sgbm_common.cuh"
#include <stdlib.h>

// =================================================================================================
// Horizontal paths
// =================================================================================================
#define HZ_THREADS 128
// columns per staging chunk: ~8 KB per warp and stage (4 columns at 8 lanes x 16 registers, 16 at 2 x 4)
template <int NREG, int LPC> struct HzK {
    static const int raw = 4096 / ((32 / LPC) * 2 * NREG * LPC);
    static const int value = (raw < 4 ? 4 : (raw > 16 ? 16 : raw)) & ~1;   // even: the column loop is unrolled in pairs
};
#define HZ_NS 3                      // chunks in flight per warp

// One lane group per image row and direction.  The cost rows are streamed through a per-warp ring
// of HZ_NS shared-memory stages filled by TMA bulk copies (one contiguous K-column segment per
// row), so HBM latency is covered by data in flight instead of registers.
// (Measured and rejected, end of round 2: the same kernel without shared memory -- the next three columns' cost vectors
// prefetched into registers by plain 128-bit loads, 96 registers, so that one of its CTAs fits on an SM BESIDE a 211 KB cost
// CTA.  Alone it is as fast as this one at 4K D=256 (2.57 ms) and slower on small frames (1080p 0.68 vs 0.46 ms); run beside
// the cost kernel of the next row band it made cost + horizontal 4.9 / 7.8 / 5.2 ms with 2 / 4 / 8 bands against 4.2 ms, with
// low stream priority 6.9 / 7.8 / 6.1 ms: sharing the SM costs the cost kernel far more than the overlap returns.)
// At 4K / D = 256 the kernel is HBM-bound (92 % of the copy peak); with few rows or few disparities it is a pure
// latency chain of W1 dependent path steps per warp, so the per-column code is kept minimal: the direction is a
// template parameter, the staged column and the output pointer advance by constants, the next column's cost vector is
// loaded before the current path step, and the GPW bulk copies of a refill are issued by GPW lanes at once.
template <int NREG, int LPC, int DIR>
__device__ __forceinline__ void horizontal_dir(const Geo &g, const uint16_t *__restrict__ C, uint16_t *__restrict__ Lh, int y0,
                                               int nrows, uint8_t *smem)
{
    constexpr int GPW = 32 / LPC;                         // lane groups (rows) per warp
    constexpr int NW = HZ_THREADS / 32;
    constexpr int HZ_K = HzK<NREG, LPC>::value;
    constexpr int DP = 2 * NREG * LPC;                    // == g.Dp for this lane mapping (sgbm_api.cu make_geo)
    constexpr uint32_t COLB = 2u * DP;                    // bytes per column vector
    constexpr int STEPB = DIR ? -(int)COLB : (int)COLB;   // walking direction in bytes, staged and in the volume
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = lane / LPC, lg = lane % LPC;
    const int W1 = g.W1;
    const int rowBase = y0 + (blockIdx.x * NW + warp) * GPW;
    const int rowEnd = y0 + nrows;
    if (rowBase >= rowEnd) return;                        // whole warp idle (warp-uniform)
    int row = rowBase + grp;
    const bool act = row < rowEnd;
    if (!act) row = rowEnd - 1;                           // keep the warp convergent, no stores
    const uint32_t P1p = sm_keep((uint32_t)g.P1 * 0x10001u), P2mP1p = sm_keep((uint32_t)(g.P2 - g.P1) * 0x10001u);
    const LaneMasks lm = lane_masks(lg, g.lanesUsed - 1);

    constexpr uint32_t stageB = (uint32_t)GPW * HZ_K * COLB;
    const uint32_t wbufA = smem_u32(smem) + (uint32_t)warp * HZ_NS * stageB;
    const SmemBar bars{smem_u32(smem) + (uint32_t)NW * HZ_NS * stageB + (uint32_t)warp * HZ_NS * 8u};
    if (lane == 0) {
        for (int i = 0; i < HZ_NS; i++) mbar_init(bars[i], 1);
        mbar_fence_init();
    }
    __syncwarp();
    const int nchunks = (W1 + HZ_K - 1) / HZ_K;
    // refill of stage ci % HZ_NS with chunk ci: lane 0 arms the barrier, then lanes 0 .. GPW-1 issue one row's copy each
    const int fr = min(rowBase + (lane < GPW ? lane : 0), rowEnd - 1);
    const uint16_t *frow = C + (size_t)fr * g.rowStride;
    auto fill = [&](int ci) {
        const int st = ci % HZ_NS;
        const int s0 = ci * HZ_K, kc = min(HZ_K, W1 - s0);
        const int xlo = DIR ? (W1 - s0 - kc) : s0;        // first (lowest) column of the chunk
        const uint32_t bytes = (uint32_t)kc * COLB;
        if (lane == 0) mbar_expect_tx(bars[st], bytes * GPW);
        __syncwarp();
        if (lane < GPW)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(wbufA + (uint32_t)st * stageB + (uint32_t)lane * HZ_K * COLB), "l"(frow + (size_t)xlo * DP), "r"(bytes),
                           "r"(bars[st].addr) : "memory");
    };
    for (int ci = 0; ci < HZ_NS && ci < nchunks; ci++) fill(ci);

    uint32_t L[NREG], m = 0;
#pragma unroll
    for (int j = 0; j < NREG; j++) L[j] = 0;              // "predecessor outside" == L = 0, m = 0 (A.4)
    // output pointer of this lane's chunk of the current column; moves by one column per step
    char *po = reinterpret_cast<char *>(Lh) + ((size_t)row * g.rowStride + (size_t)(DIR ? W1 - 1 : 0) * DP) * 2 + 16 * lg;
    const uint32_t laneA = wbufA + (uint32_t)grp * HZ_K * COLB + 16u * (uint32_t)lg;
    auto store_col = [&](int i) {                         // column i of the chunk (compile-time i: immediate offsets)
        if (act) {
            uint4 *p4 = reinterpret_cast<uint4 *>(po + i * STEPB);
#pragma unroll
            for (int k4 = 0; k4 < NREG / 4; k4++) p4[LPC * k4] = make_uint4(L[4 * k4 + 0], L[4 * k4 + 1], L[4 * k4 + 2], L[4 * k4 + 3]);
        }
    };
    for (int ci = 0; ci < nchunks; ci++) {
        const int st = ci % HZ_NS;
        mbar_wait(bars[st], (uint32_t)(ci / HZ_NS) & 1u);
        const int kc = min(HZ_K, W1 - ci * HZ_K);
        // staged column 0 of the chunk in walking order: the first for (-1,0), the last for (+1,0)
        const uint32_t ca = laneA + (uint32_t)st * stageB + (DIR ? (uint32_t)(kc - 1) * COLB : 0u);
        if (kc == HZ_K) {
            // full chunk: straight-line code, the next column's cost vector is loaded before the current path step
            uint32_t Ca[NREG], Cb[NREG];
            lds_vec<NREG, LPC>(Ca, ca);
#pragma unroll
            for (int i = 0; i < HZ_K; i += 2) {
                lds_vec<NREG, LPC>(Cb, ca + (uint32_t)((i + 1) * STEPB));
                m = path_step_m<NREG, LPC>(L, m, Ca, P1p, P2mP1p, lm);
                store_col(i);
                if (i + 2 < HZ_K) lds_vec<NREG, LPC>(Ca, ca + (uint32_t)((i + 2) * STEPB));
                m = path_step_m<NREG, LPC>(L, m, Cb, P1p, P2mP1p, lm);
                store_col(i + 1);
            }
            po += HZ_K * STEPB;
        } else {                                          // the row's last, partial chunk
            for (int i = 0; i < kc; i++) {
                uint32_t Cc[NREG];
                lds_vec<NREG, LPC>(Cc, ca + (uint32_t)(i * STEPB));
                m = path_step_m<NREG, LPC>(L, m, Cc, P1p, P2mP1p, lm);
                store_col(0);
                po += STEPB;
            }
        }
        __syncwarp();
        if (ci + HZ_NS < nchunks) fill(ci + HZ_NS);
    }
}

template <int NREG, int LPC>
__global__ void __launch_bounds__(HZ_THREADS) k_horizontal(Geo g, const uint16_t *__restrict__ C,
                                                           uint16_t *__restrict__ LhA,
                                                           uint16_t *__restrict__ LhB, int y0, int nrows)
{
    extern __shared__ __align__(128) uint8_t smem[];
    // blockIdx.y: 0 = predecessor (-1,0) -> L_hA, 1 = predecessor (+1,0) -> L_hB
    if (blockIdx.y == 0) horizontal_dir<NREG, LPC, 0>(g, C, LhA, y0, nrows, smem);
    else horizontal_dir<NREG, LPC, 1>(g, C, LhB, y0, nrows, smem);
}

// =================================================================================================
// Vertical sweep
// =================================================================================================

__device__ __forceinline__ unsigned int ld_acquire(const unsigned int *p)
{
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(unsigned int *p, unsigned int v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <int NREG, int LPC>
__device__ __forceinline__ void load_vec_cg(uint32_t (&v)[NREG], const uint16_t *col, int lg)
{
    const uint4 *p = reinterpret_cast<const uint4 *>(col) + lg;
#pragma unroll
    for (int k = 0; k < NREG / 4; k++) {
        uint4 q = __ldcg(p + LPC * k);
        v[4 * k + 0] = q.x; v[4 * k + 1] = q.y; v[4 * k + 2] = q.z; v[4 * k + 3] = q.w;
    }
}

// NDIR = 3: vertical + both diagonals (MODE_SGBM / MODE_HH sweeps);  NDIR = 1: vertical only.
//
// Row blocking (NDIR = 3).  A diagonal path moves one column per row, so during R consecutive rows
// a strip is only influenced by the R columns next to each of its borders.  Strips therefore
// exchange state once per super-step of R rows: at its end every strip publishes the (x-1)-path
// state of its last R columns and the (x+1)-path state of its first R columns; during the next
// super-step the neighbour recomputes the incoming chains itself with R-1 "halo groups" per side
// (one direction only, a shrinking triangle of (R-1)R/2 cells).  The release/acquire flag round
// trip and the fence are paid once per R rows instead of every row.
//
// Register budget: ~5*NREG live packed registers + 3*NREG prefetch => cap the CTA size per NREG.
template <int NREG> struct VertMaxThreads { static const int value = NREG >= 16 ? 384 : (NREG >= 12 ? 512 : (NREG >= 8 ? 640 : 1024)); };

// NDIR = 1 (independent columns, 128 threads per CTA) is a pure streaming kernel: four CTAs per SM keep
// more loads in flight than three at 152 registers
template <int NREG, int LPC, int NDIR>
__global__ void __launch_bounds__(NDIR == 1 ? 128 : VertMaxThreads<NREG>::value, NDIR == 1 ? 4 : 1) k_vertical(VertArgs a)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const Geo &g = a.g;
    const int Dp = g.Dp, W1 = g.W1, lastLane = g.lanesUsed - 1;
    const int R = (NDIR == 3) ? a.R : 1, HG = R - 1;
    const int strip = blockIdx.x;
    const int SWmax = a.SW;
    int xs, xe;
    if (NDIR == 3) {
        xs = (int)((long long)strip * W1 / a.nstrips);
        xe = (int)((long long)(strip + 1) * W1 / a.nstrips);
    } else {
        xs = strip * SWmax;
        xe = min(xs + SWmax, W1);
    }
    const int SW = xe - xs;
    const int gi = threadIdx.x / LPC, lg = threadIdx.x % LPC;
    const uint32_t P1p = (uint32_t)g.P1 * 0x10001u, P2mP1p = (uint32_t)(g.P2 - g.P1) * 0x10001u;

    // ---- role of this lane group: group gi <-> column xs - HG + gi (contiguous) --------------------
    const int xcol = xs - HG + gi;
    const bool inImg = xcol >= 0 && xcol < W1;
    const bool own = inImg && gi >= HG && gi < HG + SW;
    const bool haloL = (NDIR == 3) && inImg && gi < HG && strip > 0;
    const bool haloR = (NDIR == 3) && inImg && gi >= HG + SW && gi < 2 * HG + SW && strip < a.nstrips - 1;
    const int hj = haloL ? HG - 1 - gi : gi - HG - SW;   // distance-1 of a halo column from the strip
    const int slot = gi + 1;                             // slot 0 <-> column xs - R
    const int x1 = inImg ? xcol : min(max(xcol, 0), W1 - 1);

    // ---- row program ----------------------------------------------------------------------------
    int yBegin, nRows, yStep, tOut;
    const uint16_t *calt = nullptr;
    int altRows = 0;
    if (a.threeway) {
        const int n = blockIdx.y;
        const int o0 = n * a.ss, o1 = min((n + 1) * a.ss, g.H);
        if (o0 >= o1) return;
        const int s0 = max(o0 - a.ov, 0);
        yBegin = s0; nRows = o1 - s0; yStep = 1; tOut = o0 - s0;
        if (s0 > 0 && g.r > 0) { calt = a.Calt + (size_t)(n - 1) * g.r * g.rowStride; altRows = g.r; }
    } else {
        yBegin = a.backward ? g.H - 1 : 0; nRows = g.H; yStep = a.backward ? -1 : 1; tOut = 0;
    }

    // ---- shared memory (layout computed on the host, see vertical_layout) ------------------------
    // exA/exC [2][NS][Dp] u16 : previous-row state of the (x-1)/(x+1) paths, slot s <-> column xs-R+s
    // exm     [2][2][NS] u32  : their packed minima
    // ssm     [own warps][GPW][Dp] u16 : WTA scratch, one vector per lane group
    // stgC    [warps][nstg][GPW][Dp]      : TMA-staged cost rows of the warp's columns
    // stgAB   [own warps][nstg][2][GPW][Dp] : TMA-staged input volumes (L_h / S_fwd)
    constexpr int GPW = 32 / LPC;
    const int NS = a.nslots;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, grpw = lane / LPC;
    const int nstg = a.nstg, ngroups = blockDim.x / LPC;
    uint16_t *exA = reinterpret_cast<uint16_t *>(smem);
    uint16_t *exC = exA + (size_t)2 * NS * Dp;
    uint32_t *exm = reinterpret_cast<uint32_t *>(exC + (size_t)2 * NS * Dp);
    uint16_t *ssm = reinterpret_cast<uint16_t *>(smem + a.ssmOff) + (size_t)gi * Dp;
    uint16_t *stgC = reinterpret_cast<uint16_t *>(smem + a.stgCOff);       // [nstg][ngroups][Dp]
    uint16_t *stgAB = reinterpret_cast<uint16_t *>(smem + a.stgABOff);     // [nstg][nAB][SWmax][Dp]
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + a.barOff);
    (void)warp; (void)grpw; (void)ngroups;
    if (NDIR == 3) {
        uint32_t *z = reinterpret_cast<uint32_t *>(smem);
        const int nz = (int)((2 * 2 * (size_t)NS * Dp * 2 + 2 * 2 * NS * 4) / 4);
        for (int i = threadIdx.x; i < nz; i += blockDim.x) z[i] = 0;
        if (threadIdx.x == 0) {
            for (int i = 0; i < nstg; i++) mbar_init(&bars[i], 1);
            mbar_fence_init();
        }
        __syncthreads();
    }
    const size_t haloStride = (size_t)Dp + 8;            // u16 elements per published column (vector + min)
    const bool pubA = (NDIR == 3) && own && strip < a.nstrips - 1 && xcol >= xe - R;   // feeds strip+1
    const bool pubC = (NDIR == 3) && own && strip > 0 && xcol < xs + R;                // feeds strip-1
    const bool rcvL = (NDIR == 3) && gi < R && strip > 0;                              // copies column xs-R+gi
    const bool rcvR = (NDIR == 3) && gi >= R && gi < 2 * R && strip < a.nstrips - 1;   // copies column xe+(gi-R)

    uint32_t LB[NREG], mB = 0;
#pragma unroll
    for (int j = 0; j < NREG; j++) LB[j] = 0;

    const bool hasB = a.inB != nullptr;
    // NDIR = 3: the strip's columns (own + halo) are contiguous in memory, so one thread stages a
    // whole row with one TMA bulk copy per volume into a ring of nstg shared-memory stages.
    const int scol0 = xs - HG;                            // column of group 0
    const int clo = max(scol0, 0), chi = min(scol0 + ngroups, W1);
    auto fill = [&](int t) {                              // thread 0: stage row t of the row program
        const int sg = t % nstg;
        const int y = yBegin + t * yStep;
        const uint32_t bytesC = (uint32_t)(chi - clo) * Dp * 2, bytesAB = (uint32_t)SW * Dp * 2;
        mbar_expect_tx(&bars[sg], bytesC + bytesAB * (hasB ? 2u : 1u));
        bulk_g2s(stgC + ((size_t)sg * ngroups + (clo - scol0)) * Dp, a.C + (size_t)y * g.rowStride + (size_t)clo * Dp, bytesC, &bars[sg]);
        const size_t off = (size_t)y * g.rowStride + (size_t)xs * Dp;
        bulk_g2s(stgAB + (size_t)(sg * a.nAB + 0) * SWmax * Dp, a.inA + off, bytesAB, &bars[sg]);
        if (hasB) bulk_g2s(stgAB + (size_t)(sg * a.nAB + 1) * SWmax * Dp, a.inB + off, bytesAB, &bars[sg]);
    };
    // NDIR = 1: independent columns, operands of the next row are prefetched into registers.
    uint32_t Cn[NREG], An[NREG], Bn[NREG];
    auto crow_ptr = [&](int t) -> const uint16_t * {
        const int y = yBegin + t * yStep;
        return (t < altRows) ? calt + (size_t)t * g.rowStride + (size_t)x1 * Dp
                             : a.C + (size_t)y * g.rowStride + (size_t)x1 * Dp;
    };
    if (NDIR == 3) {
        if (threadIdx.x == 0) {
            for (int t = 0; t < nstg && t < nRows; t++) fill(t);
            mbar_wait(&bars[0], 0u);
        }
        __syncthreads();
    } else {
        load_vec_nc<NREG, LPC>(Cn, crow_ptr(0), lg);
        if (0 >= tOut && own) {
            load_vec<NREG, LPC>(An, a.inA + (size_t)yBegin * g.rowStride + (size_t)x1 * Dp, lg);
            if (hasB) load_vec<NREG, LPC>(Bn, a.inB + (size_t)yBegin * g.rowStride + (size_t)x1 * Dp, lg);
        }
    }

    int k = 0, sidx = 0;                                  // row inside the super-step, super-step index
    for (int t = 0; t < nRows; t++) {
        const int y = yBegin + t * yStep;
        const int pp = (t + 1) & 1, pc = t & 1;          // previous / current row parity
        // ---- super-step start: receive the neighbours' published columns ------------------------
        if (NDIR == 3 && k == 0 && t > 0) {
            if (!SGBM_DBG_HOOK(a.dbgNoSync)) {
                if (rcvL && lg == 0) while (ld_acquire(a.flagA + strip - 1) < (unsigned)sidx) { }
                if (rcvR && lg == 0) while (ld_acquire(a.flagA + strip + 1) < (unsigned)sidx) { }
            }
            __syncwarp();
            if (rcvL || rcvR) {
                const int i = rcvL ? gi : gi - R;
                const uint16_t *h = (rcvL ? a.haloA + ((size_t)(strip - 1) * 2 + ((sidx - 1) & 1)) * R * haloStride
                                          : a.haloC + ((size_t)(strip + 1) * 2 + ((sidx - 1) & 1)) * R * haloStride) +
                                    (size_t)i * haloStride;
                uint32_t v[NREG];
                load_vec_cg<NREG, LPC>(v, h, lg);
                const int sl = rcvL ? i : R + SW + i;
                store_vec<NREG, LPC>(v, (rcvL ? exA : exC) + ((size_t)pp * NS + sl) * Dp, lg);
                if (lg == 0) exm[((rcvL ? 0 : 1) * 2 + pp) * NS + sl] = __ldcg(reinterpret_cast<const unsigned int *>(h + Dp));
            }
            __syncthreads();
        }
        const bool haloAct = (haloL || haloR) && k <= HG - 1 - hj;     // chain still needed this row
        const bool doA = own || (haloL && haloAct), doC = own || (haloR && haloAct);

        uint32_t Cc[NREG], S[NREG];
        const bool outRow = t >= tOut;
        const bool warpOwn = __any_sync(0xFFFFFFFFu, own);
        if (NDIR == 3) {
            const int sg = t % nstg;                      // stage t was waited for before the last barrier
            load_vec<NREG, LPC>(Cc, stgC + ((size_t)sg * ngroups + gi) * Dp, lg);
            if (warpOwn) {
                const int oi = own ? gi - HG : 0;
                load_vec<NREG, LPC>(S, stgAB + ((size_t)(sg * a.nAB + 0) * SWmax + oi) * Dp, lg);
                if (hasB) {
                    uint32_t Bv[NREG];
                    load_vec<NREG, LPC>(Bv, stgAB + ((size_t)(sg * a.nAB + 1) * SWmax + oi) * Dp, lg);
#pragma unroll
                    for (int j = 0; j < NREG; j++) S[j] = paddmin(S[j], Bv[j], SGBM_MAX_S);
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < NREG; j++) Cc[j] = Cn[j];
            if (outRow) {
#pragma unroll
                for (int j = 0; j < NREG; j++) S[j] = hasB ? paddmin(An[j], Bn[j], SGBM_MAX_S) : An[j];
            }
            if (t + 1 < nRows) {                          // prefetch the next row of this column
                load_vec_nc<NREG, LPC>(Cn, crow_ptr(t + 1), lg);
                if (t + 1 >= tOut && own) {
                    const size_t off = (size_t)(y + yStep) * g.rowStride + (size_t)x1 * Dp;
                    load_vec<NREG, LPC>(An, a.inA + off, lg);
                    if (hasB) load_vec<NREG, LPC>(Bn, a.inB + off, lg);
                }
            }
        }
        uint32_t Ln[NREG];
        // ---- vertical path: predecessor (x, previous row), state in registers -------------------
        if (warpOwn) {
            mB = path_step<NREG, LPC>(Ln, LB, mB, Cc, P1p, P2mP1p, lg, lastLane);
#pragma unroll
            for (int j = 0; j < NREG; j++) LB[j] = Ln[j];
            if (outRow) {
#pragma unroll
                for (int j = 0; j < NREG; j++) S[j] = paddmin(S[j], Ln[j], SGBM_MAX_S);
            }
        }
        if (NDIR == 3) {
            const bool lastOfStep = (k == R - 1) && (t + 1 < nRows);   // publish after this row
            uint32_t Lp[NREG], mp;
            // ---- path with predecessor (x-1, previous row) --------------------------------------
            if (__any_sync(0xFFFFFFFFu, doA)) {
                load_vec<NREG, LPC>(Lp, exA + ((size_t)pp * NS + slot - 1) * Dp, lg);
                mp = exm[(0 * 2 + pp) * NS + slot - 1];
                const uint32_t mA = path_step<NREG, LPC>(Ln, Lp, mp, Cc, P1p, P2mP1p, lg, lastLane);
                if (doA) {
                    store_vec<NREG, LPC>(Ln, exA + ((size_t)pc * NS + slot) * Dp, lg);
                    if (lg == 0) exm[(0 * 2 + pc) * NS + slot] = mA;
                    if (pubA && lastOfStep) {
                        uint16_t *h = a.haloA + (((size_t)strip * 2 + (sidx & 1)) * R + (xcol - (xe - R))) * haloStride;
                        store_vec<NREG, LPC>(Ln, h, lg);
                        if (lg == 0) *reinterpret_cast<unsigned int *>(h + Dp) = mA;
                    }
                }
                if (outRow && own) {
#pragma unroll
                    for (int j = 0; j < NREG; j++) S[j] = paddmin(S[j], Ln[j], SGBM_MAX_S);
                }
            }
            // ---- path with predecessor (x+1, previous row) --------------------------------------
            if (__any_sync(0xFFFFFFFFu, doC)) {
                load_vec<NREG, LPC>(Lp, exC + ((size_t)pp * NS + slot + 1) * Dp, lg);
                mp = exm[(1 * 2 + pp) * NS + slot + 1];
                const uint32_t mC = path_step<NREG, LPC>(Ln, Lp, mp, Cc, P1p, P2mP1p, lg, lastLane);
                if (doC) {
                    store_vec<NREG, LPC>(Ln, exC + ((size_t)pc * NS + slot) * Dp, lg);
                    if (lg == 0) exm[(1 * 2 + pc) * NS + slot] = mC;
                    if (pubC && lastOfStep) {
                        uint16_t *h = a.haloC + (((size_t)strip * 2 + (sidx & 1)) * R + (xcol - xs)) * haloStride;
                        store_vec<NREG, LPC>(Ln, h, lg);
                        if (lg == 0) *reinterpret_cast<unsigned int *>(h + Dp) = mC;
                    }
                }
                if (outRow && own) {
#pragma unroll
                    for (int j = 0; j < NREG; j++) S[j] = paddmin(S[j], Ln[j], SGBM_MAX_S);
                }
            }
            if (lastOfStep && (pubA || pubC)) __threadfence();
        }

        if (outRow && warpOwn) {
            if (a.sout) {
                if (own) store_vec<NREG, LPC>(S, a.sout + (size_t)y * g.rowStride + (size_t)x1 * Dp, lg);
            } else {
                // ---- winner-take-all (A.5 / A.6) --------------------------------------------------
                if (a.sdbg && own) store_vec<NREG, LPC>(S, a.sdbg + (size_t)y * g.rowStride + (size_t)x1 * Dp, lg);
                {
                    const int padFrom = sgbm_pad_from(g);             // numDisparities % 8 != 0: hide the padding disparities
                    if (padFrom < 2 * NREG) mask_pad_regs<NREG>(S, lg == lastLane, padFrom);
                }
                uint32_t tm = local_min<NREG>(S);
                if (lg > lastLane) tm = SGBM_INF2;
                const uint32_t mS2 = group_min<LPC>(tm);
                const int minS = (int)(mS2 & 0xFFFFu);
                int best;
                if (!a.threeway) {
                    int idx = 0x7FFF;
#pragma unroll
                    for (int j = NREG - 1; j >= 0; j--) {
                        const uint32_t e = S[j] ^ mS2;
                        if ((e >> 16) == 0) idx = 2 * j + 1;
                        if ((e & 0xFFFFu) == 0) idx = 2 * j;
                    }
                    int dl = (lg <= lastLane && idx != 0x7FFF) ? lg * 2 * NREG + idx : 0x7FFF;
#pragma unroll
                    for (int off = LPC / 2; off >= 1; off >>= 1)
                        dl = min(dl, __shfl_xor_sync(0xFFFFFFFFu, dl, off, LPC));
                    best = (minS == 32767) ? -1 : dl;                 // first minimum (A.5)
                } else {
                    // 8-lane SIMD tie-break of the reference (A.6): per class d%8 the largest tied d,
                    // then the smallest of those.  q[i] halves hold (d/8 + 1) of classes 2i, 2i+1.
                    uint32_t q[4] = {0u, 0u, 0u, 0u};
                    const uint32_t qbase = (uint32_t)(lg * (NREG / 4));
#pragma unroll
                    for (int j = 0; j < NREG; j++) {
                        const uint32_t e = S[j] ^ mS2;
                        const uint32_t val = qbase + (uint32_t)(j >> 2) + 1u;
                        const uint32_t v2 = (((e & 0xFFFFu) == 0) ? val : 0u) | (((e >> 16) == 0) ? (val << 16) : 0u);
                        q[j & 3] = __vmaxu2(q[j & 3], v2);
                    }
                    best = 0x7FFFFFFF;
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        uint32_t v = (lg > lastLane) ? 0u : q[i];
#pragma unroll
                        for (int off = LPC / 2; off >= 1; off >>= 1)
                            v = __vmaxu2(v, __shfl_xor_sync(0xFFFFFFFFu, v, off, LPC));
                        const int lo = (int)(v & 0xFFFFu), hi = (int)(v >> 16);
                        if (lo) best = min(best, 8 * (lo - 1) + 2 * i);
                        if (hi) best = min(best, 8 * (hi - 1) + 2 * i + 1);
                    }
                }
                // S to shared scratch: sub-pixel neighbours and the masked uniqueness scan
                __syncwarp();
                store_vec<NREG, LPC>(S, ssm, lg);
                __syncwarp();
                int Sm = 0, Sp = 0;
                const bool interior = best > 0 && best < g.D - 1;
                if (lg == 0 && interior) {
                    Sm = ssm[sgbm_pos(best - 1, NREG, LPC)];
                    Sp = ssm[sgbm_pos(best + 1, NREG, LPC)];
                }
                bool reject = false;
                if (g.UR > 0) {
                    int T;
                    const int av = 100 - g.UR;
                    if (!a.threeway) {                        // S(d)*(100-UR) < minS*100  <=>  S(d) < T
                        T = av > 0 ? min((100 * minS + av - 1) / av, 32768) : (minS > 0 ? 32768 : 0);
                    } else {                                  // truncated threshold with (short) wrap (A.6)
                        const int t1 = (int)(short)((100 * minS) / av + 1);
                        T = max(t1, 0);
                    }
                    __syncwarp();
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
                    uint32_t t2 = local_min<NREG>(S2);
                    if (lg > lastLane) t2 = SGBM_INF2;
                    const int m2 = (int)(group_min<LPC>(t2) & 0xFFFFu);
                    reject = m2 < T;
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
                            dq += ((Sm - Sp) * 16 + den) / (2 * den);
                        }
                        out = dq + g.minD * 16;
                    }
                    a.raw[(size_t)y * g.W + x] = (int16_t)out;
                }
            }
        }
        if (NDIR == 3) {
            if (threadIdx.x == 0 && t + 1 < nRows) mbar_wait(&bars[(t + 1) % nstg], (uint32_t)((t + 1) / nstg) & 1u);
            __syncthreads();
            if (threadIdx.x == 0) {
                if (k == R - 1 && t + 1 < nRows) st_release(a.flagA + strip, (unsigned)(sidx + 1));
                if (t + nstg < nRows) fill(t + nstg);     // stage t % nstg is free now
            }
            if (++k == R) { k = 0; sidx++; }
        }
    }
}

// =================================================================================================
// Host side: template dispatch and launch configuration
// =================================================================================================
// Shared-memory layout of k_vertical (must match the carve-up in the kernel).
static size_t vertical_layout(VertArgs &a, int SWmax, int R, int threads, int ndir, int nstg)
{
    const Geo &g = a.g;
    const int ngroups = threads / g.lpc;
    (void)R;
    a.nslots = ngroups + 2;
    a.nstg = nstg;
    a.nAB = a.inB ? 2 : 1;
    size_t off = (ndir == 3) ? (size_t)2 * 2 * a.nslots * g.Dp * 2 + (size_t)2 * 2 * a.nslots * 4 : 0;
    off = (off + 127) & ~(size_t)127;
    a.ssmOff = (unsigned)off;
    if (!a.sout) off += (size_t)ngroups * g.Dp * 2;
    a.stgCOff = (unsigned)off;
    if (ndir == 3) off += (size_t)nstg * ngroups * g.Dp * 2;
    a.stgABOff = (unsigned)off;
    if (ndir == 3) off += (size_t)nstg * a.nAB * SWmax * g.Dp * 2;
    a.barOff = (unsigned)off;
    off += 8 * 8;
    return off;
}

template <int NREG, int LPC>
static int launch_horizontal_t(const Geo &g, const uint16_t *C, uint16_t *LhA, uint16_t *LhB, int y0, int nrows,
                               cudaStream_t st)
{
    constexpr int GPW = 32 / LPC, NW = HZ_THREADS / 32, HZ_K = HzK<NREG, LPC>::value;
    const size_t smem = (size_t)NW * HZ_NS * GPW * HZ_K * g.Dp * 2 + (size_t)NW * HZ_NS * 8;
    static unsigned long long attrDone = 0;   // one bit per device: function attributes are per device
    {
        SgbmDeviceOnce once(attrDone);
        if (once.first) {
            SGBM_CUDA_CHECK(cudaFuncSetAttribute(k_horizontal<NREG, LPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            once.done();
        }
    }
    const int rowsPerCta = NW * GPW;
    dim3 grid((nrows + rowsPerCta - 1) / rowsPerCta, 2);
    k_horizontal<NREG, LPC><<<grid, HZ_THREADS, smem, st>>>(g, C, LhA, LhB, y0, nrows);
    sgbm_count_launch(1);
    SGBM_CUDA_CHECK(cudaGetLastError());
    return 0;
}

struct VertPlan { int SW, nstrips, threads; size_t smem; };

template <int NREG, int LPC, int NDIR>
static int launch_vertical_t(VertArgs &a, int numSMs, cudaStream_t st)
{
    const Geo &g = a.g;
    auto kern = k_vertical<NREG, LPC, NDIR>;
    const SgbmKnobs &kn = sgbm_knobs();
    const int maxSmem = kn.maxSmemOptin;
    static unsigned long long attrDone = 0;   // one bit per device: function attributes are per device
    {
        SgbmDeviceOnce once(attrDone);
        if (once.first) {
            SGBM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, maxSmem));
            once.done();
        }
    }
    const int nstgWant = kn.nstg >= 2 && kn.nstg <= 4 ? kn.nstg : 3;
    if (NDIR == 1) {
        // independent columns: ordinary grid, 128 threads per CTA
        const int threads = 128;
        a.SW = threads / LPC;
        a.R = 1;
        a.nstrips = (g.W1 + a.SW - 1) / a.SW;
        const size_t smem = vertical_layout(a, a.SW, 1, threads, 1, 2);
        if (smem > (size_t)maxSmem) return sgbm_fail(-3, "vertical sweep needs %zu bytes of shared memory (max %d)", smem, maxSmem);
        dim3 grid(a.nstrips, a.threeway ? 4 : 1);
        kern<<<grid, threads, smem, st>>>(a);
        sgbm_count_launch(1);
        SGBM_CUDA_CHECK(cudaGetLastError());
        return 0;
    }
    // NDIR == 3: all strips must be co-resident (neighbour flags) -> cooperative launch, <= 1 CTA per SM.
    // R rows per super-step (env SGBM_VR overrides); every strip must own >= R columns.
    const int maxThreads = VertMaxThreads<NREG>::value;
    int R = 9;
    if (kn.vr > 0) R = kn.vr;
    if (R > 16) R = 16;
    int nstrips = numSMs, SWmax = 1, threads = 32, nstg = 2;
    size_t smem = 0;
    for (;; R--) {
        if (R < 1) return 1;                              // too wide for this kernel: the caller goes row by row
        nstrips = numSMs;
        const int minCols = R > 2 ? R : 2;                // every strip owns >= R (and >= 2) columns
        if (nstrips > g.W1 / minCols) nstrips = g.W1 / minCols;
        if (nstrips < 1) nstrips = 1;
        if (nstrips == 1 && R > 1) continue;              // a single strip has no halos
        SWmax = (g.W1 + nstrips - 1) / nstrips;
        threads = (((SWmax + 2 * (R - 1)) * LPC + 31) / 32) * 32;
        if (threads > maxThreads) continue;
        bool fits = false;
        for (nstg = nstgWant; nstg >= 2; nstg--) {
            smem = vertical_layout(a, SWmax, R, threads, 3, nstg);
            if (smem <= (size_t)maxSmem) { fits = true; break; }
        }
        if (fits) break;
    }
    a.SW = SWmax; a.R = R; a.nstrips = nstrips;
    smem = vertical_layout(a, SWmax, R, threads, 3, nstg);
    int occ = 0;
    SGBM_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
    if (occ * numSMs < a.nstrips) return sgbm_fail(-3, "vertical sweep cannot be made co-resident (%d strips, %d x %d slots)", a.nstrips, occ, numSMs);
    SGBM_CUDA_CHECK(cudaMemsetAsync(a.flagA, 0, sizeof(unsigned int) * a.nstrips, st));
    void *args[] = {&a};
    SGBM_CUDA_CHECK(cudaLaunchCooperativeKernel((void *)kern, dim3(a.nstrips), dim3(threads), args, smem, st));
    sgbm_count_launch(1);
    return 0;
}

#define SGBM_DISPATCH(NREG_, LPC_, EXPR)                                              \
    if (g.nreg == NREG_ && g.lpc == LPC_) { constexpr int NREG = NREG_; constexpr int LPC = LPC_; return EXPR; }

#ifdef SGBM_FAST_BUILD      // development builds: only the lane mappings of the BASELINE configurations (make FAST=1)
#define SGBM_DISPATCH_ALL(EXPR) SGBM_DISPATCH(4, 2, EXPR) SGBM_DISPATCH(8, 8, EXPR) SGBM_DISPATCH(12, 8, EXPR) SGBM_DISPATCH(16, 8, EXPR)
#else
#define SGBM_DISPATCH_ALL(EXPR)                                                       \
    SGBM_DISPATCH(4, 2, EXPR) SGBM_DISPATCH(4, 4, EXPR) SGBM_DISPATCH(4, 8, EXPR)     \
    SGBM_DISPATCH(4, 16, EXPR) SGBM_DISPATCH(4, 32, EXPR)                             \
    SGBM_DISPATCH(8, 2, EXPR) SGBM_DISPATCH(8, 4, EXPR) SGBM_DISPATCH(8, 8, EXPR)     \
    SGBM_DISPATCH(8, 16, EXPR) SGBM_DISPATCH(8, 32, EXPR)                             \
    SGBM_DISPATCH(12, 2, EXPR) SGBM_DISPATCH(12, 4, EXPR) SGBM_DISPATCH(12, 8, EXPR)  \
    SGBM_DISPATCH(12, 16, EXPR) SGBM_DISPATCH(12, 32, EXPR)                           \
    SGBM_DISPATCH(16, 2, EXPR) SGBM_DISPATCH(16, 4, EXPR) SGBM_DISPATCH(16, 8, EXPR)  \
    SGBM_DISPATCH(16, 16, EXPR) SGBM_DISPATCH(16, 32, EXPR)
#endif

int sgbm_launch_horizontal(const Geo &g, const uint16_t *C, uint16_t *LhA, uint16_t *LhB, int y0, int nrows,
                           cudaStream_t st)
{
    SGBM_DISPATCH_ALL((launch_horizontal_t<NREG, LPC>(g, C, LhA, LhB, y0, nrows, st)))
    return sgbm_fail(-3, "no kernel for lane mapping nreg=%d lpc=%d", g.nreg, g.lpc);
}

int sgbm_launch_sweep(const VertArgs &a, int numSMs, cudaStream_t st);   // sgbm_sweep.cu
int sgbm_launch_rowstep(const VertArgs &a, cudaStream_t st);             // sgbm_sweep.cu

int sgbm_launch_vertical(VertArgs &a, int ndir, int numSMs, cudaStream_t st)
{
    const Geo &g = a.g;
    if (ndir == 3) {
        // role-specialised sweep (sgbm_sweep.cu); the lock-step kernel below remains as the fallback for
        // geometries it cannot hold (and for A/B runs with SGBM_SWEEP=0)
        if (sgbm_knobs().rowstep) return sgbm_launch_rowstep(a, st);
        if (sgbm_knobs().sweep) {
            const int rc = sgbm_launch_sweep(a, numSMs, st);
            if (rc <= 0) return rc;
        }
#define SGBM_TRY3(NREG_, LPC_) if (g.nreg == NREG_ && g.lpc == LPC_) { const int rc = launch_vertical_t<NREG_, LPC_, 3>(a, numSMs, st); if (rc <= 0) return rc; return sgbm_launch_rowstep(a, st); }
#ifdef SGBM_FAST_BUILD
        SGBM_TRY3(4, 2) SGBM_TRY3(8, 8) SGBM_TRY3(12, 8) SGBM_TRY3(16, 8)
#else
        SGBM_TRY3(4, 2) SGBM_TRY3(4, 4) SGBM_TRY3(4, 8) SGBM_TRY3(4, 16) SGBM_TRY3(4, 32)
        SGBM_TRY3(8, 2) SGBM_TRY3(8, 4) SGBM_TRY3(8, 8) SGBM_TRY3(8, 16) SGBM_TRY3(8, 32)
        SGBM_TRY3(12, 2) SGBM_TRY3(12, 4) SGBM_TRY3(12, 8) SGBM_TRY3(12, 16) SGBM_TRY3(12, 32)
        SGBM_TRY3(16, 2) SGBM_TRY3(16, 4) SGBM_TRY3(16, 8) SGBM_TRY3(16, 16) SGBM_TRY3(16, 32)
#endif
#undef SGBM_TRY3
    } else {
        SGBM_DISPATCH_ALL((launch_vertical_t<NREG, LPC, 1>(a, numSMs, st)))
    }
    return sgbm_fail(-3, "no kernel for lane mapping nreg=%d lpc=%d", g.nreg, g.lpc);
}
