// sgbm_cost3.cu -- stage 2 of StereoSGBM.compute (main.ipynb:668) for sm_100a, third generation:
// Birchfield-Tomasi pixel cost + (2r+1)^2 block sum -> cost volume C, 1-channel input.
//
// Same arithmetic, tiling (TX columns x RB rows per CTA, rows staged by TMA bulk copies) and volume
// layout as sgbm_cost2.cu; what changed is who computes what.  k_cost2 computed the pixel costs of a
// row into a shared-memory buffer (one warp per column) and summed them in a second phase (16.6
// shared-memory wavefronts and 65 instructions per output word, two block barriers per row).  Here
// one thread owns ONE disparity pair (a packed u16x2 word) of XPT = 32 consecutive columns and walks
// along the row:
//   * the pixel cost never leaves registers: the horizontal (2r+1) window sum slides in a register
//     window (the walk is fully unrolled, so the rotating window has compile-time indices);
//   * the right-image operands of column x+1 overlap those of column x: the pair word of an odd
//     right position is assembled with one PRMT from the two neighbouring even pair words, so only the
//     even-parity pair array is staged and every plane costs one shared load per TWO columns;
//   * the left-image operands arrive pre-expanded to packed words (prefilter, 32 bytes per pixel):
//     two warp-uniform 128-bit loads per column;
//   * the vertical running sum stays in registers (crun[32]); the ring of the last 2r+1 horizontal
//     sums is thread-private shared memory (one load + one store per output word, no barrier).
// About 36 instructions and 7 shared-memory wavefronts per output word; one block barrier per row
// (stage hand-back).  Image borders (replicated pixel-cost columns, A.3) run a second instantiation
// of the walk; only the first / last thread groups of a row of tiles take it.
#include "sgbm_common.cuh"
#include <stdlib.h>
#include <string.h>

#define COST3_K 256u          // bias; multiple of 4 so that (bt_t + K) >> 2 == (bt_t >> 2) + K/4
#ifndef COST3_XPT
#define COST3_XPT 16
#endif
//         // output columns per thread

int sgbm_cost2_rpw(const Geo &g);
size_t sgbm_cost2_right_offset(const Geo &g);
size_t sgbm_cost2_leftx_offset(const Geo &g);

struct Cost3Args {
    Geo g;
    const uint4 *leftX;      // [H][W][2] packed-expanded left operands (k_prefilter2)
    const uint32_t *rpairs;  // [6][H][2][RPW] right pair words; only parity 0 is read
    int RPW;
    uint16_t *out;           // row y is written at out + (y - y0) * rowStride
    int y0, nrows;           // output rows [y0, y0 + nrows)
    int ylo;                 // vertical clamp floor (0, or the stripe start for 3WAY)
    int NXG, RB;             // thread groups along x (TX = NXG * 32), rows per band
    int NQh, nstg, nact;     // staged pair words per plane, stages, active threads (NXG * Dw)
    unsigned int stgOff, barOff, stageBytes, rpBytes;
    unsigned int one, neg1;  // 1 and 0xFFFFFFFF: opaque multipliers (see fma_mad)
};

// ring accesses go through volatile asm WITHOUT a memory clobber: the entries are thread-private, and
// the compiler stays free to hoist the (ordinary) loads of the next column's operands above them
__device__ __forceinline__ uint32_t ring_ld(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void ring_st(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v)); }

// Birchfield-Tomasi cost of one plane for two adjacent disparities, biased by K (A.2):
//   min( max(u - vhi, vlo - u, 0), max(v - uhi, ulo - v, 0) ) + K
// The left operands arrive pre-biased (uK = u + K, KmU = K - u, KmUhi = K - u_hi, UloK = u_lo + K), so each
// of the four differences is ONE multiply-add a * (+-1) + c.  The multipliers are kernel arguments the
// compiler cannot fold: that forces IMAD, which issues on the FMA pipe -- the packed min/max/permute
// instructions saturate the ALU pipe (half rate), the FMA pipe is otherwise idle in this kernel.
__device__ __forceinline__ uint32_t fma_mad(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t bt3(uint32_t uK, uint32_t KmU, uint32_t KmUhi, uint32_t UloK, uint32_t v, uint32_t vlo,
                                        uint32_t vhi, uint32_t one, uint32_t neg1)
{
    const uint32_t KK = COST3_K * 0x10001u;
    const uint32_t c1 = __vimax3_u16x2(fma_mad(vhi, neg1, uK), fma_mad(vlo, one, KmU), KK);
    const uint32_t c2 = __vimax3_u16x2(fma_mad(v, one, KmUhi), fma_mad(v, neg1, UloK), KK);
    return __vminu2(c1, c2);
}

// One row of one thread: walk XPT + 2R columns starting at tile column xg*XPT (image column x1s).
//   lrow : staged left operands of the first walked column (2 x uint4 per column)
//   er   : staged even pair words, positioned so that er[p*NQh + m] is E[k0 + m] - w for plane p
//   rg   : shared address (bytes) of this thread's ring entry of column 0 in the current slot
//   lb   : BORDER: walked columns [0, lb) lie left of the image (their cost is that of column lb)
//   rlim : BORDER: walked columns > rlim lie right of the image (cost of column rlim)
//   DWT / NTT : compile-time words per column / threads per CTA (0 = run-time): the global stores and the
//               ring accesses of the hot geometries then use immediate offsets
template <int R, int PAR, bool BORDER, int DWT, int NTT>
__device__ __forceinline__ void cost3_walk(const uint4 *__restrict__ lrow, const uint32_t *__restrict__ er, int NQh,
                                           uint32_t rg, uint32_t rstrideRt, uint32_t (&crun)[COST3_XPT], uint32_t *orow32,
                                           int DwRt, int nvalid, int lb, int rlim, uint32_t one, uint32_t neg1)
{
    const int Dw = DWT ? DWT : DwRt;
    const uint32_t rstride = NTT ? (uint32_t)NTT * 4u : rstrideRt;
    constexpr int XPT = COST3_XPT, NS = 2 * R + 1, NCOL = XPT + 2 * R;
    uint32_t win[NS];
    uint32_t A[6];
#pragma unroll
    for (int p = 0; p < 6; p++) A[p] = er[p * NQh];
    uint32_t hs = 0, lastpix = 0;
#pragma unroll
    for (int xx = 0; xx < NCOL; xx++) {
        uint32_t v[6];
        if ((xx & 1) == PAR) {                            // even right position: the staged word itself
#pragma unroll
            for (int p = 0; p < 6; p++) v[p] = A[p];
        } else {                                          // odd: (E[k+1].hi, E[k].lo)
            const int m = (xx + 1 + PAR) / 2;
#pragma unroll
            for (int p = 0; p < 6; p++) {
                const uint32_t B = er[p * NQh + m];
                v[p] = __byte_perm(B, A[p], 0x5432);
                A[p] = B;
            }
        }
        const uint4 l0 = lrow[2 * xx], l1 = lrow[2 * xx + 1];
        const uint32_t btg = bt3(l0.x, l0.y, l0.z, l0.w, v[0], v[1], v[2], one, neg1);
        const uint32_t btt = bt3(l1.x, l1.y, l1.z, l1.w, v[3], v[4], v[5], one, neg1);
        // (bt_g + K) + ((bt_t + K) >> 2): biased by K + K/4 per half; the bias cancels in the ring update
        // and is removed from the running sum once (initial value of crun)
        uint32_t pix = fma_mad((btt >> 2) & 0x3FFF3FFFu, one, btg);
        if (BORDER) {
            if (xx > rlim) pix = lastpix; else lastpix = pix;
            if (xx < lb) pix = 0;
        }
        if (xx >= NS) hs += pix - win[xx % NS]; else hs += pix;
        win[xx % NS] = pix;
        if (BORDER && xx <= R && xx > 0) {
            if (xx == lb) {                               // first column inside the image: replicate it to the left
#pragma unroll
                for (int j = 0; j < xx; j++) win[j] = pix;
                hs += (uint32_t)xx * pix;
            }
        }
        if (xx >= 2 * R) {
            const int n = xx - 2 * R;
            const uint32_t ra = rg + (uint32_t)n * rstride;
            const uint32_t old = ring_ld(ra);
            ring_st(ra, hs);
            crun[n] += hs - old;
            if (BORDER ? n < nvalid : nvalid > 0) {
                if (DWT) orow32[n * DWT] = crun[n];
                else *reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(orow32) + (uint32_t)n * (uint32_t)Dw * 4u) = crun[n];
            }
        }
    }
}

template <int R, int PAR, int DWT, int NTT>
__global__ void __launch_bounds__(NTT ? NTT : 512) k_cost3(Cost3Args a)
{
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int XPT = COST3_XPT, NS = 2 * R + 1, NCOL = XPT + 2 * R;
    const Geo &g = a.g;
    const int Dw = DWT ? DWT : g.Dp / 2, HP = g.D / 2;
    const int TX = a.NXG * XPT;
    const int x0 = blockIdx.x * TX;                      // first valid column of the tile
    const int yb = a.y0 + blockIdx.y * a.RB;             // first output row of the band
    const int yend = min(yb + a.RB, a.y0 + a.nrows);
    const int tid = threadIdx.x, nthr = NTT ? NTT : blockDim.x;
    const int nstg = a.nstg, NQh = a.NQh;
    const bool act = tid < a.nact;
    const int xg = act ? tid / Dw : 0, w = act ? tid % Dw : 0;

    // staged column range of the tile (valid columns, clamped to the image) and the right-image window
    const int xlo = max(x0 - R, 0), xhi = min(x0 + TX + R, g.W1);
    const int xa = g.minX1 + xlo;
    const int q0 = xa - g.maxD + 1;                      // first right-image pixel the tile touches
    const int iLo = (q0 >> 1) & ~3;                      // first staged pair word (16-byte aligned)
    const uint32_t leftBytes = (uint32_t)(xhi - xlo) * 32u;

    uint32_t *ring = reinterpret_cast<uint32_t *>(smem);                      // [NS][XPT][nthr]
    uint8_t *stg = smem + a.stgOff;                                           // [nstg] { E [6][NQh] u32 ; left [TX+2R] 32 B }
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + a.barOff);
    for (int i = tid; i < NS * XPT * nthr; i += nthr) ring[i] = 0;
    if (tid == 0) {
        for (int i = 0; i < nstg; i++) mbar_init(&bars[i], 1);
        mbar_fence_init();
    }
    __syncthreads();

    const int nsteps = (yend - yb) + 2 * R;
    auto fill = [&](int k, int sg) {                      // thread 0: stage source row k of the band
        const int ysrc = min(max(yb - R + k, a.ylo), g.H - 1);
        uint8_t *sb = stg + (size_t)sg * a.stageBytes;
        mbar_expect_tx(&bars[sg], 6u * (uint32_t)NQh * 4u + leftBytes);
#pragma unroll
        for (int p = 0; p < 6; p++)
            bulk_g2s(sb + (size_t)p * NQh * 4, a.rpairs + ((size_t)p * g.H + ysrc) * 2 * a.RPW + iLo, (uint32_t)NQh * 4u, &bars[sg]);
        bulk_g2s(sb + a.rpBytes, a.leftX + ((size_t)ysrc * g.W + xa) * 2, leftBytes, &bars[sg]);
    };
    if (tid == 0)
        for (int k = 0; k < nstg && k < nsteps; k++) fill(k, k);

    // this thread: natural disparity-pair index w -> word position in the volume layout
    int pos;
    {
        const int l = w / g.nreg, i = w % g.nreg;
        pos = 4 * (g.lpc * (i >> 2) + l) + (i & 3);
    }
    const int wEff = min(w, HP - 1);                      // padding words read in range and are never stored
    const int x1s = x0 - R + xg * XPT;                    // image column (valid coordinates) of walked column 0
    const int qq0 = g.minX1 + x1s - g.minD - 1;           // its pair word serves disparities (0, 1); parity == PAR
    const int eOff = (qq0 >> 1) - iLo - wEff;             // E index of that word for this thread's disparities
    const int lOff = (x1s - xlo) * 2;                     // uint4 index of its left operands
    const int lb = max(0, -x1s), rlim = g.W1 - 1 - x1s;
    const bool border = __any_sync(0xFFFFFFFFu, lb > 0 || rlim < NCOL - 1);
    int nvalid = min(XPT, g.W1 - (x0 + xg * XPT));
    if (!act || w >= HP) nvalid = 0;
    uint32_t crun[XPT];
    {
        const uint32_t bias = (uint32_t)(NS * NS) * (COST3_K + COST3_K / 4) * 0x10001u;
#pragma unroll
        for (int n = 0; n < XPT; n++) crun[n] = 0u - bias;
    }
    const uint32_t ringBase = smem_u32(ring) + (uint32_t)tid * 4u, rstride = (uint32_t)nthr * 4u;
    uint32_t *outBase = reinterpret_cast<uint32_t *>(a.out) + ((size_t)(x0 + xg * XPT) * g.Dp) / 2 + pos;

    int sg = 0, slot = 0;
    uint32_t par = 0;
    for (int k = 0; k < nsteps; k++) {
        mbar_wait(&bars[sg], par);
        const uint8_t *sb = stg + (size_t)sg * a.stageBytes;
        const uint32_t *er = reinterpret_cast<const uint32_t *>(sb) + eOff;
        const uint4 *lrow = reinterpret_cast<const uint4 *>(sb + a.rpBytes) + lOff;
        const int yout = yb + k - 2 * R;
        const int nv = k >= 2 * R ? nvalid : 0;
        uint32_t *orow32 = outBase + (k >= 2 * R ? (size_t)(yout - a.y0) * (size_t)(g.rowStride / 2) : 0);
        const uint32_t rg = ringBase + (uint32_t)(slot * XPT) * rstride;
        if (border) cost3_walk<R, PAR, true, DWT, NTT>(lrow, er, NQh, rg, rstride, crun, orow32, Dw, nv, lb, rlim, a.one, a.neg1);
        else cost3_walk<R, PAR, false, DWT, NTT>(lrow, er, NQh, rg, rstride, crun, orow32, Dw, nv, 0, NCOL, a.one, a.neg1);
        __syncthreads();                                  // everybody is done with stage sg
        if (tid == 0 && k + nstg < nsteps) fill(k + nstg, sg);
        if (++sg == nstg) { sg = 0; par ^= 1u; }
        if (++slot == NS) slot = 0;
    }
}

static bool cost3_layout(Cost3Args &a, int R, size_t maxSmem, int threads, size_t *total)
{
    const Geo &g = a.g;
    const int TXH = a.NXG * COST3_XPT + 2 * R;
    a.NQh = (((TXH + g.D) / 2 + 8) + 3) & ~3;
    a.rpBytes = 6u * (unsigned)a.NQh * 4u;
    a.stageBytes = (a.rpBytes + (unsigned)TXH * 32u + 127u) & ~127u;
    size_t off = (size_t)(2 * R + 1) * COST3_XPT * threads * 4;
    off = (off + 127) & ~(size_t)127;
    a.stgOff = (unsigned)off;
    for (a.nstg = 4; a.nstg >= 2; a.nstg--) {
        size_t end = off + (size_t)a.nstg * a.stageBytes;
        a.barOff = (unsigned)end;
        end += 8 * 4;
        if (end <= maxSmem) { *total = end; return true; }
    }
    return false;
}

template <int R, int PAR, int DWT, int NTT>
static int launch_cost3_t(Cost3Args &a, int threads, size_t smem, dim3 grid, int maxSmem, cudaStream_t st)
{
    static bool attrDone = false;
    if (!attrDone) {
        SGBM_CUDA_CHECK(cudaFuncSetAttribute(k_cost3<R, PAR, DWT, NTT>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxSmem));
        attrDone = true;
    }
    k_cost3<R, PAR, DWT, NTT><<<grid, threads, smem, st>>>(a);
    sgbm_count_launch(1);
    SGBM_CUDA_CHECK(cudaGetLastError());
    return 0;
}

// Rows [y0, y0+nrows) of the cost volume with vertical clamp floor ylo, written at out (row y0 first).
// Returns 1 when the geometry does not fit this kernel (the caller falls back to sgbm_launch_cost2).
// The HH4 rule "rows y >= H - r carry C = 0" (A.9) is applied by the caller (memset of those rows).
int sgbm_launch_cost3(const Geo &g, const uint8_t *planes, uint16_t *out, int y0, int nrows, int ylo, cudaStream_t st)
{
    if (nrows <= 0) return 0;
    if (g.cn != 1 || g.r > 5) return 1;
    static int maxSmem = -1;
    if (maxSmem < 0) {
        int dev = 0;
        SGBM_CUDA_CHECK(cudaGetDevice(&dev));
        SGBM_CUDA_CHECK(cudaDeviceGetAttribute(&maxSmem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    }
    const int Dw = g.Dp / 2, R = g.r;
    if (Dw > 512) return 1;
    Cost3Args a;
    memset(&a, 0, sizeof(a));
    a.g = g;
    a.leftX = reinterpret_cast<const uint4 *>(planes + sgbm_cost2_leftx_offset(g));
    a.rpairs = reinterpret_cast<const uint32_t *>(planes + sgbm_cost2_right_offset(g));
    a.RPW = sgbm_cost2_rpw(g);
    a.out = out; a.y0 = y0; a.nrows = nrows; a.ylo = ylo;
    a.one = 1u; a.neg1 = 0xFFFFFFFFu;
    // thread groups along x: as many as fit (<= 512 threads, <= 8 groups, not more than the image needs)
    int NXG = 512 / Dw;
    if (NXG > 16) NXG = 16;
    const int need = (g.W1 + COST3_XPT - 1) / COST3_XPT;
    if (NXG > need) NXG = need;
    if (const char *e = getenv("SGBM_COST3_NXG")) { const int v = atoi(e); if (v >= 1 && v < NXG) NXG = v; }
    size_t smem = 0;
    int threads = 0;
    bool ok = false;
    for (; NXG >= 1; NXG--) {
        a.NXG = NXG;
        threads = ((Dw * NXG + 31) / 32) * 32;
        if (cost3_layout(a, R, (size_t)maxSmem, threads, &smem)) { ok = true; break; }
    }
    if (!ok) return 1;
    a.nact = Dw * NXG;
    a.RB = 64;
    if (const char *e = getenv("SGBM_COST3_RB")) { const int v = atoi(e); if (v >= 1) a.RB = v; }
    if (a.RB > nrows) a.RB = nrows;
    const int TX = NXG * COST3_XPT;
    // the staged right rows must stay inside the padded parity rows of the prefilter output
    if (((g.W - 1) >> 1) + a.NQh + 4 > a.RPW) return 1;
    dim3 grid((g.W1 + TX - 1) / TX, (nrows + a.RB - 1) / a.RB);
    const int par = (g.minX1 - R - g.minD - 1) & 1;       // parity of the first walked column's right position
    // blockSize 3 / 5 / 7 at the lane mappings of numDisparities = 128 / 192 / 256: compile-time strides
#define COST3_HOT(RR, DW_, NT_)                                                                                    \
    if (R == RR && Dw == DW_ && threads == NT_)                                                                    \
        return par ? launch_cost3_t<RR, 1, DW_, NT_>(a, threads, smem, grid, maxSmem, st)                          \
                   : launch_cost3_t<RR, 0, DW_, NT_>(a, threads, smem, grid, maxSmem, st);
#if COST3_XPT == 16
    COST3_HOT(1, 64, 512) COST3_HOT(1, 96, 480) COST3_HOT(1, 128, 512)
    COST3_HOT(2, 64, 512) COST3_HOT(2, 96, 480) COST3_HOT(2, 128, 512)
    COST3_HOT(3, 64, 512) COST3_HOT(3, 96, 480) COST3_HOT(3, 128, 512)
#else
    COST3_HOT(1, 64, 256) COST3_HOT(1, 96, 288) COST3_HOT(1, 128, 256)
    COST3_HOT(2, 64, 256) COST3_HOT(2, 96, 288) COST3_HOT(2, 128, 256)
    COST3_HOT(3, 64, 256) COST3_HOT(3, 96, 192) COST3_HOT(3, 128, 256)
#endif
#undef COST3_HOT
#define COST3_CASE(RR)                                                                                  \
    if (R == RR) return par ? launch_cost3_t<RR, 1, 0, 0>(a, threads, smem, grid, maxSmem, st)          \
                            : launch_cost3_t<RR, 0, 0, 0>(a, threads, smem, grid, maxSmem, st);
    COST3_CASE(0) COST3_CASE(1) COST3_CASE(2) COST3_CASE(3) COST3_CASE(4) COST3_CASE(5)
#undef COST3_CASE
    return 1;
}
