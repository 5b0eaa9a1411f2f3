// sgbm_cost3.cu -- stage 2 of StereoSGBM.compute (main.ipynb:668) for sm_100a, third generation:
// Birchfield-Tomasi pixel cost + (2r+1)^2 block sum -> cost volume C, 1-channel input.
//
// Same arithmetic, tiling (TX columns x RB rows per CTA, rows staged by TMA bulk copies) and volume
// layout as sgbm_cost2.cu; what changed is who computes what.  k_cost2 computed the pixel costs of a
// row into a shared-memory buffer (one warp per column) and summed them in a second phase (16.6
// shared-memory wavefronts and 65 instructions per output word, two block barriers per row).  Here
// one thread owns TWO adjacent disparity pairs (packed u16x2 words 2j and 2j+1 = disparities 4j..4j+3)
// of XPT = 12 consecutive columns and walks along the row:
//   * the pixel cost never leaves registers: the horizontal (2r+1) window sum slides in a register
//     window (the walk is fully unrolled, so the rotating window has compile-time indices);
//   * right-image operands are a stream: the pair word of an odd right position is assembled with one
//     PRMT from the two neighbouring even pair words, so only the even-parity pair array is staged;
//     the operands of word 2j+1 at column x are those of word 2j at column x-2, so one aligned 64-bit
//     shared load per plane feeds FOUR columns of BOTH words (the prefilter shifts the array by one
//     word when needed so that every thread's stream starts 8-byte aligned);
//   * the left-image operands arrive pre-biased and pre-expanded to packed words (prefilter, 32 bytes
//     per pixel): two warp-uniform 128-bit loads per column, shared by the thread's two words;
//   * each of the four Birchfield-Tomasi differences is one IMAD (FMA pipe); the packed min / max /
//     permute instructions are ALU-pipe only and run at half rate, so the split keeps both pipes busy;
//   * the vertical running sum stays in registers; the ring of the last 2r+1 horizontal sums is
//     thread-private shared memory (one 64-bit load + store per two output words, no barrier);
//   * one 64-bit global store per two output words.
// One block barrier per row (stage hand-back).  Image borders (replicated pixel-cost columns, A.3) run a
// second instantiation of the walk; only the first / last thread groups of a row of tiles take it.
#include "sgbm_common.cuh"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define COST3_K 256u          // bias; multiple of 4 so that (bt_t + K) >> 2 == (bt_t >> 2) + K/4
#ifndef COST3_XPT
#define COST3_XPT 12
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
    int NQh, nstg, nact;     // staged pair words per plane, stages, active threads (NXG * Dw / 2)
    unsigned int stgOff, barOff, stageBytes, rpBytes;
    unsigned int one, neg1;  // 1 and 0xFFFFFFFF: opaque multipliers (see fma_mad)
    int eshift;              // words the prefilter shifted the even pair array by (sgbm_cost3_eshift)
};

// ring accesses go through volatile asm WITHOUT a memory clobber: the entries are thread-private, and
// the compiler stays free to hoist the (ordinary) loads of the next columns' operands above them
__device__ __forceinline__ uint2 ring_ld(uint32_t addr)
{
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void ring_st(uint32_t addr, uint32_t x, uint32_t y)
{
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(x), "r"(y));
}

// The left operands arrive pre-biased (uK = u + K, KmU = K - u, KmUhi = K - u_hi, UloK = u_lo + K), so each
// of the four differences is ONE multiply-add a * (+-1) + c.  The multipliers are kernel arguments the
// compiler cannot fold: that forces IMAD, which issues on the FMA pipe.
// Birchfield-Tomasi cost of one plane for two adjacent disparities, biased by K (A.2):
//   min( max(u - vhi, vlo - u, 0), max(v - uhi, ulo - v, 0) ) + K
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
//   er2  : the thread's right-operand stream T[-1], T[0], T[1], ... as aligned pairs: er2[p*NQh2 + i] =
//          (T[2i-1], T[2i]) of plane p, where T[m] = E[k0 + m - 2j] (even pair words, k0 = column 0's)
//   rg   : shared address (bytes) of this thread's ring entry of column 0 in the current slot
//   lb   : BORDER: walked columns [0, lb) lie left of the image (their cost is that of column lb)
//   rlim : BORDER: walked columns > rlim lie right of the image (cost of column rlim)
//   DWT / NTT : compile-time words per column / threads per CTA (0 = run-time): the global stores and the
//               ring accesses of the hot geometries then use immediate offsets
template <int R, int PAR, bool BORDER, int DWT, int NTT>
__device__ __forceinline__ void cost3_walk(const uint4 *__restrict__ lrow, const uint2 *__restrict__ er2, int NQh2,
                                           uint32_t rg, uint32_t rstrideRt, uint32_t (&crun0)[COST3_XPT],
                                           uint32_t (&crun1)[COST3_XPT], uint2 *orow, int DwRt, int nvalid, int lb, int rlim,
                                           uint32_t one, uint32_t neg1)
{
    constexpr int XPT = COST3_XPT, NS = 2 * R + 1, NCOL = XPT + 2 * R;
    constexpr int MMAX = (NCOL + PAR) / 2;                // largest stream index the walk touches
    constexpr int NPAIR = (MMAX + 1) / 2 + 1;             // aligned pairs i = 0 .. (MMAX+1)/2
    constexpr int LA = 3;                                 // pairs are loaded LA columns before their first use
    const int Dh = DWT ? DWT / 2 : DwRt / 2;              // uint2 per column of the volume
    const uint32_t rstride = NTT ? (uint32_t)NTT * 8u : rstrideRt;
    uint32_t T[6][2 * NPAIR];                             // T[p][m + 1]
    uint32_t Vh[3][6];                                    // operands of columns c, c-1, c-2
    uint32_t win0[NS], win1[NS];
    uint32_t hs0 = 0, hs1 = 0, last0 = 0, last1 = 0;
#pragma unroll
    for (int i = 0; i < NPAIR; i++) {                     // pairs needed before the loop's prefetch window opens
        const int need = (4 * i - 3 - PAR) > -2 ? (4 * i - 3 - PAR) : -2;
        if (need - LA <= -2) {
#pragma unroll
            for (int p = 0; p < 6; p++) {
                const uint2 q = er2[p * NQh2 + i];
                T[p][2 * i] = q.x; T[p][2 * i + 1] = q.y;
            }
        }
    }
#pragma unroll
    for (int c = -2; c < NCOL; c++) {
#pragma unroll
        for (int i = 0; i < NPAIR; i++) {
            const int need = (4 * i - 3 - PAR) > -2 ? (4 * i - 3 - PAR) : -2;
            if (need - LA == c && need - LA > -2) {
#pragma unroll
                for (int p = 0; p < 6; p++) {
                    const uint2 q = er2[p * NQh2 + i];
                    T[p][2 * i] = q.x; T[p][2 * i + 1] = q.y;
                }
            }
        }
        uint32_t (&V)[6] = Vh[(c + 3) % 3];
        if ((c & 1) == PAR) {                             // even right position: the staged word itself
            const int m = (c + PAR) / 2;
#pragma unroll
            for (int p = 0; p < 6; p++) V[p] = T[p][m + 1];
        } else {                                          // odd: (E[k+1].hi, E[k].lo)
            const int mm = (c + 1 + PAR) / 2;
#pragma unroll
            for (int p = 0; p < 6; p++) V[p] = __byte_perm(T[p][mm + 1], T[p][mm], 0x5432);
        }
        if (c < 0) continue;
        const int xx = c;
        const uint32_t (&V1)[6] = Vh[(c + 1) % 3];        // word 2j+1: the operands of column c-2
        const uint4 l0 = lrow[2 * xx], l1 = lrow[2 * xx + 1];
        // (bt_g + K) + ((bt_t + K) >> 2): biased by K + K/4 per half; the bias cancels in the ring update
        // and is removed from the running sum once (initial value of crun)
        uint32_t pix0, pix1;
        {
            const uint32_t btg = bt3(l0.x, l0.y, l0.z, l0.w, V[0], V[1], V[2], one, neg1);
            const uint32_t btt = bt3(l1.x, l1.y, l1.z, l1.w, V[3], V[4], V[5], one, neg1);
            pix0 = fma_mad((btt >> 2) & 0x3FFF3FFFu, one, btg);
        }
        {
            const uint32_t btg = bt3(l0.x, l0.y, l0.z, l0.w, V1[0], V1[1], V1[2], one, neg1);
            const uint32_t btt = bt3(l1.x, l1.y, l1.z, l1.w, V1[3], V1[4], V1[5], one, neg1);
            pix1 = fma_mad((btt >> 2) & 0x3FFF3FFFu, one, btg);
        }
        if (BORDER) {
            if (xx > rlim) { pix0 = last0; pix1 = last1; } else { last0 = pix0; last1 = pix1; }
            if (xx < lb) { pix0 = 0; pix1 = 0; }
        }
        if (xx >= NS) { hs0 += pix0 - win0[xx % NS]; hs1 += pix1 - win1[xx % NS]; }
        else { hs0 += pix0; hs1 += pix1; }
        win0[xx % NS] = pix0; win1[xx % NS] = pix1;
        if (BORDER && xx <= R && xx > 0) {
            if (xx == lb) {                               // first column inside the image: replicate it to the left
#pragma unroll
                for (int j = 0; j < xx; j++) { win0[j] = pix0; win1[j] = pix1; }
                hs0 += (uint32_t)xx * pix0; hs1 += (uint32_t)xx * pix1;
            }
        }
        if (xx >= 2 * R) {
            const int n = xx - 2 * R;
            const uint32_t ra = rg + (uint32_t)n * rstride;
            const uint2 old = ring_ld(ra);
            ring_st(ra, hs0, hs1);
            crun0[n] += hs0 - old.x;
            crun1[n] += hs1 - old.y;
            if (BORDER ? n < nvalid : nvalid > 0) {
                if (DWT) orow[n * (DWT / 2)] = make_uint2(crun0[n], crun1[n]);
                else *reinterpret_cast<uint2 *>(reinterpret_cast<char *>(orow) + (uint32_t)n * (uint32_t)Dh * 8u) = make_uint2(crun0[n], crun1[n]);
            }
        }
    }
}

template <int R, int PAR, int DWT, int NTT>
__global__ void __launch_bounds__(NTT ? NTT : 448) k_cost3(Cost3Args a)
{
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int XPT = COST3_XPT, NS = 2 * R + 1, NCOL = XPT + 2 * R;
    const Geo &g = a.g;
    const int Dw = DWT ? DWT : g.Dp / 2, Dh = Dw / 2, HP = g.D / 2;
    const int TX = a.NXG * XPT;
    const int x0 = blockIdx.x * TX;                      // first valid column of the tile
    const int yb = a.y0 + blockIdx.y * a.RB;             // first output row of the band
    const int yend = min(yb + a.RB, a.y0 + a.nrows);
    const int tid = threadIdx.x, nthr = NTT ? NTT : blockDim.x;
    const int nstg = a.nstg, NQh = a.NQh;
    const bool act = tid < a.nact;
    const int xg = act ? tid / Dh : 0, w0 = act ? 2 * (tid % Dh) : 0;   // natural words w0, w0 + 1

    // staged column range of the tile (valid columns, clamped to the image) and the right-image window
    const int xlo = max(x0 - R, 0), xhi = min(x0 + TX + R, g.W1);
    const int xa = g.minX1 + xlo;
    const int q0 = xa - g.maxD + 1;                      // first right-image pixel the tile touches
    const int iLo = (q0 >> 1) & ~3;                      // first staged pair word (16-byte aligned)
    const uint32_t leftBytes = (uint32_t)(xhi - xlo) * 32u;

    uint2 *ring = reinterpret_cast<uint2 *>(smem);                            // [NS][XPT][nthr]
    uint8_t *stg = smem + a.stgOff;                                           // [nstg] { E [6][NQh] u32 ; left [TX+2R] 32 B }
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + a.barOff);
    for (int i = tid; i < NS * XPT * nthr; i += nthr) ring[i] = make_uint2(0u, 0u);
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

    // natural disparity-pair index -> word position in the volume layout (w0 even: w0 + 1 is the next word)
    int pos;
    {
        const int l = w0 / g.nreg, i = w0 % g.nreg;
        pos = 4 * (g.lpc * (i >> 2) + l) + (i & 3);
    }
    const int wEff = min(w0, HP - 2);                     // padding words read in range and are never stored
    const int x1s = x0 - R + xg * XPT;                    // image column (valid coordinates) of walked column 0
    const int qq0 = g.minX1 + x1s - g.minD - 1;           // its pair word serves disparities (0, 1); parity == PAR
    // staged word index of T[-1] = E[k0 - w0 - 1]; even by construction (a.eshift, see sgbm_cost3_eshift)
    const int s1 = (qq0 >> 1) + a.eshift - iLo - wEff - 1;
    const int lOff = (x1s - xlo) * 2;                     // uint4 index of the first walked column's left operands
    const int lb = max(0, -x1s), rlim = g.W1 - 1 - x1s;
    const bool border = __any_sync(0xFFFFFFFFu, lb > 0 || rlim < NCOL - 1);
    int nvalid = min(XPT, g.W1 - (x0 + xg * XPT));
    if (!act || w0 >= HP) nvalid = 0;
    uint32_t crun0[XPT], crun1[XPT];
    {
        const uint32_t bias = (uint32_t)(NS * NS) * (COST3_K + COST3_K / 4) * 0x10001u;
#pragma unroll
        for (int n = 0; n < XPT; n++) { crun0[n] = 0u - bias; crun1[n] = 0u - bias; }
    }
    const uint32_t ringBase = smem_u32(ring) + (uint32_t)tid * 8u, rstride = (uint32_t)nthr * 8u;
    uint2 *outBase = reinterpret_cast<uint2 *>(a.out) + ((size_t)(x0 + xg * XPT) * g.Dp) / 4 + pos / 2;

    int sg = 0, slot = 0;
    uint32_t par = 0;
    for (int k = 0; k < nsteps; k++) {
        mbar_wait(&bars[sg], par);
        const uint8_t *sb = stg + (size_t)sg * a.stageBytes;
        const uint2 *er2 = reinterpret_cast<const uint2 *>(sb) + (s1 >> 1);
        const uint4 *lrow = reinterpret_cast<const uint4 *>(sb + a.rpBytes) + lOff;
        const int yout = yb + k - 2 * R;
        const int nv = k >= 2 * R ? nvalid : 0;
        uint2 *orow = outBase + (k >= 2 * R ? (size_t)(yout - a.y0) * (size_t)(g.rowStride / 4) : 0);
        const uint32_t rg = ringBase + (uint32_t)(slot * XPT) * rstride;
        if (border) cost3_walk<R, PAR, true, DWT, NTT>(lrow, er2, NQh / 2, rg, rstride, crun0, crun1, orow, Dw, nv, lb, rlim, a.one, a.neg1);
        else cost3_walk<R, PAR, false, DWT, NTT>(lrow, er2, NQh / 2, rg, rstride, crun0, crun1, orow, Dw, nv, 0, NCOL, a.one, a.neg1);
        // One block barrier per row hands the stage back.  (Per-stage "empty" mbarriers instead were measured
        // SLOWER, 2.06 -> 2.20 ms at 4K D=256: warps that drift apart execute different parts of the 20 KB
        // straight-line walk and miss in the instruction cache.)
        __syncthreads();
        if (tid == 0 && k + nstg < nsteps) fill(k + nstg, sg);
        if (++sg == nstg) { sg = 0; par ^= 1u; }
        if (++slot == NS) slot = 0;
    }
}

static bool cost3_layout(Cost3Args &a, int R, size_t maxSmem, int threads, int minStages, size_t *total)
{
    const Geo &g = a.g;
    const int TXH = a.NXG * COST3_XPT + 2 * R;
    a.NQh = (((TXH + g.D) / 2 + 10) + 3) & ~3;
    a.rpBytes = 6u * (unsigned)a.NQh * 4u;
    a.stageBytes = (a.rpBytes + (unsigned)TXH * 32u + 127u) & ~127u;
    size_t off = (size_t)(2 * R + 1) * COST3_XPT * threads * 8;
    off = (off + 127) & ~(size_t)127;
    a.stgOff = (unsigned)off;
    for (a.nstg = 4; a.nstg >= minStages; a.nstg--) {
        size_t end = off + (size_t)a.nstg * a.stageBytes;
        a.barOff = (unsigned)end;
        end += 8 * 4;
        if (end <= maxSmem) { *total = end; return true; }
    }
    return false;
}

// Words by which k_prefilter2 shifts the even pair array so that every thread's operand stream starts on
// an 8-byte boundary: the staged index of T[-1] is k0 + shift - iLo - 2j - 1 with k0 = (first walked
// column's right position) >> 1, whose parity is the same for every tile and thread group (tile widths
// and XPT are multiples of 4 columns), and iLo a multiple of 4.
int sgbm_cost3_eshift(const Geo &g) { return (((g.minX1 - g.r - g.minD - 1) >> 1) + 1) & 1; }

// Geometry of the launch; false when this kernel does not hold the geometry (caller uses sgbm_cost2.cu).
static bool cost3_plan(const Geo &g, int maxSmem, Cost3Args &a, int *threadsOut, size_t *smemOut)
{
    if (g.cn != 1 || g.r > 5 || (g.D & 3)) return false;
    const int Dw = g.Dp / 2, Dh = Dw / 2, R = g.r;
    if (Dh > 448) return false;
    memset(&a, 0, sizeof(a));
    a.g = g;
    a.RPW = sgbm_cost2_rpw(g);
    a.one = 1u; a.neg1 = 0xFFFFFFFFu;
    a.eshift = sgbm_cost3_eshift(g);
    // thread groups along x: as many as fit (<= 448 threads, <= 16 groups, not more than the image needs)
    int cap = 448 / Dh;
    if (cap > 16) cap = 16;
    if (cap < 1) cap = 1;
    const int need = (g.W1 + COST3_XPT - 1) / COST3_XPT;
    if (cap > need) cap = need;
    if (const int v = sgbm_knobs().cost3NXG) { if (v >= 1 && v < cap) cap = v; }
    for (int minStages = 3; minStages >= 2; minStages--)
        for (int NXG = cap; NXG >= 1; NXG--) {
            a.NXG = NXG;
            const int threads = ((Dh * NXG + 31) / 32) * 32;
            if (cost3_layout(a, R, (size_t)maxSmem, threads, minStages, smemOut)) {
                a.nact = Dh * NXG;
                *threadsOut = threads;
                // the staged right rows must stay inside the padded parity rows of the prefilter output
                return ((g.W - 1) >> 1) + a.NQh + 6 <= a.RPW;
            }
        }
    return false;
}

static int cost3_max_smem(int *out)
{
    *out = sgbm_knobs().maxSmemOptin;
    return 0;
}

// 1 when k_cost3 holds this geometry (then the prefilter must apply sgbm_cost3_eshift), 0 when not, < 0 on error
int sgbm_cost3_supported(const Geo &g)
{
    int maxSmem = 0, rc = cost3_max_smem(&maxSmem);
    if (rc) return rc;
    Cost3Args a;
    int threads = 0;
    size_t smem = 0;
    return cost3_plan(g, maxSmem, a, &threads, &smem) ? 1 : 0;
}

template <int R, int PAR, int DWT, int NTT>
static int launch_cost3_t(Cost3Args &a, int threads, size_t smem, dim3 grid, int maxSmem, cudaStream_t st)
{
    static unsigned long long attrDone = 0;   // one bit per device: function attributes are per device
    {
        SgbmDeviceOnce once(attrDone);
        if (once.first) {
            SGBM_CUDA_CHECK(cudaFuncSetAttribute(k_cost3<R, PAR, DWT, NTT>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxSmem));
            once.done();
        }
    }
    k_cost3<R, PAR, DWT, NTT><<<grid, threads, smem, st>>>(a);
    sgbm_count_launch(1);
    SGBM_CUDA_CHECK(cudaGetLastError());
    return 0;
}

// Rows [y0, y0+nrows) of the cost volume with vertical clamp floor ylo, written at out (row y0 first).
// Returns 1 when the geometry does not fit this kernel (see sgbm_cost3_supported).
// The HH4 rule "rows y >= H - r carry C = 0" (A.9) is applied by the caller (memset of those rows).
int sgbm_launch_cost3(const Geo &g, const uint8_t *planes, uint16_t *out, int y0, int nrows, int ylo, cudaStream_t st)
{
    if (nrows <= 0) return 0;
    int maxSmem = 0, rc = cost3_max_smem(&maxSmem);
    if (rc) return rc;
    Cost3Args a;
    int threads = 0;
    size_t smem = 0;
    if (!cost3_plan(g, maxSmem, a, &threads, &smem)) return 1;
    const int Dw = g.Dp / 2, R = g.r;
    a.leftX = reinterpret_cast<const uint4 *>(planes + sgbm_cost2_leftx_offset(g));
    a.rpairs = reinterpret_cast<const uint32_t *>(planes + sgbm_cost2_right_offset(g));
    a.out = out; a.y0 = y0; a.nrows = nrows; a.ylo = ylo;
    // rows per band: every CTA spends RB + 2R row steps (plus ~2 for set-up); pick the band height whose
    // number of waves (one CTA per SM) times that cost is smallest
    const int TX = a.NXG * COST3_XPT;
    const int tilesX = (g.W1 + TX - 1) / TX;
    const int numSMs = sgbm_knobs().numSMs;
    {
        long long best = -1;
        for (int rb = 16; rb <= 192; rb++) {
            if (rb > nrows && rb != 16) break;
            const int rbe = rb < nrows ? rb : nrows;
            const int bands = (nrows + rbe - 1) / rbe;
            const long long waves = ((long long)tilesX * bands + numSMs - 1) / numSMs;
            const long long cost = waves * (rbe + 2 * R + 2);
            if (best < 0 || cost < best) { best = cost; a.RB = rbe; }
        }
    }
    if (const int v = sgbm_knobs().cost3RB) { if (v >= 1) a.RB = v; }
    if (a.RB > nrows) a.RB = nrows;
    dim3 grid(tilesX, (nrows + a.RB - 1) / a.RB);
    if (sgbm_knobs().cost3Pad > 0 && smem + (size_t)sgbm_knobs().cost3Pad <= (size_t)maxSmem) smem += (size_t)sgbm_knobs().cost3Pad;
    const int par = (g.minX1 - R - g.minD - 1) & 1;       // parity of the first walked column's right position
    if (sgbm_knobs().verbose)
        fprintf(stderr, "cost3: R=%d Dw=%d NXG=%d threads=%d smem=%zu nstg=%d RB=%d grid=%dx%d par=%d eshift=%d\n", R, Dw, a.NXG,
                threads, smem, a.nstg, a.RB, grid.x, grid.y, par, a.eshift);
    // blockSize 3 / 5 / 7 at the lane mappings of numDisparities = 128 / 192 / 256: compile-time strides
#define COST3_HOT(RR, DW_, NT_)                                                                                    \
    if (R == RR && Dw == DW_ && threads == NT_)                                                                    \
        return par ? launch_cost3_t<RR, 1, DW_, NT_>(a, threads, smem, grid, maxSmem, st)                          \
                   : launch_cost3_t<RR, 0, DW_, NT_>(a, threads, smem, grid, maxSmem, st);
#ifndef COST3_NO_HOT
    COST3_HOT(1, 64, 448) COST3_HOT(1, 96, 448) COST3_HOT(1, 128, 448)
    COST3_HOT(2, 64, 416) COST3_HOT(2, 96, 384) COST3_HOT(2, 128, 384)
    COST3_HOT(3, 64, 288) COST3_HOT(3, 96, 288) COST3_HOT(3, 128, 256)
#endif
#undef COST3_HOT
#define COST3_CASE(RR)                                                                                  \
    if (R == RR) return par ? launch_cost3_t<RR, 1, 0, 0>(a, threads, smem, grid, maxSmem, st)          \
                            : launch_cost3_t<RR, 0, 0, 0>(a, threads, smem, grid, maxSmem, st);
    COST3_CASE(0) COST3_CASE(1) COST3_CASE(2) COST3_CASE(3) COST3_CASE(4) COST3_CASE(5)
#undef COST3_CASE
    return 1;
}
