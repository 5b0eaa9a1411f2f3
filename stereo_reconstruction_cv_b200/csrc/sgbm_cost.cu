// sgbm_cost.cu -- stages 1+2 of StereoSGBM.compute (main.ipynb:668) for sm_100a:
//   k_prefilter : x-Sobel prefilter + preFilterCap clip, raw plane, half-sample intervals (A.1, A.2)
//   k_cost      : Birchfield-Tomasi pixel cost + (2r+1)^2 block sum -> cost volume C  (A.2, A.3)
// All arithmetic is exact integer; the volume layout is described in sgbm_common.cuh.
#include "sgbm_common.cuh"

// ------------------------------------------------------------------------------------------------
// Prefilter planes.  For every image (0 = left, 1 = right) and channel c, six u8 planes of H*W:
//   0: g   1: g_lo   2: g_hi   3: t   4: t_lo   5: t_hi
// g = clipped x-Sobel + ftzero (wraps to u8), t = raw intensity; both are forced to ftzero in
// columns 0 and W-1 (A.1).  lo/hi = min/max over {p, (p+p[x-1])/2, (p+p[x+1])/2} (A.2).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int pf_g(const uint8_t *img, long long pitch, int cn, int c, int W, int H, int x,
                                    int y, int ftzero)
{
    if (x <= 0 || x >= W - 1) return ftzero & 0xFF;
    const uint8_t *r0 = img + (long long)y * pitch;
    const uint8_t *rm = img + (long long)max(y - 1, 0) * pitch;
    const uint8_t *rp = img + (long long)min(y + 1, H - 1) * pitch;
    int xa = (x + 1) * cn + c, xb = (x - 1) * cn + c;
    int v = 2 * ((int)r0[xa] - (int)r0[xb]) + ((int)rm[xa] - (int)rm[xb]) + ((int)rp[xa] - (int)rp[xb]);
    v = min(max(v, -ftzero), ftzero) + ftzero;
    return v & 0xFF;
}
__device__ __forceinline__ int pf_t(const uint8_t *img, long long pitch, int cn, int c, int W, int x, int y,
                                    int ftzero)
{
    if (x <= 0 || x >= W - 1) return ftzero & 0xFF;
    return img[(long long)y * pitch + x * cn + c];
}

__global__ void k_prefilter(const uint8_t *left, const uint8_t *right, long long pitch, int W, int H, int cn,
                            int ftzero, uint8_t *planes)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    int ic = blockIdx.z;                 // image * cn + channel
    if (x >= W) return;
    int im = ic / cn, c = ic % cn;
    const uint8_t *img = im ? right : left;
    size_t plane = (size_t)W * H;
    uint8_t *out = planes + (size_t)ic * 6 * plane + (size_t)y * W + x;
    int g0 = pf_g(img, pitch, cn, c, W, H, x, y, ftzero);
    int t0 = pf_t(img, pitch, cn, c, W, x, y, ftzero);
    int glo = g0, ghi = g0, tlo = t0, thi = t0;
    if (x > 0) {
        int g1 = (g0 + pf_g(img, pitch, cn, c, W, H, x - 1, y, ftzero)) >> 1;
        int t1 = (t0 + pf_t(img, pitch, cn, c, W, x - 1, y, ftzero)) >> 1;
        glo = min(glo, g1); ghi = max(ghi, g1); tlo = min(tlo, t1); thi = max(thi, t1);
    }
    if (x < W - 1) {
        int g1 = (g0 + pf_g(img, pitch, cn, c, W, H, x + 1, y, ftzero)) >> 1;
        int t1 = (t0 + pf_t(img, pitch, cn, c, W, x + 1, y, ftzero)) >> 1;
        glo = min(glo, g1); ghi = max(ghi, g1); tlo = min(tlo, t1); thi = max(thi, t1);
    }
    out[0 * plane] = (uint8_t)g0;  out[1 * plane] = (uint8_t)glo; out[2 * plane] = (uint8_t)ghi;
    out[3 * plane] = (uint8_t)t0;  out[4 * plane] = (uint8_t)tlo; out[5 * plane] = (uint8_t)thi;
}

// ------------------------------------------------------------------------------------------------
// Cost volume.  One CTA owns TX valid columns x all disparities and walks down a band of rows,
// keeping the running vertical block sum in registers and the last 2r+1 horizontal sums in a
// shared-memory ring:   C(y) = C(y-1) + hsum(y+r) - hsum(y-r-1)     (A.3, clamped rows/columns).
// Per source row:  A) stage left/right prefiltered values in shared memory (right side as packed
// reversed pairs so that one 32-bit load yields the operands of two adjacent disparities),
// B) pixel costs for TX+2r columns (packed u16x2, biased by K so differences stay non-negative),
//    one warp per column, lanes over disparity pairs,
// C) each thread owns one 32-bit word (two disparities) of XPT consecutive columns: sliding
//    horizontal window sum, ring update, running vertical sum, coalesced 32-bit stores of C.
// blockDim = Dw * NXG with Dw = Dp/2 words per column; TX = NXG * XPT.
// ------------------------------------------------------------------------------------------------
#define COST_K 256u          // bias; multiple of 4 so that (bt_t + K) >> 2 == (bt_t >> 2) + K/4

struct CostArgs {
    Geo g;
    const uint8_t *planes;   // k_prefilter output
    uint16_t *out;           // row y is written at out + (y - y0) * rowStride
    int y0, nrows;           // output rows [y0, y0 + nrows)
    int ylo;                 // vertical clamp floor (0, or the stripe start for 3WAY)
    int NXG, RB;             // thread groups along x, rows per band
    int zeroTail;            // HH4 quirk (A.9): rows y >= H - r get C = 0
};

template <int NREG, int XPT>
__global__ void __launch_bounds__(512) k_cost(CostArgs a)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const Geo &g = a.g;
    const int r = g.r, D = g.D, Dw = g.Dp / 2, HP = D / 2;
    const int TX = a.NXG * XPT, TXH = TX + 2 * r;
    const int x0 = blockIdx.x * TX;                      // first valid column of the tile
    const int yb = a.y0 + blockIdx.y * a.RB;             // first output row of the band
    const int yend = min(yb + a.RB, a.y0 + a.nrows);
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;
    const int cn = g.cn;

    // image-coordinate range of the tile incl. halo (clamped to the valid range)
    const int xa = g.minX1 + min(max(x0 - r, 0), g.W1 - 1);
    const int xb = g.minX1 + min(max(x0 + TX - 1 + r, 0), g.W1 - 1);
    const int q0 = xa - g.maxD + 1;                      // first right-image column needed (pair base)
    const int NQ = (xb - xa) + D - 1;                    // pair entries q0 .. q0+NQ-1
    const int NQh = (TXH + D) / 2 + 2;                   // entries per parity array (tile independent)

    // shared memory carve-up
    uint32_t *ring = reinterpret_cast<uint32_t *>(smem);                     // [(2r+1)][TX][Dw]
    uint32_t *pixbuf = ring + (size_t)(2 * r + 1) * TX * Dw;                  // [TXH][Dw]
    uint32_t *rp = pixbuf + (size_t)TXH * Dw;                                 // [cn][6][2][NQh]
    uint8_t *lv = reinterpret_cast<uint8_t *>(rp + (size_t)cn * 6 * 2 * NQh); // [cn][6][TXH]

    const size_t plane = (size_t)g.W * g.H;
    const int nslots = 2 * r + 1;
    for (int i = tid; i < nslots * TX * Dw; i += nthr) ring[i] = 0;

    // phase C ownership: word w of columns xg*XPT .. xg*XPT+XPT-1
    const int w = tid % Dw, xg = tid / Dw;
    uint32_t crun[XPT];
#pragma unroll
    for (int n = 0; n < XPT; n++) crun[n] = 0;
    const uint32_t KK = COST_K * 0x10001u;
    const uint32_t KSUB = (COST_K + COST_K / 4) * 0x10001u;
    const int planeStrideW = 2 * NQh;                    // words between right-side planes

    const int nsteps = (yend - yb) + 2 * r;
    for (int k = 0; k < nsteps; k++) {
        const int ysrc = min(max(yb - r + k, a.ylo), g.H - 1);
        // ---- A: stage this source row --------------------------------------------------------
        for (int c = 0; c < cn; c++) {
            const uint8_t *pl = a.planes + (size_t)(0 * cn + c) * 6 * plane + (size_t)ysrc * g.W;
            const uint8_t *pr = a.planes + (size_t)(1 * cn + c) * 6 * plane + (size_t)ysrc * g.W + q0;
            for (int xx = tid; xx < TXH; xx += nthr) {
                const int x = g.minX1 + min(max(x0 - r + xx, 0), g.W1 - 1);
#pragma unroll
                for (int p = 0; p < 6; p++) lv[(c * 6 + p) * TXH + xx] = pl[(size_t)p * plane + x];
            }
            for (int qi = tid; qi < NQ; qi += nthr) {                       // 0 <= q, q+1 <= W-1 (A.2 range)
                uint32_t *dst = rp + (size_t)(c * 6) * planeStrideW + (qi & 1) * NQh + (qi >> 1);
#pragma unroll
                for (int p = 0; p < 6; p++) {
                    const uint8_t *src = pr + (size_t)p * plane + qi;
                    dst[p * planeStrideW] = (uint32_t)src[1] | ((uint32_t)src[0] << 16);   // lo = v(q+1), hi = v(q)
                }
            }
        }
        __syncthreads();
        // ---- B: pixel costs for the TXH columns ----------------------------------------------
        for (int c = 0; c < cn; c++) {
            for (int xx = warp; xx < TXH; xx += nwarp) {
                const int x = g.minX1 + min(max(x0 - r + xx, 0), g.W1 - 1);
                const int qtop = x - g.minD - 1 - q0;            // pair entry index for dr = 0
                const uint8_t *lc = lv + (c * 6) * TXH + xx;
                uint32_t uK[2], Ku[2], KmUhi[2], UloK[2];
#pragma unroll
                for (int p = 0; p < 2; p++) {
                    const uint32_t u = lc[(3 * p + 0) * TXH], ulo = lc[(3 * p + 1) * TXH], uhi = lc[(3 * p + 2) * TXH];
                    uK[p] = (u + COST_K) * 0x10001u; Ku[p] = (COST_K - u) * 0x10001u;
                    KmUhi[p] = (COST_K - uhi) * 0x10001u; UloK[p] = (ulo + COST_K) * 0x10001u;
                }
                const uint32_t *rc = rp + (size_t)(c * 6) * planeStrideW + (qtop & 1) * NQh + (qtop >> 1) - lane;
                uint32_t *pcol = pixbuf + (size_t)xx * Dw;
                for (int pi = lane; pi < HP; pi += 32, rc -= 32) {
                    uint32_t bt[2];
#pragma unroll
                    for (int p = 0; p < 2; p++) {
                        const uint32_t v2 = rc[(3 * p + 0) * planeStrideW];
                        const uint32_t vlo2 = rc[(3 * p + 1) * planeStrideW];
                        const uint32_t vhi2 = rc[(3 * p + 2) * planeStrideW];
                        const uint32_t c1 = __vimax3_u16x2(uK[p] - vhi2, vlo2 + Ku[p], KK);      // max(u-vhi, vlo-u, 0) + K
                        const uint32_t c2 = __vimax3_u16x2(v2 + KmUhi[p], UloK[p] - v2, KK);     // max(v-uhi, ulo-v, 0) + K
                        bt[p] = __vminu2(c1, c2);
                    }
                    // (bt_g + K) + ((bt_t + K) >> 2) - (K + K/4)
                    const uint32_t val = bt[0] + ((bt[1] >> 2) & 0x3FFF3FFFu) - KSUB;
                    const int l = pi / NREG, m = pi - l * NREG;
                    uint32_t *dst = pcol + 4 * (g.lpc * (m >> 2) + l) + (m & 3);
                    if (c == 0) *dst = val; else *dst += val;
                }
            }
        }
        __syncthreads();
        // ---- C: sliding horizontal sum, ring, running vertical sum, store ----------------------
        if (xg < a.NXG) {
            const int slot = k % nslots;
            const int yout = yb + k - 2 * r;
            const bool emit = (k >= 2 * r);
            const bool zero = a.zeroTail && r > 0 && yout >= g.H - r;
            const int xbase = xg * XPT;
            const uint32_t *pb = pixbuf + (size_t)xbase * Dw + w;            // column xbase-r of the tile (halo offset r)
            uint32_t *rg = ring + ((size_t)slot * TX + xbase) * Dw + w;
            uint32_t *orow32 = reinterpret_cast<uint32_t *>(a.out) +
                               (emit ? ((size_t)(yout - a.y0) * g.rowStride + (size_t)(x0 + xbase) * g.Dp) / 2 + w : 0);
            uint32_t hs = 0;
            for (int i = 0; i <= 2 * r; i++) hs += pb[(size_t)i * Dw];
#pragma unroll
            for (int n = 0; n < XPT; n++) {
                if (n > 0) hs += pb[(size_t)(n + 2 * r) * Dw] - pb[(size_t)(n - 1) * Dw];
                const uint32_t old = rg[(size_t)n * Dw];
                rg[(size_t)n * Dw] = hs;
                crun[n] = crun[n] + hs - old;
                if (emit && x0 + xbase + n < g.W1) orow32[(size_t)n * Dw] = zero ? 0u : crun[n];
            }
        }
        // the next iteration's first __syncthreads orders ring/pixbuf reuse
    }
}

static size_t cost_smem_bytes(const Geo &g, int TX)
{
    int r = g.r, TXH = TX + 2 * r, Dw = g.Dp / 2;
    int NQh = (TXH + g.D) / 2 + 2;
    size_t b = (size_t)(2 * r + 1) * TX * Dw * 4 + (size_t)TXH * Dw * 4 + (size_t)g.cn * 6 * 2 * NQh * 4 +
               (size_t)g.cn * 6 * TXH;
    return (b + 15) & ~(size_t)15;
}

int sgbm_launch_prefilter(const Geo &g, const uint8_t *left, const uint8_t *right, long long pitch,
                          uint8_t *planes, cudaStream_t st)
{
    dim3 grid((g.W + 255) / 256, g.H, 2 * g.cn);
    k_prefilter<<<grid, 256, 0, st>>>(left, right, pitch, g.W, g.H, g.cn, g.ftzero, planes);
    sgbm_count_launch(1);
    SGBM_CUDA_CHECK(cudaGetLastError());
    return 0;
}

template <int NREG, int XPT>
static int launch_cost_t(CostArgs &a, int threads, size_t smem, dim3 grid, int maxSmem, cudaStream_t st)
{
    static unsigned long long attrDone = 0;   // one bit per device: function attributes are per device
    {
        SgbmDeviceOnce once(attrDone);
        if (once.first) {
            SGBM_CUDA_CHECK(cudaFuncSetAttribute(k_cost<NREG, XPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxSmem));
            once.done();
        }
    }
    k_cost<NREG, XPT><<<grid, threads, smem, st>>>(a);
    sgbm_count_launch(1);
    SGBM_CUDA_CHECK(cudaGetLastError());
    return 0;
}

// Rows [y0, y0+nrows) of the cost volume with vertical clamp floor ylo, written at out (row y0 first).
int sgbm_launch_cost(const Geo &g, const uint8_t *planes, uint16_t *out, int y0, int nrows, int ylo,
                     int zeroTail, cudaStream_t st)
{
    if (nrows <= 0) return 0;
    const int maxSmem = sgbm_knobs().maxSmemOptin;
    const int Dw = g.Dp / 2;
    if (Dw > 512) return sgbm_fail(-3, "cost kernel: numDisparities too large (Dp=%d)", g.Dp);
    // threads = Dw * NXG (<= 512, >= 128 when possible); TX = NXG * XPT
    int NXG = 1;
    while (Dw * NXG * 2 <= 256) NXG *= 2;
    int threads = ((Dw * NXG + 31) / 32) * 32;
    int XPT = 16;
    while (XPT > 4 && cost_smem_bytes(g, NXG * XPT) > (size_t)maxSmem / 2) XPT >>= 1;
    while (NXG > 1 && cost_smem_bytes(g, NXG * XPT) > (size_t)maxSmem) { NXG >>= 1; threads = ((Dw * NXG + 31) / 32) * 32; }
    if (cost_smem_bytes(g, NXG * XPT) > (size_t)maxSmem)
        return sgbm_fail(-3, "cost kernel: blockSize/numDisparities too large for shared memory (r=%d, Dp=%d)", g.r, g.Dp);
    const int TX = NXG * XPT;
    CostArgs a;
    a.g = g; a.planes = planes; a.out = out; a.y0 = y0; a.nrows = nrows; a.ylo = ylo; a.NXG = NXG;
    a.RB = nrows < 64 ? nrows : 64;
    a.zeroTail = zeroTail;
    dim3 grid((g.W1 + TX - 1) / TX, (nrows + a.RB - 1) / a.RB);
    const size_t smem = cost_smem_bytes(g, TX);
#define COST_CASE(NR, XP) if (g.nreg == NR && XPT == XP) return launch_cost_t<NR, XP>(a, threads, smem, grid, maxSmem, st);
    COST_CASE(4, 16) COST_CASE(4, 8) COST_CASE(4, 4)
    COST_CASE(8, 16) COST_CASE(8, 8) COST_CASE(8, 4)
    COST_CASE(12, 16) COST_CASE(12, 8) COST_CASE(12, 4)
    COST_CASE(16, 16) COST_CASE(16, 8) COST_CASE(16, 4)
#undef COST_CASE
    return sgbm_fail(-3, "cost kernel: no instantiation for nreg=%d xpt=%d", g.nreg, XPT);
}
