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
// C) horizontal sum, ring update, running sum, coalesced 32-bit stores of C.
// ------------------------------------------------------------------------------------------------
#define COST_THREADS 256
#define COST_NIT 16          // max packed items per thread: TX * Dp/2 <= COST_THREADS * COST_NIT
#define COST_K 256u          // bias; multiple of 4 so that (bt_t + K) >> 2 == (bt_t >> 2) + K/4

struct CostArgs {
    Geo g;
    const uint8_t *planes;   // k_prefilter output
    uint16_t *out;           // row y is written at out + (y - y0) * rowStride
    int y0, nrows;           // output rows [y0, y0 + nrows)
    int ylo;                 // vertical clamp floor (0, or the stripe start for 3WAY)
    int TX, RB;              // tile width (valid columns), rows per band
    int zeroTail;            // HH4 quirk (A.9): rows y >= H - r get C = 0
};

__global__ void __launch_bounds__(COST_THREADS) k_cost(CostArgs a)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const Geo &g = a.g;
    const int r = g.r, TX = a.TX, TXH = TX + 2 * r, D = g.D, Dp = g.Dp, Dw = Dp / 2;   // Dw: u32 words/column
    const int x0 = blockIdx.x * TX;                      // first valid column of the tile
    const int yb = a.y0 + blockIdx.y * a.RB;             // first output row of the band
    const int yend = min(yb + a.RB, a.y0 + a.nrows);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = COST_THREADS / 32;
    const int cn = g.cn;

    // image-coordinate range of the tile incl. halo (clamped to the valid range)
    const int xa = g.minX1 + min(max(x0 - r, 0), g.W1 - 1);
    const int xb = g.minX1 + min(max(x0 + TX - 1 + r, 0), g.W1 - 1);
    const int q0 = xa - g.maxD + 1;                      // first right-image column needed (pair base)
    const int NQ = (xb - xa) + D - 1;                    // pair entries q0 .. q0+NQ-1
    const int NQh = (NQ + 1) / 2 + 1;                    // entries per parity array

    // shared memory carve-up
    uint32_t *ring = reinterpret_cast<uint32_t *>(smem);                     // [(2r+1)][TX][Dw]
    uint32_t *pixbuf = ring + (size_t)(2 * r + 1) * TX * Dw;                  // [TXH][Dw]
    uint32_t *rp = pixbuf + (size_t)TXH * Dw;                                 // [cn][6][2][NQh]
    uint8_t *lv = reinterpret_cast<uint8_t *>(rp + (size_t)cn * 6 * 2 * NQh); // [cn][6][TXH]

    const size_t plane = (size_t)g.W * g.H;
    const int nslots = 2 * r + 1;
    for (int i = tid; i < nslots * TX * Dw; i += COST_THREADS) ring[i] = 0;

    uint32_t crun[COST_NIT];
#pragma unroll
    for (int n = 0; n < COST_NIT; n++) crun[n] = 0;
    const int nitems = TX * Dw;
    const uint32_t KK = COST_K * 0x10001u;

    const int nsteps = (yend - yb) + 2 * r;
    for (int k = 0; k < nsteps; k++) {
        const int ysrc = min(max(yb - r + k, a.ylo), g.H - 1);
        // ---- A: stage this source row --------------------------------------------------------
        for (int c = 0; c < cn; c++) {
            const uint8_t *pl = a.planes + (size_t)(0 * cn + c) * 6 * plane + (size_t)ysrc * g.W;
            const uint8_t *pr = a.planes + (size_t)(1 * cn + c) * 6 * plane + (size_t)ysrc * g.W;
            for (int i = tid; i < 6 * TXH; i += COST_THREADS) {
                int p = i / TXH, xx = i % TXH;
                int x = g.minX1 + min(max(x0 - r + xx, 0), g.W1 - 1);
                lv[(c * 6 + p) * TXH + xx] = pl[(size_t)p * plane + x];
            }
            for (int i = tid; i < 6 * NQ; i += COST_THREADS) {
                int p = i / NQ, qi = i % NQ;
                int q = q0 + qi;                                      // 0 <= q, q+1 <= W-1 (A.2 range)
                const uint8_t *src = pr + (size_t)p * plane + q;
                uint32_t v = (uint32_t)src[1] | ((uint32_t)src[0] << 16);   // lo = v(q+1), hi = v(q)
                rp[((c * 6 + p) * 2 + (qi & 1)) * NQh + (qi >> 1)] = v;
            }
        }
        __syncthreads();
        // ---- B: pixel costs for the TXH columns ----------------------------------------------
        for (int xx = warp; xx < TXH; xx += nwarp) {
            const int x = g.minX1 + min(max(x0 - r + xx, 0), g.W1 - 1);
            const int qtop = x - g.minD - 1 - q0;            // pair entry index for dr = 0
            const int par = qtop & 1;
            const int itop = qtop >> 1;
            for (int pi = lane; pi < D / 2; pi += 32) {
                uint32_t acc = 0;
                for (int c = 0; c < cn; c++) {
                    const uint8_t *lc = lv + (c * 6) * TXH + xx;
                    const uint32_t *rc = rp + (size_t)(c * 6) * 2 * NQh + par * NQh + (itop - pi);
                    uint32_t bt[2];
#pragma unroll
                    for (int p = 0; p < 2; p++) {
                        uint32_t u = lc[(3 * p + 0) * TXH], ulo = lc[(3 * p + 1) * TXH], uhi = lc[(3 * p + 2) * TXH];
                        uint32_t v2 = rc[(size_t)(3 * p + 0) * 2 * NQh];
                        uint32_t vlo2 = rc[(size_t)(3 * p + 1) * 2 * NQh];
                        uint32_t vhi2 = rc[(size_t)(3 * p + 2) * 2 * NQh];
                        uint32_t A = (u + COST_K) * 0x10001u - vhi2;          // u - vhi + K
                        uint32_t B = vlo2 + (COST_K - u) * 0x10001u;          // vlo - u + K
                        uint32_t c1 = __vimax3_u16x2(A, B, KK);
                        uint32_t A2 = v2 + (COST_K - uhi) * 0x10001u;         // v - uhi + K
                        uint32_t B2 = (ulo + COST_K) * 0x10001u - v2;         // ulo - v + K
                        uint32_t c2 = __vimax3_u16x2(A2, B2, KK);
                        bt[p] = __vminu2(c1, c2);                             // bt + K
                    }
                    // (bt_g + K) + ((bt_t + K) >> 2) - (K + K/4)
                    acc += bt[0] + ((bt[1] >> 2) & 0x3FFF3FFFu) - (COST_K + COST_K / 4) * 0x10001u;
                }
                const int l = pi / g.nreg, m = pi % g.nreg;
                pixbuf[(size_t)xx * Dw + 4 * (g.lpc * (m >> 2) + l) + (m & 3)] = acc;
            }
        }
        __syncthreads();
        // ---- C: horizontal sum, ring, running vertical sum, store ------------------------------
        const int slot = k % nslots;
        const int yout = yb + k - 2 * r;
        const bool emit = (k >= 2 * r);
        const bool zero = a.zeroTail && r > 0 && yout >= g.H - r;
        uint32_t *orow32 = reinterpret_cast<uint32_t *>(a.out) +
                           (emit ? ((size_t)(yout - a.y0) * g.rowStride + (size_t)x0 * Dp) / 2 : 0);
#pragma unroll
        for (int n = 0; n < COST_NIT; n++) {
            int it = tid + n * COST_THREADS;
            if (it < nitems) {
                int x = it / Dw, w = it - x * Dw;
                const uint32_t *pb = pixbuf + (size_t)x * Dw + w;
                uint32_t hs = 0;
                for (int i = 0; i <= 2 * r; i++) hs += pb[(size_t)i * Dw];
                uint32_t *rg = ring + ((size_t)slot * TX + x) * Dw + w;
                uint32_t old = *rg;
                *rg = hs;
                crun[n] = crun[n] + hs - old;
                if (emit && x0 + x < g.W1) orow32[it] = zero ? 0u : crun[n];
            }
        }
        // the next iteration's first __syncthreads orders ring/pixbuf reuse
    }
}

size_t sgbm_cost_smem_bytes(const Geo &g, int TX)
{
    int r = g.r, TXH = TX + 2 * r, Dw = g.Dp / 2;
    int NQ = TXH + g.D, NQh = (NQ + 1) / 2 + 1;
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

// Rows [y0, y0+nrows) of the cost volume with vertical clamp floor ylo, written at out (row y0 first).
int sgbm_launch_cost(const Geo &g, const uint8_t *planes, uint16_t *out, int y0, int nrows, int ylo,
                     int zeroTail, cudaStream_t st)
{
    if (nrows <= 0) return 0;
    static int maxSmem = -1;
    if (maxSmem < 0) {
        int dev = 0;
        SGBM_CUDA_CHECK(cudaGetDevice(&dev));
        SGBM_CUDA_CHECK(cudaDeviceGetAttribute(&maxSmem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        SGBM_CUDA_CHECK(cudaFuncSetAttribute(k_cost, cudaFuncAttributeMaxDynamicSharedMemorySize, maxSmem));
    }
    int TX = 32;
    while (TX > 1 && (sgbm_cost_smem_bytes(g, TX) > (size_t)maxSmem / 2 || TX * (g.Dp / 2) > COST_THREADS * COST_NIT))
        TX >>= 1;
    if (sgbm_cost_smem_bytes(g, TX) > (size_t)maxSmem || TX * (g.Dp / 2) > COST_THREADS * COST_NIT)
        return sgbm_fail(-3, "cost kernel: blockSize/numDisparities too large for shared memory (r=%d, Dp=%d)", g.r, g.Dp);
    CostArgs a;
    a.g = g; a.planes = planes; a.out = out; a.y0 = y0; a.nrows = nrows; a.ylo = ylo; a.TX = TX;
    a.RB = nrows < 64 ? nrows : 64;
    a.zeroTail = zeroTail;
    dim3 grid((g.W1 + TX - 1) / TX, (nrows + a.RB - 1) / a.RB);
    k_cost<<<grid, COST_THREADS, sgbm_cost_smem_bytes(g, TX), st>>>(a);
    sgbm_count_launch(1);
    SGBM_CUDA_CHECK(cudaGetLastError());
    return 0;
}
