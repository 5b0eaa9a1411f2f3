// sgbm_cost2.cu -- stages 1+2 of StereoSGBM.compute (main.ipynb:668) for sm_100a, second generation:
//   k_prefilter2 : x-Sobel prefilter + preFilterCap clip, raw plane, half-sample intervals (A.1, A.2).
//                  Left image: six u8 planes per channel with a 16-byte aligned row pitch.
//                  Right image: the same six planes pre-expanded to packed "pair words"
//                  (v(q+1) | v(q) << 16), split by the parity of q, so that the cost kernel stages
//                  them with plain TMA bulk copies and one 32-bit shared load yields the operands of
//                  two adjacent disparities.
//   k_cost2      : Birchfield-Tomasi pixel cost + (2r+1)^2 block sum -> cost volume C (A.2, A.3).
// Same arithmetic and volume layout as sgbm_cost.cu (exact integers; sgbm_common.cuh); what changed
// is the data movement: no per-row scalar staging loops (TMA bulk copies into a ring of stages, armed
// through mbarriers), pixel costs kept in natural disparity order with bank padding (no index math
// in the inner loop), 16 warps per CTA.
#include "sgbm_common.cuh"
#include <stdlib.h>
#include <string.h>

#define COST2_K 256u          // bias; multiple of 4 so that (bt_t + K) >> 2 == (bt_t >> 2) + K/4

// ---- geometry of the prefilter outputs (shared with the workspace layout in sgbm_api.cu) ------------
int sgbm_cost2_left_pitch(const Geo &g) { return ((g.W + 15) & ~15) + 16; }
int sgbm_cost2_rpw(const Geo &g) { return (g.W / 2 + g.D / 2 + g.r + 160 + 3) & ~3; }   // words per parity row
static size_t cost2_right_bytes(const Geo &g) { return (((size_t)g.cn * 6 * g.H * 2 * sgbm_cost2_rpw(g) * 4 + 256) + 255) & ~(size_t)255; }
// third section (sgbm_cost3.cu, 1-channel only): the left operands of every pixel pre-expanded to packed
// u16x2 words, 32 bytes per pixel: { u + K, K - u, K - u_hi, u_lo + K } for the gradient and the raw plane
static size_t cost2_leftx_bytes(const Geo &g) { return g.cn == 1 ? (size_t)g.H * g.W * 32 + 4096 : 0; }
size_t sgbm_cost2_right_offset(const Geo &g)
{
    const size_t left = (size_t)g.cn * 6 * g.H * sgbm_cost2_left_pitch(g) + 1024;   // + over-read of the last staged row
    return (left + 255) & ~(size_t)255;
}
size_t sgbm_cost2_leftx_offset(const Geo &g) { return sgbm_cost2_right_offset(g) + cost2_right_bytes(g); }
size_t sgbm_cost2_planes_bytes(const Geo &g) { return sgbm_cost2_leftx_offset(g) + cost2_leftx_bytes(g); }
static size_t cost2_right_offset(const Geo &g) { return sgbm_cost2_right_offset(g); }

__device__ __forceinline__ int pf2_g(const uint8_t *img, long long pitch, int cn, int c, int W, int H, int x, int y, int ftzero)
{
    if (x <= 0 || x >= W - 1) return ftzero & 0xFF;
    const uint8_t *r0 = img + (long long)y * pitch;
    const uint8_t *rm = img + (long long)max(y - 1, 0) * pitch;
    const uint8_t *rp = img + (long long)min(y + 1, H - 1) * pitch;
    const int xa = (x + 1) * cn + c, xb = (x - 1) * cn + c;
    int v = 2 * ((int)r0[xa] - (int)r0[xb]) + ((int)rm[xa] - (int)rm[xb]) + ((int)rp[xa] - (int)rp[xb]);
    v = min(max(v, -ftzero), ftzero) + ftzero;
    return v & 0xFF;
}
__device__ __forceinline__ int pf2_t(const uint8_t *img, long long pitch, int cn, int c, int W, int x, int y, int ftzero)
{
    if (x <= 0 || x >= W - 1) return ftzero & 0xFF;
    return img[(long long)y * pitch + x * cn + c];
}
// the six plane values of pixel x (x clamped into the row by the caller): g, glo, ghi, t, tlo, thi
__device__ __forceinline__ void pf2_six(const uint8_t *img, long long pitch, int cn, int c, int W, int H, int x, int y,
                                        int ftzero, int (&o)[6])
{
    const int g0 = pf2_g(img, pitch, cn, c, W, H, x, y, ftzero), t0 = pf2_t(img, pitch, cn, c, W, x, y, ftzero);
    int glo = g0, ghi = g0, tlo = t0, thi = t0;
    if (x > 0) {
        const int g1 = (g0 + pf2_g(img, pitch, cn, c, W, H, x - 1, y, ftzero)) >> 1;
        const int t1 = (t0 + pf2_t(img, pitch, cn, c, W, x - 1, y, ftzero)) >> 1;
        glo = min(glo, g1); ghi = max(ghi, g1); tlo = min(tlo, t1); thi = max(thi, t1);
    }
    if (x < W - 1) {
        const int g1 = (g0 + pf2_g(img, pitch, cn, c, W, H, x + 1, y, ftzero)) >> 1;
        const int t1 = (t0 + pf2_t(img, pitch, cn, c, W, x + 1, y, ftzero)) >> 1;
        glo = min(glo, g1); ghi = max(ghi, g1); tlo = min(tlo, t1); thi = max(thi, t1);
    }
    o[0] = g0; o[1] = glo; o[2] = ghi; o[3] = t0; o[4] = tlo; o[5] = thi;
}

__global__ void k_prefilter2(const uint8_t *left, const uint8_t *right, long long pitch, int W, int H, int cn, int ftzero,
                             uint8_t *leftP, int PL, uint32_t *rpairs, int RPW, uint4 *leftX, int eshift, int only3)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const int ic = blockIdx.z;                 // image * cn + channel
    if (x >= W) return;
    const int im = ic / cn, c = ic % cn;
    int a[6];
    if (im == 0) {
        pf2_six(left, pitch, cn, c, W, H, x, y, ftzero, a);
        if (!only3) {
#pragma unroll
            for (int p = 0; p < 6; p++) leftP[((size_t)(c * 6 + p) * H + y) * PL + x] = (uint8_t)a[p];
        }
        if (leftX) {                           // cn == 1: packed-expanded operands for k_cost3
            uint4 *o = leftX + ((size_t)y * W + x) * 2;
            // per plane (g, t): u + K, K - u, K - u_hi, u_lo + K   (K = 256, the bias of k_cost3)
            o[0] = make_uint4((a[0] + 256u) * 0x10001u, (256u - a[0]) * 0x10001u, (256u - a[2]) * 0x10001u, (a[1] + 256u) * 0x10001u);
            o[1] = make_uint4((a[3] + 256u) * 0x10001u, (256u - a[3]) * 0x10001u, (256u - a[5]) * 0x10001u, (a[4] + 256u) * 0x10001u);
        }
    } else {
        if (only3 && (x & 1)) return;          // k_cost3 reads the even-parity pair words only
        int b[6];
        pf2_six(right, pitch, cn, c, W, H, x, y, ftzero, a);
        pf2_six(right, pitch, cn, c, W, H, min(x + 1, W - 1), y, ftzero, b);
#pragma unroll
        for (int p = 0; p < 6; p++)            // pair word of q = x: lo = v(q+1), hi = v(q)
            rpairs[(((size_t)(c * 6 + p) * H + y) * 2 + (x & 1)) * RPW + (x >> 1) + ((x & 1) ? 0 : eshift)] =
                (uint32_t)b[p] | ((uint32_t)a[p] << 16);  // eshift: k_cost3 wants the even array one word to the right
    }
}

// Prefilter for the k_cost3 path (1-channel input): one block per 256-pixel row segment and image.  The
// gradient / raw values are computed once per pixel into shared memory, the half-sample intervals from
// those, and only what k_cost3 reads is written: 32 bytes of packed operands per left pixel, one pair
// word per EVEN right position and plane (threads map to pair indices, so the stores are dense).
#define PF3_TX 256
__global__ void __launch_bounds__(PF3_TX) k_prefilter3(const uint8_t *__restrict__ left, const uint8_t *__restrict__ right,
                                                       long long pitch, int W, int H, int ftzero, uint4 *__restrict__ leftX,
                                                       uint32_t *__restrict__ rpairs, int RPW, int eshift)
{
    __shared__ uint8_t sgt[2][PF3_TX + 4];                // g, t at x0-1 .. x0+TX+1
    const int tid = threadIdx.x, x0 = blockIdx.x * PF3_TX, y = blockIdx.y, im = blockIdx.z;
    const uint8_t *img = im ? right : left;
    const uint8_t *r0 = img + (long long)y * pitch, *rm = img + (long long)max(y - 1, 0) * pitch,
                  *rp = img + (long long)min(y + 1, H - 1) * pitch;
    for (int i = tid; i < PF3_TX + 3; i += PF3_TX) {
        const int x = x0 - 1 + i;
        int gv = ftzero & 0xFF, tv = ftzero & 0xFF;       // columns 0 and W-1 (and outside) carry ftzero (A.1)
        if (x > 0 && x < W - 1) {
            const int v = 2 * ((int)r0[x + 1] - (int)r0[x - 1]) + ((int)rm[x + 1] - (int)rm[x - 1]) + ((int)rp[x + 1] - (int)rp[x - 1]);
            gv = (min(max(v, -ftzero), ftzero) + ftzero) & 0xFF;
            tv = r0[x];
        }
        sgt[0][i] = (uint8_t)gv; sgt[1][i] = (uint8_t)tv;
    }
    __syncthreads();
    // value, half-sample minimum and maximum (A.2) of plane q at tile position i (image column x0 + i), straight from the
    // staged g / t values: the six plane values of a pixel never go through shared memory (the kernel is bound by
    // instruction issue, not by memory: 200 instructions per pixel before this, most of them byte-sized shared accesses)
    auto six = [&](int q, int i, uint32_t &v0o, uint32_t &loo, uint32_t &hio) {
        const int x = x0 + i;
        const int v0 = sgt[q][i + 1];
        int lo = v0, hi = v0;
        if (x > 0) { const int v1 = (v0 + sgt[q][i]) >> 1; lo = min(lo, v1); hi = max(hi, v1); }
        if (x < W - 1) { const int v1 = (v0 + sgt[q][i + 2]) >> 1; lo = min(lo, v1); hi = max(hi, v1); }
        v0o = (uint32_t)v0; loo = (uint32_t)lo; hio = (uint32_t)hi;
    };
    if (im == 0) {
        const int x = x0 + tid;
        if (x < W) {
            uint32_t a0, a1, a2, a3, a4, a5;
            six(0, tid, a0, a1, a2);
            six(1, tid, a3, a4, a5);
            uint4 *o = leftX + ((size_t)y * W + x) * 2;
            o[0] = make_uint4((a0 + 256u) * 0x10001u, (256u - a0) * 0x10001u, (256u - a2) * 0x10001u, (a1 + 256u) * 0x10001u);
            o[1] = make_uint4((a3 + 256u) * 0x10001u, (256u - a3) * 0x10001u, (256u - a5) * 0x10001u, (a4 + 256u) * 0x10001u);
        }
    } else if (tid < PF3_TX / 2) {
        const int i = 2 * tid, x = x0 + i;                // even right position: lo = v(x+1), hi = v(x)
        if (x < W) {
            const int i1 = x + 1 < W ? i + 1 : i;
            uint32_t e[6], o1[6];
            six(0, i, e[0], e[1], e[2]);
            six(1, i, e[3], e[4], e[5]);
            six(0, i1, o1[0], o1[1], o1[2]);
            six(1, i1, o1[3], o1[4], o1[5]);
#pragma unroll
            for (int p = 0; p < 6; p++)
                rpairs[((size_t)p * H + y) * 2 * RPW + (x >> 1) + eshift] = o1[p] | (e[p] << 16);
        }
    }
}

int sgbm_launch_prefilter2(const Geo &g, const uint8_t *left, const uint8_t *right, long long pitch, uint8_t *planes,
                           int eshift, int only3, cudaStream_t st)
{
    if (only3 && g.cn == 1) {
        dim3 grid3((g.W + PF3_TX - 1) / PF3_TX, g.H, 2);
        k_prefilter3<<<grid3, PF3_TX, 0, st>>>(left, right, pitch, g.W, g.H, g.ftzero,
                                               reinterpret_cast<uint4 *>(planes + sgbm_cost2_leftx_offset(g)),
                                               reinterpret_cast<uint32_t *>(planes + cost2_right_offset(g)), sgbm_cost2_rpw(g), eshift);
        sgbm_count_launch(1);
        SGBM_CUDA_CHECK(cudaGetLastError());
        return 0;
    }
    dim3 grid((g.W + 255) / 256, g.H, 2 * g.cn);
    k_prefilter2<<<grid, 256, 0, st>>>(left, right, pitch, g.W, g.H, g.cn, g.ftzero, planes, sgbm_cost2_left_pitch(g),
                                       reinterpret_cast<uint32_t *>(planes + cost2_right_offset(g)), sgbm_cost2_rpw(g),
                                       g.cn == 1 ? reinterpret_cast<uint4 *>(planes + sgbm_cost2_leftx_offset(g)) : nullptr, eshift, only3);
    sgbm_count_launch(1);
    SGBM_CUDA_CHECK(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Cost volume.  One CTA owns TX valid columns x all disparities and walks down a band of rows,
// keeping the running vertical block sum in registers and the last 2r+1 horizontal sums in a
// shared-memory ring:   C(y) = C(y-1) + hsum(y+r) - hsum(y-r-1)     (A.3, clamped rows/columns).
// Per source row:
//   stage : thread 0 issued TMA bulk copies of the row's left planes and right pair words nstg-1
//           rows ahead; everybody waits on the stage's mbarrier.
//   B     : pixel costs for TX+2r columns, one warp per column, lanes over disparity pairs in
//           natural order (packed u16x2, biased by K so differences stay non-negative).
//   C     : each thread owns one 32-bit word of the output layout (two disparities) of XPT consecutive
//           columns: sliding horizontal window sum, ring update, running vertical sum, coalesced
//           32-bit stores of C.
// blockDim = Dw * NXG with Dw = Dp/2 words per column; TX = NXG * XPT.
// ------------------------------------------------------------------------------------------------
struct Cost2Args {
    Geo g;
    const uint8_t *leftP;    // [cn][6][H][PL] u8
    const uint32_t *rpairs;  // [cn][6][H][2][RPW]
    int PL, RPW;
    uint16_t *out;           // row y is written at out + (y - y0) * rowStride
    int y0, nrows;           // output rows [y0, y0 + nrows)
    int ylo;                 // vertical clamp floor (0, or the stripe start for 3WAY)
    int NXG, RB;             // thread groups along x, rows per band
    int zeroTail;            // HH4 quirk (A.9): rows y >= H - r get C = 0
    int NQh, LVW, nstg;
    unsigned int pixOff, stgOff, descOff, barOff;
};

template <int NREG, int LPC, int XPT>
__global__ void __launch_bounds__(512) k_cost2(Cost2Args a)
{
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int Dw = NREG * LPC;                       // 32-bit words per column of the volume
    constexpr int PADW = ((NREG / 4) % 2 == 0) ? 4 : 0;  // bank padding of the natural-order pixel buffer
    constexpr int PIXS = Dw + (Dw / NREG) * PADW;
    const Geo &g = a.g;
    const int r = g.r, HP = g.D / 2;
    const int TX = a.NXG * XPT, TXH = TX + 2 * r;
    const int x0 = blockIdx.x * TX;                      // first valid column of the tile
    const int yb = a.y0 + blockIdx.y * a.RB;             // first output row of the band
    const int yend = min(yb + a.RB, a.y0 + a.nrows);
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;
    const int cn = g.cn;
    const int NQh = a.NQh, LVW = a.LVW, nstg = a.nstg;

    // image-coordinate range of the tile incl. halo (clamped to the valid range)
    const int xa = g.minX1 + min(max(x0 - r, 0), g.W1 - 1);
    const int xaA = xa & ~15;                            // 16-byte aligned start of the staged left rows
    const int q0 = xa - g.maxD + 1;                      // first right-image pixel the tile touches
    const int iLo = (q0 >> 1) & ~3;                      // first staged pair index (per parity), 16-byte aligned

    // shared memory carve-up
    uint32_t *ring = reinterpret_cast<uint32_t *>(smem);                      // [(2r+1)][TX][Dw]
    uint32_t *pixbuf = reinterpret_cast<uint32_t *>(smem + a.pixOff);         // [TXH][PIXS]
    uint8_t *stg = smem + a.stgOff;                                           // [nstg] { rp [cn][6][2][NQh] u32 ; lv [cn][6][LVW] u8 }
    int2 *coldesc = reinterpret_cast<int2 *>(smem + a.descOff);               // [TXH] row-invariant column offsets
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + a.barOff);
    const unsigned int rpBytes = (unsigned)cn * 6 * 2 * NQh * 4, lvBytes = (unsigned)cn * 6 * LVW;
    const unsigned int stageBytes = rpBytes + ((lvBytes + 15) & ~15u);

    const int nslots = 2 * r + 1;
    for (int i = tid; i < nslots * TX * Dw; i += nthr) ring[i] = 0;
    for (int i = tid; i < TXH * PIXS; i += nthr) pixbuf[i] = 0;
    for (int xx = tid; xx < TXH; xx += nthr) {
        const int x = g.minX1 + min(max(x0 - r + xx, 0), g.W1 - 1);
        const int qq = x - g.minD - 1;                   // pair (qq + 1, qq) serves disparities (0, 1)
        coldesc[xx] = make_int2(x - xaA, (qq & 1) * NQh + ((qq >> 1) - iLo));
    }
    if (tid == 0) {
        for (int i = 0; i < nstg; i++) mbar_init(&bars[i], 1);
        mbar_fence_init();
    }
    __syncthreads();

    const int nsteps = (yend - yb) + 2 * r;
    auto fill = [&](int k, int sg) {                      // thread 0: stage source row k of the band
        const int ysrc = min(max(yb - r + k, a.ylo), g.H - 1);
        uint8_t *sb = stg + (size_t)sg * stageBytes;
        mbar_expect_tx(&bars[sg], (unsigned)cn * 6 * (2 * NQh * 4 + LVW));
        for (int cp = 0; cp < cn * 6; cp++) {
            const uint32_t *src = a.rpairs + ((size_t)cp * g.H + ysrc) * 2 * a.RPW + iLo;
            bulk_g2s(sb + (size_t)(cp * 2 + 0) * NQh * 4, src, (unsigned)NQh * 4, &bars[sg]);
            bulk_g2s(sb + (size_t)(cp * 2 + 1) * NQh * 4, src + a.RPW, (unsigned)NQh * 4, &bars[sg]);
            bulk_g2s(sb + rpBytes + (size_t)cp * LVW, a.leftP + ((size_t)cp * g.H + ysrc) * a.PL + xaA, (unsigned)LVW, &bars[sg]);
        }
    };
    if (tid == 0)
        for (int k = 0; k < nstg && k < nsteps; k++) fill(k, k);

    // phase C ownership: output word `pos` of columns xg*XPT .. xg*XPT+XPT-1.  Natural disparity-pair
    // index of that word: lane l = chunk % LPC holds pairs l*NREG + 4*(chunk / LPC) + e.
    const int pos = tid % Dw, xg = tid / Dw;
    int psw;
    {
        const int ch = pos >> 2, e = pos & 3;
        const int l = ch % LPC, k4 = ch / LPC;
        const int wnat = l * NREG + 4 * k4 + e;
        psw = wnat + (wnat / NREG) * PADW;
    }
    uint32_t crun[XPT];
#pragma unroll
    for (int n = 0; n < XPT; n++) crun[n] = 0;
    const uint32_t KK = COST2_K * 0x10001u;
    const uint32_t KSUB = (COST2_K + COST2_K / 4) * 0x10001u;
    const int lsw = lane + (lane / NREG) * PADW;         // swizzled position of natural pair `lane`
    constexpr int itStep = 32 + (32 / NREG) * PADW;      // ... and the step for +32 pairs (NREG divides 32 or PADW == 0)
    const int xbase = xg * XPT;
    const uint32_t *pb = pixbuf + xbase * PIXS + psw;    // phase C: column xbase-r of the tile (halo offset r)
    const int planeStride = 2 * NQh;

    int sg = 0, slot = 0;
    uint32_t par = 0;
    for (int k = 0; k < nsteps; k++) {
        mbar_wait(&bars[sg], par);
        const uint32_t *rp = reinterpret_cast<const uint32_t *>(stg + (size_t)sg * stageBytes);
        const uint8_t *lv = stg + (size_t)sg * stageBytes + rpBytes;
        // ---- B: pixel costs for the TXH columns ----------------------------------------------
        for (int c = 0; c < cn; c++) {
            for (int xx = warp; xx < TXH; xx += nwarp) {
                const int2 cd = coldesc[xx];
                const uint8_t *lc = lv + (c * 6) * LVW + cd.x;
                uint32_t uK[2], Ku[2], KmUhi[2], UloK[2];
#pragma unroll
                for (int p = 0; p < 2; p++) {
                    const uint32_t u = lc[(3 * p + 0) * LVW], ulo = lc[(3 * p + 1) * LVW], uhi = lc[(3 * p + 2) * LVW];
                    uK[p] = (u + COST2_K) * 0x10001u; Ku[p] = (COST2_K - u) * 0x10001u;
                    KmUhi[p] = (COST2_K - uhi) * 0x10001u; UloK[p] = (ulo + COST2_K) * 0x10001u;
                }
                const uint32_t *rc = rp + (c * 6) * planeStride + cd.y - lane;
                uint32_t *pcol = pixbuf + xx * PIXS + lsw;
#pragma unroll 4
                for (int pi = lane; pi < HP; pi += 32, rc -= 32, pcol += itStep) {
                    uint32_t bt[2];
#pragma unroll
                    for (int p = 0; p < 2; p++) {
                        const uint32_t v2 = rc[(3 * p + 0) * planeStride];
                        const uint32_t vlo2 = rc[(3 * p + 1) * planeStride];
                        const uint32_t vhi2 = rc[(3 * p + 2) * planeStride];
                        const uint32_t c1 = __vimax3_u16x2(uK[p] - vhi2, vlo2 + Ku[p], KK);      // max(u-vhi, vlo-u, 0) + K
                        const uint32_t c2 = __vimax3_u16x2(v2 + KmUhi[p], UloK[p] - v2, KK);     // max(v-uhi, ulo-v, 0) + K
                        bt[p] = __vminu2(c1, c2);
                    }
                    // (bt_g + K) + ((bt_t + K) >> 2) - (K + K/4)
                    const uint32_t val = bt[0] + ((bt[1] >> 2) & 0x3FFF3FFFu) - KSUB;
                    if (c == 0) *pcol = val; else *pcol += val;
                }
            }
        }
        __syncthreads();
        if (tid == 0 && k + nstg < nsteps) fill(k + nstg, sg);  // stage sg is free again
        // ---- C: sliding horizontal sum, ring, running vertical sum, store ----------------------
        if (xg < a.NXG) {
            const int yout = yb + k - 2 * r;
            const bool emit = (k >= 2 * r);
            const bool zero = a.zeroTail && r > 0 && yout >= g.H - r;
            uint32_t *rg = ring + (slot * TX + xbase) * Dw + pos;
            uint32_t *orow32 = reinterpret_cast<uint32_t *>(a.out) +
                               (emit ? ((size_t)(yout - a.y0) * g.rowStride + (size_t)(x0 + xbase) * g.Dp) / 2 + pos : 0);
            const int nvalid = emit ? g.W1 - (x0 + xbase) : 0;   // columns of this thread that exist
            uint32_t hs = 0;
            for (int i = 0; i <= 2 * r; i++) hs += pb[i * PIXS];
            const uint32_t *pn = pb + 2 * r * PIXS;
#pragma unroll
            for (int n = 0; n < XPT; n++) {
                if (n > 0) hs += pn[n * PIXS] - pb[(n - 1) * PIXS];
                const uint32_t old = rg[n * Dw];
                rg[n * Dw] = hs;
                crun[n] = crun[n] + hs - old;
                if (n < nvalid) orow32[n * Dw] = zero ? 0u : crun[n];
            }
        }
        __syncthreads();                                  // pixbuf is rewritten by the next row's phase B
        if (++sg == nstg) { sg = 0; par ^= 1u; }
        if (++slot == nslots) slot = 0;
    }
}

static bool cost2_layout(Cost2Args &a, int TX, size_t maxSmem, size_t *total)
{
    const Geo &g = a.g;
    const int r = g.r, TXH = TX + 2 * r, Dw = g.Dp / 2;
    const int padw = ((g.nreg / 4) % 2 == 0) ? 4 : 0;
    const int pixs = Dw + (Dw / g.nreg) * padw;
    a.NQh = (((TXH + g.D) / 2 + 8) + 3) & ~3;
    a.LVW = (TXH + 15 + 15) & ~15;
    size_t off = (size_t)(2 * r + 1) * TX * Dw * 4;
    a.pixOff = (unsigned)off; off += (size_t)TXH * pixs * 4;
    off = (off + 15) & ~(size_t)15;
    a.descOff = (unsigned)off; off += (size_t)TXH * 8;
    off = (off + 127) & ~(size_t)127;
    a.stgOff = (unsigned)off;
    const size_t rpBytes = (size_t)g.cn * 6 * 2 * a.NQh * 4, lvBytes = ((size_t)g.cn * 6 * a.LVW + 15) & ~(size_t)15;
    for (a.nstg = 3; a.nstg >= 2; a.nstg--) {
        size_t end = off + (size_t)a.nstg * (rpBytes + lvBytes);
        a.barOff = (unsigned)end;
        end += 8 * 4;
        if (end <= maxSmem) { *total = end; return true; }
    }
    return false;
}

template <int NREG, int LPC, int XPT>
static int launch_cost2_t(Cost2Args &a, int threads, size_t smem, dim3 grid, int maxSmem, cudaStream_t st)
{
    static unsigned long long attrDone = 0;   // one bit per device: function attributes are per device
    {
        SgbmDeviceOnce once(attrDone);
        if (once.first) {
            SGBM_CUDA_CHECK(cudaFuncSetAttribute(k_cost2<NREG, LPC, XPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxSmem));
            once.done();
        }
    }
    k_cost2<NREG, LPC, XPT><<<grid, threads, smem, st>>>(a);
    sgbm_count_launch(1);
    SGBM_CUDA_CHECK(cudaGetLastError());
    return 0;
}

// Rows [y0, y0+nrows) of the cost volume with vertical clamp floor ylo, written at out (row y0 first).
// Returns 1 when the geometry does not fit this kernel (caller falls back to sgbm_launch_cost).
int sgbm_launch_cost2(const Geo &g, const uint8_t *planes, uint16_t *out, int y0, int nrows, int ylo, int zeroTail,
                      cudaStream_t st)
{
    if (nrows <= 0) return 0;
    const int maxSmem = sgbm_knobs().maxSmemOptin;
    const int Dw = g.Dp / 2;
    if (Dw > 512) return 1;
    if (32 % g.nreg != 0 && ((g.nreg / 4) % 2 == 0)) return 1;      // padded natural order needs NREG | 32 (12: no padding)
    Cost2Args a;
    memset(&a, 0, sizeof(a));
    a.g = g;
    a.leftP = planes; a.PL = sgbm_cost2_left_pitch(g);
    a.rpairs = reinterpret_cast<const uint32_t *>(planes + cost2_right_offset(g)); a.RPW = sgbm_cost2_rpw(g);
    a.out = out; a.y0 = y0; a.nrows = nrows; a.ylo = ylo; a.zeroTail = zeroTail;
    // threads = Dw * NXG (<= 512); TX = NXG * XPT
    int NXG = 1;
    while (Dw * NXG * 2 <= 512 && NXG * 16 < 256) NXG *= 2;
    const int XPT = 16;
    size_t smem = 0;
    bool ok = false;
    for (;;) {
        ok = cost2_layout(a, NXG * XPT, (size_t)maxSmem, &smem);
        if (ok || NXG == 1) break;
        NXG >>= 1;
    }
    if (!ok) return 1;
    const int threads = ((Dw * NXG + 31) / 32) * 32;
    a.NXG = NXG;
    a.RB = nrows < 64 ? nrows : 64;
    const int TX = NXG * XPT;
    // the staged right rows must stay inside the padded parity rows of the prefilter output
    if (((g.W - 1) >> 1) + a.NQh + 4 > a.RPW) return 1;
    dim3 grid((g.W1 + TX - 1) / TX, (nrows + a.RB - 1) / a.RB);
#define COST2_CASE(NR, LP) if (g.nreg == NR && g.lpc == LP) return launch_cost2_t<NR, LP, 16>(a, threads, smem, grid, maxSmem, st);
    // the lane mappings make_geo produces for numDisparities <= 512 (lpc <= 16); others use sgbm_cost.cu
    COST2_CASE(16, 8) COST2_CASE(16, 16) COST2_CASE(12, 2) COST2_CASE(12, 4) COST2_CASE(12, 8) COST2_CASE(12, 16)
    COST2_CASE(8, 2) COST2_CASE(8, 4) COST2_CASE(8, 8) COST2_CASE(8, 16) COST2_CASE(4, 2) COST2_CASE(4, 4) COST2_CASE(4, 8)
    COST2_CASE(4, 16)
#undef COST2_CASE
    return 1;
}
