/*
 * sgbm_oracle.c -- CPU restatement of the dense-stereo hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the checker, never the product: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The shipped path is the CUDA
 * library in stereo_reconstruction_cv_b200/csrc and it fails loudly when that is missing.
 *
 * What it restates.  The reference (rafayaamirgull/stereo_reconstruction_cv) runs the path
 * through third-party OpenCV, which is NOT vendored under /root/reference:
 *     main.ipynb:655-666   cv2.StereoSGBM_create(...)
 *     main.ipynb:668       stereo.compute(imgL, imgR)
 *     main.ipynb:697       cv2.reprojectImageTo3D(disparity_map, Q)
 * Dependency: opencv-python==4.11.0.86 (environment.yml:89-90); the binary installed in this
 * image is opencv-python-headless 4.13.0.92.  The arithmetic followed here is the published
 * semi-global matching algorithm (Hirschmueller 2008; Birchfield-Tomasi 1998 pixel cost) in the
 * exact fixed-point form specified in SURVEY.md Appendix A (A.0 .. A.9), which was established
 * by black-box probes against that binary.  Section tags "A.n" below cite that appendix.
 *
 * Parity pin: tests/test_oracle.py checks this file against (i) the live cv2 binary when it is
 * importable and (ii) the committed golden vectors under tests/golden/ that were generated from
 * cv2 by tests/golden/make_golden.py.  0 mismatching pixels is the bar.
 *
 * Plain C99, no SIMD, single thread; written for clarity (full cost volume in memory), so it is
 * meant for small / medium images (<= ~1 Mpixel x 128 disparities).
 */
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#define ORACLE_MODE_SGBM      0
#define ORACLE_MODE_HH        1
#define ORACLE_MODE_SGBM_3WAY 2
#define ORACLE_MODE_HH4       3

typedef struct {
    int minDisparity, numDisparities, blockSize, P1, P2, disp12MaxDiff, preFilterCap,
        uniquenessRatio, speckleWindowSize, speckleRange, mode;
} oracle_params;

static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int iclamp(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* ---- A.0 effective parameters ------------------------------------------------------------ */
typedef struct {
    int minD, D, maxD, r, P1, P2, UR, DMD, ftzero, INV, minX1, maxX1, W1, mode;
} eff_params;

static int effective(const oracle_params *p, int W, eff_params *e)
{
    e->mode = p->mode;
    e->minD = p->minDisparity;
    e->D = p->numDisparities;
    e->maxD = e->minD + e->D;
    if (p->mode == ORACLE_MODE_SGBM_3WAY)
        e->r = (p->blockSize > 0 ? p->blockSize : 3) / 2;
    else
        e->r = (p->blockSize > 0 ? p->blockSize : 5) / 2;
    e->P1 = p->P1 > 0 ? p->P1 : 2;
    e->P2 = imax(p->P2 > 0 ? p->P2 : 5, e->P1 + 1);
    e->UR = p->uniquenessRatio >= 0 ? p->uniquenessRatio : 10;
    e->DMD = p->disp12MaxDiff > 0 ? p->disp12MaxDiff : 1;
    e->ftzero = imax(p->preFilterCap, 15) | 1;
    e->INV = (e->minD - 1) * 16;
    e->minX1 = imax(e->maxD, 0);
    e->maxX1 = W + imin(e->minD, 0);
    e->W1 = e->maxX1 - e->minX1;
    if (e->D <= 0) return -1;
    if (!(W - (e->minD + e->D) > p->blockSize / 2)) return -2;   /* cv2.error site [P15] */
    if (e->W1 <= 0) return -3;                                   /* own validation   [P18] */
    return 0;
}

/* ---- A.1 prefilter: planes g (clipped x-Sobel) and t (raw, borders forced to ftzero) ------ */
static void prefilter_plane(const uint8_t *img, int W, int H, ptrdiff_t pitch, int cn, int c,
                            int ftzero, uint8_t *g, uint8_t *t)
{
    for (int y = 0; y < H; y++) {
        const uint8_t *r0 = img + (ptrdiff_t)y * pitch;
        const uint8_t *rm = img + (ptrdiff_t)imax(y - 1, 0) * pitch;
        const uint8_t *rp = img + (ptrdiff_t)imin(y + 1, H - 1) * pitch;
        uint8_t *gy = g + (size_t)y * W, *ty = t + (size_t)y * W;
        for (int x = 0; x < W; x++) {
            if (x == 0 || x == W - 1) {
                gy[x] = (uint8_t)ftzero;
                ty[x] = (uint8_t)ftzero;
                continue;
            }
            int d0 = (int)r0[(x + 1) * cn + c] - (int)r0[(x - 1) * cn + c];
            int dm = (int)rm[(x + 1) * cn + c] - (int)rm[(x - 1) * cn + c];
            int dp = (int)rp[(x + 1) * cn + c] - (int)rp[(x - 1) * cn + c];
            int v = iclamp(2 * d0 + dm + dp, -ftzero, ftzero) + ftzero;
            gy[x] = (uint8_t)v;                    /* u8() wraps mod 256 [P4] */
            ty[x] = r0[x * cn + c];
        }
    }
}

/* half-sample interval of a plane row (A.2): pmin/pmax over {p, (p+p[x-1])/2, (p+p[x+1])/2} */
static void interval_row(const uint8_t *p, int W, uint8_t *lo, uint8_t *hi)
{
    for (int x = 0; x < W; x++) {
        int v = p[x];
        int vl = x > 0 ? (v + p[x - 1]) / 2 : v;
        int vr = x < W - 1 ? (v + p[x + 1]) / 2 : v;
        lo[x] = (uint8_t)imin(v, imin(vl, vr));
        hi[x] = (uint8_t)imax(v, imax(vl, vr));
    }
}

static inline int bt_cost(int u, int ulo, int uhi, int v, int vlo, int vhi)
{
    int c0 = imax(0, imax(u - vhi, vlo - u));
    int c1 = imax(0, imax(v - uhi, ulo - v));
    return imin(c0, c1);
}

/* ---- A.2 + horizontal half of A.3: hsum[y][x1][d] = sum_{i=-r..r} pix(clamp(x+i), y, d) ---- */
static void pixel_cost_hsum(const uint8_t *left, const uint8_t *right, int W, int H,
                            ptrdiff_t pitchL, ptrdiff_t pitchR, int cn, const eff_params *e,
                            int16_t *hsum /* H*W1*D */)
{
    const int D = e->D, W1 = e->W1, r = e->r;
    size_t plane = (size_t)W * H;
    uint8_t *gL = malloc(plane), *tL = malloc(plane), *gR = malloc(plane), *tR = malloc(plane);
    uint8_t *lo = malloc((size_t)W * 8), *hi = lo + (size_t)W * 4;
    int *pix = malloc(sizeof(int) * (size_t)W1 * D);
    memset(hsum, 0, sizeof(int16_t) * (size_t)H * W1 * D);
    int *acc = calloc((size_t)H * W1 * D, sizeof(int));   /* int accumulator across channels */
    for (int c = 0; c < cn; c++) {
        prefilter_plane(left, W, H, pitchL, cn, c, e->ftzero, gL, tL);
        prefilter_plane(right, W, H, pitchR, cn, c, e->ftzero, gR, tR);
        for (int y = 0; y < H; y++) {
            const uint8_t *pl[2] = { gL + (size_t)y * W, tL + (size_t)y * W };
            const uint8_t *pr[2] = { gR + (size_t)y * W, tR + (size_t)y * W };
            uint8_t *Llo[2] = { lo, lo + W }, *Lhi[2] = { hi, hi + W };
            uint8_t *Rlo[2] = { lo + 2 * W, lo + 3 * W }, *Rhi[2] = { hi + 2 * W, hi + 3 * W };
            for (int k = 0; k < 2; k++) {
                interval_row(pl[k], W, Llo[k], Lhi[k]);
                interval_row(pr[k], W, Rlo[k], Rhi[k]);
            }
            for (int x1 = 0; x1 < W1; x1++) {
                int x = x1 + e->minX1;
                for (int d = 0; d < D; d++) {
                    int xr = x - (d + e->minD);
                    int cg = bt_cost(pl[0][x], Llo[0][x], Lhi[0][x], pr[0][xr], Rlo[0][xr], Rhi[0][xr]);
                    int ct = bt_cost(pl[1][x], Llo[1][x], Lhi[1][x], pr[1][xr], Rlo[1][xr], Rhi[1][xr]);
                    pix[(size_t)x1 * D + d] = cg + (ct >> 2);
                }
            }
            int *hy = acc + (size_t)y * W1 * D;
            for (int x1 = 0; x1 < W1; x1++)
                for (int i = -r; i <= r; i++) {
                    int xs = iclamp(x1 + i, 0, W1 - 1);      /* clamp to the VALID range (A.3) */
                    const int *ps = pix + (size_t)xs * D;
                    int *hd = hy + (size_t)x1 * D;
                    for (int d = 0; d < D; d++) hd[d] += ps[d];
                }
        }
    }
    for (size_t i = 0; i < (size_t)H * W1 * D; i++) hsum[i] = (int16_t)acc[i];
    free(acc); free(pix); free(lo); free(gL); free(tL); free(gR); free(tR);
}

/* ---- vertical half of A.3: C(x,y,d) = sum_{j=-r..r} hsum(x, clamp(y+j, ylo, H-1), d) ------- */
static void block_cost_row(const int16_t *hsum, int H, int W1, int D, int r, int y, int ylo,
                           int16_t *Crow /* W1*D */)
{
    size_t n = (size_t)W1 * D;
    int *acc = calloc(n, sizeof(int));
    for (int j = -r; j <= r; j++) {
        int ys = iclamp(y + j, ylo, H - 1);
        const int16_t *h = hsum + (size_t)ys * n;
        for (size_t i = 0; i < n; i++) acc[i] += h[i];
    }
    for (size_t i = 0; i < n; i++) Crow[i] = (int16_t)acc[i];
    free(acc);
}

/* ---- A.4 one path step: Lout(d) = C(d) + min(Lp(d), Lp(d-1)+P1, Lp(d+1)+P1, m+P2) - m ------ */
static inline void path_step(const int16_t *C, const int *Lp /* NULL => predecessor outside */,
                             int D, int P1, int P2, int *Lout)
{
    if (!Lp) { for (int d = 0; d < D; d++) Lout[d] = C[d]; return; }
    int m = Lp[0];
    for (int d = 1; d < D; d++) m = imin(m, Lp[d]);
    for (int d = 0; d < D; d++) {
        int v = imin(Lp[d], m + P2);
        if (d > 0) v = imin(v, Lp[d - 1] + P1);
        if (d < D - 1) v = imin(v, Lp[d + 1] + P1);
        Lout[d] = C[d] + v - m;
    }
}

/* Accumulate the path costs of a set of directions over rows [y0,y1) into S (int32, not yet
 * saturated).  dirs: list of (dx,dy) PREDECESSOR offsets.  Cvol rows are indexed from y0.    */
static void aggregate_paths(const int16_t *Cvol, int y0, int y1, int W1, int D, int P1, int P2,
                            const int (*dirs)[2], int ndirs, int *S /* (y1-y0)*W1*D */)
{
    size_t rowN = (size_t)W1 * D;
    int *Lprev = malloc(sizeof(int) * rowN), *Lcur = malloc(sizeof(int) * rowN);
    for (int k = 0; k < ndirs; k++) {
        int dx = dirs[k][0], dy = dirs[k][1];
        int ystart = dy > 0 ? y1 - 1 : y0, yend = dy > 0 ? y0 - 1 : y1, ystep = dy > 0 ? -1 : 1;
        for (int y = ystart; y != yend; y += ystep) {
            const int16_t *Cy = Cvol + (size_t)(y - y0) * rowN;
            int *Sy = S + (size_t)(y - y0) * rowN;
            int xstart = dx > 0 ? W1 - 1 : 0, xend = dx > 0 ? -1 : W1, xstep = dx > 0 ? -1 : 1;
            if (dy != 0) { xstart = 0; xend = W1; xstep = 1; }
            for (int x = xstart; x != xend; x += xstep) {
                int qx = x + dx, qy = y + dy;
                const int *Lp = NULL;
                if (qx >= 0 && qx < W1 && qy >= y0 && qy < y1)
                    Lp = (dy == 0 ? Lcur : Lprev) + (size_t)qx * D;
                int *Lo = Lcur + (size_t)x * D;
                path_step(Cy + (size_t)x * D, Lp, D, P1, P2, Lo);
                for (int d = 0; d < D; d++) Sy[(size_t)x * D + d] += Lo[d];
            }
            int *t = Lprev; Lprev = Lcur; Lcur = t;
        }
    }
    free(Lprev); free(Lcur);
}

/* ---- A.5 / A.6 winner-take-all for one row (x from W1-1 down to 0) + LR check -------------- */
static void wta_row(const int *Srow /* W1*D, unsaturated int */, int W, const eff_params *e,
                    int16_t *disp /* W */, int *disp2, int *disp2cost /* W scratch */)
{
    const int D = e->D, W1 = e->W1, minD = e->minD, INV = e->INV, UR = e->UR;
    const int threeway = e->mode == ORACLE_MODE_SGBM_3WAY;
    int16_t *Sv = malloc(sizeof(int16_t) * D);
    for (int x = 0; x < W; x++) { disp[x] = (int16_t)INV; disp2[x] = INV; disp2cost[x] = 32767; }
    for (int x1 = W1 - 1; x1 >= 0; x1--) {
        const int *Si = Srow + (size_t)x1 * D;
        int minS = 32767, best = -1;
        for (int d = 0; d < D; d++) Sv[d] = (int16_t)imin(Si[d], 32767);   /* A.4 saturation */
        if (!threeway) {
            for (int d = 0; d < D; d++) if (Sv[d] < minS) { minS = Sv[d]; best = d; }
            int reject = 0;
            for (int d = 0; d < D && !reject; d++)
                if (Sv[d] * (100 - UR) < minS * 100 && abs(best - d) > 1) reject = 1;
            if (reject) continue;
        } else {
            /* 8-lane SIMD tie-break (A.6): per residue class d%8 the LARGEST tied d, then the
             * smallest of those */
            for (int d = 0; d < D; d++) if (Sv[d] < minS) minS = Sv[d];
            {
                int cand = -1;
                for (int l = 0; l < 8 && l < D; l++) {
                    int lb = -1;
                    for (int d = l; d < D; d += 8) if (Sv[d] == minS) lb = d;
                    if (lb >= 0 && (cand < 0 || lb < cand)) cand = lb;
                }
                best = cand;                       /* all-saturated rows are an ordinary tie here */
            }
            if (UR > 0) {
                int thr = (100 * minS) / (100 - UR);           /* trunc (C division) */
                int16_t t1 = (int16_t)(thr + 1);               /* (short) wrap        */
                int reject = 0;
                for (int d = 0; d < D && !reject; d++)
                    if (Sv[d] < t1 && abs(d - best) > 1) reject = 1;
                if (reject) continue;
            }
        }
        int x = x1 + e->minX1;
        int x2 = x - best - minD;
        if (x2 >= 0 && x2 < W && disp2cost[x2] > minS) { disp2cost[x2] = minS; disp2[x2] = best + minD; }
        int dq;
        if (best > 0 && best < D - 1) {
            int den = imax(Sv[best - 1] + Sv[best + 1] - 2 * Sv[best], 1);
            dq = best * 16 + ((Sv[best - 1] - Sv[best + 1]) * 16 + den) / (2 * den);  /* C trunc */
        } else dq = best * 16;
        disp[x] = (int16_t)(dq + minD * 16);
    }
    for (int x = e->minX1; x < e->maxX1; x++) {
        int d1 = disp[x];
        if (d1 == INV) continue;
        int _d = d1 >> 4, d_ = (d1 + 15) >> 4;
        int _x = x - _d, x_ = x - d_;
        if (0 <= _x && _x < W && disp2[_x] >= minD && abs(disp2[_x] - _d) > e->DMD &&
            0 <= x_ && x_ < W && disp2[x_] >= minD && abs(disp2[x_] - d_) > e->DMD)
            disp[x] = (int16_t)INV;
    }
    free(Sv);
}

/* ---- A.7 post filters ---------------------------------------------------------------------- */
void oracle_median3x3_i16(const int16_t *src, int16_t *dst, int W, int H)
{
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            int16_t v[9]; int n = 0;
            for (int j = -1; j <= 1; j++)
                for (int i = -1; i <= 1; i++)
                    v[n++] = src[(size_t)iclamp(y + j, 0, H - 1) * W + iclamp(x + i, 0, W - 1)];
            for (int a = 1; a < 9; a++) {               /* insertion sort */
                int16_t k = v[a]; int b = a - 1;
                while (b >= 0 && v[b] > k) { v[b + 1] = v[b]; b--; }
                v[b + 1] = k;
            }
            dst[(size_t)y * W + x] = v[4];
        }
}

static int uf_find(int *p, int i) { while (p[i] != i) { p[i] = p[p[i]]; i = p[i]; } return i; }

/* connected components over 4-neighbour edges (both != newVal, |a-b| <= maxDiff); components of
 * size <= maxSpeckleSize become newVal.  Union-find, scan-order independent (A.7, [P12]).      */
void oracle_filter_speckles_i16(int16_t *img, int W, int H, int newVal, int maxSpeckleSize, int maxDiff)
{
    size_t n = (size_t)W * H;
    int *par = malloc(sizeof(int) * n), *cnt = calloc(n, sizeof(int));
    for (size_t i = 0; i < n; i++) par[i] = (int)i;
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            size_t i = (size_t)y * W + x;
            if (img[i] == newVal) continue;
            if (x + 1 < W && img[i + 1] != newVal && abs(img[i] - img[i + 1]) <= maxDiff) {
                int a = uf_find(par, (int)i), b = uf_find(par, (int)i + 1);
                if (a != b) par[imax(a, b)] = imin(a, b);
            }
            if (y + 1 < H && img[i + W] != newVal && abs(img[i] - img[i + W]) <= maxDiff) {
                int a = uf_find(par, (int)i), b = uf_find(par, (int)(i + W));
                if (a != b) par[imax(a, b)] = imin(a, b);
            }
        }
    for (size_t i = 0; i < n; i++) if (img[i] != newVal) cnt[uf_find(par, (int)i)]++;
    for (size_t i = 0; i < n; i++)
        if (img[i] != newVal && cnt[uf_find(par, (int)i)] <= maxSpeckleSize) img[i] = (int16_t)newVal;
    free(par); free(cnt);
}

/* ---- the whole compute(): A.0 .. A.7.  Returns 0 or a negative validation code -------------- */
int oracle_sgbm_compute(const oracle_params *p, const uint8_t *left, const uint8_t *right,
                        int W, int H, int cn, ptrdiff_t pitchL, ptrdiff_t pitchR,
                        int16_t *disp_out /* H*W */,
                        int16_t *dbg_C /* optional H*W1*D (ylo=0 volume), may be NULL */,
                        int16_t *dbg_S /* optional H*W1*D saturated S, may be NULL     */,
                        int16_t *dbg_raw /* optional H*W: disparity before median/speckle */)
{
    eff_params e;
    int rc = effective(p, W, &e);
    if (rc) return rc;
    const int D = e.D, W1 = e.W1, r = e.r;
    size_t rowN = (size_t)W1 * D, volN = rowN * H;
    int16_t *hsum = malloc(sizeof(int16_t) * volN);
    pixel_cost_hsum(left, right, W, H, pitchL, pitchR, cn, &e, hsum);
    int16_t *raw = malloc(sizeof(int16_t) * (size_t)W * H);
    int *d2 = malloc(sizeof(int) * W * 2);

    if (e.mode != ORACLE_MODE_SGBM_3WAY) {
        int16_t *C = malloc(sizeof(int16_t) * volN);
        for (int y = 0; y < H; y++) {
            if (e.mode == ORACLE_MODE_HH4 && r > 0 && y >= H - r)
                memset(C + (size_t)y * rowN, 0, sizeof(int16_t) * rowN);      /* A.9 quirk */
            else
                block_cost_row(hsum, H, W1, D, r, y, 0, C + (size_t)y * rowN);
        }
        if (dbg_C) memcpy(dbg_C, C, sizeof(int16_t) * volN);
        int *S = calloc(volN, sizeof(int));
        static const int d_sgbm[5][2] = { {-1,0}, {-1,-1}, {0,-1}, {1,-1}, {1,0} };
        static const int d_hh[8][2] = { {-1,0}, {-1,-1}, {0,-1}, {1,-1}, {1,0}, {1,1}, {0,1}, {-1,1} };
        static const int d_hh4[4][2] = { {-1,0}, {1,0}, {0,-1}, {0,1} };
        if (e.mode == ORACLE_MODE_SGBM) aggregate_paths(C, 0, H, W1, D, e.P1, e.P2, d_sgbm, 5, S);
        else if (e.mode == ORACLE_MODE_HH) aggregate_paths(C, 0, H, W1, D, e.P1, e.P2, d_hh, 8, S);
        else aggregate_paths(C, 0, H, W1, D, e.P1, e.P2, d_hh4, 4, S);
        if (dbg_S) for (size_t i = 0; i < volN; i++) dbg_S[i] = (int16_t)imin(S[i], 32767);
        for (int y = 0; y < H; y++)
            wta_row(S + (size_t)y * rowN, W, &e, raw + (size_t)y * W, d2, d2 + W);
        free(S); free(C);
    } else {
        /* A.6: four fixed stripes, each recomputed from s0 = max(n*ss - ov, 0) */
        static const int d_3way[3][2] = { {-1,0}, {0,-1}, {1,0} };
        int ss = (H + 3) / 4;
        int ov = (p->blockSize / 2 + 1) + (int)ceil(0.1 * ss);
        for (int n = 0; n < 4; n++) {
            int o0 = n * ss, o1 = imin((n + 1) * ss, H);
            if (o0 >= o1) continue;
            int s0 = imax(o0 - ov, 0);
            size_t rows = (size_t)(o1 - s0);
            int16_t *C = malloc(sizeof(int16_t) * rows * rowN);
            for (int y = s0; y < o1; y++)
                block_cost_row(hsum, H, W1, D, r, y, s0, C + (size_t)(y - s0) * rowN);
            if (dbg_C && n == 0) memcpy(dbg_C, C, sizeof(int16_t) * rows * rowN);
            int *S = calloc(rows * rowN, sizeof(int));
            aggregate_paths(C, s0, o1, W1, D, e.P1, e.P2, d_3way, 3, S);
            for (int y = o0; y < o1; y++) {
                if (dbg_S)
                    for (size_t i = 0; i < rowN; i++)
                        dbg_S[(size_t)y * rowN + i] = (int16_t)imin(S[(size_t)(y - s0) * rowN + i], 32767);
                wta_row(S + (size_t)(y - s0) * rowN, W, &e, raw + (size_t)y * W, d2, d2 + W);
            }
            free(S); free(C);
        }
    }
    if (dbg_raw) memcpy(dbg_raw, raw, sizeof(int16_t) * (size_t)W * H);
    oracle_median3x3_i16(raw, disp_out, W, H);                                  /* always (A.7) */
    if (p->speckleWindowSize > 0)
        oracle_filter_speckles_i16(disp_out, W, H, e.INV, p->speckleWindowSize, 16 * p->speckleRange);
    free(d2); free(raw); free(hsum);
    return 0;
}

/* ---- A.8 reprojectImageTo3D (handleMissingValues=False, float32 output) -------------------- */
void oracle_reproject_f32(const float *disp, int W, int H, const double *Q /* 16, row major */,
                          float *xyz /* H*W*3 */)
{
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            double d = (double)disp[(size_t)y * W + x];
            double h[4];
            for (int k = 0; k < 4; k++)
                h[k] = ((Q[4 * k + 0] * x + Q[4 * k + 1] * y) + Q[4 * k + 2] * d) + Q[4 * k + 3];
            double iw = 1.0 / h[3];
            float *o = xyz + ((size_t)y * W + x) * 3;
            for (int k = 0; k < 3; k++) o[k] = (float)((double)(float)h[k] * iw);
        }
}
