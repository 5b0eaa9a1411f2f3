// sgbm_post.cu -- per-pixel tail of the path for sm_100a (all HBM-bound, one thread per pixel):
//   k_init_wta     : raw disparity = INV, disp2 keys = never hit, before a frame's winner-take-all
//   k_pad_cost     : constant cost of the padding disparities when numDisparities % 8 != 0
//   k_lr_median    : disp12MaxDiff left-right consistency check (A.5) + the always-on 3x3 median (A.7) in one tiled pass
//   k_lrcheck      : the LR check alone (debug hook: sgbm_debug_fetch of the raw disparity)
//   k_median3x3    : the 3x3 median alone (public medianBlur3)
//   k_cc_*         : cv2.filterSpeckles as connected components (lock-free union-find) (A.7)
//   k_disp_to_float: .astype(float32)/16 and positivity mask               (main.ipynb:668-670)
//   k_reproject    : cv2.reprojectImageTo3D in fp64, bit exact             (main.ipynb:697, A.8)
//   k_compact_*    : finite/positive mask + ordered gather to XYZ/RGB      (main.ipynb:726-737)
#include "sgbm_common.cuh"

// Start of a frame's winner-take-all: raw disparity = INV everywhere, disp2 keys = "never hit".  One kernel, 128-bit stores.
__global__ void k_init_wta(int16_t *raw, unsigned int *d2key, size_t n, int INV)
{
    const size_t i8 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
    if (i8 >= n) return;
    const unsigned int v2 = ((unsigned int)INV & 0xFFFFu) * 0x10001u;
    if (i8 + 8 <= n) {
        *reinterpret_cast<uint4 *>(raw + i8) = make_uint4(v2, v2, v2, v2);
        const uint4 ones = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
        reinterpret_cast<uint4 *>(d2key + i8)[0] = ones;
        reinterpret_cast<uint4 *>(d2key + i8)[1] = ones;
    } else {
        for (size_t i = i8; i < n; i++) { raw[i] = (int16_t)INV; d2key[i] = 0xFFFFFFFFu; }
    }
}

// numDisparities % 8 != 0: the padding disparities d in [D, Dc) of every column of `nrows` rows of the cost volume get the
// constant that makes them inert in the path step (sgbm_api.cu: pad_cost_value).  Rare configurations: plain 16-bit stores.
__global__ void k_pad_cost(uint16_t *C, long long ncols, int Dp, int D, int Dc, int nreg, int lpc, int value)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int npad = Dc - D;
    if (i >= ncols * npad) return;
    const long long col = i / npad;
    const int d = D + (int)(i - col * npad);
    C[col * Dp + sgbm_pos(d, nreg, lpc)] = (uint16_t)value;
}

// ---- LR check ---------------------------------------------------------------------------------
// d2key[y][x2] = (cost << 16) | (0xFFFF - x1) of the winning left pixel, 0xFFFFFFFF if never hit.
// disp2[x2] = (x1 + minX1) - x2 for a hit, INV (the x16-scaled marker, [P2]) otherwise.
__device__ __forceinline__ int disp2_at(const unsigned int *krow, int p, int minX1, int INV)
{
    unsigned int k = krow[p];
    if (k == 0xFFFFFFFFu) return INV;
    return (int)(0xFFFFu - (k & 0xFFFFu)) + minX1 - p;
}

__global__ void k_lrcheck(int16_t *raw, const unsigned int *d2key, int W, int H, int minX1, int maxX1,
                          int minD, int INV, int DMD)
{
    int x = minX1 + blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x >= maxX1) return;
    int d1 = raw[(size_t)y * W + x];
    if (d1 == INV) return;
    const unsigned int *krow = d2key + (size_t)y * W;
    int _d = d1 >> 4, d_ = (d1 + 15) >> 4;
    int _x = x - _d, x_ = x - d_;
    bool bad0 = false, bad1 = false;
    if (0 <= _x && _x < W) { int v = disp2_at(krow, _x, minX1, INV); bad0 = v >= minD && abs(v - _d) > DMD; }
    if (0 <= x_ && x_ < W) { int v = disp2_at(krow, x_, minX1, INV); bad1 = v >= minD && abs(v - d_) > DMD; }
    if (bad0 && bad1) raw[(size_t)y * W + x] = (int16_t)INV;
}

// ---- 3x3 median, replicate border -----------------------------------------------------------------
__device__ __forceinline__ void cswap(int &a, int &b) { int t = min(a, b); b = max(a, b); a = t; }
// the same on two signed 16-bit values per register
__device__ __forceinline__ void pcswap(uint32_t &a, uint32_t &b) { const uint32_t t = __vmins2(a, b); b = __vmaxs2(a, b); a = t; }

__global__ void k_median3x3(const int16_t *src, int16_t *dst, int W, int H, long long dstPitchElems)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x >= W) return;
    int xm = max(x - 1, 0), xp = min(x + 1, W - 1);
    const int16_t *r0 = src + (size_t)max(y - 1, 0) * W, *r1 = src + (size_t)y * W, *r2 = src + (size_t)min(y + 1, H - 1) * W;
    int p0 = r0[xm], p1 = r0[x], p2 = r0[xp], p3 = r1[xm], p4 = r1[x], p5 = r1[xp], p6 = r2[xm], p7 = r2[x], p8 = r2[xp];
    // 19-exchange median-of-9 network
    cswap(p1, p2); cswap(p4, p5); cswap(p7, p8); cswap(p0, p1); cswap(p3, p4); cswap(p6, p7);
    cswap(p1, p2); cswap(p4, p5); cswap(p7, p8); cswap(p0, p3); cswap(p5, p8); cswap(p4, p7);
    cswap(p3, p6); cswap(p1, p4); cswap(p2, p5); cswap(p4, p7); cswap(p4, p2); cswap(p6, p4);
    cswap(p4, p2);
    dst[(size_t)y * dstPitchElems + x] = (int16_t)p4;
}

// ---- LR check + 3x3 median in one pass (the frame pipeline; the two kernels above serve the debug hook and the public
// medianBlur3) ----------------------------------------------------------------------------------------------------------
// A CTA produces a LM_TW x LM_TH tile of the median: it first applies the LR check to the (LM_TW + 2) x (LM_TH + 2) raw
// pixels the tile's windows touch (replicate border = the checked value of the clamped pixel) into shared memory, then takes
// the medians from there: the raw image is read once and the checked image never exists in HBM.
#define LM_TW 128
#define LM_TH 16
#define LM_THREADS 256
__device__ __forceinline__ int lr_checked(const int16_t *raw, const unsigned int *d2key, int W, int x, int y, int minX1, int maxX1,
                                          int minD, int INV, int DMD)
{
    const int d1 = raw[(size_t)y * W + x];
    if (d1 == INV || x < minX1 || x >= maxX1) return d1;
    const unsigned int *krow = d2key + (size_t)y * W;
    const int _d = d1 >> 4, d_ = (d1 + 15) >> 4;
    const int _x = x - _d, x_ = x - d_;
    bool bad0 = false, bad1 = false;
    if (0 <= _x && _x < W) { const int v = disp2_at(krow, _x, minX1, INV); bad0 = v >= minD && abs(v - _d) > DMD; }
    if (0 <= x_ && x_ < W) { const int v = disp2_at(krow, x_, minX1, INV); bad1 = v >= minD && abs(v - d_) > DMD; }
    return (bad0 && bad1) ? INV : d1;
}

__global__ void __launch_bounds__(LM_THREADS) k_lr_median(const int16_t *raw, const unsigned int *d2key, int16_t *dst, int W, int H,
                                                           long long dstPitchElems, int minX1, int maxX1, int minD, int INV, int DMD)
{
    __shared__ __align__(16) int16_t tile[LM_TH + 2][LM_TW + 2 + 2];
    const int x0 = blockIdx.x * LM_TW, y0 = blockIdx.y * LM_TH;
    for (int idx = threadIdx.x; idx < (LM_TH + 2) * (LM_TW + 2); idx += LM_THREADS) {
        const int ty = idx / (LM_TW + 2), tx = idx - ty * (LM_TW + 2);
        const int gy = min(max(y0 + ty - 1, 0), H - 1), gx = min(max(x0 + tx - 1, 0), W - 1);
        tile[ty][tx] = (int16_t)lr_checked(raw, d2key, W, gx, gy, minX1, maxX1, minD, INV, DMD);
    }
    __syncthreads();
    // thread -> two adjacent columns x four rows.  The two columns' 3x3 windows are sorted together as packed signed 16-bit
    // pairs (low half: column cx, high half: column cx + 1): one 19-exchange network of packed min / max per two pixels.
    const int cx = (threadIdx.x & 63) * 2, ry = (threadIdx.x >> 6) * 4;
    uint32_t P[6][3];                                     // tile rows ry .. ry + 5, window columns 0 .. 2
#pragma unroll
    for (int t = 0; t < 6; t++) {
        const uint32_t *w = reinterpret_cast<const uint32_t *>(&tile[ry + t][cx]);      // cx is even, rows are 4-byte aligned
        const uint32_t w0 = w[0], w1 = w[1];
        P[t][0] = w0; P[t][1] = __byte_perm(w0, w1, 0x5432); P[t][2] = w1;
    }
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const int y = y0 + ry + r;
        if (y >= H) break;
        uint32_t p0 = P[r][0], p1 = P[r][1], p2 = P[r][2], p3 = P[r + 1][0], p4 = P[r + 1][1], p5 = P[r + 1][2], p6 = P[r + 2][0],
                 p7 = P[r + 2][1], p8 = P[r + 2][2];
        pcswap(p1, p2); pcswap(p4, p5); pcswap(p7, p8); pcswap(p0, p1); pcswap(p3, p4); pcswap(p6, p7);
        pcswap(p1, p2); pcswap(p4, p5); pcswap(p7, p8); pcswap(p0, p3); pcswap(p5, p8); pcswap(p4, p7);
        pcswap(p3, p6); pcswap(p1, p4); pcswap(p2, p5); pcswap(p4, p7); pcswap(p4, p2); pcswap(p6, p4);
        pcswap(p4, p2);
        const int x = x0 + cx;
        if (x < W) dst[(size_t)y * dstPitchElems + x] = (int16_t)(p4 & 0xFFFFu);
        if (x + 1 < W) dst[(size_t)y * dstPitchElems + x + 1] = (int16_t)(p4 >> 16);
    }
}

// ---- speckle filter: connected components by union-find ----------------------------------------
__device__ __forceinline__ int uf_find(const int *L, int i)
{
    int p = ((volatile const int *)L)[i];
    while (p != i) { i = p; p = ((volatile const int *)L)[i]; }
    return i;
}
// Find with path halving.  Parents only ever decrease (union by smaller index), so shortening a link with
// atomicMin can never undo a concurrent union; the result is the same root, later finds are shorter.
__device__ __forceinline__ int uf_find_halve(int *L, int i)
{
    int p = ((volatile int *)L)[i];
    while (p != i) {
        const int gp = ((volatile int *)L)[p];
        if (gp != p) atomicMin(&L[i], gp);
        i = p; p = gp;
    }
    return i;
}
__device__ __forceinline__ void uf_union(int *L, int a, int b)
{
    bool done;
    do {
        a = uf_find_halve(L, a);
        b = uf_find_halve(L, b);
        if (a < b) { int old = atomicMin(&L[b], a); done = (old == b); b = old; }
        else if (b < a) { int old = atomicMin(&L[a], b); done = (old == a); a = old; }
        else done = true;
    } while (!done);
}

// Run-based labelling.  A "run" is a maximal horizontal segment of connected pixels; every pixel
// is labelled with the index of its run's first pixel (no atomics: segmented max-scan per row),
// so the union-find only has to merge runs of adjacent rows, and sizes are added per run.
#define CC_THREADS 256
__device__ __forceinline__ bool cc_link(int a, int b, int newVal, int maxDiff)
{
    return a != newVal && b != newVal && abs(a - b) <= maxDiff;
}

// one CTA per row: label[i] = first pixel of the run containing i (or -1), size[i] = 0
__global__ void __launch_bounds__(CC_THREADS) k_cc_runs(const int16_t *img, int *label, int *size, int W, int newVal, int maxDiff)
{
    __shared__ int wmax[CC_THREADS / 32];
    __shared__ int carry;
    const int y = blockIdx.x;
    const int16_t *row = img + (size_t)y * W;
    if (threadIdx.x == 0) carry = -1;
    __syncthreads();
    for (int x0 = 0; x0 < W; x0 += CC_THREADS) {
        const int x = x0 + threadIdx.x;
        int start = -1, v = newVal;
        if (x < W) {
            v = row[x];
            const bool linked = x > 0 && cc_link(row[x - 1], v, newVal, maxDiff);
            if (!linked) start = x;                       // run starts here (also for invalid pixels)
        }
        int m = start;                                    // inclusive max-scan of run starts
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xFFFFFFFFu, m, o);
            if ((threadIdx.x & 31) >= o) m = max(m, t);
        }
        if ((threadIdx.x & 31) == 31) wmax[threadIdx.x >> 5] = m;
        __syncthreads();
        int pre = carry;
        for (int w = 0; w < (int)(threadIdx.x >> 5); w++) pre = max(pre, wmax[w]);
        m = max(m, pre);
        if (x < W) {
            const size_t i = (size_t)y * W + x;
            label[i] = (v != newVal) ? y * W + m : -1;
            size[i] = 0;
        }
        __syncthreads();
        if (threadIdx.x == CC_THREADS - 1) carry = m;
        __syncthreads();
    }
}

// The three kernels below handle CC_PX consecutive pixels of a row per thread: their work per pixel is a few dependent
// global loads (labels, roots, sizes), so with one pixel per thread they were bound by load latency at ~25 % issue
// utilisation; four independent chains per thread hide it (4K: merge 0.082 / count 0.067 / apply 0.041 ms before).
#define CC_PX 4

// vertical merges: one union per contact between a run of row y and a run of row y+1
__global__ void k_cc_merge(const int16_t *img, int *label, int W, int H, int newVal, int maxDiff)
{
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * CC_PX;
    const int y = blockIdx.y;
    if (x0 >= W || y + 1 >= H) return;
    const int16_t *ra = img + (size_t)y * W, *rb = ra + W;
    int a[CC_PX + 1], b[CC_PX + 1];                       // columns x0 - 1 .. x0 + CC_PX - 1
#pragma unroll
    for (int k = 0; k <= CC_PX; k++) {
        const int x = x0 - 1 + k;
        const bool in = x >= 0 && x < W;
        a[k] = in ? ra[x] : newVal;
        b[k] = in ? rb[x] : newVal;
    }
    bool todo[CC_PX];
    int la[CC_PX], lb[CC_PX];
#pragma unroll
    for (int k = 0; k < CC_PX; k++) {
        const int x = x0 + k;
        bool t = x < W && cc_link(a[k + 1], b[k + 1], newVal, maxDiff);
        // the same two runs already touched at x-1?
        if (t && x > 0 && cc_link(a[k], a[k + 1], newVal, maxDiff) && cc_link(b[k], b[k + 1], newVal, maxDiff) &&
            cc_link(a[k], b[k], newVal, maxDiff)) t = false;
        todo[k] = t;
        la[k] = lb[k] = 0;
        if (t) { la[k] = label[y * W + x]; lb[k] = label[(y + 1) * W + x]; }
    }
#pragma unroll
    for (int k = 0; k < CC_PX; k++)
        if (todo[k]) uf_union(label, la[k], lb[k]);
}

// run ends add their run length to the root's size and flatten the run start's label
__global__ void k_cc_count(const int16_t *img, int *label, int *size, int W, int H, int newVal, int maxDiff)
{
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * CC_PX;
    const int y = blockIdx.y;
    if (x0 >= W) return;
    const int16_t *row = img + (size_t)y * W;
    int v[CC_PX + 2];                                     // columns x0 - 1 .. x0 + CC_PX
#pragma unroll
    for (int k = 0; k < CC_PX + 2; k++) {
        const int x = x0 - 1 + k;
        v[k] = (x >= 0 && x < W) ? row[x] : newVal;
    }
    int s[CC_PX];
    bool end[CC_PX];
#pragma unroll
    for (int k = 0; k < CC_PX; k++) {
        const int x = x0 + k, i = y * W + x;
        const int c = v[k + 1];
        end[k] = x < W && c != newVal && !cc_link(c, v[k + 2], newVal, maxDiff);          // (v past the row is newVal: no link)
        s[k] = 0;
        if (end[k]) {
            const bool runStart = !cc_link(v[k], c, newVal, maxDiff);                      // (v before the row is newVal)
            s[k] = runStart ? i : label[i];               // non-start pixels keep pointing at their run start
        }
    }
#pragma unroll
    for (int k = 0; k < CC_PX; k++) {
        if (!end[k]) continue;
        const int i = y * W + x0 + k;
        const int r = uf_find(label, s[k]);
        atomicAdd(&size[r], i - s[k] + 1);
        if (r != s[k]) atomicMin(&label[s[k]], r);        // flatten: k_cc_apply reaches the root in two loads
    }
}

__global__ void k_cc_apply(int16_t *img, const int *label, const int *size, int n, int newVal, int maxSize)
{
    const int i0 = (blockIdx.x * blockDim.x + threadIdx.x) * CC_PX;
    if (i0 >= n) return;
    int s[CC_PX], p[CC_PX];
#pragma unroll
    for (int k = 0; k < CC_PX; k++) s[k] = i0 + k < n ? label[i0 + k] : -1;
#pragma unroll
    for (int k = 0; k < CC_PX; k++) p[k] = s[k] >= 0 ? label[s[k]] : -1;          // first hop of all four chains at once
#pragma unroll
    for (int k = 0; k < CC_PX; k++) {
        if (s[k] < 0) continue;
        int r = s[k], q = p[k];
        while (q != r) { r = q; q = label[r]; }           // labels are final here
        s[k] = r;
    }
#pragma unroll
    for (int k = 0; k < CC_PX; k++) p[k] = s[k] >= 0 ? size[s[k]] : 0x7FFFFFFF;
#pragma unroll
    for (int k = 0; k < CC_PX; k++)
        if (p[k] <= maxSize) img[i0 + k] = (int16_t)newVal;
}

// ---- float conversion + reprojection --------------------------------------------------------------
__device__ __forceinline__ float disp_float_masked(int d16)
{
    float f = (float)d16 / 16.0f;                 // exact
    return f * ((f > 0.0f) ? 1.0f : 0.0f);        // numpy: disparity_map * mask  (-0.0 for negatives)
}

__global__ void k_disp_to_float(const int16_t *d, float *out, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = disp_float_masked(d[i]);
}

struct QMat { double q[16]; };

// h = Q * (x, y, d, 1)^T left to right in fp64 without contraction; out = f32( f64(f32(h_c)) * (1/h_3) )
__device__ __forceinline__ void reproject_px(const QMat &Q, int x, int y, double d, float &X, float &Y, float &Z)
{
    double h[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        double s = __dadd_rn(__dmul_rn(Q.q[4 * k + 0], (double)x), __dmul_rn(Q.q[4 * k + 1], (double)y));
        s = __dadd_rn(s, __dmul_rn(Q.q[4 * k + 2], d));
        h[k] = __dadd_rn(s, Q.q[4 * k + 3]);
    }
    double iw = __ddiv_rn(1.0, h[3]);
    X = __double2float_rn(__dmul_rn((double)__double2float_rn(h[0]), iw));
    Y = __double2float_rn(__dmul_rn((double)__double2float_rn(h[1]), iw));
    Z = __double2float_rn(__dmul_rn((double)__double2float_rn(h[2]), iw));
}

template <typename T>
__global__ void k_reproject(const T *disp, QMat Q, int W, int H, float *xyz, uint8_t *valid)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x >= W) return;
    size_t i = (size_t)y * W + x;
    double d = (double)disp[i];
    float X, Y, Z;
    reproject_px(Q, x, y, d, X, Y, Z);
    xyz[3 * i + 0] = X; xyz[3 * i + 1] = Y; xyz[3 * i + 2] = Z;
    if (valid) valid[i] = (isfinite(X) && d > 0.0) ? 1 : 0;
}

// ---- reprojectImageTo3D with handleMissingValues / ddepth (A.8) ---------------------------------------
// cv2: Z = 10000 where |d - min(disp)| <= FLT_EPSILON (d and the minimum as doubles); ddepth CV_16S / CV_32S
// round the float32 result half-to-even like cvRound (cvtss2si: anything that does not fit an int32,
// NaN and +-inf included, becomes INT_MIN) and CV_16S then saturates to [-32768, 32767].
// The minimum goes through an order-preserving key so that one atomicMin does floats of either sign.
__device__ __forceinline__ unsigned int order_key_f32(float f)
{
    const unsigned int u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float order_key_inv_f32(unsigned int k)
{
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}
template <typename T>
__global__ void k_disp_min(const T *disp, size_t n, unsigned int *keyOut)
{
    unsigned int k = 0xFFFFFFFFu;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float f = (float)disp[i];                 // cv2 converts integer disparities to float32 first
        if (!(f != f)) k = min(k, order_key_f32(f));    // minMaxIdx ignores nothing, but NaN never compares smaller
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) k = min(k, __shfl_xor_sync(0xFFFFFFFFu, k, o));
    if ((threadIdx.x & 31) == 0) atomicMin(keyOut, k);
}
__device__ __forceinline__ int cv_round_f32(float v)
{
    const float r = rintf(v);                           // half to even, like cvtss2si under the default MXCSR
    return (r >= -2147483648.0f && r < 2147483648.0f) ? (int)r : (int)0x80000000;
}
// DD: 5 = float32, 3 = int16, 4 = int32 (cv2 depth codes)
template <typename T, int DD>
__global__ void k_reproject_ex(const T *disp, QMat Q, int W, int H, void *out, const unsigned int *minKey)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x >= W) return;
    size_t i = (size_t)y * W + x;
    const double d = (double)(float)disp[i];
    float v[3];
    reproject_px(Q, x, y, d, v[0], v[1], v[2]);
    if (minKey && fabs(d - (double)order_key_inv_f32(*minKey)) <= 1.1920928955078125e-07) v[2] = 10000.0f;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        if (DD == 5) ((float *)out)[3 * i + c] = v[c];
        else if (DD == 4) ((int *)out)[3 * i + c] = cv_round_f32(v[c]);
        else { const int r = cv_round_f32(v[c]); ((int16_t *)out)[3 * i + c] = (int16_t)max(-32768, min(32767, r)); }
    }
}

// ---- fused tail: /16, mask, reproject, validity, ordered compaction ---------------------------------
#define CP_THREADS 256
#define CP_ITEMS 4                      // pixels per thread, CP_THREADS*CP_ITEMS pixels per block

__global__ void k_compact_count(const int16_t *disp, QMat Q, int W, int H, unsigned int *blockCount)
{
    __shared__ unsigned int wsum[CP_THREADS / 32];
    size_t n = (size_t)W * H;
    size_t base = (size_t)blockIdx.x * CP_THREADS * CP_ITEMS + (size_t)threadIdx.x * CP_ITEMS;
    unsigned int cnt = 0;
#pragma unroll
    for (int k = 0; k < CP_ITEMS; k++) {
        size_t i = base + k;
        if (i < n) {
            float f = disp_float_masked(disp[i]);
            float X, Y, Z;
            reproject_px(Q, (int)(i % W), (int)(i / W), (double)f, X, Y, Z);
            cnt += (isfinite(X) && f > 0.0f) ? 1u : 0u;
        }
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int s = 0;
        for (int w = 0; w < CP_THREADS / 32; w++) s += wsum[w];
        blockCount[blockIdx.x] = s;
    }
}

// single-CTA exclusive scan of the block counts (nblocks <= a few 10^4)
__global__ void k_compact_scan(const unsigned int *blockCount, unsigned long long *blockOffset, int nblocks,
                               unsigned long long *total)
{
    __shared__ unsigned long long carry;
    __shared__ unsigned long long wtot[32];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int b0 = 0; b0 < nblocks; b0 += blockDim.x) {
        int i = b0 + threadIdx.x;
        unsigned long long v = i < nblocks ? blockCount[i] : 0, incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if ((threadIdx.x & 31) >= o) incl += t;
        }
        if ((threadIdx.x & 31) == 31) wtot[threadIdx.x >> 5] = incl;
        __syncthreads();
        unsigned long long woff = 0;
        for (int w = 0; w < (int)(threadIdx.x >> 5); w++) woff += wtot[w];
        if (i < nblocks) blockOffset[i] = carry + woff + incl - v;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry += woff + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

__global__ void k_compact_scatter(const int16_t *disp, QMat Q, int W, int H, const uint8_t *bgr, int bgrCn,
                                  long long bgrPitch, const unsigned long long *blockOffset, float *xyz,
                                  uint8_t *rgb)
{
    __shared__ unsigned int wsum[CP_THREADS / 32];
    size_t n = (size_t)W * H;
    size_t base = (size_t)blockIdx.x * CP_THREADS * CP_ITEMS + (size_t)threadIdx.x * CP_ITEMS;
    float P[CP_ITEMS][3];
    bool ok[CP_ITEMS];
    unsigned int cnt = 0;
#pragma unroll
    for (int k = 0; k < CP_ITEMS; k++) {
        size_t i = base + k;
        ok[k] = false;
        if (i < n) {
            float f = disp_float_masked(disp[i]);
            reproject_px(Q, (int)(i % W), (int)(i / W), (double)f, P[k][0], P[k][1], P[k][2]);
            ok[k] = isfinite(P[k][0]) && f > 0.0f;
            cnt += ok[k] ? 1u : 0u;
        }
    }
    unsigned int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned int t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if ((threadIdx.x & 31) >= o) incl += t;
    }
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = incl;
    __syncthreads();
    unsigned int woff = 0;
    for (int w = 0; w < (int)(threadIdx.x >> 5); w++) woff += wsum[w];
    unsigned long long pos = blockOffset[blockIdx.x] + woff + incl - cnt;
#pragma unroll
    for (int k = 0; k < CP_ITEMS; k++) {
        if (ok[k]) {
            size_t i = base + k;
            xyz[3 * pos + 0] = P[k][0]; xyz[3 * pos + 1] = P[k][1]; xyz[3 * pos + 2] = P[k][2];
            if (rgb) {
                int x = (int)(i % W), y = (int)(i / W);
                const uint8_t *s = bgr + (size_t)y * bgrPitch + (size_t)x * bgrCn;
                if (bgrCn == 3) { rgb[3 * pos + 0] = s[2]; rgb[3 * pos + 1] = s[1]; rgb[3 * pos + 2] = s[0]; }
                else { rgb[3 * pos + 0] = s[0]; rgb[3 * pos + 1] = s[0]; rgb[3 * pos + 2] = s[0]; }
            }
            pos++;
        }
    }
}

// =================================================================================================
// host launchers
// =================================================================================================
int sgbm_launch_pad_cost(const Geo &g, uint16_t *C, int nrows, int value, cudaStream_t st)
{
    const int Dc = (g.D + 7) & ~7;
    if (Dc == g.D || nrows <= 0) return 0;
    const long long ncols = (long long)g.W1 * nrows, n = ncols * (Dc - g.D);
    k_pad_cost<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(C, ncols, g.Dp, g.D, Dc, g.nreg, g.lpc, value);
    sgbm_count_launch(1);
    SGBM_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int sgbm_launch_init_wta(int16_t *raw, unsigned int *d2key, size_t n, int INV, cudaStream_t st)
{
    const size_t threads = (n + 7) / 8;
    k_init_wta<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(raw, d2key, n, INV);
    sgbm_count_launch(1);
    SGBM_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int sgbm_launch_lr_median(const Geo &g, const int16_t *raw, const unsigned int *d2key, int16_t *dst, long long dstPitchElems,
                          cudaStream_t st)
{
    dim3 grid((g.W + LM_TW - 1) / LM_TW, (g.H + LM_TH - 1) / LM_TH);
    k_lr_median<<<grid, LM_THREADS, 0, st>>>(raw, d2key, dst, g.W, g.H, dstPitchElems, g.minX1, g.maxX1, g.minD, g.INV, g.DMD);
    sgbm_count_launch(1);
    SGBM_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int sgbm_launch_lrcheck(const Geo &g, int16_t *raw, const unsigned int *d2key, cudaStream_t st)
{
    dim3 grid((g.W1 + 255) / 256, g.H);
    k_lrcheck<<<grid, 256, 0, st>>>(raw, d2key, g.W, g.H, g.minX1, g.maxX1, g.minD, g.INV, g.DMD);
    sgbm_count_launch(1);
    SGBM_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int sgbm_launch_median(const int16_t *src, int16_t *dst, int W, int H, long long dstPitchElems, cudaStream_t st)
{
    dim3 grid((W + 255) / 256, H);
    k_median3x3<<<grid, 256, 0, st>>>(src, dst, W, H, dstPitchElems);
    sgbm_count_launch(1);
    SGBM_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int sgbm_launch_speckles(int16_t *img, int W, int H, int newVal, int maxSize, int maxDiff, void *scratch,
                         cudaStream_t st)
{
    int n = W * H;
    int *label = (int *)scratch, *size = label + n;
    k_cc_runs<<<H, CC_THREADS, 0, st>>>(img, label, size, W, newVal, maxDiff);
    dim3 grid((W + 256 * CC_PX - 1) / (256 * CC_PX), H);
    k_cc_merge<<<grid, 256, 0, st>>>(img, label, W, H, newVal, maxDiff);
    k_cc_count<<<grid, 256, 0, st>>>(img, label, size, W, H, newVal, maxDiff);
    k_cc_apply<<<(n + 256 * CC_PX - 1) / (256 * CC_PX), 256, 0, st>>>(img, label, size, n, newVal, maxSize);
    sgbm_count_launch(4);
    SGBM_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int sgbm_launch_disp_to_float(const int16_t *d, float *out, size_t n, cudaStream_t st)
{
    k_disp_to_float<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d, out, n);
    sgbm_count_launch(1);
    SGBM_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int sgbm_launch_reproject(const void *disp, int isFloat, const double *Qh, int W, int H, float *xyz,
                          uint8_t *valid, cudaStream_t st)
{
    QMat Q;
    for (int i = 0; i < 16; i++) Q.q[i] = Qh[i];
    dim3 grid((W + 255) / 256, H);
    if (isFloat) k_reproject<float><<<grid, 256, 0, st>>>((const float *)disp, Q, W, H, xyz, valid);
    else k_reproject<int16_t><<<grid, 256, 0, st>>>((const int16_t *)disp, Q, W, H, xyz, valid);
    sgbm_count_launch(1);
    SGBM_CUDA_CHECK(cudaGetLastError());
    return 0;
}

template <typename T>
static int launch_reproject_ex_t(const T *disp, const QMat &Q, int W, int H, int handleMissing, int ddepth, void *out,
                                 unsigned int *scratch, cudaStream_t st)
{
    dim3 grid((W + 255) / 256, H);
    unsigned int *key = handleMissing ? scratch : nullptr;
    if (key) {
        SGBM_CUDA_CHECK(cudaMemsetAsync(key, 0xFF, 4, st));
        const size_t n = (size_t)W * H;
        const unsigned blocks = (unsigned)((n + 256 * 8 - 1) / (256 * 8));
        k_disp_min<T><<<blocks < 4096u ? blocks : 4096u, 256, 0, st>>>(disp, n, key);
        sgbm_count_launch(1);
    }
    if (ddepth == 3) k_reproject_ex<T, 3><<<grid, 256, 0, st>>>(disp, Q, W, H, out, key);
    else if (ddepth == 4) k_reproject_ex<T, 4><<<grid, 256, 0, st>>>(disp, Q, W, H, out, key);
    else k_reproject_ex<T, 5><<<grid, 256, 0, st>>>(disp, Q, W, H, out, key);
    sgbm_count_launch(1);
    SGBM_CUDA_CHECK(cudaGetLastError());
    return 0;
}

// dispDepth / ddepth: cv2 depth codes (0 = uint8, 3 = int16, 4 = int32, 5 = float32)
int sgbm_launch_reproject_ex(const void *disp, int dispDepth, const double *Qh, int W, int H, int handleMissing, int ddepth,
                             void *out, void *scratch, cudaStream_t st)
{
    QMat Q;
    for (int i = 0; i < 16; i++) Q.q[i] = Qh[i];
    unsigned int *sc = (unsigned int *)scratch;
    switch (dispDepth) {
    case 0: return launch_reproject_ex_t((const uint8_t *)disp, Q, W, H, handleMissing, ddepth, out, sc, st);
    case 3: return launch_reproject_ex_t((const int16_t *)disp, Q, W, H, handleMissing, ddepth, out, sc, st);
    case 4: return launch_reproject_ex_t((const int *)disp, Q, W, H, handleMissing, ddepth, out, sc, st);
    case 5: return launch_reproject_ex_t((const float *)disp, Q, W, H, handleMissing, ddepth, out, sc, st);
    }
    return sgbm_fail(-1, "unsupported disparity depth %d", dispDepth);
}

size_t sgbm_compact_scratch_bytes(int W, int H)
{
    size_t nb = ((size_t)W * H + CP_THREADS * CP_ITEMS - 1) / (CP_THREADS * CP_ITEMS);
    return nb * (sizeof(unsigned int) + sizeof(unsigned long long)) + 64;
}

int sgbm_launch_compact(const int16_t *disp, const double *Qh, int W, int H, const uint8_t *bgr, int bgrCn,
                        long long bgrPitch, float *xyz, uint8_t *rgb, unsigned long long *nOut, void *scratch,
                        cudaStream_t st)
{
    QMat Q;
    for (int i = 0; i < 16; i++) Q.q[i] = Qh[i];
    int nb = (int)(((size_t)W * H + CP_THREADS * CP_ITEMS - 1) / (CP_THREADS * CP_ITEMS));
    unsigned long long *off = (unsigned long long *)scratch;
    unsigned int *cnt = (unsigned int *)(off + nb);
    k_compact_count<<<nb, CP_THREADS, 0, st>>>(disp, Q, W, H, cnt);
    k_compact_scan<<<1, 1024, 0, st>>>(cnt, off, nb, nOut);
    k_compact_scatter<<<nb, CP_THREADS, 0, st>>>(disp, Q, W, H, bgr, bgrCn, bgrPitch, off, xyz, rgb);
    sgbm_count_launch(3);
    SGBM_CUDA_CHECK(cudaGetLastError());
    return 0;
}
