// sgbm_rectify.cu -- the rectification warp that precedes the dense-stereo path on every frame
// (SURVEY.md 8(f) n1):
//     mapL1, mapL2 = cv2.initUndistortRectifyMap(K0, None, R1, P1, image_size, cv2.CV_32F)   main.ipynb:496-497
//     imgL_rect    = cv2.remap(imgL, mapL1, mapL2, interpolation=cv2.INTER_LINEAR)           main.ipynb:499-500
// (same calls at gui.py:160-164).  Both are restated from their observable arithmetic (validated
// against the cv2 binary, oracle/rectify.py):
//   * map: [x y w]^T = (P[:, :3] * R)^-1 [u v 1]^T in fp64, map = f * (x / w) + c, cast to fp32.  The
//     reference implementation walks a row in blocks of 8 pixels (its AVX-512 code path): the block
//     base is accumulated sequentially (base += 8 * ir), the lanes add l * ir.  The kernel keeps
//     exactly that association order (no FMA contraction), which makes the maps bit-identical on hosts
//     whose OpenCV dispatches to AVX-512; on other hosts OpenCV's own result differs from this one
//     by one fp32 ulp in ~1e-6 of the entries.
//   * remap INTER_LINEAR, BORDER_CONSTANT(0), 8-bit: coordinates are rounded to 1/32 pixel
//     (round-half-even of map * 32), the four taps are weighted with the integers
//     32 * (32 - fx | fx) * (32 - fy | fy) (sum 2^15) and the result is (acc + 2^14) >> 15.  Exact.
#include "sgbm_common.cuh"

// one thread per (row, lane): 8 lanes walk the row block by block with a sequentially accumulated base
__global__ void k_rectify_map(int W, int H, double ir0, double ir1, double ir2, double ir3, double ir4, double ir5, double ir6,
                              double ir7, double ir8, double fx, double fy, double u0, double v0, float *map1, float *map2)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = t >> 3, l = t & 7;
    if (y >= H) return;
    const double yd = (double)y;
    double bx = __dadd_rn(__dmul_rn(yd, ir1), ir2), by = __dadd_rn(__dmul_rn(yd, ir4), ir5), bw = __dadd_rn(__dmul_rn(yd, ir7), ir8);
    const double lx = __dmul_rn(ir0, (double)l), ly = __dmul_rn(ir3, (double)l), lw = __dmul_rn(ir6, (double)l);
    const double sx = __dmul_rn(8.0, ir0), sy = __dmul_rn(8.0, ir3), sw = __dmul_rn(8.0, ir6);
    const int nb = W / 8;
    float *o1 = map1 + (size_t)y * W, *o2 = map2 + (size_t)y * W;
    for (int b = 0; b < nb; b++) {
        const double w = __ddiv_rn(1.0, __dadd_rn(bw, lw));
        const double x = __dmul_rn(__dadd_rn(bx, lx), w), yy = __dmul_rn(__dadd_rn(by, ly), w);
        o1[b * 8 + l] = (float)__dadd_rn(__dmul_rn(fx, x), u0);
        o2[b * 8 + l] = (float)__dadd_rn(__dmul_rn(fy, yy), v0);
        bx = __dadd_rn(bx, sx); by = __dadd_rn(by, sy); bw = __dadd_rn(bw, sw);
    }
    if (l == 0) {                                          // scalar tail of the row, continuing from the last base
        for (int j = nb * 8; j < W; j++) {
            const double w = __ddiv_rn(1.0, bw);
            o1[j] = (float)__dadd_rn(__dmul_rn(fx, __dmul_rn(bx, w)), u0);
            o2[j] = (float)__dadd_rn(__dmul_rn(fy, __dmul_rn(by, w)), v0);
            bx = __dadd_rn(bx, ir0); by = __dadd_rn(by, ir3); bw = __dadd_rn(bw, ir6);
        }
    }
}

// cv::invert of a 3x3 double matrix (closed form: determinant + adjugate, the order OpenCV uses)
static bool inv3x3(const double *S, double *t)
{
    double d = S[0] * (S[4] * S[8] - S[5] * S[7]) - S[1] * (S[3] * S[8] - S[5] * S[6]) + S[2] * (S[3] * S[7] - S[4] * S[6]);
    if (d == 0.0) return false;
    d = 1.0 / d;
    t[0] = (S[4] * S[8] - S[5] * S[7]) * d; t[1] = (S[2] * S[7] - S[1] * S[8]) * d; t[2] = (S[1] * S[5] - S[2] * S[4]) * d;
    t[3] = (S[5] * S[6] - S[3] * S[8]) * d; t[4] = (S[0] * S[8] - S[2] * S[6]) * d; t[5] = (S[2] * S[3] - S[0] * S[5]) * d;
    t[6] = (S[3] * S[7] - S[4] * S[6]) * d; t[7] = (S[1] * S[6] - S[0] * S[7]) * d; t[8] = (S[0] * S[4] - S[1] * S[3]) * d;
    return true;
}

int sgbm_launch_rectify_map(const double *K, const double *R, const double *P, int pcols, int W, int H, float *map1,
                            float *map2, cudaStream_t st)
{
    double A[9], iR[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            if (R) {
                double s = 0.0;
                for (int k = 0; k < 3; k++) s += P[i * pcols + k] * R[k * 3 + j];
                A[i * 3 + j] = s;
            } else {
                A[i * 3 + j] = P[i * pcols + j];
            }
        }
    if (!inv3x3(A, iR)) return sgbm_fail(-1, "newCameraMatrix * R is singular");
    const int threads = 128, total = H * 8;
    k_rectify_map<<<(total + threads - 1) / threads, threads, 0, st>>>(W, H, iR[0], iR[1], iR[2], iR[3], iR[4], iR[5], iR[6], iR[7],
                                                                        iR[8], K[0], K[4], K[2], K[5], map1, map2);
    sgbm_count_launch(1);
    SGBM_CUDA_CHECK(cudaGetLastError());
    return 0;
}

// dst(y, x) = bilinear sample of src at (map1, map2), 1/32-pixel fixed point, constant border 0
template <int CN>
__global__ void k_remap_linear(const uint8_t *__restrict__ src, int sw, int sh, long long spitch, const float *__restrict__ map1,
                               const float *__restrict__ map2, int W, int H, uint8_t *__restrict__ dst, long long dpitch)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const size_t mi = (size_t)y * W + x;
    // cvRound(map * INTER_TAB_SIZE): fp32 product, round half to even; NaN / overflow follow cvtss2si (INT_MIN)
    const float px = __fmul_rn(map1[mi], 32.0f), py = __fmul_rn(map2[mi], 32.0f);
    const int sx = (px >= -2147483648.0f && px < 2147483648.0f) ? __float2int_rn(px) : (int)0x80000000;
    const int sy = (py >= -2147483648.0f && py < 2147483648.0f) ? __float2int_rn(py) : (int)0x80000000;
    const int fx = sx & 31, fy = sy & 31;
    const int ix = min(max(sx >> 5, -32768), 32767), iy = min(max(sy >> 5, -32768), 32767);    // saturate_cast<short>
    int w00 = (32 - fx) * (32 - fy) * 32, w01 = fx * (32 - fy) * 32, w10 = (32 - fx) * fy * 32, w11 = fx * fy * 32;
    if ((fx | fy) == 0) { w00 = 32767; w11 = 1; }          // saturate_cast<short>(32768) and its sum fix-up
    const bool x0 = ix >= 0 && ix < sw, x1 = ix + 1 >= 0 && ix + 1 < sw, y0 = iy >= 0 && iy < sh, y1 = iy + 1 >= 0 && iy + 1 < sh;
    const uint8_t *r0 = src + (long long)(y0 ? iy : 0) * spitch, *r1 = src + (long long)(y1 ? iy + 1 : 0) * spitch;
    const int c0 = (x0 ? ix : 0) * CN, c1 = (x1 ? ix + 1 : 0) * CN;
    uint8_t *o = dst + (long long)y * dpitch + x * CN;
#pragma unroll
    for (int c = 0; c < CN; c++) {
        const int a = (y0 && x0 ? r0[c0 + c] : 0) * w00 + (y0 && x1 ? r0[c1 + c] : 0) * w01 + (y1 && x0 ? r1[c0 + c] : 0) * w10 +
                      (y1 && x1 ? r1[c1 + c] : 0) * w11;
        o[c] = (uint8_t)min(max((a + 16384) >> 15, 0), 255);
    }
}

int sgbm_launch_remap_linear(const uint8_t *src, int sw, int sh, int cn, long long spitch, const float *map1, const float *map2, int W,
                             int H, uint8_t *dst, long long dpitch, cudaStream_t st)
{
    dim3 grid((W + 255) / 256, H);
    if (cn == 1) k_remap_linear<1><<<grid, 256, 0, st>>>(src, sw, sh, spitch, map1, map2, W, H, dst, dpitch);
    else if (cn == 3) k_remap_linear<3><<<grid, 256, 0, st>>>(src, sw, sh, spitch, map1, map2, W, H, dst, dpitch);
    else return sgbm_fail(-1, "remap: channels must be 1 or 3 (got %d)", cn);
    sgbm_count_launch(1);
    SGBM_CUDA_CHECK(cudaGetLastError());
    return 0;
}
