/*
 * sgbm_b200.h -- C ABI of the B200-native dense-stereo engine (libsgbm_b200.so).
 *
 * The reference (rafayaamirgull/stereo_reconstruction_cv) has no FFI of its own for this path:
 * it reaches the arithmetic through the cv2 Python binding.  Each entry point below names the
 * reference call it replaces:
 *
 *   sgbm_create / sgbm_set_params / sgbm_destroy  <- cv2.StereoSGBM_create(...)      main.ipynb:655-666
 *   sgbm_compute / sgbm_compute_host              <- stereo.compute(imgL, imgR)      main.ipynb:668
 *   sgbm_disp_to_float                            <- .astype(float32)/16, mask > 0   main.ipynb:668-670
 *   sgbm_reproject_f32 / sgbm_reproject_i16       <- cv2.reprojectImageTo3D(d, Q)    main.ipynb:697
 *   sgbm_reproject_compact                        <- finite/positive mask + gather   main.ipynb:726-737
 *   sgbm_filter_speckles / sgbm_median3x3         <- cv2.filterSpeckles / medianBlur (stages of compute)
 *
 * Conventions: plain pointers and sizes, no C++/torch types.  Every function returns 0 on
 * success or a negative SGBM_E_* code; sgbm_last_error() returns a thread-local message.  No
 * exception crosses the ABI.  Unless the name ends in _host, image/disparity pointers are DEVICE
 * pointers and all work is enqueued on `cuda_stream` (a cudaStream_t passed as void*) without a
 * host synchronisation; the caller owns all buffers, the handle owns its workspace.
 *
 * Threads and streams.  Calls on ONE handle are serialised by the library (a mutex around the enqueue)
 * and ordered on the device: a call waits, on its own stream, for the end of the handle's previous call,
 * whatever stream that ran on, because the workspace belongs to the handle.  Distinct handles are
 * independent and may be driven from different threads, streams and devices concurrently.
 *
 * Environment knobs (SGBM_*; tuning and A/B tests) are read once, by sgbm_create, and stay with the
 * handle; nothing reads the environment afterwards.  The release library has no result-changing hook.
 */
#ifndef SGBM_B200_H
#define SGBM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SGBM_MODE_SGBM       0
#define SGBM_MODE_HH         1
#define SGBM_MODE_SGBM_3WAY  2
#define SGBM_MODE_HH4        3

#define SGBM_OK                 0
#define SGBM_E_INVALID_ARG     -1   /* null pointer, bad channels/pitch, unsupported value          */
#define SGBM_E_BAD_SIZE        -2   /* cv2.error analogue: W - (minD+D) <= blockSize/2, W1 <= 0     */
#define SGBM_E_UNSUPPORTED     -3   /* parameter outside what this build implements                */
#define SGBM_E_CUDA            -4   /* CUDA runtime error (message has the cudaError string)       */
#define SGBM_E_NOMEM           -5

/* The 11 integers of cv2.StereoSGBM_create, same names and meaning (main.ipynb:655-666). */
typedef struct sgbm_params {
    int minDisparity;
    int numDisparities;   /* > 0, <= 1024; a multiple of 8, or any value >= 4 outside MODE_SGBM_3WAY (cv2 accepts those too) */
    int blockSize;
    int P1;
    int P2;
    int disp12MaxDiff;
    int preFilterCap;
    int uniquenessRatio;
    int speckleWindowSize;
    int speckleRange;
    int mode;
} sgbm_params;

typedef struct sgbm_handle sgbm_handle;

const char *sgbm_last_error(void);
const char *sgbm_version(void);

/* Number of SMs / device name of the current CUDA device (diagnostics for bench.py). */
int sgbm_device_info(int *sm_count, int *cc_major, int *cc_minor, char *name, int name_len);

/* A handle belongs to the CUDA device that is current when it is created (its workspace, streams and
 * events live there): make the same device current for every later call on it, or the call returns
 * SGBM_E_INVALID_ARG.  A process may hold handles on several devices. */
int sgbm_create(const sgbm_params *p, sgbm_handle **out);
int sgbm_destroy(sgbm_handle *h);
int sgbm_set_params(sgbm_handle *h, const sgbm_params *p);
int sgbm_get_params(const sgbm_handle *h, sgbm_params *p);

/* Bytes of device workspace the handle allocates for calls with `batch` W x H frames (volumes C, L_h,
 * S, ... of every frame it keeps in flight: one for batch = 1, up to three side by side for batches). */
int sgbm_workspace_bytes(const sgbm_handle *h, int W, int H, int channels, int batch, size_t *out);

/*
 * Disparity for `batch` independent rectified pairs.  To the caller everything is ordered on
 * `cuda_stream`; inside, a batch runs up to four frames side by side (each on its share of the SMs,
 * on internal streams forked from and joined to `cuda_stream`) when the frames are small enough for
 * that to pay, and two frames in flight otherwise.  left/right: uint8, `channels` in {1,3} interleaved, row pitch `pitch_bytes`,
 * frame stride = pitch_bytes * H.  disp_out: int16 (disparity x16, invalid = (minD-1)*16), row
 * pitch out_pitch_bytes, frame stride out_pitch_bytes * H.  Result is bit-identical to
 * cv2.StereoSGBM.compute inside the parity domain documented in DESIGN.md.
 */
int sgbm_compute(sgbm_handle *h, const uint8_t *left, const uint8_t *right, int W, int H,
                 int channels, ptrdiff_t pitch_bytes, int batch, int16_t *disp_out,
                 ptrdiff_t out_pitch_bytes, void *cuda_stream);

/*
 * Same with HOST pointers: H2D, compute, D2H, stream synchronise (what the numpy call site
 * main.ipynb:668 needs).  Page-locked buffers with dense rows are DMA'd directly, pageable ones are
 * staged through the handle's pinned slots; with batch > 1 three streams overlap the copies of
 * neighbouring frames with the kernels.
 */
int sgbm_compute_host(sgbm_handle *h, const uint8_t *left, const uint8_t *right, int W, int H,
                      int channels, ptrdiff_t pitch_bytes, int batch, int16_t *disp_out,
                      ptrdiff_t out_pitch_bytes);

/* f = disp/16.0f, then f *= (f > 0)   (main.ipynb:668-670).  Device pointers, dense rows. */
int sgbm_disp_to_float(const int16_t *disp_x16, int W, int H, float *out, void *cuda_stream);

/*
 * cv2.reprojectImageTo3D(disp, Q, handleMissingValues=False): xyz[y][x][0..2] float32.
 * Q: 16 doubles, row major, HOST pointer.  _i16 uses the integer value as is (no /16), like cv2.
 * valid_or_null (optional, W*H bytes): 1 where X is finite and disp > 0 (main.ipynb:726-730).
 */
int sgbm_reproject_f32(const float *disp, const double *Q, int W, int H, float *xyz,
                       uint8_t *valid_or_null, void *cuda_stream);
int sgbm_reproject_i16(const int16_t *disp, const double *Q, int W, int H, float *xyz,
                       uint8_t *valid_or_null, void *cuda_stream);

/*
 * The full cv2.reprojectImageTo3D(disparity, Q, handleMissingValues, ddepth) surface.  disp_depth and
 * ddepth are cv2 depth codes: 0 = CV_8U, 3 = CV_16S, 4 = CV_32S, 5 = CV_32F (ddepth -1 = CV_32F).
 * Integer disparities are converted to float32 and used as they are (no /16), like cv2.
 * handle_missing_values: Z = 10000 where |d - min(disparity)| <= FLT_EPSILON; needs scratch16 (16 bytes
 * of device memory).  Integer ddepth: cvRound (half to even; what does not fit an int32, NaN and +-inf
 * included, becomes INT_MIN), CV_16S saturates.  out: H x W x 3 of the ddepth type.
 */
int sgbm_reproject_ex(const void *disp, int disp_depth, const double *Q, int W, int H, int handle_missing_values,
                      int ddepth, void *out, void *scratch16, void *cuda_stream);

/*
 * Fused tail of the notebook (main.ipynb:668-670, 697, 726-737): int16 disparity x16 ->
 * /16, mask > 0, reproject, keep finite X and d > 0, gather XYZ (float32 N x 3) and RGB (uint8
 * N x 3, taken from the interleaved 3-channel image `bgr` with channel order swapped, or from a
 * 1-channel image replicated when bgr_channels == 1; may be NULL).  Points are written in
 * row-major pixel order (same order as numpy boolean indexing).  n_out: DEVICE pointer to one
 * unsigned long long receiving N.  xyz_out / rgb_out must hold W*H points.
 */
int sgbm_reproject_compact(const int16_t *disp_x16, const double *Q, int W, int H,
                           const uint8_t *bgr, int bgr_channels, ptrdiff_t bgr_pitch_bytes,
                           float *xyz_out, uint8_t *rgb_out, unsigned long long *n_out,
                           void *scratch, size_t scratch_bytes, void *cuda_stream);
int sgbm_reproject_compact_scratch_bytes(int W, int H, size_t *out);

/* cv2.filterSpeckles(img, newVal, maxSpeckleSize, maxDiff) in place on a dense int16 image.
 * scratch: W*H*8 bytes of device memory. */
int sgbm_filter_speckles(int16_t *img, int W, int H, int newVal, int maxSpeckleSize, int maxDiff,
                         void *scratch, size_t scratch_bytes, void *cuda_stream);
/* cv2.medianBlur(src, 3) for int16, replicate border; src != dst. */
int sgbm_median3x3(const int16_t *src, int16_t *dst, int W, int H, void *cuda_stream);

/*
 * Rectification warp in front of the path, device pointers:
 *   sgbm_init_rectify_map  <- cv2.initUndistortRectifyMap(K, None, R, P, size, CV_32F)   main.ipynb:496-497, gui.py:160-161
 *   sgbm_remap_linear_u8   <- cv2.remap(img, map1, map2, cv2.INTER_LINEAR)               main.ipynb:499-500, gui.py:163-164
 * K: 9 doubles (row major, HOST).  dist_or_null: n_dist distortion coefficients, must all be zero
 * (the reference passes None).  R_or_null: 9 doubles or NULL (identity).  P: 3 x p_cols doubles
 * (p_cols 3 or 4; only the left 3x3 block is used).  map1/map2: W*H float32 each (x and y
 * coordinates, the CV_32FC1 pair cv2 returns).  remap: 8-bit, 1 or 3 interleaved channels,
 * INTER_LINEAR in 1/32-pixel fixed point, BORDER_CONSTANT with value 0; bit-identical to cv2.remap.
 */
int sgbm_init_rectify_map(const double *K, const double *dist_or_null, int n_dist, const double *R_or_null,
                          const double *P, int p_cols, int W, int H, float *map1, float *map2, void *cuda_stream);
int sgbm_remap_linear_u8(const uint8_t *src, int src_w, int src_h, int channels, ptrdiff_t src_pitch_bytes,
                         const float *map1, const float *map2, int W, int H, uint8_t *dst, ptrdiff_t dst_pitch_bytes,
                         void *cuda_stream);

/*
 * Page-locked host memory for results.  A result array that lives in page-locked memory receives the
 * device -> host copy of sgbm_compute_host directly (no staging copy, no page faults of a freshly
 * allocated array): the Python binding hands out the arrays `stereo.compute(imgL, imgR)` returns
 * (main.ipynb:668) from a recycling pool of such blocks.  cudaHostAlloc(portable) / cudaFreeHost.
 */
int sgbm_host_alloc(size_t bytes, void **out);
int sgbm_host_free(void *p);

/*
 * Test hooks (used by tests/ only): copy an internal stage of the LAST frame computed by `h` to
 * a host buffer in canonical [y][x1][d] int16 order.  which: 0 = block cost C, 1 = aggregated S
 * (only after sgbm_debug_keep(h,1) was set before compute), 2 = raw disparity before median.
 */
int sgbm_debug_keep(sgbm_handle *h, int on);
int sgbm_debug_fetch(sgbm_handle *h, int which, void *host_dst, size_t bytes);
/* Host logic only, no device needed: the strip / ring schedule the sweep planner picks for frames of this size on a GPU
 * with `num_sms` SMs and `max_smem_bytes` of shared memory per CTA.  out16 = { found, strips, columns per strip, rows per
 * super-step, chain batches, rows per ring stage, S slots, cost stages, input stages, warps of role V, of each diagonal
 * role, of role W, W row groups, W warps per row, threads, shared-memory bytes }.  tests/test_host_logic.py. */
int sgbm_debug_sweep_plan(const sgbm_params *p, int W, int H, int channels, int num_sms, int max_smem_bytes, int w_role,
                          int input_volumes, int *out16);


/*
 * Deferred error of the asynchronous entry points.  The sweep kernels bound every hand-off wait; a wait
 * that makes no progress for ~2 s (a protocol error) makes the kernel drain instead of hanging the GPU.
 * sgbm_compute reports that for an earlier frame at its next call, sgbm_compute_host before it returns,
 * and sgbm_status on demand (after the caller synchronised its stream).  0 = ok.  No counterpart in cv2:
 * StereoSGBM.compute (main.ipynb:668) is synchronous.
 */
int sgbm_status(sgbm_handle *h);

/*
 * Measurement hooks for bench.py.  sgbm_kernel_launches: number of CUDA kernels this library has
 * launched in the calling process so far.  sgbm_profile_enable(h,1): record CUDA events around
 * every stage of compute() on the compute stream; sgbm_profile_read synchronises that stream and
 * returns, per stage, its name (32 bytes each), the accumulated device milliseconds, the number
 * of times the stage ran and the number of kernels it launched, then resets the accumulators.
 */
unsigned long long sgbm_kernel_launches(void);
int sgbm_profile_enable(sgbm_handle *h, int on);
int sgbm_profile_read(sgbm_handle *h, char *names32, double *total_ms, int *runs, int *kernels,
                      int max_stages, int *n_stages);

/* Integer-pipe microbenchmark (packed 16-bit DPX ops): measured lane-ops/s for the roofline. */
int sgbm_microbench_int16(int which, double *giga_lane_ops_per_s);

#ifdef __cplusplus
}
#endif
#endif /* SGBM_B200_H */
