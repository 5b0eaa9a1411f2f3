// sgbm_api.cu -- the C ABI of libsgbm_b200.so (see include/sgbm_b200.h): parameter handling
// (effective values), validation, workspace management and the per-frame kernel
// schedule that replaces cv2.StereoSGBM.compute / reprojectImageTo3D at main.ipynb:655-668, 697.
#include "../../include/sgbm_b200.h"
#include "sgbm_common.cuh"
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <mutex>
#include <new>
#include <vector>

// ---- launchers implemented in the other translation units -------------------------------------
int sgbm_launch_prefilter(const Geo &g, const uint8_t *left, const uint8_t *right, long long pitch, uint8_t *planes, cudaStream_t st);
int sgbm_launch_cost(const Geo &g, const uint8_t *planes, uint16_t *out, int y0, int nrows, int ylo, int zeroTail, cudaStream_t st);
int sgbm_launch_prefilter2(const Geo &g, const uint8_t *left, const uint8_t *right, long long pitch, uint8_t *planes, int eshift, int only3, cudaStream_t st);
int sgbm_cost3_supported(const Geo &g);
int sgbm_cost3_eshift(const Geo &g);
int sgbm_launch_cost2(const Geo &g, const uint8_t *planes, uint16_t *out, int y0, int nrows, int ylo, int zeroTail, cudaStream_t st);
size_t sgbm_cost2_planes_bytes(const Geo &g);
int sgbm_launch_cost3(const Geo &g, const uint8_t *planes, uint16_t *out, int y0, int nrows, int ylo, cudaStream_t st);
int sgbm_launch_horizontal(const Geo &g, const uint16_t *C, uint16_t *LhA, uint16_t *LhB, int y0, int nrows, cudaStream_t st);
int sgbm_launch_vertical(VertArgs &a, int ndir, int numSMs, cudaStream_t st);
bool sgbm_sweep_fits(const Geo &g, int numSMs, int mode);
bool sgbm_sweep_plan_debug(const Geo &g, int numSMs, int maxSmem, int wrole, int nAB, int *out);
int sgbm_launch_lrcheck(const Geo &g, int16_t *raw, const unsigned int *d2key, cudaStream_t st);
int sgbm_launch_pad_cost(const Geo &g, uint16_t *C, int nrows, int value, cudaStream_t st);
int sgbm_launch_init_wta(int16_t *raw, unsigned int *d2key, size_t n, int INV, cudaStream_t st);
int sgbm_launch_lr_median(const Geo &g, const int16_t *raw, const unsigned int *d2key, int16_t *dst, long long dstPitchElems, cudaStream_t st);
int sgbm_launch_median(const int16_t *src, int16_t *dst, int W, int H, long long dstPitchElems, cudaStream_t st);
int sgbm_launch_speckles(int16_t *img, int W, int H, int newVal, int maxSize, int maxDiff, void *scratch, cudaStream_t st);
int sgbm_launch_disp_to_float(const int16_t *d, float *out, size_t n, cudaStream_t st);
int sgbm_launch_reproject(const void *disp, int isFloat, const double *Qh, int W, int H, float *xyz, uint8_t *valid, cudaStream_t st);
size_t sgbm_compact_scratch_bytes(int W, int H);
int sgbm_launch_reproject_ex(const void *disp, int dispDepth, const double *Qh, int W, int H, int handleMissing, int ddepth,
                             void *out, void *scratch, cudaStream_t st);
int sgbm_launch_compact(const int16_t *disp, const double *Qh, int W, int H, const uint8_t *bgr, int bgrCn, long long bgrPitch,
                        float *xyz, uint8_t *rgb, unsigned long long *nOut, void *scratch, cudaStream_t st);
int sgbm_run_microbench(int which, double *out);
int sgbm_launch_rectify_map(const double *K, const double *R, const double *P, int pcols, int W, int H, float *map1, float *map2, cudaStream_t st);
int sgbm_launch_remap_linear(const uint8_t *src, int sw, int sh, int cn, long long spitch, const float *map1, const float *map2, int W,
                             int H, uint8_t *dst, long long dpitch, cudaStream_t st);

// ---- error reporting ----------------------------------------------------------------------------
static thread_local char g_err[512] = "";

int sgbm_fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
int sgbm_fail_cuda(cudaError_t e, const char *what, const char *file, int line)
{
    snprintf(g_err, sizeof(g_err), "CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
    return SGBM_E_CUDA;
}

// ---- handle -------------------------------------------------------------------------------------
static std::atomic<unsigned long long> g_launches{0};
void sgbm_count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

enum { ST_START = 0, ST_PREFILTER, ST_COST, ST_COST_ALT, ST_HORIZONTAL, ST_INIT, ST_VERT_FWD, ST_VERT_WTA, ST_LRCHECK,
       ST_MEDIAN, ST_SPECKLE, ST_COUNT };
static const char *const kStageNames[ST_COUNT] = {"start", "prefilter", "cost", "cost_alt", "horizontal", "init",
                                                  "vertical_fwd", "vertical_wta", "lrcheck", "median", "speckle"};
struct ProfMark { int stage; cudaEvent_t ev; unsigned long long launches; };
#define SGBM_MAX_LANES 4             // frames a batch call runs side by side (each on numSMs / lanes SMs)
#define SGBM_MAX_SLOTS (2 * SGBM_MAX_LANES)
#define SGBM_MAX_BANDS 8

struct sgbm_handle {
    sgbm_params p{};
    SgbmKnobs knobs{};                  // environment knobs + device facts, read once in sgbm_create
    std::mutex mu;                      // calls on one handle are serialised (and ordered by evLast on the device)
    cudaEvent_t evLast = nullptr;       // end of the handle's previous call: the next call's stream waits for it
    bool evLastValid = false;
    int numSMs = 0;
    int device = 0;                     // the device that was current at sgbm_create: workspace, streams and events live there
    // device workspace (grown on demand); lane 1 exists only while batches run two frames side by side
    void *ws[SGBM_MAX_LANES] = {};
    size_t wsBytes[SGBM_MAX_LANES] = {};
    int lanesWanted = 0;                // SGBM_LANES (1 switches the side-by-side batch schedule off); 0 = by geometry, see lanes_for
    cudaStream_t laneStream[SGBM_MAX_LANES] = {};   // [0] unused: lane 0 runs on the caller's / the handle's stream
    cudaEvent_t evFork = nullptr, evJoin[SGBM_MAX_LANES] = {};
    // row bands inside one frame: the horizontal kernel of band b runs beside the cost kernel of band b + 1
    cudaStream_t bandStream[SGBM_MAX_LANES][SGBM_MAX_BANDS] = {};
    cudaEvent_t evBand[SGBM_MAX_LANES][SGBM_MAX_BANDS] = {}, evBandJoin[SGBM_MAX_LANES][SGBM_MAX_BANDS] = {};
    int bandsWanted = -1;               // SGBM_BANDS; -1 = decide by frame size
    // pinned + device staging for the _host entry point: two slots so that the host copies and the
    // PCIe transfers of frame b+1 / b-1 overlap the kernels of frame b
    void *hostIn[SGBM_MAX_SLOTS] = {}, *hostOut[SGBM_MAX_SLOTS] = {}, *devIn[SGBM_MAX_SLOTS] = {}, *devOut[SGBM_MAX_SLOTS] = {};
    size_t hostInBytes[SGBM_MAX_SLOTS] = {}, hostOutBytes[SGBM_MAX_SLOTS] = {}, devInBytes[SGBM_MAX_SLOTS] = {}, devOutBytes[SGBM_MAX_SLOTS] = {};
    cudaStream_t ownStream = nullptr, inStream = nullptr, outStream = nullptr;
    cudaEvent_t evIn[SGBM_MAX_SLOTS] = {}, evComp[SGBM_MAX_SLOTS] = {}, evOut[SGBM_MAX_SLOTS] = {};
    unsigned int *watch = nullptr;   // pinned copy of the sweep watchdog words, [lanes][8]
    unsigned int *watchDev[SGBM_MAX_LANES] = {};
    // debug
    int keep = 0;
    Geo lastGeo{};
    uint16_t *lastC = nullptr, *lastS = nullptr;
    int16_t *lastRaw = nullptr;
    int lastValid = 0;
    // profiling
    int prof = 0;
    cudaStream_t profStream = nullptr;
    std::vector<cudaEvent_t> evPool;
    size_t evUsed = 0;
    std::vector<ProfMark> marks;
};

static int prof_mark(sgbm_handle *h, int stage, cudaStream_t st)
{
    if (!h->prof) return 0;
    if (h->evUsed == h->evPool.size()) {
        cudaEvent_t e;
        SGBM_CUDA_CHECK(cudaEventCreate(&e));
        h->evPool.push_back(e);
    }
    cudaEvent_t e = h->evPool[h->evUsed++];
    SGBM_CUDA_CHECK(cudaEventRecord(e, st));
    h->marks.push_back({stage, e, g_launches.load()});
    h->profStream = st;
    return 0;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---- knobs of the running call ---------------------------------------------------------------------
static thread_local const SgbmKnobs *g_knobs = nullptr;
static const SgbmKnobs g_defaultKnobs{};
const SgbmKnobs &sgbm_knobs() { return g_knobs ? *g_knobs : g_defaultKnobs; }
std::mutex &sgbm_setup_mutex() { static std::mutex m; return m; }
struct KnobScope {
    const SgbmKnobs *prev;
    explicit KnobScope(const sgbm_handle *h) : prev(g_knobs) { g_knobs = &h->knobs; }
    ~KnobScope() { g_knobs = prev; }
};

static int env_int(const char *name, int dflt)
{
    const char *e = getenv(name);
    return e && *e ? atoi(e) : dflt;
}

// The one place the environment is read: sgbm_create.
static void read_knobs(SgbmKnobs &k)
{
    k.nreg = env_int("SGBM_NREG", 0);
    k.vr = env_int("SGBM_VR", 0);
    k.sweepK = env_int("SGBM_SWEEP_K", 0); k.sweepNSC = env_int("SGBM_SWEEP_NSC", 0);
    k.sweepNSI = env_int("SGBM_SWEEP_NSI", 0); k.sweepNWW = env_int("SGBM_SWEEP_NWW", 0);
    k.sweepWRG = env_int("SGBM_SWEEP_WRG", 0);
    k.sweepPF = env_int("SGBM_SWEEP_PF", -1);
    k.sweepW = env_int("SGBM_SWEEP_W", 1) != 0;
    k.sweep = env_int("SGBM_SWEEP", 1) != 0;
    k.rowstep = env_int("SGBM_ROWSTEP", 0) != 0;
    k.cost2 = env_int("SGBM_COST2", 1) != 0;
    k.cost3 = env_int("SGBM_COST3", 1) != 0;
    k.cost3NXG = env_int("SGBM_COST3_NXG", 0); k.cost3RB = env_int("SGBM_COST3_RB", 0);
    k.cost3Pad = env_int("SGBM_COST3_PAD", 0);
    k.nstg = env_int("SGBM_NSTG", 0);
    k.sweepSat = env_int("SGBM_SWEEP_SAT", 0) != 0;
    k.sweepRPS = env_int("SGBM_SWEEP_RPS", 0);
    k.hhSplit = env_int("SGBM_HH_SPLIT", 1) != 0;
    k.verbose = env_int("SGBM_VERBOSE", 0) != 0;
#ifdef SGBM_DEBUG_HOOKS
    k.dbgNoSync = getenv("SGBM_DBG_NOSYNC") ? 1 : 0;
    k.dbgStall = getenv("SGBM_DBG_STALL") ? 1 : 0;
#endif
#ifdef SGBM_SWEEP_TRACING
    if (const char *e = getenv("SGBM_SWEEP_TRACE")) { strncpy(k.tracePath, e, sizeof(k.tracePath) - 1); }
#endif
}

static int check_device(const sgbm_handle *h)
{
    int dev = -1;
    SGBM_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev != h->device)
        return sgbm_fail(SGBM_E_INVALID_ARG, "handle was created on CUDA device %d but device %d is current", h->device, dev);
    return 0;
}

// Geometry handed to the prefilter / cost kernels.  They work on disparity PAIRS; for an odd numDisparities they compute
// one disparity more (d = D reads the right image at x - maxD >= 0 for every valid x, so the operand is in range) and
// k_pad_cost overwrites it together with the other padding disparities.  Everything else about the geometry -- valid
// range, lane mapping, strides -- is the true one.
static Geo cost_geo(const Geo &g)
{
    Geo c = g;
    if (g.D & 1) { c.D = g.D + 1; c.maxD = g.maxD + 1; }
    return c;
}

// Cost of a padding disparity (numDisparities % 8 != 0), or -1 when no value works.  It has to act as +infinity in the path
// step (A.4) without leaving 16 bits:  L_pad = C_pad + min(..) - m  lies in [C_pad, C_pad + P2], so C_pad + P2 <= 65535; it
// must never be the minimum over d nor the better neighbour of d = D - 1:  C_pad + P1 >= m + P2 with m <= cMax + P2.
// Hence C_pad = 65535 - P2, possible when cMax + 3 P2 - P1 <= 65535.  What the sums S hold at those disparities does not
// matter (they may wrap): every winner-take-all masks them (mask_pad_regs).
static int pad_cost_value(const Geo &g)
{
    const long long pixMax = (long long)(2 * g.ftzero < 255 ? 2 * g.ftzero : 255) + 63;
    const long long cMax = (long long)g.cn * (2 * g.r + 1) * (2 * g.r + 1) * pixMax;
    const long long v = 65535 - g.P2;
    return v >= cMax + 2ll * g.P2 - g.P1 ? (int)v : -1;
}

// Effective parameters and geometry (A.0) + lane mapping.  Returns 0 or an error code.
static int make_geo(const sgbm_params &p, int W, int H, int cn, Geo &g)
{
    const SgbmKnobs &kn = sgbm_knobs();
    memset(&g, 0, sizeof(g));
    if (W <= 0 || H <= 0) return sgbm_fail(SGBM_E_INVALID_ARG, "empty image %dx%d", W, H);
    if (cn != 1 && cn != 3) return sgbm_fail(SGBM_E_INVALID_ARG, "channels must be 1 or 3 (got %d)", cn);
    if (p.mode < 0 || p.mode > 3) return sgbm_fail(SGBM_E_INVALID_ARG, "unknown mode %d", p.mode);
    g.W = W; g.H = H; g.cn = cn; g.mode = p.mode;
    g.minD = p.minDisparity; g.D = p.numDisparities; g.maxD = g.minD + g.D;
    if (g.D <= 0) return sgbm_fail(SGBM_E_BAD_SIZE, "numDisparities must be > 0 (got %d)", g.D);
    if (p.mode == SGBM_MODE_SGBM_3WAY) g.r = (p.blockSize > 0 ? p.blockSize : 3) / 2;
    else g.r = (p.blockSize > 0 ? p.blockSize : 5) / 2;
    g.P1 = p.P1 > 0 ? p.P1 : 2;
    g.P2 = p.P2 > 0 ? p.P2 : 5;
    if (g.P2 < g.P1 + 1) g.P2 = g.P1 + 1;
    g.UR = p.uniquenessRatio >= 0 ? p.uniquenessRatio : 10;
    g.DMD = p.disp12MaxDiff > 0 ? p.disp12MaxDiff : 1;
    g.ftzero = (p.preFilterCap > 15 ? p.preFilterCap : 15) | 1;
    g.INV = (g.minD - 1) * 16;
    g.minX1 = g.maxD > 0 ? g.maxD : 0;
    g.maxX1 = W + (g.minD < 0 ? g.minD : 0);
    g.W1 = g.maxX1 - g.minX1;
    // cv2.error site stereosgbm.cpp:511 [P15] and the empty-valid-range case cv2 crashes on [P18]
    if (!(W - (g.minD + g.D) > p.blockSize / 2) || g.W1 <= 0)
        return sgbm_fail(SGBM_E_BAD_SIZE, "image width %d too small for minDisparity=%d numDisparities=%d blockSize=%d",
                         W, g.minD, g.D, p.blockSize);
    if (g.D > 1024) return sgbm_fail(SGBM_E_UNSUPPORTED, "numDisparities must be <= 1024 (got %d)", g.D);
    // numDisparities that are not a multiple of 8 (cv2 documents % 16 but accepts anything; SURVEY 8(c), [P16]): computed
    // on volumes as wide as the next multiple of 8 whose padding disparities carry a large constant cost (pad_cost_value).
    // Not for MODE_SGBM_3WAY (cv2's own SIMD tail makes its results irregular there, A.6) and not below 4 (the masked
    // uniqueness scan needs a disparity outside the winner's window).  Odd values: see cost_geo.
    const int Dc = (g.D + 7) & ~7;
    if (Dc != g.D) {
        if (p.mode == SGBM_MODE_SGBM_3WAY)
            return sgbm_fail(SGBM_E_UNSUPPORTED, "MODE_SGBM_3WAY needs numDisparities %% 8 == 0 (got %d)", g.D);
        if (g.D < 4)
            return sgbm_fail(SGBM_E_UNSUPPORTED, "numDisparities must be >= 4 (got %d)", g.D);
    }
    if (g.P2 > 32767) return sgbm_fail(SGBM_E_UNSUPPORTED, "P2 must be <= 32767 (got %d)", g.P2);
    if (g.UR > 100 || (p.mode == SGBM_MODE_SGBM_3WAY && g.UR >= 100))
        return sgbm_fail(SGBM_E_UNSUPPORTED, "uniquenessRatio %d is not supported", g.UR);
    if (g.INV < -32768 || g.maxD * 16 > 32767 || g.minD * 16 < -32768)
        return sgbm_fail(SGBM_E_UNSUPPORTED, "disparity range does not fit the int16 x16 output");
    if (g.W1 > 65535) return sgbm_fail(SGBM_E_UNSUPPORTED, "valid width %d > 65535", g.W1);
    // lane mapping: D = 2*nreg*lanesUsed, lpc = pow2 >= lanesUsed (>= 2); maximise lanesUsed/lpc
    static const int prefBig[4] = {16, 12, 8, 4}, prefSmall[4] = {8, 12, 16, 4};
    const int *pref = Dc >= 192 ? prefBig : prefSmall;
    const int forced = kn.nreg;
    double bestEff = -1;
    for (int i = 0; i < 4; i++) {
        int nreg = pref[i];
        if (forced && nreg != forced) continue;
        if (Dc % (2 * nreg)) continue;
        int lanes = Dc / (2 * nreg);
        if (lanes > 32) continue;
        int lpc = 2;
        while (lpc < lanes) lpc <<= 1;
        double eff = (double)lanes / lpc;
        if (eff > bestEff + 1e-9) { bestEff = eff; g.nreg = nreg; g.lpc = lpc; g.lanesUsed = lanes; }
    }
    if (bestEff < 0) return sgbm_fail(SGBM_E_UNSUPPORTED, "no lane mapping for numDisparities=%d", g.D);
    g.Dp = 2 * g.nreg * g.lpc;
    g.lpcShift = 0;
    while ((1 << g.lpcShift) < g.lpc) g.lpcShift++;
    g.rowStride = (long long)g.W1 * g.Dp;
    if (Dc != g.D && pad_cost_value(g) < 0)
        return sgbm_fail(SGBM_E_UNSUPPORTED, "numDisparities %d (not a multiple of 8) needs cMax + 3*P2 - P1 <= 65535 (blockSize %d, P2 %d)",
                         g.D, 2 * g.r + 1, g.P2);
    return 0;
}

struct WsLayout {
    size_t planes, C, LhA, LhB, Calt, raw, d2key, med, speck, haloA, haloC, flags, watch, rowState, sdbg, total;
};

static void ws_layout(const Geo &g, const sgbm_params &p, int numSMs, int keep, WsLayout &L)
{
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    size_t vol = (size_t)g.rowStride * g.H * 2;
    L.watch = take(64);       // hand-off watchdog of the sweep (sgbm_sweep.cu); first, so its place never moves
    {   // prefilter output: the larger of the two generations' formats (sgbm_cost.cu / sgbm_cost2.cu)
        const size_t v1 = (size_t)2 * g.cn * 6 * g.W * g.H, v2 = sgbm_cost2_planes_bytes(cost_geo(g));
        L.planes = take(v1 > v2 ? v1 : v2);
    }
    L.C = take(vol);
    L.LhA = take(vol);
    L.LhB = take(vol);
    L.Calt = take(p.mode == SGBM_MODE_SGBM_3WAY ? (size_t)3 * (g.r > 0 ? g.r : 1) * g.rowStride * 2 : 16);
    L.raw = take((size_t)g.W * g.H * 2);
    L.d2key = take((size_t)g.W * g.H * 4);
    L.med = take((size_t)g.W * g.H * 2);
    L.speck = take((size_t)g.W * g.H * 8);
    int maxStrips = g.W1 < numSMs ? g.W1 : numSMs;
    const int maxR = 16;
    L.haloA = take((size_t)maxStrips * 4 * maxR * (g.Dp + 8) * 2);    // 4 super-step slots (sgbm_sweep.cu)
    L.haloC = take((size_t)maxStrips * 4 * maxR * (g.Dp + 8) * 2);
    L.flags = take((size_t)2 * maxStrips * 64 * 4);                   // one flag per halo ring entry
    L.rowState = take((size_t)2 * 3 * g.W1 * (g.Dp + 8) * 2);         // row-at-a-time fallback (sgbm_sweep.cu)
    L.sdbg = take(keep ? vol : 16);
    L.total = off;
}

extern "C" const char *sgbm_last_error(void) { return g_err; }
extern "C" const char *sgbm_version(void) { return "sgbm_b200 0.1 (sm_100a)"; }

extern "C" int sgbm_device_info(int *sm_count, int *cc_major, int *cc_minor, char *name, int name_len)
{
    int dev = 0;
    SGBM_CUDA_CHECK(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    SGBM_CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    if (name && name_len > 0) { strncpy(name, prop.name, name_len - 1); name[name_len - 1] = 0; }
    return 0;
}

extern "C" int sgbm_create(const sgbm_params *p, sgbm_handle **out)
{
    if (!p || !out) return sgbm_fail(SGBM_E_INVALID_ARG, "null argument");
    if (p->mode < 0 || p->mode > 3) return sgbm_fail(SGBM_E_INVALID_ARG, "unknown mode %d", p->mode);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return sgbm_fail(SGBM_E_CUDA, "no CUDA device available (%s): this engine has no CPU fallback", cudaGetErrorString(e));
    sgbm_handle *h = new (std::nothrow) sgbm_handle();
    if (!h) return sgbm_fail(SGBM_E_NOMEM, "out of host memory");
    h->p = *p;
    int dev = 0;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess ||
        (e = cudaDeviceGetAttribute(&h->numSMs, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) {
        delete h;
        return sgbm_fail_cuda(e, "querying the device", __FILE__, __LINE__);
    }
    h->device = dev;
    read_knobs(h->knobs);
    h->knobs.device = dev;
    h->knobs.numSMs = h->numSMs;
    if ((e = cudaDeviceGetAttribute(&h->knobs.maxSmemOptin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&h->evLast, cudaEventDisableTiming)) != cudaSuccess) {
        delete h;
        return sgbm_fail_cuda(e, "setting up the handle", __FILE__, __LINE__);
    }
    { const int v = env_int("SGBM_SM_LIMIT", 0); if (v >= 1 && v < h->numSMs) h->numSMs = v; }
    if (getenv("SGBM_LANES")) { const int v = env_int("SGBM_LANES", 0); h->lanesWanted = v < 1 ? 1 : (v > SGBM_MAX_LANES ? SGBM_MAX_LANES : v); }
    if (getenv("SGBM_BANDS")) { const int v = env_int("SGBM_BANDS", 0); h->bandsWanted = v < 1 ? 1 : (v > SGBM_MAX_BANDS ? SGBM_MAX_BANDS : v); }
    *out = h;
    return 0;
}

extern "C" int sgbm_destroy(sgbm_handle *h)
{
    if (!h) return 0;
    for (int i = 0; i < SGBM_MAX_LANES; i++) {
        if (h->ws[i]) cudaFree(h->ws[i]);
        if (h->laneStream[i]) cudaStreamDestroy(h->laneStream[i]);
        if (h->evJoin[i]) cudaEventDestroy(h->evJoin[i]);
        for (int b = 0; b < SGBM_MAX_BANDS; b++) {
            if (h->bandStream[i][b]) cudaStreamDestroy(h->bandStream[i][b]);
            if (h->evBand[i][b]) cudaEventDestroy(h->evBand[i][b]);
            if (h->evBandJoin[i][b]) cudaEventDestroy(h->evBandJoin[i][b]);
        }
    }
    if (h->watch) cudaFreeHost(h->watch);
    if (h->evFork) cudaEventDestroy(h->evFork);
    if (h->evLast) cudaEventDestroy(h->evLast);
    for (int i = 0; i < SGBM_MAX_SLOTS; i++) {
        if (h->devIn[i]) cudaFree(h->devIn[i]);
        if (h->devOut[i]) cudaFree(h->devOut[i]);
        if (h->hostIn[i]) cudaFreeHost(h->hostIn[i]);
        if (h->hostOut[i]) cudaFreeHost(h->hostOut[i]);
        if (h->evIn[i]) cudaEventDestroy(h->evIn[i]);
        if (h->evComp[i]) cudaEventDestroy(h->evComp[i]);
        if (h->evOut[i]) cudaEventDestroy(h->evOut[i]);
    }
    if (h->ownStream) cudaStreamDestroy(h->ownStream);
    if (h->inStream) cudaStreamDestroy(h->inStream);
    if (h->outStream) cudaStreamDestroy(h->outStream);
    for (cudaEvent_t e : h->evPool) cudaEventDestroy(e);
    delete h;
    return 0;
}

extern "C" int sgbm_set_params(sgbm_handle *h, const sgbm_params *p)
{
    if (!h || !p) return sgbm_fail(SGBM_E_INVALID_ARG, "null argument");
    if (p->mode < 0 || p->mode > 3) return sgbm_fail(SGBM_E_INVALID_ARG, "unknown mode %d", p->mode);
    h->p = *p;
    h->lastValid = 0;
    return 0;
}
extern "C" int sgbm_get_params(const sgbm_handle *h, sgbm_params *p)
{
    if (!h || !p) return sgbm_fail(SGBM_E_INVALID_ARG, "null argument");
    *p = h->p;
    return 0;
}

static int lanes_for(const sgbm_handle *h, const Geo &g, int batch, int *sweepSMs);

// Device workspace the handle allocates for frames of this size: one workspace per frame it keeps in
// flight (batch > 1: up to SGBM_MAX_LANES, see lanes_for).  Needs the handle's device to be current.
extern "C" int sgbm_workspace_bytes(const sgbm_handle *h, int W, int H, int channels, int batch, size_t *out)
{
    if (!h || !out) return sgbm_fail(SGBM_E_INVALID_ARG, "null argument");
    if (batch <= 0) return sgbm_fail(SGBM_E_INVALID_ARG, "batch must be >= 1");
    KnobScope ks(h);
    Geo g;
    int rc = make_geo(h->p, W, H, channels, g);
    if (rc) return rc;
    WsLayout L;
    ws_layout(g, h->p, h->numSMs, h->keep, L);
    int sweepSMs = 0;
    *out = L.total * (size_t)lanes_for(h, g, batch, &sweepSMs);
    return 0;
}

static int ensure_ws(sgbm_handle *h, int lane, size_t bytes, cudaStream_t st)
{
    if (h->wsBytes[lane] >= bytes) return 0;
    if (h->ws[lane]) {
        SGBM_CUDA_CHECK(cudaStreamSynchronize(st));
        SGBM_CUDA_CHECK(cudaFree(h->ws[lane]));
        h->ws[lane] = nullptr; h->wsBytes[lane] = 0;
    }
    cudaError_t e = cudaMalloc(&h->ws[lane], bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return sgbm_fail(SGBM_E_NOMEM, "cudaMalloc of %zu workspace bytes failed: %s", bytes, cudaGetErrorString(e));
    }
    h->wsBytes[lane] = bytes;
    // zero it once: the padding words of the volumes are never written, and the sweep watchdog is sticky
    SGBM_CUDA_CHECK(cudaMemsetAsync(h->ws[lane], 0, bytes, st));
    return 0;
}

// Reports (once) a sweep hand-off that timed out in a frame that has completed; see sgbm_sweep.cu.
static int check_watch(sgbm_handle *h, cudaStream_t st)
{
    if (!h->watch) return 0;
    for (int lane = 0; lane < SGBM_MAX_LANES; lane++) {
        unsigned int *w = h->watch + 8 * lane;
        if (w[0] == 0u) continue;
        const unsigned strip = w[1], warp = w[2], id = w[3], row = w[4];
        w[0] = 0u;
        if (h->watchDev[lane]) cudaMemsetAsync(h->watchDev[lane], 0, 32, st);
        return sgbm_fail(SGBM_E_CUDA, "sweep hand-off timed out (strip %u, warp %u, wait %u, row %u): the frame's disparity is invalid",
                         strip, warp, id, row);
    }
    return 0;
}

// Frames side by side.  Small frames, each on numSMs / lanes SMs: the sweeps are bound by per-row hand-off latency, not by
// throughput, when their strips are narrow (1440p and below), so two or three narrower launches finish
// their frames in little more than the time of one (720p D=128: 84 -> 133 -> 179 GDE/s with 1 / 2 / 3
// lanes; four lanes are slower again).  As many lanes as the persistent sweep still holds the geometry for.
static int lanes_for(const sgbm_handle *h, const Geo &g, int batch, int *sweepSMs)
{
    *sweepSMs = h->numSMs;
    // three lanes by default (four were slower at 720p D=128: 172 vs 180 GDE/s); four when a column is only two lanes
    // (numDisparities <= 32): those frames leave the GPU emptiest
    const int want = h->lanesWanted > 0 ? h->lanesWanted : (g.lpc <= 2 ? 4 : 3);
    if (batch < 2 || want < 2 || h->prof || h->keep) return 1;
    if (h->p.mode == SGBM_MODE_SGBM || h->p.mode == SGBM_MODE_HH) {
        int lanes = want < batch ? want : batch;
        for (; lanes >= 2; lanes--)
            if (h->numSMs / lanes >= 1 && sgbm_sweep_fits(g, h->numSMs / lanes, h->p.mode)) {
                *sweepSMs = h->numSMs / lanes;
                return lanes;
            }
    }
    // Large frames (or the modes without persistent sweeps): two frames in flight on two streams, every
    // kernel on the whole GPU.  The sweeps of the two frames take turns, the other kernels and all the
    // launch tails overlap (4K D=256: MODE_HH 180 -> 188 GDE/s, MODE_SGBM_3WAY 285 -> 317 GDE/s).
    return 2;
}

static int ensure_lane_streams(sgbm_handle *h, int lanes)
{
    if (!h->evFork) SGBM_CUDA_CHECK(cudaEventCreateWithFlags(&h->evFork, cudaEventDisableTiming));
    for (int i = 1; i < lanes; i++) {
        if (h->laneStream[i]) continue;
        SGBM_CUDA_CHECK(cudaStreamCreateWithFlags(&h->laneStream[i], cudaStreamNonBlocking));
        SGBM_CUDA_CHECK(cudaEventCreateWithFlags(&h->evJoin[i], cudaEventDisableTiming));
    }
    return 0;
}

static int ensure_band_streams(sgbm_handle *h, int lane, int bands)
{
    int lo = 0, hi = 0;
    SGBM_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    for (int b = 0; b < bands; b++) {
        if (h->bandStream[lane][b]) continue;
        // high priority: a band's few horizontal CTAs are placed as soon as they are ready, the cost kernel
        // of the following band fills the rest of the GPU
        SGBM_CUDA_CHECK(cudaStreamCreateWithPriority(&h->bandStream[lane][b], cudaStreamNonBlocking, hi));
        SGBM_CUDA_CHECK(cudaEventCreateWithFlags(&h->evBand[lane][b], cudaEventDisableTiming));
        SGBM_CUDA_CHECK(cudaEventCreateWithFlags(&h->evBandJoin[lane][b], cudaEventDisableTiming));
    }
    return 0;
}

// One frame: the kernel schedule for each mode.
static int compute_frame(sgbm_handle *h, int lane, int sweepSMs, const Geo &g, const WsLayout &L, const uint8_t *left,
                         const uint8_t *right, long long pitch, int16_t *out, long long outPitchElems, cudaStream_t st)
{
    const sgbm_params &p = h->p;
    uint8_t *base = (uint8_t *)h->ws[lane];
    uint8_t *planes = base + L.planes;
    uint16_t *C = (uint16_t *)(base + L.C), *LhA = (uint16_t *)(base + L.LhA), *LhB = (uint16_t *)(base + L.LhB);
    uint16_t *Calt = (uint16_t *)(base + L.Calt);
    int16_t *raw = (int16_t *)(base + L.raw), *med = (int16_t *)(base + L.med);
    unsigned int *d2key = (unsigned int *)(base + L.d2key);
    int rc;
    if (!h->watch) {
        SGBM_CUDA_CHECK(cudaMallocHost((void **)&h->watch, 32 * SGBM_MAX_LANES));
        memset(h->watch, 0, 32 * SGBM_MAX_LANES);
    }
    h->watchDev[lane] = (unsigned int *)(base + L.watch);
    if ((rc = check_watch(h, st))) return rc;               // a previous frame's sweep gave up
    if ((rc = prof_mark(h, ST_START, st))) return rc;
    // second-generation prefilter + cost kernels (sgbm_cost2.cu); the first generation stays as the
    // fallback for geometries the new kernel does not hold (rc == 1) and for A/B runs (SGBM_COST2=0)
    bool cost2 = h->knobs.cost2 != 0, cost3 = h->knobs.cost3 != 0;
    int bands = 1, bandRows = g.H;
    const int padCost = (g.D & 7) ? pad_cost_value(g) : -1;      // numDisparities % 8 != 0: see pad_cost_value
    const Geo gc = cost_geo(g);
    if (!cost2) cost3 = false;
    if (cost3) {
        rc = sgbm_cost3_supported(gc);
        if (rc < 0) return rc;
        cost3 = rc == 1;
    }
    if (cost2) {
        if ((rc = sgbm_launch_prefilter2(gc, left, right, pitch, planes, cost3 ? sgbm_cost3_eshift(gc) : 0, cost3 ? 1 : 0, st))) return rc;
        if ((rc = prof_mark(h, ST_PREFILTER, st))) return rc;
        // third generation (sgbm_cost3.cu): register-resident pixel costs; 1-channel, blockSize <= 11
        // Row bands (large frames, whole-GPU schedule): a row of the horizontal paths needs only its own cost
        // row, so band b's horizontal kernel runs on a high-priority side stream beside band b + 1's cost kernel.  Both
        // lean on the shared-memory data pipe and a 211 KB cost CTA leaves no room for a horizontal CTA on its SM, so what
        // overlaps are the tails of the waves.  Six bands (end of round 2, 4K D=256): MODE_HH 10.23 -> 9.90 ms, 3WAY
        // 6.97 -> 6.76 ms; 2 / 3 / 4 / 5 / 8 bands: 10.23 / 9.95 / 9.92 / 9.92 / 10.04 ms; at 1080p bands only cost
        // (1.86 -> 2.06 ms with two), so smaller frames keep one.
        if (cost3) {
            const long long elems = (long long)g.W1 * g.H * g.Dp;
            // (2560x1440 D=256 MODE_HH: 5.58 / 5.33 / 5.31 / 5.35 / 5.46 ms with 1 / 2 / 3 / 4 / 6 bands; 4K D=128 and D=192 do not care)
            const int autoBands = sweepSMs != h->numSMs ? 1 : (elems >= (1ll << 30) ? 6 : (g.nreg >= 16 && elems >= (3ll << 28) ? 3 : 1));
            bands = h->bandsWanted > 0 ? h->bandsWanted : autoBands;
            if (p.mode == SGBM_MODE_HH4) bands = 1;                    // (its zeroed last rows are written after the cost kernel)
            bandRows = ((g.H + bands - 1) / bands + 15) / 16 * 16;
            bands = (g.H + bandRows - 1) / bandRows;
            if (bands > 1 && (rc = ensure_band_streams(h, lane, bands))) return rc;
        }
        for (int b = 0; cost3 && b < bands; b++) {
            const int y0 = b * bandRows, nr = g.H - y0 < bandRows ? g.H - y0 : bandRows;
            if ((rc = sgbm_launch_cost3(gc, planes, C + (size_t)y0 * g.rowStride, y0, nr, 0, st)))
                return rc < 0 ? rc : sgbm_fail(SGBM_E_UNSUPPORTED, "cost kernel geometry changed between plan and launch");
            if (padCost >= 0 && (rc = sgbm_launch_pad_cost(g, C + (size_t)y0 * g.rowStride, nr, padCost, st))) return rc;
            if (bands > 1) {
                cudaStream_t bs = h->bandStream[lane][b];
                SGBM_CUDA_CHECK(cudaEventRecord(h->evBand[lane][b], st));
                SGBM_CUDA_CHECK(cudaStreamWaitEvent(bs, h->evBand[lane][b], 0));
                if ((rc = sgbm_launch_horizontal(g, C, LhA, LhB, y0, nr, bs))) return rc;
                SGBM_CUDA_CHECK(cudaEventRecord(h->evBandJoin[lane][b], bs));
            }
        }
        if (cost3 && p.mode == SGBM_MODE_HH4 && g.r > 0) {            // A.9: the last r rows carry C = 0
            const int nz = g.r < g.H ? g.r : g.H;
            SGBM_CUDA_CHECK(cudaMemsetAsync(C + (size_t)(g.H - nz) * g.rowStride, 0, (size_t)nz * g.rowStride * 2, st));
            if (padCost >= 0 && (rc = sgbm_launch_pad_cost(g, C + (size_t)(g.H - nz) * g.rowStride, nz, padCost, st))) return rc;
        }
        if (!cost3) {
            rc = sgbm_launch_cost2(gc, planes, C, 0, g.H, 0, p.mode == SGBM_MODE_HH4, st);
            if (rc < 0) return rc;
            if (rc == 1) cost2 = false;
        }
    }
    if (!cost2) {
        if ((rc = sgbm_launch_prefilter(gc, left, right, pitch, planes, st))) return rc;
        if ((rc = prof_mark(h, ST_PREFILTER, st))) return rc;
        if ((rc = sgbm_launch_cost(gc, planes, C, 0, g.H, 0, p.mode == SGBM_MODE_HH4, st))) return rc;
    }
    if (!cost3 && padCost >= 0 && (rc = sgbm_launch_pad_cost(g, C, g.H, padCost, st))) return rc;
    if ((rc = prof_mark(h, ST_COST, st))) return rc;
    const int ss = (g.H + 3) / 4;
    int ov = 0;
    if (p.mode == SGBM_MODE_SGBM_3WAY) {
        ov = (p.blockSize / 2 + 1) + (int)ceil(0.1 * (double)ss);    // same double expression as the reference (A.6)
        for (int n = 1; n < 4 && g.r > 0; n++) {
            int o0 = n * ss;
            if (o0 >= g.H) break;
            int s0 = o0 - ov > 0 ? o0 - ov : 0;
            if (s0 == 0) continue;
            int nr = g.r < g.H - s0 ? g.r : g.H - s0;
            uint16_t *dst = Calt + (size_t)(n - 1) * g.r * g.rowStride;
            rc = cost3 ? sgbm_launch_cost3(gc, planes, dst, s0, nr, s0, st)
                 : cost2 ? sgbm_launch_cost2(gc, planes, dst, s0, nr, s0, 0, st) : sgbm_launch_cost(gc, planes, dst, s0, nr, s0, 0, st);
            if (rc) return rc < 0 ? rc : sgbm_fail(SGBM_E_UNSUPPORTED, "cost kernel geometry changed between launches");
        }
    }
    if (p.mode == SGBM_MODE_SGBM_3WAY && (rc = prof_mark(h, ST_COST_ALT, st))) return rc;
    if (bands > 1) {
        for (int b = 0; b < bands; b++) SGBM_CUDA_CHECK(cudaStreamWaitEvent(st, h->evBandJoin[lane][b], 0));
    } else if ((rc = sgbm_launch_horizontal(g, C, LhA, LhB, 0, g.H, st))) return rc;
    if ((rc = prof_mark(h, ST_HORIZONTAL, st))) return rc;
    if ((rc = sgbm_launch_init_wta(raw, d2key, (size_t)g.W * g.H, g.INV, st))) return rc;
    if ((rc = prof_mark(h, ST_INIT, st))) return rc;

    VertArgs a;
    memset(&a, 0, sizeof(a));
    a.g = g; a.C = C; a.Calt = Calt; a.raw = raw; a.d2key = d2key;
    a.haloA = (uint16_t *)(base + L.haloA); a.haloC = (uint16_t *)(base + L.haloC);
    a.flagA = (unsigned int *)(base + L.flags); a.flagC = a.flagA + (g.W1 < h->numSMs ? g.W1 : h->numSMs);
    a.ss = ss; a.ov = ov;
    a.watchDev = h->watchDev[lane]; a.watchHost = h->watch + 8 * lane;
    a.rowState = (uint16_t *)(base + L.rowState);
    a.dbgNoSync = SGBM_DBG_HOOK(h->knobs.dbgNoSync);
    a.sdbg = h->keep ? (uint16_t *)(base + L.sdbg) : nullptr;
    switch (p.mode) {
    case SGBM_MODE_SGBM:
        a.inA = LhA; a.inB = LhB; a.sout = nullptr; a.backward = 0;
        if ((rc = sgbm_launch_vertical(a, 3, sweepSMs, st))) return rc;
        break;
    case SGBM_MODE_HH:
        // S_fwd overwrites LhA in place.  hhSplit: the forward sweep (HBM-bound) adds only L_hA, the backward
        // sweep (issue-bound, DRAM mostly idle) reads L_hB beside S_fwd -- same bytes in total, better balance.
        a.inA = LhA; a.inB = h->knobs.hhSplit ? nullptr : LhB; a.sout = LhA; a.backward = 0;
        if ((rc = sgbm_launch_vertical(a, 3, sweepSMs, st))) return rc;
        if ((rc = prof_mark(h, ST_VERT_FWD, st))) return rc;
        a.inA = LhA; a.inB = h->knobs.hhSplit ? LhB : nullptr; a.sout = nullptr; a.backward = 1;
        if ((rc = sgbm_launch_vertical(a, 3, sweepSMs, st))) return rc;
        break;
    case SGBM_MODE_SGBM_3WAY:
        a.inA = LhA; a.inB = LhB; a.sout = nullptr; a.threeway = 1;
        if ((rc = sgbm_launch_vertical(a, 1, h->numSMs, st))) return rc;
        break;
    case SGBM_MODE_HH4:
        a.inA = LhA; a.inB = LhB; a.sout = LhA; a.backward = 0;
        if ((rc = sgbm_launch_vertical(a, 1, h->numSMs, st))) return rc;
        if ((rc = prof_mark(h, ST_VERT_FWD, st))) return rc;
        a.inA = LhA; a.inB = nullptr; a.sout = nullptr; a.backward = 1;
        if ((rc = sgbm_launch_vertical(a, 1, h->numSMs, st))) return rc;
        break;
    }
    if ((rc = prof_mark(h, ST_VERT_WTA, st))) return rc;
    // LR check and 3x3 median are one kernel; with the debug hook on, the stand-alone LR check runs first so that
    // sgbm_debug_fetch(2) finds the checked raw disparity (the check is idempotent)
    if (h->keep && (rc = sgbm_launch_lrcheck(g, raw, d2key, st))) return rc;
    if ((rc = prof_mark(h, ST_LRCHECK, st))) return rc;
    const bool dense = outPitchElems == g.W;
    const bool speck = p.speckleWindowSize > 0;
    int16_t *mdst = (speck && !dense) ? med : out;
    long long mpitch = (speck && !dense) ? g.W : outPitchElems;
    if ((rc = sgbm_launch_lr_median(g, raw, d2key, mdst, mpitch, st))) return rc;
    if ((rc = prof_mark(h, ST_MEDIAN, st))) return rc;
    if (speck) {
        if ((rc = sgbm_launch_speckles(mdst, g.W, g.H, g.INV, p.speckleWindowSize, 16 * p.speckleRange, base + L.speck, st))) return rc;
        if (!dense)
            SGBM_CUDA_CHECK(cudaMemcpy2DAsync(out, (size_t)outPitchElems * 2, med, (size_t)g.W * 2, (size_t)g.W * 2, g.H,
                                              cudaMemcpyDeviceToDevice, st));
    }
    if (speck && (rc = prof_mark(h, ST_SPECKLE, st))) return rc;
    h->lastGeo = g; h->lastC = C; h->lastS = a.sdbg; h->lastRaw = raw; h->lastValid = 1;
    return 0;
}

extern "C" int sgbm_compute(sgbm_handle *h, const uint8_t *left, const uint8_t *right, int W, int H, int channels,
                            ptrdiff_t pitch_bytes, int batch, int16_t *disp_out, ptrdiff_t out_pitch_bytes,
                            void *cuda_stream)
{
    if (!h || !left || !right || !disp_out) return sgbm_fail(SGBM_E_INVALID_ARG, "null argument");
    if (batch <= 0) return sgbm_fail(SGBM_E_INVALID_ARG, "batch must be >= 1");
    if (int rcDev = check_device(h)) return rcDev;
    if (pitch_bytes < (ptrdiff_t)W * channels || out_pitch_bytes < (ptrdiff_t)W * 2 || (out_pitch_bytes & 1))
        return sgbm_fail(SGBM_E_INVALID_ARG, "bad pitch (in %td, out %td) for width %d", pitch_bytes, out_pitch_bytes, W);
    // One call at a time per handle; on the device this call is ordered behind the handle's previous call
    // whatever stream that ran on (the workspace, the lane / band streams and the watchdog words are the
    // handle's, not the stream's).
    std::lock_guard<std::mutex> lock(h->mu);
    KnobScope ks(h);
    Geo g;
    int rc = make_geo(h->p, W, H, channels, g);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    if (h->evLastValid) SGBM_CUDA_CHECK(cudaStreamWaitEvent(st, h->evLast, 0));
    WsLayout L;
    ws_layout(g, h->p, h->numSMs, h->keep, L);
    int sweepSMs = h->numSMs;
    int lanes = lanes_for(h, g, batch, &sweepSMs);
    if ((rc = ensure_ws(h, 0, L.total, st))) return rc;
    for (int i = 1; i < lanes; i++)
        if (ensure_ws(h, i, L.total, st)) { lanes = 1; sweepSMs = h->numSMs; break; }   // no memory for a second workspace
    if (lanes > 1) {
        // frames b % lanes != 0 run on internal streams forked from the caller's stream and joined before
        // returning: to the caller everything is still ordered on `cuda_stream`
        if ((rc = ensure_lane_streams(h, lanes))) return rc;
        SGBM_CUDA_CHECK(cudaEventRecord(h->evFork, st));
        for (int i = 1; i < lanes; i++) SGBM_CUDA_CHECK(cudaStreamWaitEvent(h->laneStream[i], h->evFork, 0));
    }
    for (int b = 0; b < batch; b++) {
        const int lane = b % lanes;
        rc = compute_frame(h, lane, sweepSMs, g, L, left + (size_t)b * pitch_bytes * H, right + (size_t)b * pitch_bytes * H,
                           pitch_bytes, (int16_t *)((uint8_t *)disp_out + (size_t)b * out_pitch_bytes * H), out_pitch_bytes / 2,
                           lane ? h->laneStream[lane] : st);
        if (rc) break;
    }
    for (int i = 1; i < lanes; i++) {
        SGBM_CUDA_CHECK(cudaEventRecord(h->evJoin[i], h->laneStream[i]));
        SGBM_CUDA_CHECK(cudaStreamWaitEvent(st, h->evJoin[i], 0));
    }
    SGBM_CUDA_CHECK(cudaEventRecord(h->evLast, st));
    h->evLastValid = true;
    return rc;
}

static int ensure_buf(void **p, size_t *have, size_t need, bool pinned)
{
    if (*have >= need) return 0;
    if (*p) { if (pinned) cudaFreeHost(*p); else cudaFree(*p); *p = nullptr; *have = 0; }
    cudaError_t e = pinned ? cudaMallocHost(p, need) : cudaMalloc(p, need);
    if (e != cudaSuccess) { cudaGetLastError(); return sgbm_fail(SGBM_E_NOMEM, "staging allocation of %zu bytes failed: %s", need, cudaGetErrorString(e)); }
    *have = need;
    return 0;
}

static int compute_host_impl(sgbm_handle *h, const uint8_t *left, const uint8_t *right, int W, int H, int channels,
                             ptrdiff_t pitch_bytes, int batch, int16_t *disp_out, ptrdiff_t out_pitch_bytes);

extern "C" int sgbm_compute_host(sgbm_handle *h, const uint8_t *left, const uint8_t *right, int W, int H, int channels,
                                 ptrdiff_t pitch_bytes, int batch, int16_t *disp_out, ptrdiff_t out_pitch_bytes)
{
    const int rc = compute_host_impl(h, left, right, W, H, channels, pitch_bytes, batch, disp_out, out_pitch_bytes);
    if (rc && h) {
        // an error in the middle of a batch: nothing may still be reading or writing the caller's
        // (possibly page-locked, directly DMA'd) buffers once this call has returned
        char keep[sizeof(g_err)];
        memcpy(keep, g_err, sizeof(keep));
        cudaStream_t all[3 + SGBM_MAX_LANES] = {h->inStream, h->ownStream, h->outStream};
        for (int i = 1; i < SGBM_MAX_LANES; i++) all[2 + i] = h->laneStream[i];
        for (cudaStream_t s : all)
            if (s) cudaStreamSynchronize(s);
        cudaGetLastError();
        memcpy(g_err, keep, sizeof(keep));
    }
    return rc;
}

static int compute_host_impl(sgbm_handle *h, const uint8_t *left, const uint8_t *right, int W, int H, int channels,
                             ptrdiff_t pitch_bytes, int batch, int16_t *disp_out, ptrdiff_t out_pitch_bytes)
{
    if (!h || !left || !right || !disp_out) return sgbm_fail(SGBM_E_INVALID_ARG, "null argument");
    if (batch <= 0) return sgbm_fail(SGBM_E_INVALID_ARG, "batch must be >= 1");
    if (int rcDev = check_device(h)) return rcDev;
    if (pitch_bytes < (ptrdiff_t)W * channels || out_pitch_bytes < (ptrdiff_t)W * 2)
        return sgbm_fail(SGBM_E_INVALID_ARG, "bad pitch");
    std::lock_guard<std::mutex> lock(h->mu);
    KnobScope ks(h);
    Geo g;
    int rc = make_geo(h->p, W, H, channels, g);
    if (rc) return rc;
    if (!h->ownStream) {
        SGBM_CUDA_CHECK(cudaStreamCreateWithFlags(&h->ownStream, cudaStreamNonBlocking));
        SGBM_CUDA_CHECK(cudaStreamCreateWithFlags(&h->inStream, cudaStreamNonBlocking));
        SGBM_CUDA_CHECK(cudaStreamCreateWithFlags(&h->outStream, cudaStreamNonBlocking));
        for (int i = 0; i < SGBM_MAX_SLOTS; i++) {
            SGBM_CUDA_CHECK(cudaEventCreateWithFlags(&h->evIn[i], cudaEventDisableTiming));
            SGBM_CUDA_CHECK(cudaEventCreateWithFlags(&h->evComp[i], cudaEventDisableTiming));
            SGBM_CUDA_CHECK(cudaEventCreateWithFlags(&h->evOut[i], cudaEventDisableTiming));
        }
    }
    cudaStream_t st = h->ownStream;
    if (h->evLastValid) SGBM_CUDA_CHECK(cudaEventSynchronize(h->evLast));         // a device-pointer call still in flight
    const size_t rowIn = (size_t)W * channels, frameIn = rowIn * H, frameOut = (size_t)W * H * 2;
    WsLayout L;
    ws_layout(g, h->p, h->numSMs, h->keep, L);
    int sweepSMs = h->numSMs;
    int lanes = lanes_for(h, g, batch, &sweepSMs);
    if ((rc = ensure_ws(h, 0, L.total, st))) return rc;
    for (int i = 1; i < lanes; i++)
        if (ensure_ws(h, i, L.total, st)) { lanes = 1; sweepSMs = h->numSMs; break; }   // no memory for a second workspace
    // staging slots: two per lane, so that every lane always has its next frame queued behind the running one
    const int nslots = batch > 1 ? (2 * lanes < batch ? 2 * lanes : batch) : 1;
    // Page-locked caller buffers (cudaHostAlloc / cudaHostRegister, e.g. torch pin_memory) with dense rows
    // are DMA'd directly; pageable ones are staged through the handle's pinned slots by the host thread.
    auto pinned = [](const void *p) {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
        return at.type == cudaMemoryTypeHost;
    };
    const bool directIn = (size_t)pitch_bytes == rowIn && pinned(left) && pinned(right);
    const bool directOut = (size_t)out_pitch_bytes == (size_t)W * 2 && pinned(disp_out);
    for (int i = 0; i < nslots; i++) {
        if (!directIn && (rc = ensure_buf(&h->hostIn[i], &h->hostInBytes[i], 2 * frameIn, true))) return rc;
        if (!directOut && (rc = ensure_buf(&h->hostOut[i], &h->hostOutBytes[i], frameOut, true))) return rc;
        if ((rc = ensure_buf(&h->devIn[i], &h->devInBytes[i], 2 * frameIn, false))) return rc;
        if ((rc = ensure_buf(&h->devOut[i], &h->devOutBytes[i], frameOut, false))) return rc;
    }
    if (lanes > 1) {                                      // the other lanes compute on their own streams and workspaces
        if ((rc = ensure_lane_streams(h, lanes))) return rc;
        SGBM_CUDA_CHECK(cudaStreamSynchronize(st));       // (their first-use memsets were enqueued on st)
    }
    // Three kinds of streams: H2D, kernels (one per lane), D2H; while the kernels of frame b run, later
    // frames go in and earlier ones come out (and, for pageable buffers, the host copies them into / out
    // of the other staging slots).
    auto drain = [&](int b) -> int {                      // wait for frame b's D2H and hand the rows to the caller
        const int sl = b % nslots;
        SGBM_CUDA_CHECK(cudaEventSynchronize(h->evOut[sl]));
        if (directOut) return 0;
        uint8_t *o = (uint8_t *)disp_out + (size_t)b * out_pitch_bytes * H;
        if ((size_t)out_pitch_bytes == (size_t)W * 2) memcpy(o, h->hostOut[sl], frameOut);
        else
            for (int y = 0; y < H; y++) memcpy(o + (size_t)y * out_pitch_bytes, (uint8_t *)h->hostOut[sl] + (size_t)y * W * 2, (size_t)W * 2);
        return 0;
    };
    for (int b = 0; b < batch; b++) {
        const int sl = b % nslots;
        if (b >= nslots && (rc = drain(b - nslots))) return rc;     // frees slot sl (its kernels and D2H are complete)
        const uint8_t *l = left + (size_t)b * pitch_bytes * H, *r = right + (size_t)b * pitch_bytes * H;
        if (directIn) {
            SGBM_CUDA_CHECK(cudaMemcpyAsync(h->devIn[sl], l, frameIn, cudaMemcpyHostToDevice, h->inStream));
            SGBM_CUDA_CHECK(cudaMemcpyAsync((uint8_t *)h->devIn[sl] + frameIn, r, frameIn, cudaMemcpyHostToDevice, h->inStream));
        } else {
            uint8_t *hi = (uint8_t *)h->hostIn[sl];
            if ((size_t)pitch_bytes == rowIn) {
                memcpy(hi, l, frameIn);
                memcpy(hi + frameIn, r, frameIn);
            } else {
                for (int y = 0; y < H; y++) {
                    memcpy(hi + (size_t)y * rowIn, l + (size_t)y * pitch_bytes, rowIn);
                    memcpy(hi + frameIn + (size_t)y * rowIn, r + (size_t)y * pitch_bytes, rowIn);
                }
            }
            SGBM_CUDA_CHECK(cudaMemcpyAsync(h->devIn[sl], h->hostIn[sl], 2 * frameIn, cudaMemcpyHostToDevice, h->inStream));
        }
        const int lane = b % lanes;
        cudaStream_t cs = lane ? h->laneStream[lane] : st;
        SGBM_CUDA_CHECK(cudaEventRecord(h->evIn[sl], h->inStream));
        SGBM_CUDA_CHECK(cudaStreamWaitEvent(cs, h->evIn[sl], 0));
        rc = compute_frame(h, lane, sweepSMs, g, L, (const uint8_t *)h->devIn[sl],
                           (const uint8_t *)h->devIn[sl] + frameIn, (long long)rowIn, (int16_t *)h->devOut[sl], W, cs);
        if (rc) return rc;
        SGBM_CUDA_CHECK(cudaEventRecord(h->evComp[sl], cs));
        SGBM_CUDA_CHECK(cudaStreamWaitEvent(h->outStream, h->evComp[sl], 0));
        void *dst = directOut ? (void *)((uint8_t *)disp_out + (size_t)b * frameOut) : h->hostOut[sl];
        SGBM_CUDA_CHECK(cudaMemcpyAsync(dst, h->devOut[sl], frameOut, cudaMemcpyDeviceToHost, h->outStream));
        SGBM_CUDA_CHECK(cudaEventRecord(h->evOut[sl], h->outStream));
        // (the next H2D into this slot is enqueued only after drain() has seen this frame's D2H complete, so
        // it cannot overtake these kernels; the H2D stream itself never waits for kernels)
    }
    for (int b = batch > nslots ? batch - nslots : 0; b < batch; b++)
        if ((rc = drain(b))) return rc;
    SGBM_CUDA_CHECK(cudaStreamSynchronize(st));
    for (int i = 1; i < lanes; i++) SGBM_CUDA_CHECK(cudaStreamSynchronize(h->laneStream[i]));
    return check_watch(h, st);
}

extern "C" int sgbm_status(sgbm_handle *h)
{
    if (!h) return sgbm_fail(SGBM_E_INVALID_ARG, "null handle");
    return check_watch(h, h->ownStream);
}

extern "C" int sgbm_disp_to_float(const int16_t *disp_x16, int W, int H, float *out, void *cuda_stream)
{
    if (!disp_x16 || !out || W <= 0 || H <= 0) return sgbm_fail(SGBM_E_INVALID_ARG, "bad argument");
    return sgbm_launch_disp_to_float(disp_x16, out, (size_t)W * H, (cudaStream_t)cuda_stream);
}

extern "C" int sgbm_reproject_f32(const float *disp, const double *Q, int W, int H, float *xyz, uint8_t *valid_or_null,
                                  void *cuda_stream)
{
    if (!disp || !Q || !xyz || W <= 0 || H <= 0) return sgbm_fail(SGBM_E_INVALID_ARG, "bad argument");
    return sgbm_launch_reproject(disp, 1, Q, W, H, xyz, valid_or_null, (cudaStream_t)cuda_stream);
}
extern "C" int sgbm_reproject_i16(const int16_t *disp, const double *Q, int W, int H, float *xyz, uint8_t *valid_or_null,
                                  void *cuda_stream)
{
    if (!disp || !Q || !xyz || W <= 0 || H <= 0) return sgbm_fail(SGBM_E_INVALID_ARG, "bad argument");
    return sgbm_launch_reproject(disp, 0, Q, W, H, xyz, valid_or_null, (cudaStream_t)cuda_stream);
}

extern "C" int sgbm_reproject_ex(const void *disp, int disp_depth, const double *Q, int W, int H, int handle_missing_values,
                                 int ddepth, void *out, void *scratch16, void *cuda_stream)
{
    if (!disp || !Q || !out || W <= 0 || H <= 0) return sgbm_fail(SGBM_E_INVALID_ARG, "bad argument");
    if (disp_depth != 0 && disp_depth != 3 && disp_depth != 4 && disp_depth != 5)
        return sgbm_fail(SGBM_E_INVALID_ARG, "disparity must be uint8, int16, int32 or float32 (cv2: stereo_geom.cpp:17)");
    if (ddepth == -1) ddepth = 5;
    if (ddepth != 3 && ddepth != 4 && ddepth != 5)
        return sgbm_fail(SGBM_E_INVALID_ARG, "ddepth must be -1, CV_16S, CV_32S or CV_32F");
    if (handle_missing_values && !scratch16) return sgbm_fail(SGBM_E_INVALID_ARG, "handleMissingValues needs 16 bytes of device scratch");
    return sgbm_launch_reproject_ex(disp, disp_depth, Q, W, H, handle_missing_values ? 1 : 0, ddepth, out, scratch16,
                                    (cudaStream_t)cuda_stream);
}

extern "C" int sgbm_reproject_compact_scratch_bytes(int W, int H, size_t *out)
{
    if (!out || W <= 0 || H <= 0) return sgbm_fail(SGBM_E_INVALID_ARG, "bad argument");
    *out = sgbm_compact_scratch_bytes(W, H);
    return 0;
}
extern "C" int sgbm_reproject_compact(const int16_t *disp_x16, const double *Q, int W, int H, const uint8_t *bgr,
                                      int bgr_channels, ptrdiff_t bgr_pitch_bytes, float *xyz_out, uint8_t *rgb_out,
                                      unsigned long long *n_out, void *scratch, size_t scratch_bytes, void *cuda_stream)
{
    if (!disp_x16 || !Q || !xyz_out || !n_out || !scratch || W <= 0 || H <= 0) return sgbm_fail(SGBM_E_INVALID_ARG, "bad argument");
    if (scratch_bytes < sgbm_compact_scratch_bytes(W, H)) return sgbm_fail(SGBM_E_INVALID_ARG, "scratch too small");
    if (rgb_out && (!bgr || (bgr_channels != 1 && bgr_channels != 3))) return sgbm_fail(SGBM_E_INVALID_ARG, "bad colour image");
    return sgbm_launch_compact(disp_x16, Q, W, H, bgr, bgr_channels, (long long)bgr_pitch_bytes, xyz_out,
                               rgb_out ? rgb_out : nullptr, n_out, scratch, (cudaStream_t)cuda_stream);
}

extern "C" int sgbm_filter_speckles(int16_t *img, int W, int H, int newVal, int maxSpeckleSize, int maxDiff, void *scratch,
                                    size_t scratch_bytes, void *cuda_stream)
{
    if (!img || !scratch || W <= 0 || H <= 0) return sgbm_fail(SGBM_E_INVALID_ARG, "bad argument");
    if (scratch_bytes < (size_t)W * H * 8) return sgbm_fail(SGBM_E_INVALID_ARG, "scratch too small (need W*H*8)");
    return sgbm_launch_speckles(img, W, H, newVal, maxSpeckleSize, maxDiff, scratch, (cudaStream_t)cuda_stream);
}
extern "C" int sgbm_median3x3(const int16_t *src, int16_t *dst, int W, int H, void *cuda_stream)
{
    if (!src || !dst || src == dst || W <= 0 || H <= 0) return sgbm_fail(SGBM_E_INVALID_ARG, "bad argument");
    return sgbm_launch_median(src, dst, W, H, W, (cudaStream_t)cuda_stream);
}

extern "C" int sgbm_init_rectify_map(const double *K, const double *dist_or_null, int n_dist, const double *R_or_null,
                                     const double *P, int p_cols, int W, int H, float *map1, float *map2, void *cuda_stream)
{
    if (!K || !P || !map1 || !map2 || W <= 0 || H <= 0) return sgbm_fail(SGBM_E_INVALID_ARG, "bad argument");
    if (p_cols != 3 && p_cols != 4) return sgbm_fail(SGBM_E_INVALID_ARG, "newCameraMatrix must be 3x3 or 3x4");
    for (int i = 0; dist_or_null && i < n_dist; i++)
        if (dist_or_null[i] != 0.0)
            return sgbm_fail(SGBM_E_UNSUPPORTED, "non-zero distortion coefficients are not implemented (the reference passes None)");
    return sgbm_launch_rectify_map(K, R_or_null, P, p_cols, W, H, map1, map2, (cudaStream_t)cuda_stream);
}

extern "C" int sgbm_remap_linear_u8(const uint8_t *src, int src_w, int src_h, int channels, ptrdiff_t src_pitch_bytes,
                                    const float *map1, const float *map2, int W, int H, uint8_t *dst, ptrdiff_t dst_pitch_bytes,
                                    void *cuda_stream)
{
    if (!src || !map1 || !map2 || !dst || src_w <= 0 || src_h <= 0 || W <= 0 || H <= 0) return sgbm_fail(SGBM_E_INVALID_ARG, "bad argument");
    if (src_pitch_bytes < (ptrdiff_t)src_w * channels || dst_pitch_bytes < (ptrdiff_t)W * channels) return sgbm_fail(SGBM_E_INVALID_ARG, "bad pitch");
    if (src_w > 32767 || src_h > 32767) return sgbm_fail(SGBM_E_UNSUPPORTED, "source larger than 32767 pixels (cv2 asserts the same for fixed-point remap)");
    return sgbm_launch_remap_linear(src, src_w, src_h, channels, (long long)src_pitch_bytes, map1, map2, W, H, dst,
                                    (long long)dst_pitch_bytes, (cudaStream_t)cuda_stream);
}

extern "C" int sgbm_host_alloc(size_t bytes, void **out)
{
    if (!out || bytes == 0) return sgbm_fail(SGBM_E_INVALID_ARG, "bad argument");
    *out = nullptr;
    cudaError_t e = cudaHostAlloc(out, bytes, cudaHostAllocPortable);
    if (e != cudaSuccess) {
        cudaGetLastError();
        *out = nullptr;
        return sgbm_fail(SGBM_E_NOMEM, "cudaHostAlloc of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
    }
    return 0;
}
extern "C" int sgbm_host_free(void *p)
{
    if (!p) return 0;
    SGBM_CUDA_CHECK(cudaFreeHost(p));
    return 0;
}

extern "C" int sgbm_debug_sweep_plan(const sgbm_params *p, int W, int H, int channels, int num_sms, int max_smem_bytes, int w_role,
                                     int input_volumes, int *out16)
{
    if (!p || !out16 || num_sms < 1 || max_smem_bytes < 1 || input_volumes < 1 || input_volumes > 2)
        return sgbm_fail(SGBM_E_INVALID_ARG, "bad argument");
    Geo g;
    const int rc = make_geo(*p, W, H, channels, g);       // (default knobs: no handle, no device)
    if (rc) return rc;
    sgbm_sweep_plan_debug(g, num_sms, max_smem_bytes, w_role, input_volumes, out16);
    return 0;
}

extern "C" int sgbm_debug_keep(sgbm_handle *h, int on)
{
    if (!h) return sgbm_fail(SGBM_E_INVALID_ARG, "null handle");
    h->keep = on ? 1 : 0;
    h->lastValid = 0;
    return 0;
}

extern "C" int sgbm_debug_fetch(sgbm_handle *h, int which, void *host_dst, size_t bytes)
{
    if (!h || !host_dst) return sgbm_fail(SGBM_E_INVALID_ARG, "null argument");
    if (!h->lastValid) return sgbm_fail(SGBM_E_INVALID_ARG, "no frame has been computed");
    const Geo &g = h->lastGeo;
    SGBM_CUDA_CHECK(cudaDeviceSynchronize());
    if (which == 2) {
        size_t need = (size_t)g.W * g.H * 2;
        if (bytes < need) return sgbm_fail(SGBM_E_INVALID_ARG, "buffer too small");
        SGBM_CUDA_CHECK(cudaMemcpy(host_dst, h->lastRaw, need, cudaMemcpyDeviceToHost));
        return 0;
    }
    const uint16_t *src = which == 0 ? h->lastC : h->lastS;
    if (!src) return sgbm_fail(SGBM_E_INVALID_ARG, "stage %d was not kept (call sgbm_debug_keep first)", which);
    size_t need = (size_t)g.H * g.W1 * g.D * 2;
    if (bytes < need) return sgbm_fail(SGBM_E_INVALID_ARG, "buffer too small");
    std::vector<uint16_t> tmp((size_t)g.rowStride * g.H);
    SGBM_CUDA_CHECK(cudaMemcpy(tmp.data(), src, tmp.size() * 2, cudaMemcpyDeviceToHost));
    uint16_t *dst = (uint16_t *)host_dst;
    std::vector<int> pos(g.D);
    for (int d = 0; d < g.D; d++) pos[d] = sgbm_pos(d, g.nreg, g.lpc);
    for (size_t c = 0; c < (size_t)g.H * g.W1; c++)
        for (int d = 0; d < g.D; d++) dst[c * g.D + d] = tmp[c * g.Dp + pos[d]];
    return 0;
}

extern "C" int sgbm_microbench_int16(int which, double *giga_lane_ops_per_s)
{
    if (!giga_lane_ops_per_s) return sgbm_fail(SGBM_E_INVALID_ARG, "null argument");
    return sgbm_run_microbench(which, giga_lane_ops_per_s);
}

extern "C" unsigned long long sgbm_kernel_launches(void) { return g_launches.load(); }

extern "C" int sgbm_profile_enable(sgbm_handle *h, int on)
{
    if (!h) return sgbm_fail(SGBM_E_INVALID_ARG, "null handle");
    h->prof = on ? 1 : 0;
    h->marks.clear();
    h->evUsed = 0;
    return 0;
}

extern "C" int sgbm_profile_read(sgbm_handle *h, char *names32, double *total_ms, int *runs, int *kernels, int max_stages,
                                 int *n_stages)
{
    if (!h || !names32 || !total_ms || !runs || !kernels || !n_stages) return sgbm_fail(SGBM_E_INVALID_ARG, "null argument");
    if (max_stages < ST_COUNT - 1) return sgbm_fail(SGBM_E_INVALID_ARG, "need room for %d stages", ST_COUNT - 1);
    if (!h->marks.empty()) SGBM_CUDA_CHECK(cudaEventSynchronize(h->marks.back().ev));
    double ms[ST_COUNT] = {0};
    int cnt[ST_COUNT] = {0}, ker[ST_COUNT] = {0};
    for (size_t i = 1; i < h->marks.size(); i++) {
        const ProfMark &a = h->marks[i - 1], &b = h->marks[i];
        if (b.stage == ST_START) continue;
        float t = 0;
        SGBM_CUDA_CHECK(cudaEventElapsedTime(&t, a.ev, b.ev));
        ms[b.stage] += t; cnt[b.stage]++; ker[b.stage] += (int)(b.launches - a.launches);
    }
    int n = 0;
    for (int s = 1; s < ST_COUNT; s++) {
        if (!cnt[s]) continue;
        strncpy(names32 + 32 * n, kStageNames[s], 31);
        names32[32 * n + 31] = 0;
        total_ms[n] = ms[s]; runs[n] = cnt[s]; kernels[n] = ker[s];
        n++;
    }
    *n_stages = n;
    h->marks.clear();
    h->evUsed = 0;
    return 0;
}
