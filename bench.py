#!/usr/bin/env python
"""bench.py -- throughput of the dense-stereo hot path (BASELINE.json metric: MDE/s = W*H*D / s / 1e6).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg1..cfg5] [--impl reference] [--no-extras]

Headline (the one JSON line's top-level keys): cfg3 = BASELINE.json configs[2], the configuration the
north-star target is quoted on -- synthetic 3840x2160 pair, D=256, MODE_HH, speckle filter + LR check on.
A step is one stereo pair per GPU (weak scaling: every rank processes its own frames, no collective on
the data path).

  value    : whole-job MDE/s with the inputs already resident in HBM (CUDA events, max over ranks)
  e2e      : the same metric through the reference-facing call with HOST buffers (C ABI host entry point:
             H2D and D2H inside the timed region), median of >= 3 repetitions; e2e_single_call is the
             notebook's literal call -- one stereo.compute(imgL, imgR) on pageable numpy arrays (main.ipynb:668)
             -- cold (first call of a new object: workspace allocation included) and warm
  roofline : dominant kernel's interface bytes / its event-timed duration vs the measured HBM peak, plus the
             whole frame against (a) the bytes this design moves and (b) SURVEY 8(d)'s compulsory bytes
  cpu_baseline : the reference's own implementation (cv2.StereoSGBM) on a bounded sample, same run

The default run also carries every other BASELINE config in `workloads` (cfg1, cfg2, cfg4, cfg5: value,
ms, e2e, roofline, cpu_baseline each; cfg1 / cfg2 / cfg4 also as batched calls that run frames side by side) and the 512-pair batch of configs[3] in `batch512`: 512 1080p pairs in
host memory, split over the N ranks by sharding.shard_range (strong scaling), each rank one compute_batch call.

--impl reference times cv2.StereoSGBM (the oracle port when cv2 is missing) on the box's host cores.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (W, H, D, mode, modename, BASELINE.json config index)
    "cfg1": (3840, 2160, 16, 0, "MODE_SGBM", 0),      # the notebook's literal call (main.ipynb:655-666, 781) at the dataset's size
    "cfg2": (1280, 720, 128, 0, "MODE_SGBM", 1),
    "cfg3": (3840, 2160, 256, 1, "MODE_HH", 2),
    "cfg4": (1920, 1080, 192, 0, "MODE_SGBM", 3),
    "cfg5": (3840, 2160, 256, 2, "MODE_SGBM_3WAY", 4),
}
PARAMS = dict(minDisparity=0, blockSize=5, P1=200, P2=800, disp12MaxDiff=1, preFilterCap=63, uniquenessRatio=10,
              speckleWindowSize=100, speckleRange=32)
# per-workload overrides: cfg1 uses the notebook's parameters (blockSize 11, 3-channel P1 / P2 recipe)
WORKLOAD_PARAMS = {"cfg1": dict(blockSize=11, P1=8 * 3 * 11 ** 2, P2=32 * 3 * 11 ** 2)}
NOTEBOOK_Q = np.array([[1, 0, 0, -1909.9754], [0, 1, 0, -1057.74529], [0, 0, 0, 2045.48384], [0, 0, -1, 0]],
                      np.float64)                                          # main.ipynb:598-607
BATCH512_PAIRS = 512
BATCH512_DISTINCT = 16           # distinct synthetic pairs (seeds 0..15), tiled x32 to 512 frames (SURVEY 8(d) cfg4 allows it)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU during the timed region (pynvml)."""

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake": 0x80}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.02)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join()
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


_PAIR_CACHE = {}


def make_inputs(W, H, D, seed):
    from synth import make_pair
    key = (W, H, D, seed)
    if key not in _PAIR_CACHE:
        l, r, _ = make_pair(W, H, D, seed=seed)
        _PAIR_CACHE[key] = (l, r)
    return _PAIR_CACHE[key]


def params_for(workload):
    return dict(PARAMS, **WORKLOAD_PARAMS.get(workload, {}))


def config_for(workload):
    """The workload as named by BASELINE.json -- identical in the product and the reference arm."""
    W, H, D, mode, modename, idx = WORKLOADS[workload]
    p = params_for(workload)
    return {"workload": "%s (BASELINE.json configs[%d]): synthetic %dx%d rectified pair, numDisparities=%d, blockSize=%d, P1=%d, "
                        "P2=%d, %s, uniquenessRatio=10, disp12MaxDiff=1, speckle filter (100, 32) + LR check on%s"
                        % (workload, idx, W, H, D, p["blockSize"], p["P1"], p["P2"], modename,
                           ", + /16, mask, reprojectImageTo3D(Q of main.ipynb:598-607), mask + compaction to XYZ" if workload == "cfg5" else ""),
            "generator": "SURVEY.md 8(d) synthetic pair generator, seeds 0..",
            "l2": "no flush: every frame streams cost / path volumes far larger than the 126 MB L2",
            "parallelism": "frames sharded over the GPUs, one process per GPU, no collective on the data path"}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own implementation (cv2.StereoSGBM on host cores)
# ------------------------------------------------------------------------------------------------
def _ref_compute_fn(workload, threads_inside=1):
    W, H, D, mode, modename, _ = WORKLOADS[workload]
    from oracle import cv2_ref
    kw = dict(params_for(workload), numDisparities=D, mode=mode)
    if cv2_ref.available():
        import cv2
        cv2.setNumThreads(int(threads_inside))

        def fn(l, r):
            st = cv2.StereoSGBM_create(**kw)
            d = st.compute(l, r)
            if workload == "cfg5":                          # the notebook's tail (main.ipynb:668-670, 697, 726-737)
                f = d.astype(np.float32) / 16.0
                f = f * (f > 0).astype(np.float32)
                xyz = cv2.reprojectImageTo3D(f, NOTEBOOK_Q)
                m = ~np.isnan(xyz[:, :, 0]) & ~np.isinf(xyz[:, :, 0]) & (f > 0)
                return xyz[m]
            return d
        return fn, "reference", "cv2 %s" % cv2.__version__
    import oracle
    p = oracle.OracleParams(**kw)
    return (lambda l, r: oracle.compute(p, l, r)), "port", "oracle C port"


def cpu_sample_rows(workload):
    # bounded sample: a full-width band of the workload's frame.  MODE_HH holds C and S whole (~ 4 bytes * W * rows * D
    # per call), so the band also bounds host memory; 3WAY always cuts its input into 4 stripes, so cfg5 takes whole frames.
    W, H, D, mode, _, _ = WORKLOADS[workload]
    if workload == "cfg5":
        return H
    return min(H, 270 if D >= 192 else 360)


def cpu_baseline_single(workload):
    """One call of the reference implementation on a bounded sample, the way the reference itself runs it: SGBM / HH are
    single-threaded inside OpenCV; 3WAY (cfg5) uses OpenCV's thread pool, so it gets cv2.setNumThreads(os.cpu_count())."""
    W, H, D, mode, modename, _ = WORKLOADS[workload]
    ncpu = os.cpu_count() or 1
    inside = ncpu if mode == 2 else 1
    fn, kind, what = _ref_compute_fn(workload, inside)
    rows = cpu_sample_rows(workload)
    l, r = make_inputs(W, H, D, 0)
    y0 = (H - rows) // 2
    lb, rb = np.ascontiguousarray(l[y0:y0 + rows]), np.ascontiguousarray(r[y0:y0 + rows])
    t0 = time.perf_counter()
    fn(lb, rb)
    dt = time.perf_counter() - t0
    return {"value": float(W) * rows * D / dt / 1e6, "unit": "MDE/s", "cores": min(inside, 4) if mode == 2 else 1, "kind": kind,
            "sample": "one call on %s of the workload frame, numDisparities=%d %s, cv2.setNumThreads(%d)%s (%s), %.2f s"
                      % ("the whole %dx%d frame" % (W, H) if rows == H else "a %dx%d band (rows %d..%d)" % (W, rows, y0, y0 + rows),
                         D, modename, inside, " -- 3WAY runs its 4 fixed stripes on the pool" if mode == 2 else
                         " -- SGBM / HH are single-threaded inside OpenCV", what, dt)}


def run_reference(args):
    """The reference arm: cv2.StereoSGBM on all the host cores it can use, on the headline workload's config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    workload = args.workload or "cfg3"
    W, H, D, mode, modename, _ = WORKLOADS[workload]
    fn, kind, what = _ref_compute_fn(workload, 1)
    rows = cpu_sample_rows(workload)
    ncpu = os.cpu_count() or 1
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 32 << 30
    per_call = 6 * W * rows * D + (64 << 20)
    threads = int(max(1, min(ncpu, 64, (avail // 4) // per_call)))
    l, r = make_inputs(W, H, D, 0)
    bands = []
    for t in range(threads):
        y0 = (t * 97) % max(H - rows, 1)
        bands.append((np.ascontiguousarray(l[y0:y0 + rows]), np.ascontiguousarray(r[y0:y0 + rows])))
    from concurrent.futures import ThreadPoolExecutor
    pool = ThreadPoolExecutor(threads)

    def step():
        list(pool.map(lambda b: fn(b[0], b[1]), bands))
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    evals = float(W) * rows * D * threads * args.steps
    value = evals / dt / 1e6
    # one WHOLE frame, once, single call (what the notebook's stereo.compute does): the band -> frame factor
    whole = None
    if not args.no_whole_frame:
        need = 4 * W * H * D * (2 if mode in (1, 3) else 1) + (1 << 30)
        if avail > need * 1.5:
            t1 = time.perf_counter()
            fn(l, r)
            dw = time.perf_counter() - t1
            one = cpu_baseline_single(workload)
            whole = {"seconds": dw, "value": float(W) * H * D / dw / 1e6, "unit": "MDE/s", "cores": 1,
                     "band_single_call_value": one["value"],
                     "band_to_frame_factor": (float(W) * H * D / dw / 1e6) / one["value"],
                     "note": "one whole %dx%d frame in one call (main.ipynb:668) vs the same call on a %d-row band: the bands of the "
                             "timed steps are kinder to cv2 than whole frames by the inverse of this factor" % (W, H, rows)}
        else:
            whole = {"skipped": "not enough host memory for a whole-frame call (~%.0f GB needed)" % (need / 2 ** 30)}
    sample = "%d threads x one %dx%d band per step, cv2.setNumThreads(1) inside each call; %s" % (threads, W, rows, what)
    out = {"metric": "MDE/s", "value": value, "unit": "MDE/s", "impl": "reference", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "int16", "data": "synthetic",
           "config": config_for(workload),
           "cpu_baseline": {"value": value, "unit": "MDE/s", "cores": threads, "kind": kind, "sample": sample},
           "whole_frame_single_call": whole,
           "host": {"cpu_count": ncpu, "available_gb": round(avail / 2 ** 30, 1)},
           "e2e": {"value": value, "unit": "MDE/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    emit(out)


# ------------------------------------------------------------------------------------------------
# product arm
# ------------------------------------------------------------------------------------------------
class Ctx:
    pass


def algorithmic_bytes(workload):
    """compulsory HBM bytes of one frame -- 2 u8 images in, int16 disparity out (4 B / pixel); MODE_HH
    additionally spills the forward S once (4 B per cost-volume element); cfg5 adds XYZ out (12 B / pixel)."""
    W, H, D, mode, _, _ = WORKLOADS[workload]
    b = 4.0 * W * H
    if mode in (1, 3):
        b += 4.0 * (W - D) * H * D
    if workload == "cfg5":
        b += 12.0 * W * H
    return b


def design_bytes(workload, hh_split):
    """Bytes this design's kernels move per frame through their interfaces: C written once and
    read by every path kernel, L_hA / L_hB written and read once, S_fwd spilled once in MODE_HH."""
    W, H, D, mode, _, _ = WORKLOADS[workload]
    elems = float(W - D) * H * D
    per = {0: 16, 1: 22, 2: 16, 3: 22}[mode]
    return per * elems + 4.0 * W * H


def stage_bytes_per_elem(stage, mode, hh_split):
    if stage in ("cost", "cost_alt"):
        return 2
    if stage == "horizontal":
        return 8
    if stage == "vertical_fwd":
        return (6 if hh_split else 8) if mode == 1 else 6
    if stage == "vertical_wta":
        if mode == 1:
            return 6 if hh_split else 4
        return 4 if mode == 3 else 6
    return 0


def run_workload(cx, workload, steps, warmup, fps=None, want_e2e=True, want_cpu=True, e2e_reps=3):
    """One BASELINE config: device-resident throughput (CUDA events, max over ranks), per-stage kernel times,
    end-to-end through the host API, roofline and the CPU baseline.  Returns the dict that goes into the JSON line."""
    import ctypes as C
    torch, dist, sg, L = cx.torch, cx.dist, cx.sg, cx.L
    W, H, D, mode, modename, _ = WORKLOADS[workload]
    dev, world, rank = cx.dev, cx.world, cx.rank
    with_reproject = workload == "cfg5"
    pool = max(1, min(cx.pool, steps + warmup))
    frames = [make_inputs(W, H, D, seed=rank * pool + i) for i in range(pool)]
    lts = [torch.from_numpy(f[0]).to(dev) for f in frames]
    rts = [torch.from_numpy(f[1]).to(dev) for f in frames]
    st = sg.StereoSGBM_create(numDisparities=D, mode=mode, **params_for(workload))
    # Frames per step: the 4K workloads are one pair per step.  The video-sized ones (cfg2 / cfg4, "batch of ... pairs")
    # are reported both ways: one pair per call (single-frame latency) and six pairs per call, which lets the engine run
    # two or three of them side by side, each on its share of the SMs (sgbm_compute batch schedule).
    if fps is None:
        fps = 1
    out = torch.empty((H, W) if fps == 1 else (fps, H, W), dtype=torch.int16, device=dev)
    if fps > 1:
        lb = [torch.stack([lts[(i + k) % pool] for k in range(fps)]) for i in range(pool)]
        rb = [torch.stack([rts[(i + k) % pool] for k in range(fps)]) for i in range(pool)]
    counter = [0]

    def step_device():
        i = counter[0] % pool
        counter[0] += 1
        if fps > 1:
            st.compute(lb[i], rb[i], out)
            return None
        st.compute(lts[i], rts[i], out)
        if with_reproject:
            return sg.reprojectCompact(out, NOTEBOOK_Q)
        return None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def read_stages():
        names = C.create_string_buffer(32 * 16)
        tot = (C.c_double * 16)()
        runs = (C.c_int * 16)()
        kers = (C.c_int * 16)()
        n = C.c_int()
        check(L, L.sgbm_profile_read(st._h, names, tot, runs, kers, 16, C.byref(n)))
        check(L, L.sgbm_profile_enable(st._h, 0))
        res = {}
        for i in range(n.value):
            nm = names.raw[32 * i:32 * i + 32].split(b"\0")[0].decode()
            res[nm] = {"ms": tot[i] / max(runs[i], 1), "kernels": kers[i] // max(runs[i], 1)}
        return res

    for _ in range(max(warmup, 3)):
        step_device()
    barrier()
    # ---- timed region: inputs resident in HBM, CUDA events on the launching stream ----------------
    # (per-stage events are recorded inside the timed region for one-frame steps; with frames side by side the
    # stages of the lanes overlap, so they are timed frame by frame in a pass of their own)
    if fps == 1:
        check(L, L.sgbm_profile_enable(st._h, 1))
    launches0 = L.sgbm_kernel_launches()
    sampler = ClockSampler(cx.local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(steps):
        step_device()
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    launches = int(L.sgbm_kernel_launches() - launches0)
    if fps == 1:
        stages = read_stages()
    else:
        check(L, L.sgbm_profile_enable(st._h, 1))
        single = torch.empty((H, W), dtype=torch.int16, device=dev)
        for i in range(min(steps, pool)):
            st.compute(lts[i], rts[i], single)
        stages = read_stages()
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    evals_frame = float(W) * H * D
    value = evals_frame * fps * steps * world / (ms_max * 1e-3) / 1e6
    res = {"value": value, "unit": "MDE/s", "ms_per_step": ms_max / steps, "frames_per_step": fps,
           "frames_per_s": steps * fps * world / (ms_max * 1e-3), "steps": steps, "gpu_launches": launches,
           "stages_ms": {k: round(v["ms"], 4) for k, v in stages.items()}, "clocks": clocks, "config": config_for(workload)}
    hh_split = bool(int(os.environ.get("SGBM_HH_SPLIT", "1")))

    # ---- end to end through the reference-facing call with HOST buffers ------------------------------
    if want_e2e:
        e2e_steps = max(2, min(steps, 16))

        def pinned_like(shape, dtype):                     # page-locked numpy array (the contract's "pinned host memory")
            return torch.empty(shape, dtype=dtype, pin_memory=True).numpy()
        hl = pinned_like((e2e_steps, H, W), torch.uint8)
        hr = pinned_like((e2e_steps, H, W), torch.uint8)
        for i in range(e2e_steps):
            hl[i], hr[i] = frames[i % pool]
        hout = pinned_like((e2e_steps, H, W), torch.int16)
        d2h_cloud = [0]

        def e2e_once():
            if not with_reproject:
                st.compute_batch(hl, hr, hout)
                return hout[0]
            first = None
            for i in range(e2e_steps):                          # the cloud is read back frame by frame
                d = st.compute(torch.from_numpy(hl[i]).to(dev, non_blocking=True), torch.from_numpy(hr[i]).to(dev, non_blocking=True))
                pts, _ = sg.reprojectCompact(d, NOTEBOOK_Q, to_host=True)       # numpy view of a pinned buffer
                d2h_cloud[0] = int(pts.size) * 4
                if first is None:
                    first = d.cpu().numpy()
            return first

        got = e2e_once()                                   # warm (staging buffers, second workspace of the batch schedule)
        reps = []
        for _ in range(max(e2e_reps, 3)):
            barrier()
            t0 = time.perf_counter()
            got = e2e_once()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            tt = torch.tensor([dt], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            reps.append(float(tt.item()))
        dt_med = float(np.median(reps))
        chk = torch.empty((H, W), dtype=torch.int16, device=dev)
        st.compute(lts[0], rts[0], chk)
        same = bool((torch.from_numpy(np.ascontiguousarray(got)).to(dev) == chk).all().item())
        res["e2e"] = {"value": evals_frame * e2e_steps * world / dt_med / 1e6, "unit": "MDE/s",
                      "h2d_bytes_per_step": int(2 * W * H), "d2h_bytes_per_step": int(d2h_cloud[0] if with_reproject else 2 * W * H),
                      "steps": e2e_steps, "repetitions_s": [round(x, 5) for x in reps], "statistic": "median of %d" % len(reps),
                      "matches_device_path": same,
                      "api": ("StereoSGBM.compute + reprojectCompact(to_host) per pair" if with_reproject else
                              "StereoSGBM.compute_batch(%d pairs, page-locked numpy in/out): the batched host entry point keeps two "
                              "or three frames in flight, so it can exceed a one-pair-per-step device figure" % e2e_steps)}
        # the notebook's literal call: ONE stereo.compute(imgL, imgR) on pageable numpy arrays (main.ipynb:668), rank 0
        if rank == 0:
            pl, pr = np.array(frames[0][0], copy=True), np.array(frames[0][1], copy=True)      # pageable
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            st2 = sg.StereoSGBM_create(numDisparities=D, mode=mode, **params_for(workload))
            d_cold = st2.compute(pl, pr)
            cold = time.perf_counter() - t0
            warm = []
            for _ in range(7):        # results come from a recycling pool of page-locked blocks: the first calls of a new size allocate
                t0 = time.perf_counter()
                d_warm = st2.compute(pl, pr)
                warm.append(time.perf_counter() - t0)
            res["e2e_single_call"] = {"api": "StereoSGBM_create(...).compute(imgL, imgR), pageable numpy in, new numpy array out (main.ipynb:655-668)",
                                      "cold_ms": cold * 1e3, "warm_ms": float(np.median(warm)) * 1e3, "warm_calls_ms": [round(x * 1e3, 2) for x in warm],
                                      "cold_includes": "object creation, %.1f GB workspace cudaMalloc + zeroing, pinned staging allocation"
                                                       % (st2.workspaceBytes(W, H) / 1e9),
                                      "warm_value": evals_frame / float(np.median(warm)) / 1e6, "unit": "MDE/s",
                                      "matches_device_path": bool(np.array_equal(d_cold, chk.cpu().numpy()) and np.array_equal(d_warm, d_cold))}
            del st2

    # ---- roofline -------------------------------------------------------------------------------------
    if rank == 0:
        peaks, peak_kind = measured_peaks()
        elems = float(W - D) * H * D
        dom = max(stages, key=lambda k: stages[k]["ms"]) if stages else None
        traffic = None
        try:                                                 # dram bytes per launch from the committed `ncu --set full` capture
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            traffic = tj.get(workload, {}).get(dom)
        except Exception:
            pass
        if dom:
            bpe = stage_bytes_per_elem(dom, mode, hh_split)
            dom_ms = stages[dom]["ms"]
            ach = bpe * elems / (dom_ms * 1e-3) / 1e9
            frame_ms = ms_max / steps / fps
            res["roofline"] = {
                "bound": "hbm", "kernel": dom, "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": ach / peaks["hbm_gbs"], "traffic": traffic, "peak_source": peak_kind, "kernel_ms": dom_ms,
                # share of the step for one-frame steps; with frames side by side the kernels of the lanes
                # overlap, so the share is taken of one frame's summed kernel time
                "share_of_step": dom_ms / (ms / steps) if fps == 1 else dom_ms / sum(v["ms"] for v in stages.values()),
                "algorithmic_bytes_per_launch": bpe * elems, "bytes_per_element": bpe,
                "frame": {"design_bytes": design_bytes(workload, hh_split),
                          "design_bytes_frac": design_bytes(workload, hh_split) / (frame_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                          "algorithmic_bytes": algorithmic_bytes(workload),
                          "algorithmic_frac": algorithmic_bytes(workload) / (frame_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                          "note": "design_bytes = what this design's kernels move through their interfaces (C materialised and "
                                  "re-read by every path kernel); algorithmic_bytes = SURVEY 8(d)'s compulsory traffic"}}
            try:
                mix = sg.microbench_int16(6)               # G lane-ops/s of the path-step instruction mix
                ops_per_elem = {"horizontal": 2 * 3.3, "vertical_fwd": 3 * 3.3, "vertical_wta": 3 * 3.3 + 1.5,
                                "cost": 13.6, "cost_alt": 13.6}.get(dom, 0.0)
                if mode == 2 and dom == "vertical_wta":
                    ops_per_elem = 3.3 + 6.0
                a_ach = ops_per_elem * elems / (dom_ms * 1e-3) / 1e9
                res["alu_roofline"] = {"kernel": dom, "achieved": a_ach, "peak": mix, "unit": "G lane-ops/s (packed u16x2 mix, "
                                       "microbenchmarked)", "frac": a_ach / mix, "ops_per_elem": ops_per_elem}
            except Exception as ex:                          # pragma: no cover
                res["alu_roofline"] = {"error": str(ex)}
        if want_cpu:
            res["cpu_baseline"] = cpu_baseline_single(workload)
    del st
    torch.cuda.empty_cache()
    return res


def run_batch512(cx):
    """BASELINE configs[3]: 512 1920x1080 pairs (D=192, MODE_SGBM) in HOST memory, split over the ranks by
    sharding.shard_range (strong scaling), each rank ONE compute_batch call on its block -- H2D, kernels, D2H included."""
    torch, dist, sg = cx.torch, cx.dist, cx.sg
    from stereo_reconstruction_cv_b200 import sharding
    W, H, D, mode, modename, _ = WORKLOADS["cfg4"]
    start, stop = sharding.shard_range(BATCH512_PAIRS, cx.world, cx.rank)
    n = stop - start
    distinct = [make_inputs(W, H, D, seed=s) for s in range(BATCH512_DISTINCT)]
    hl = torch.empty((max(n, 1), H, W), dtype=torch.uint8, pin_memory=True).numpy()
    hr = torch.empty((max(n, 1), H, W), dtype=torch.uint8, pin_memory=True).numpy()
    hout = torch.empty((max(n, 1), H, W), dtype=torch.int16, pin_memory=True).numpy()
    for i in range(n):
        hl[i], hr[i] = distinct[(start + i) % BATCH512_DISTINCT]
    st = sg.StereoSGBM_create(numDisparities=D, mode=mode, **params_for("cfg4"))
    if n:
        st.compute_batch(hl[:min(n, 6)], hr[:min(n, 6)], hout[:min(n, 6)])       # warm: workspaces, staging, kernel attributes
    if cx.world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if n:
        st.compute_batch(hl[:n], hr[:n], hout[:n])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    tt = torch.tensor([dt], dtype=torch.float64, device=cx.dev)
    if cx.world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt = float(tt.item())
    # every distinct pair's result must equal the one-pair device path (rank 0 checks its first 16)
    ok = True
    if cx.rank == 0 and n:
        for i in range(min(n, BATCH512_DISTINCT)):
            d = st.compute(torch.from_numpy(hl[i]).to(cx.dev), torch.from_numpy(hr[i]).to(cx.dev)).cpu().numpy()
            ok = ok and bool(np.array_equal(d, hout[i]))
    del st
    return {"pairs": BATCH512_PAIRS, "n_gpus": cx.world, "pairs_per_rank": [b - a for a, b in sharding.shard_ranges(BATCH512_PAIRS, cx.world)],
            "seconds": dt, "frames_per_s": BATCH512_PAIRS / dt, "value": float(W) * H * D * BATCH512_PAIRS / dt / 1e6, "unit": "MDE/s",
            "scaling": "strong", "h2d_bytes": int(2 * W * H) * BATCH512_PAIRS, "d2h_bytes": int(2 * W * H) * BATCH512_PAIRS,
            "matches_device_path": ok,
            "workload": "512 synthetic 1920x1080 pairs (%d distinct, seeds 0..%d, tiled), D=192, MODE_SGBM, page-locked host arrays in "
                        "and out, block-partitioned by sharding.shard_range, one StereoSGBM.compute_batch call per rank"
                        % (BATCH512_DISTINCT, BATCH512_DISTINCT - 1)}


def run_gather(cx):
    """The optional NCCL exchange of SURVEY 8(e): every rank's compacted cfg5 cloud (main.ipynb:726-737) gathered on rank 0."""
    torch, dist, sg = cx.torch, cx.dist, cx.sg
    from stereo_reconstruction_cv_b200 import sharding
    W, H, D, mode, _, _ = WORKLOADS["cfg5"]
    l, r = make_inputs(W, H, D, seed=cx.rank)
    st = sg.StereoSGBM_create(numDisparities=D, mode=mode, **params_for("cfg5"))
    d = st.compute(torch.from_numpy(l).to(cx.dev), torch.from_numpy(r).to(cx.dev))
    col = torch.from_numpy(np.stack([l, r, l], -1)).to(cx.dev)
    xyz, rgb = sg.reprojectCompact(d, NOTEBOOK_Q, col)
    sharding.gather_point_cloud(xyz, rgb, dst=0)            # warm (NCCL channels)
    times = []
    gx = None
    for _ in range(3):
        if cx.world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        gx, gc = sharding.gather_point_cloud(xyz, rgb, dst=0)
        e1.record()
        torch.cuda.synchronize()
        tt = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=cx.dev)
        if cx.world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        times.append(float(tt.item()))
    cnt = torch.tensor([xyz.shape[0]], dtype=torch.int64, device=cx.dev)
    if cx.world > 1:
        dist.all_reduce(cnt)
    total = int(cnt.item())
    ms = float(np.median(times))
    out = {"api": "sharding.gather_point_cloud (all_gather of counts + gather of padded buffers, torch.distributed %s)"
                  % ("nccl" if cx.world > 1 else "none: single rank"),
           "points_total": total, "bytes_gathered": total * 15, "ms": ms, "gb_per_s": total * 15 / (ms * 1e-3) / 1e9 if ms > 0 else None,
           "off_the_headline": True}
    if cx.rank == 0 and gx is not None:
        out["rank0_points"] = int(gx.shape[0])
        out["complete"] = bool(gx.shape[0] == total)
    return out


def run_product(args):
    import torch
    import torch.distributed as dist

    import stereo_reconstruction_cv_b200 as sg
    from stereo_reconstruction_cv_b200 import _lib

    cx = Ctx()
    cx.torch, cx.dist, cx.sg = torch, dist, sg
    cx.world = int(os.environ.get("WORLD_SIZE", "1"))
    cx.rank = int(os.environ.get("RANK", "0"))
    cx.local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(cx.local)
    if cx.world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", cx.local))
    cx.dev = torch.device("cuda", cx.local)
    cx.L = _lib.lib()
    cx.pool = args.pool
    headline = args.workload or "cfg3"
    extras = not args.no_extras and args.workload is None
    W, H, D, mode, modename, _ = WORKLOADS[headline]
    fps = args.frames_per_step if args.frames_per_step > 0 else 1
    t_all = time.perf_counter()
    main = run_workload(cx, headline, args.steps, args.warmup, fps=fps)
    outj = {"metric": "MDE/s", "value": main["value"], "unit": "MDE/s", "n_gpus": cx.world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": main["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int16", "data": "synthetic",
            "config": main["config"],
            "details": {"frames_per_step": main["frames_per_step"], "frames_per_s": main["frames_per_s"],
                        "frame_pool": "%d distinct synthetic pairs per GPU (seeds rank*%d..), resident in HBM, round-robin" % (cx.pool, cx.pool),
                        "hh_split": os.environ.get("SGBM_HH_SPLIT")},
            "e2e": main.get("e2e"), "e2e_single_call": main.get("e2e_single_call"),
            "gpu_launches": main["gpu_launches"], "clocks": main["clocks"], "stages_ms": main["stages_ms"],
            "roofline": main.get("roofline"), "alu_roofline": main.get("alu_roofline"), "cpu_baseline": main.get("cpu_baseline")}
    if extras:
        # the other BASELINE configs, shorter runs; cfg2 / cfg4 both one pair per call and six pairs per call
        wl = {}
        sub_steps = max(5, min(args.steps, 10))
        for name, f in (("cfg1", 1), ("cfg1_batch8", 8), ("cfg2", 1), ("cfg2_batch6", 6), ("cfg4", 1), ("cfg4_batch6", 6), ("cfg5", 1)):
            base = name.split("_")[0]
            r = run_workload(cx, base, sub_steps, 3, fps=f, want_e2e=(f == 1), want_cpu=(f == 1), e2e_reps=3)
            r.pop("config", None)
            r["workload"] = config_for(base)["workload"] + ("; %d pairs per compute() call, run side by side" % f if f > 1 else "")
            wl[name] = r
        outj["workloads"] = wl
        outj["batch512"] = run_batch512(cx)
        outj["gather"] = run_gather(cx)
    elif args.workload == "cfg5" and cx.world > 1:
        outj["gather"] = run_gather(cx)
    outj["bench_wall_s"] = round(time.perf_counter() - t_all, 1)
    if cx.rank == 0:
        emit(outj)
    if cx.world > 1:
        dist.barrier()
        dist.destroy_process_group()


def check(L, rc):
    if rc != 0:
        raise RuntimeError(L.sgbm_last_error().decode())


_REAL_STDOUT = None


def emit(obj):
    """The one JSON line of the contract, on the process's real stdout."""
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(obj) + "\n")
    out.flush()


def main():
    global _REAL_STDOUT
    # Libraries chat on stdout (NCCL prints its version banner there): keep the real stdout for the JSON
    # line and send everything else written to fd 1 to stderr.
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="product", choices=["product", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS),
                    help="headline workload (default cfg3, with every other BASELINE config in `workloads`)")
    ap.add_argument("--no-extras", action="store_true", help="headline workload only (profiling runs)")
    ap.add_argument("--no-whole-frame", action="store_true", help="reference arm: skip the one whole-frame call")
    ap.add_argument("--pool", type=int, default=4, help="distinct synthetic pairs per GPU visited round-robin")
    ap.add_argument("--frames-per-step", type=int, default=0, help="pairs per compute() call of the headline workload (default 1)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_product(args)


if __name__ == "__main__":
    main()
