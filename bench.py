#!/usr/bin/env python
"""bench.py -- throughput of the dense-stereo hot path (BASELINE.json metric: MDE/s = W*H*D / s / 1e6).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg3|cfg2|cfg4|cfg5] [--impl reference]

A step is one stereo pair of the workload per GPU (weak scaling: every rank processes its own
frame, no collective on the data path).  Default workload cfg3 = BASELINE.json configs[2], the
configuration the north-star target is quoted on: synthetic 3840x2160 pair, D=256, MODE_HH,
speckle filter + LR check enabled.

  value    : whole-job MDE/s with the inputs already resident in HBM (CUDA events, max over ranks)
  e2e      : the same metric through the reference-facing call StereoSGBM.compute(numpy, numpy)
             (C ABI host entry point: pinned staging, H2D and D2H inside the timed region)
  roofline : dominant kernel's algorithmic HBM bytes / its event-timed duration vs measured peak
  alu      : same kernel against the measured packed-int16 issue rate (the binding roofline)
  cpu_baseline : the reference's own implementation (cv2.StereoSGBM) on a bounded sample

--impl reference times cv2.StereoSGBM (falls back to the C oracle port when cv2 is missing) on
the box's host cores with one compute() per thread.
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

# algorithmic HBM bytes per (x1, y, d) cost-volume element of each stage's interface
STAGE_BYTES_PER_ELEM = {"cost": 2, "cost_alt": 2, "horizontal": 8, "vertical_fwd": 8, "vertical_wta": None}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback"


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
            self._stop.wait(0.1)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join()
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


def make_inputs(W, H, D, seed):
    from stereo_reconstruction_cv_b200.synth import make_pair
    l, r, _ = make_pair(W, H, D, seed=seed)
    return l, r


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own implementation (cv2.StereoSGBM on host cores)
# ------------------------------------------------------------------------------------------------
def params_for(workload):
    return dict(PARAMS, **WORKLOAD_PARAMS.get(workload, {}))


def _ref_compute_fn(D, mode, workload="cfg3"):
    from oracle import cv2_ref
    kw = dict(params_for(workload), numDisparities=D, mode=mode)
    if cv2_ref.available():
        import cv2
        cv2.setNumThreads(1)

        def fn(l, r):
            st = cv2.StereoSGBM_create(**kw)
            return st.compute(l, r)
        return fn, "reference", "cv2 %s" % cv2.__version__
    import oracle
    p = oracle.OracleParams(**kw)
    return (lambda l, r: oracle.compute(p, l, r)), "port", "oracle C port"


def cpu_sample_rows(H, D, mode):
    # bounded sample: a full-width band of the workload's frame (MODE_HH holds C and S whole:
    # ~ 4 bytes * W * rows * D per call, so the band also bounds host memory)
    return min(H, 270 if D >= 192 else 360)


def run_reference(args, W, H, D, mode, modename):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    fn, kind, what = _ref_compute_fn(D, mode, args.workload)
    rows = cpu_sample_rows(H, D, mode)
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
    sample = "%d threads x one %dx%d band (D=%d, %s) of the workload per step; %s" % (threads, W, rows, D, modename, what)
    out = {"metric": "MDE/s", "value": value, "unit": "MDE/s", "impl": "reference", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "int16", "data": "synthetic",
           "config": {"workload": "%s: synthetic %dx%d pair, D=%d, %s, speckle+LR on (bounded sample: %s)"
                      % (args.workload, W, H, D, modename, sample)},
           "cpu_baseline": {"value": value, "unit": "MDE/s", "cores": threads, "kind": kind, "sample": sample},
           "e2e": {"value": value, "unit": "MDE/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    emit(out)


def cpu_baseline_single(W, H, D, mode, modename, workload="cfg3"):
    fn, kind, what = _ref_compute_fn(D, mode, workload)
    rows = cpu_sample_rows(H, D, mode)
    l, r = make_inputs(W, H, D, 0)
    y0 = (H - rows) // 2
    lb, rb = np.ascontiguousarray(l[y0:y0 + rows]), np.ascontiguousarray(r[y0:y0 + rows])
    t0 = time.perf_counter()
    fn(lb, rb)
    dt = time.perf_counter() - t0
    return {"value": float(W) * rows * D / dt / 1e6, "unit": "MDE/s", "cores": 1, "kind": kind,
            "sample": "one %dx%d band (rows %d..%d) of the workload frame, D=%d %s, single call (%s; SGBM/HH are "
                      "single-threaded in OpenCV), %.1f s" % (W, rows, y0, y0 + rows, D, modename, what, dt)}


# ------------------------------------------------------------------------------------------------
# product arm
# ------------------------------------------------------------------------------------------------
def run_product(args, W, H, D, mode, modename):
    import ctypes as C

    import torch
    import torch.distributed as dist

    import stereo_reconstruction_cv_b200 as sg
    from stereo_reconstruction_cv_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    L = _lib.lib()

    # Every rank owns its own frames (block sharding of the batch, sharding.shard_range): a pool of
    # `pool` distinct synthetic pairs per rank, resident in HBM, visited round-robin by the steps.
    pool = max(1, min(args.pool, args.steps + max(args.warmup, 3)))
    frames = [make_inputs(W, H, D, seed=rank * pool + i) for i in range(pool)]
    l, r = frames[0]
    lts = [torch.from_numpy(f[0]).to(dev) for f in frames]
    rts = [torch.from_numpy(f[1]).to(dev) for f in frames]
    st = sg.StereoSGBM_create(numDisparities=D, mode=mode, **params_for(args.workload))
    with_reproject = args.workload == "cfg5"
    # Frames per step: the video-sized workloads (BASELINE cfg2 / cfg4, "batch of ... pairs") hand the engine
    # six pairs per call, which lets it run two or three of them side by side, each on its share of the
    # SMs (sgbm_compute batch schedule); the 4K workloads are one pair per step.
    fps = args.frames_per_step if args.frames_per_step > 0 else (6 if (H <= 1080 and not with_reproject) else 1)
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

    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    # ---- timed region: inputs resident in HBM, CUDA events on the launching stream ----------------
    # (per-stage events are recorded inside the timed region for one-frame steps; with two frames side by
    # side the stages of the two lanes overlap, so they are timed frame by frame in a pass of their own)
    if fps == 1:
        check(L, L.sgbm_profile_enable(st._h, 1))
    launches0 = L.sgbm_kernel_launches()
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
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
        for i in range(min(args.steps, pool)):
            st.compute(lts[i], rts[i], single)
        stages = read_stages()
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    evals_frame = float(W) * H * D
    evals_step = evals_frame * fps
    value = evals_step * args.steps * world / (ms_max * 1e-3) / 1e6

    # ---- end to end through the reference-facing call with HOST buffers ------------------------------
    # One call of the public host API per timed region: e2e_steps frames from HOST numpy arrays to HOST
    # int16 disparities (compute_batch -> sgbm_compute_host: pinned staging, H2D, kernels, D2H all inside).
    # cfg5 additionally reprojects + compacts on the device and reads the point cloud back per frame.
    e2e_steps = max(2, min(args.steps, 16))

    def pinned_like(shape, dtype):                         # page-locked numpy array (the contract's "pinned host memory")
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
        for i in range(e2e_steps):                              # the cloud is read back frame by frame
            d = st.compute(torch.from_numpy(hl[i]).to(dev, non_blocking=True), torch.from_numpy(hr[i]).to(dev, non_blocking=True))
            pts, _ = sg.reprojectCompact(d, NOTEBOOK_Q, to_host=True)       # numpy view of a pinned buffer
            d2h_cloud[0] = int(pts.size) * 4
            if first is None:
                first = d.cpu().numpy()
        return first

    res = e2e_once()
    barrier()
    t0 = time.perf_counter()
    res = e2e_once()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = evals_frame * e2e_steps * world / float(t.item()) / 1e6
    chk = torch.empty((H, W), dtype=torch.int16, device=dev)
    st.compute(lts[0], rts[0], chk)
    same = bool((torch.from_numpy(res).to(dev) == chk).all().item())

    if rank == 0:
        peaks, peak_kind = measured_peaks()
        W1 = W - D
        elems = float(W1) * H * D
        # dominant stage by device time
        dom = max(stages, key=lambda k: stages[k]["ms"]) if stages else None
        bpe = {"cost": 2, "cost_alt": 2, "horizontal": 8, "vertical_fwd": 8,
               "vertical_wta": 4 if mode in (1, 3) else 6}.get(dom, 0)
        roof = None
        alu = None
        traffic = None
        try:                                                 # dram bytes per launch from the committed `ncu --set full` capture
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            traffic = tj.get(args.workload, {}).get(dom)
        except Exception:
            pass
        if dom:
            dom_ms = stages[dom]["ms"]
            ach = bpe * elems / (dom_ms * 1e-3) / 1e9
            roof = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": ach / peaks["hbm_gbs"], "traffic": traffic, "peak_source": peak_kind,
                    "kernel_ms": dom_ms,
                    # share of the step for one-frame steps; with frames side by side the kernels of the lanes
                    # overlap, so the share is taken of one frame's summed kernel time
                    "share_of_step": dom_ms / (ms / args.steps) if fps == 1 else dom_ms / sum(v["ms"] for v in stages.values()),
                    "algorithmic_bytes_per_launch": bpe * elems}
            # whole pipeline: sum of the stages' algorithmic bytes over the step time
            pipe_bpe = {0: 16, 1: 22, 2: 16, 3: 22}[mode]
            roof["pipeline"] = {"algorithmic_bytes_per_step": pipe_bpe * elems,
                                "achieved": pipe_bpe * elems * fps / (ms / args.steps * 1e-3) / 1e9,
                                "frac": pipe_bpe * elems * fps / (ms / args.steps * 1e-3) / 1e9 / peaks["hbm_gbs"]}
            try:
                mix = sg.microbench_int16(6)               # G lane-ops/s of the path-step instruction mix
                ops_per_elem = {"horizontal": 2 * 3.3, "vertical_fwd": 3 * 3.3, "vertical_wta": 3 * 3.3 + 1.5,
                                "cost": 13.6, "cost_alt": 13.6}.get(dom, 0.0)
                if mode == 2 and dom == "vertical_wta":
                    ops_per_elem = 3.3 + 6.0
                a_ach = ops_per_elem * elems / (dom_ms * 1e-3) / 1e9
                alu = {"kernel": dom, "achieved": a_ach, "peak": mix, "unit": "G lane-ops/s (packed u16x2 mix, "
                       "microbenchmarked)", "frac": a_ach / mix, "ops_per_elem": ops_per_elem}
            except Exception as ex:                          # pragma: no cover
                alu = {"error": str(ex)}
        cpu = cpu_baseline_single(W, H, D, mode, modename, args.workload) if world == 1 or rank == 0 else None
        outj = {"metric": "MDE/s", "value": value, "unit": "MDE/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_max / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "int16", "data": "synthetic",
                "config": {"workload": "%s (BASELINE.json configs[%d]): %s synthetic %dx%d rectified pair%s per GPU per step, "
                                       "D=%d, blockSize=%d, %s, speckle filter + LR check on%s"
                                       % (args.workload, WORKLOADS[args.workload][5], "one" if fps == 1 else str(fps), W, H,
                                          "" if fps == 1 else "s (one batched call)", D, params_for(args.workload)["blockSize"], modename,
                                          ", + fused reprojectImageTo3D/compaction" if with_reproject else ""),
                           "frames_per_step": fps,
                           "frames_per_s": args.steps * fps * world / (ms_max * 1e-3),
                           "l2": "no flush: each step streams the %.1f GB cost/path volumes (>> 126 MB L2)"
                                 % (3 * elems * 2 / 1e9),
                           "frame_pool": "%d distinct synthetic pairs per GPU (seeds rank*%d..), resident in HBM, round-robin" % (pool, pool),
                           "parallelism": "frames sharded over %d GPU(s), no collective" % world},
                "e2e": {"value": e2e_value, "unit": "MDE/s", "h2d_bytes_per_step": int(2 * W * H),
                        "d2h_bytes_per_step": int(d2h_cloud[0] if with_reproject else 2 * W * H), "steps": e2e_steps, "matches_device_path": same,
                        "api": ("StereoSGBM.compute + reprojectCompact(to_host) per pair" if with_reproject else
                                "StereoSGBM.compute_batch(%d pairs, page-locked numpy in/out): the batched host entry point keeps "
                                "two or three frames in flight, so it can exceed a one-pair-per-step device figure" % e2e_steps)},
                "gpu_launches": launches, "clocks": clocks, "stages_ms": {k: round(v["ms"], 4) for k, v in stages.items()},
                "roofline": roof, "alu_roofline": alu, "cpu_baseline": cpu}
        emit(outj)
    if world > 1:
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
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--pool", type=int, default=4, help="distinct synthetic pairs per GPU visited round-robin")
    ap.add_argument("--frames-per-step", type=int, default=0, help="pairs per compute() call (0: 6 for <= 1080p, else 1)")
    args = ap.parse_args()
    W, H, D, mode, modename, _ = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, W, H, D, mode, modename)
    else:
        run_product(args, W, H, D, mode, modename)


if __name__ == "__main__":
    main()
