#!/usr/bin/env python
"""Development: device-resident time per frame of a few (size, numDisparities, mode) combinations outside bench.py's
workloads.   python tools/time_modes.py   (knobs come from the environment, as everywhere)"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import stereo_reconstruction_cv_b200 as sg  # noqa: E402
from synth import make_pair  # noqa: E402

CASES = [(1920, 1080, 192, 1), (2560, 1440, 256, 1), (1280, 720, 128, 1), (1920, 1080, 64, 0), (1280, 720, 64, 1), (3840, 2160, 128, 0),
         (3840, 2160, 128, 1), (3840, 2160, 192, 1)]
for (W, H, D, mode) in CASES:
    l, r, _ = make_pair(W, H, D, seed=1)
    lt, rt = torch.from_numpy(l).cuda(), torch.from_numpy(r).cuda()
    st = sg.StereoSGBM_create(minDisparity=0, numDisparities=D, blockSize=5, P1=200, P2=800, disp12MaxDiff=1, preFilterCap=63,
                              uniquenessRatio=10, speckleWindowSize=100, speckleRange=32, mode=mode)
    for _ in range(3):
        st.compute(lt, rt)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 10
    for _ in range(n):
        st.compute(lt, rt)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print("%dx%d D=%d mode=%d: %.3f ms  %.1f GDE/s" % (W, H, D, mode, ms, W * H * D / ms / 1e6), flush=True)
