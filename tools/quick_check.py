#!/usr/bin/env python
"""Development check for FAST=1 builds (lane mappings of numDisparities 16 / 128 / 192 / 256 only):
the CUDA path against the C oracle on multi-strip images, every mode, repeated frames.
  python tools/quick_check.py [--full]     (--full adds the cv2 digests of tests/golden at full size)"""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from oracle import OracleParams  # noqa: E402
import stereo_reconstruction_cv_b200 as sg  # noqa: E402
from synth import make_noise_pair, make_pair  # noqa: E402

bad = 0
t0 = time.time()
for (W, H, D, bs, P1, P2) in [(1700, 90, 16, 11, 2904, 11616), (1500, 80, 128, 5, 200, 800), (1900, 70, 192, 5, 200, 800),
                              (2300, 64, 256, 5, 200, 800), (700, 60, 16, 3, 72, 288), (900, 50, 256, 9, 648, 2592)]:
    for mode in (0, 1, 2, 3):
        for kind in ("synth", "noise"):
            l, r = make_pair(W, H, D, seed=W + mode)[:2] if kind == "synth" else make_noise_pair(W, H, seed=W + mode)
            p = OracleParams(0, D, bs, P1, P2, 1, 63, 10, 100, 32, mode)
            ref = oracle.compute(p, l, r)
            st = sg.StereoSGBM_create(**p.__dict__)
            for rep in range(3):
                n = int((st.compute(l, r) != ref).sum())
                if n:
                    bad += 1
                    print("MISMATCH", W, H, D, bs, mode, kind, rep, n, flush=True)
print("oracle cases done, bad =", bad, "%.1fs" % (time.time() - t0), flush=True)
if "--full" in sys.argv:
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_digests.json")))["digests"]
    for name in ("cfg2_1280x720_D128_SGBM", "cfg2_1280x720_D128_HH", "cfg4_1920x1080_D192_SGBM_seed0", "cfg3_3840x2160_D256_HH",
                 "cfg5_3840x2160_D256_3WAY"):
        d = g[name]
        l, r, _ = make_pair(d["W"], d["H"], d["D"], seed=d["seed"])
        st = sg.StereoSGBM_create(minDisparity=0, numDisparities=d["D"], blockSize=5, P1=200, P2=800, disp12MaxDiff=1,
                                  preFilterCap=63, uniquenessRatio=10, speckleWindowSize=100, speckleRange=32, mode=d["mode"])
        for rep in range(2):
            ok = hashlib.sha256(st.compute(l, r).tobytes()).hexdigest() == d["disp_sha256"]
            if not ok:
                bad += 1
            print(name, rep, "ok" if ok else "MISMATCH", flush=True)
    import cv2
    full = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_full.json")))
    l = cv2.imread(os.path.join(ROOT, "tests/golden/dataset/d3_img1.jpg"), 0)
    r = cv2.imread(os.path.join(ROOT, "tests/golden/dataset/d3_img2.jpg"), 0)
    for mode in (0, 1, 2):
        st = sg.StereoSGBM_create(minDisparity=0, numDisparities=16, blockSize=11, P1=2904, P2=11616, disp12MaxDiff=1, preFilterCap=63,
                                  uniquenessRatio=10, speckleWindowSize=100, speckleRange=32, mode=mode)
        ok = hashlib.sha256(st.compute(l, r).tobytes()).hexdigest() == full["notebook_call"]["d3_m%d" % mode]["disp_sha256"]
        bad += 0 if ok else 1
        print("notebook d3 mode", mode, "ok" if ok else "MISMATCH", flush=True)
print("QUICK CHECK", "PASSED" if bad == 0 else "FAILED (%d)" % bad)
sys.exit(1 if bad else 0)
