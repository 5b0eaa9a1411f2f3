#!/usr/bin/env python
"""Development: per-row time stamps of one strip of the last sweep of a workload (needs the TRACE=1 build
under build/trace, see tools/sweep_trace.py).   python tools/run_trace.py cfg2 [mode]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402,F401
from stereo_reconstruction_cv_b200 import _lib  # noqa: E402

_lib._LIB = _lib.load(os.path.join(ROOT, "build", "trace", "pkg", "libsgbm_b200.so"))
import stereo_reconstruction_cv_b200 as sg  # noqa: E402
from synth import make_pair  # noqa: E402

CFG = {"cfg1": (3840, 2160, 16, 0, dict(blockSize=11, P1=2904, P2=11616)), "cfg2": (1280, 720, 128, 0, {}),
       "cfg3": (3840, 2160, 256, 1, {}), "cfg4": (1920, 1080, 192, 0, {})}
name = sys.argv[1]
W, H, D, mode, over = CFG[name]
if len(sys.argv) > 2:
    mode = int(sys.argv[2])
kw = dict(minDisparity=0, numDisparities=D, blockSize=5, P1=200, P2=800, disp12MaxDiff=1, preFilterCap=63, uniquenessRatio=10,
          speckleWindowSize=100, speckleRange=32, mode=mode)
kw.update(over)
l, r, _ = make_pair(W, H, D, seed=0)
lt, rt = torch.from_numpy(l).cuda(), torch.from_numpy(r).cuda()
warm = sg.StereoSGBM_create(**kw)
for _ in range(3):
    warm.compute(lt, rt)
torch.cuda.synchronize()
path = os.path.join(ROOT, "gpurun_out", "trace_%s_m%d.bin" % (name, mode))
os.environ["SGBM_SWEEP_TRACE"] = path
os.environ["SGBM_VERBOSE"] = "1"
st = sg.StereoSGBM_create(**kw)
st.compute(lt, rt)
torch.cuda.synchronize()
print("==== %s mode %d" % (name, mode), flush=True)
subprocess.call([sys.executable, os.path.join(ROOT, "tools", "sweep_trace.py"), path, str(H)])
os.remove(path)
