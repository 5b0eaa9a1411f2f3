#!/usr/bin/env python
"""Reads the clock64 trace of one strip of a sweep (TRACE=1 build, SGBM_SWEEP_TRACE=<file>):
  python tools/sweep_trace.py FILE H
Stamps per row: role V/A/C idx 0 loop top, 1 cost stage ready, 2 path step done (+cost stage released),
3 S slot ready, 4 S slot written + arrive, 5 end; producer (role 3) idx 0 top, 1 stage free, 2 copies issued;
W (role 3) idx 4 top, 5 slot full, 6 WTA done."""
import sys

import numpy as np

H = int(sys.argv[2])
a = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(H, 4, 8).astype(np.int64)
t0 = a[a > 0].min()
lo, hi = H // 3, H // 3 + min(400, H // 2)
x = a[lo:hi] - t0
print("rows %d..%d; total cycles/row %.0f" % (lo, hi, (a[hi - 1, 2, 4] - a[lo, 2, 4]) / (hi - 1 - lo)))
for r, nm in enumerate("VAC"):
    seg = np.diff(x[:, r, :6], axis=1)
    per = np.diff(x[:, r, 0])
    print("%s period mean %.0f p90 %.0f | wait cost %.0f, path step %.0f, wait S slot %.0f, S update %.0f, tail %.0f, loop %.0f"
          % (nm, per.mean(), np.percentile(per, 90), seg[:, 0].mean(), seg[:, 1].mean(), seg[:, 2].mean(), seg[:, 3].mean(),
             seg[:, 4].mean(), (x[1:, r, 0] - x[:-1, r, 5]).mean()))
p = x[:, 3, :3]
print("P period %.0f | wait free stage %.0f, issue %.0f" % (np.diff(p[:, 0]).mean(), (p[:, 1] - p[:, 0]).mean(), (p[:, 2] - p[:, 1]).mean()))
if (a[lo:hi, 3, 4] > 0).all():
    w = x[:, 3, 4:7]
    print("W period %.0f | wait slot %.0f, WTA %.0f" % (np.diff(w[:, 0]).mean(), (w[:, 1] - w[:, 0]).mean(), (w[:, 2] - w[:, 1]).mean()))
    print("lags: V.write->A.ready %.0f  A.write->C.ready %.0f  C.write->W.ready %.0f  W.free(row)->V.slot ready(row+K?) see below"
          % ((x[:, 1, 3] - x[:, 0, 4]).mean(), (x[:, 2, 3] - x[:, 1, 4]).mean(), (x[:, 3, 5] - x[:, 2, 4]).mean()))
    for K in (2, 3, 4, 5):
        print("   V.t4(row+%d) - W.t6(row) = %.0f" % (K, (x[K:, 0, 4] - x[:-K, 3, 6]).mean()))
print("cost stage: P.issue(row) -> V.ready(row) %.0f" % ((x[:, 0, 1] - x[:, 3, 2]).mean()))
