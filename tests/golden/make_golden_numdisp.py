"""Golden vectors for numDisparities that are NOT a multiple of 8 (cv2 documents % 16 but accepts any positive value;
SURVEY.md 8(c), probe P16), from the reference's own implementation:   python tests/golden/make_golden_numdisp.py
cv2.StereoSGBM_create(...).compute (main.ipynb:655-668) on small synthetic pairs -> golden_numdisp.npz + .json."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
import cv2  # noqa: E402

from oracle import OracleParams, cv2_ref  # noqa: E402
from synth import make_noise_pair, make_pair  # noqa: E402

arrays, meta = {}, {"cv2_version": cv2.__version__, "cases": {}}
for D in (4, 5, 7, 12, 20, 21, 27, 44, 100):
    for mode in (0, 1):
        for kind in ("synth", "noise"):
            W, H = 96 + D + (D % 7) * 8, 40
            l, r = make_pair(W, H, max(D, 8), seed=D)[:2] if kind == "synth" else make_noise_pair(W, H, seed=D)
            bs, minD = (5, 0) if kind == "synth" else (3, -2)
            p = OracleParams(minD, D, bs, 8 * bs * bs, 32 * bs * bs, 1, 63, 10, 40, 16, mode)
            name = "D%d_m%d_%s" % (D, mode, kind)
            ref = cv2_ref.compute(p, l, r)
            for _ in range(3):                            # the same call gives the same answer (MODE_SGBM / MODE_HH are single-threaded)
                assert np.array_equal(ref, cv2_ref.compute(p, l, r)), name
            arrays[name + "__disp"] = ref                 # inputs are regenerated from the recipe (W, H, kind, seed = D)
            meta["cases"][name] = dict(p.__dict__, W=W, H=H, kind=kind, seed=D)
np.savez_compressed(os.path.join(HERE, "golden_numdisp.npz"), **arrays)
json.dump(meta, open(os.path.join(HERE, "golden_numdisp.json"), "w"), indent=1)
print(len(meta["cases"]), "cases,", os.path.getsize(os.path.join(HERE, "golden_numdisp.npz")), "bytes")
