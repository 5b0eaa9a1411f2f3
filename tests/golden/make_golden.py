"""Generate the golden vectors under tests/golden/ from the reference's own implementation.

Run in the build container (needs cv2 and /root/reference):   python tests/golden/make_golden.py
The reference calls cv2.StereoSGBM_create(...).compute (main.ipynb:655-668) and
cv2.reprojectImageTo3D (main.ipynb:697); this script calls exactly those on small inputs
(synthetic pairs, pure-noise pairs and crops of the reference's dataset/d1..d3 images) and stores
inputs + outputs in golden_small.npz.  For the full-size BASELINE.json configs it stores only
SHA-256 digests of input and output (golden_digests.json).
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
import cv2  # noqa: E402

from oracle import OracleParams, cv2_ref  # noqa: E402
from synth import make_noise_pair, make_pair  # noqa: E402

REF = "/root/reference/dataset"
NOTEBOOK_Q = np.array([[1, 0, 0, -1909.9754], [0, 1, 0, -1057.74529], [0, 0, 0, 2045.48384],
                       [0, 0, -1, 0]], np.float64)                      # main.ipynb:598-607


def notebook_params(ndisp=16, mindis=0, mode=0):                        # main.ipynb:655-666
    return OracleParams(mindis, ndisp, 11, 8 * 3 * 11 ** 2, 32 * 3 * 11 ** 2, 1, 63, 10, 100, 32, mode)


def std_params(D, mode, bs=5, minD=0):
    return OracleParams(minD, D, bs, 8 * bs * bs, 32 * bs * bs, 1, 63, 10, 100, 32, mode)


def crop(ds, x0, y0, w, h):
    l = cv2.imread(f"{REF}/{ds}/img1.jpg", cv2.IMREAD_GRAYSCALE)[y0:y0 + h, x0:x0 + w].copy()
    r = cv2.imread(f"{REF}/{ds}/img2.jpg", cv2.IMREAD_GRAYSCALE)[y0:y0 + h, x0:x0 + w].copy()
    return l, r


def small_cases():
    cases = []
    for mode in (0, 1, 2, 3):
        l, r, _ = make_pair(160, 64, 32, seed=10 + mode)
        cases.append((f"synth_m{mode}", std_params(32, mode), l, r))
        l, r = make_noise_pair(96, 48, seed=20 + mode)
        cases.append((f"noise_m{mode}", std_params(16, mode, bs=3), l, r))
        l, r = crop("d1", 1500, 900, 200, 72)
        cases.append((f"d1crop_notebook_m{mode}", notebook_params(16, 0, mode), l, r))
        l, r = crop("d2", 800, 500, 176, 64)
        cases.append((f"d2crop_minD_m{mode}", OracleParams(-8, 48, 7, 100, 1500, 2, 31, 15, 50, 2, mode), l, r))
        l, r = crop("d3", 2000, 1000, 144, 56)
        cases.append((f"d3crop_defaults_m{mode}", OracleParams(3, 32, 0, 0, 0, 0, 0, 0, 0, 0, mode), l, r))
    l, r, _ = make_pair(128, 48, 16, seed=5)
    l3 = np.stack([l, np.roll(l, 1, 1), l[::-1].copy()], -1)
    r3 = np.stack([r, np.roll(r, 1, 1), r[::-1].copy()], -1)
    cases.append(("synth_3ch_m0", std_params(16, 0), l3, r3))
    return cases


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    cv2.setNumThreads(0)
    out = {}
    meta = {"cv2_version": cv2.__version__, "cases": {}}
    for name, p, l, r in small_cases():
        disp = cv2_ref.compute(p, l, r)
        out[name + "__left"] = l
        out[name + "__right"] = r
        out[name + "__disp"] = disp
        meta["cases"][name] = {k: int(v) for k, v in p.__dict__.items()}
    # post filters + reprojection on one map
    l, r, _ = make_pair(192, 96, 32, seed=3)
    raw = cv2_ref.compute(OracleParams(0, 32, 5, 200, 800, 1, 63, 10, 0, 0, 0), l, r)
    out["post__in"] = raw
    out["post__median"] = cv2.medianBlur(raw, 3)
    sp = raw.copy()
    cv2.filterSpeckles(sp, -16, 60, 32)
    out["post__speckle_60_32"] = sp
    dispf = raw.astype(np.float32) / 16.0
    dispf = dispf * (dispf > 0).astype(np.float32)                       # main.ipynb:668-670
    out["reproj__disp_f32"] = dispf
    out["reproj__Q"] = NOTEBOOK_Q
    out["reproj__xyz"] = cv2.reprojectImageTo3D(dispf, NOTEBOOK_Q)      # main.ipynb:697
    Qg = np.random.default_rng(7).normal(size=(4, 4))
    out["reproj__Q_general"] = Qg
    out["reproj__xyz_general"] = cv2.reprojectImageTo3D(dispf, Qg)
    out["reproj__xyz_i16"] = cv2.reprojectImageTo3D(raw, NOTEBOOK_Q)    # int16 used as is (A.8)
    np.savez_compressed(os.path.join(HERE, "golden_small.npz"), **out)

    # digests of mid/full-size configs (inputs regenerated from the seed at test time)
    dig = {}
    for name, (W, H, D, mode) in {"cfg2_1280x720_D128_SGBM": (1280, 720, 128, 0),
                                  "mid_640x360_D64_SGBM": (640, 360, 64, 0),
                                  "mid_640x360_D64_HH": (640, 360, 64, 1),
                                  "mid_640x360_D64_3WAY": (640, 360, 64, 2),
                                  "cfg2_1280x720_D128_HH": (1280, 720, 128, 1),
                                  "cfg2_1280x720_D128_3WAY": (1280, 720, 128, 2)}.items():
        l, r, _ = make_pair(W, H, D, seed=0)
        p = std_params(D, mode)
        disp = cv2_ref.compute(p, l, r)
        dig[name] = {"W": W, "H": H, "D": D, "mode": mode, "seed": 0, "left_sha256": sha(l),
                     "right_sha256": sha(r), "disp_sha256": sha(disp),
                     "valid_fraction": float((disp >= 0).mean())}
    if "--full" in sys.argv:
        for name, (W, H, D, mode) in {"cfg3_3840x2160_D256_HH": (3840, 2160, 256, 1),
                                      "cfg5_3840x2160_D256_3WAY": (3840, 2160, 256, 2),
                                      "cfg4_1920x1080_D192_SGBM_seed0": (1920, 1080, 192, 0)}.items():
            l, r, _ = make_pair(W, H, D, seed=0)
            disp = cv2_ref.compute(std_params(D, mode), l, r)
            dig[name] = {"W": W, "H": H, "D": D, "mode": mode, "seed": 0, "left_sha256": sha(l),
                         "right_sha256": sha(r), "disp_sha256": sha(disp),
                         "valid_fraction": float((disp >= 0).mean())}
    else:
        old = os.path.join(HERE, "golden_digests.json")
        if os.path.exists(old):
            for k, v in json.load(open(old))["digests"].items():
                dig.setdefault(k, v)
    meta["digests"] = dig
    json.dump(meta, open(os.path.join(HERE, "golden_digests.json"), "w"), indent=1, sort_keys=True)
    print("wrote", len(out), "arrays;", len(dig), "digests")


if __name__ == "__main__":
    main()
