"""Golden vectors for the rectification warp, produced by the reference's own
implementation: cv2.initUndistortRectifyMap + cv2.remap exactly as called at main.ipynb:496-500,
with the notebook's intrinsics (main.ipynb:24-26) and a crop of dataset/d1.

Run in the build container (needs cv2 and /root/reference):  python tests/golden/make_golden_rectify.py
"""
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
K0 = np.array([[1733.74, 0, 792.27], [0, 1733.74, 541.89], [0, 0, 1]], np.float64)        # main.ipynb:24-26


def main():
    rng = np.random.default_rng(7)
    out = {}
    cases = []
    # (a) notebook-like: small rotation, new projection matrix, odd size (scalar tail of the row loop)
    R, _ = cv2.Rodrigues(np.array([0.012, -0.021, 0.004]))
    P = np.array([[1700.0, 0, 800.0, 0], [0, 1700.0, 540.0, 0], [0, 0, 1, 0]])
    cases.append(("nb", K0, R, P, (203, 61)))
    # (b) identity rotation, 3x3 P, width a multiple of 8
    cases.append(("id", K0, None, P[:, :3].copy(), (160, 48)))
    # (c) stronger rotation -> samples leave the image (constant border)
    R2, _ = cv2.Rodrigues(np.array([0.2, 0.15, -0.3]))
    cases.append(("far", K0 / np.array([[8.0], [8.0], [1.0]]), R2, P / np.array([[8.0], [8.0], [1.0]]), (131, 77)))
    src = cv2.imread("/root/reference/dataset/d1/img1.jpg", cv2.IMREAD_GRAYSCALE)[900:1000, 1500:1720].copy()
    srcc = cv2.imread("/root/reference/dataset/d1/img1.jpg", cv2.IMREAD_COLOR)[900:1000, 1500:1720].copy()
    out["src_gray"], out["src_bgr"] = src, srcc
    names = []
    for name, K, R, P, size in cases:
        m1, m2 = cv2.initUndistortRectifyMap(K, None, R, P, size, cv2.CV_32F)
        out[name + "_K"], out[name + "_P"], out[name + "_size"] = K, P, np.array(size)
        out[name + "_R"] = R if R is not None else np.zeros((0, 0))
        out[name + "_map1"], out[name + "_map2"] = m1, m2
        # shift the maps into the crop so that most samples are inside
        mx = m1 - m1.min() + np.float32(3.25)
        my = m2 - m2.min() - np.float32(1.5)
        out[name + "_mx"], out[name + "_my"] = mx, my
        out[name + "_remap_gray"] = cv2.remap(src, mx, my, interpolation=cv2.INTER_LINEAR)
        out[name + "_remap_bgr"] = cv2.remap(srcc, mx, my, interpolation=cv2.INTER_LINEAR)
        names.append(name)
    # random maps: sub-pixel grid coverage, exact integers, far outside
    mx = rng.uniform(-4, 224, (40, 57)).astype(np.float32)
    my = rng.uniform(-4, 104, (40, 57)).astype(np.float32)
    mx[:5] = np.round(mx[:5])
    my[:5] = np.round(my[:5])
    mx[5:8] *= 300
    my[8:10] *= -200
    out["rand_mx"], out["rand_my"] = mx, my
    out["rand_remap_gray"] = cv2.remap(src, mx, my, interpolation=cv2.INTER_LINEAR)
    out["rand_remap_bgr"] = cv2.remap(srcc, mx, my, interpolation=cv2.INTER_LINEAR)
    out["names"] = np.array(names)
    out["cv2_version"] = np.array(cv2.__version__)
    np.savez_compressed(os.path.join(HERE, "golden_rectify.npz"), **out)
    print("wrote golden_rectify.npz:", {k: getattr(v, "shape", None) for k, v in out.items() if k.endswith("remap_gray")})


if __name__ == "__main__":
    main()
