"""Full-size golden digests of the reference's OWN calls, produced by the reference's implementation (cv2).

Run in the build container (needs cv2 and /root/reference):   python tests/golden/make_golden_full.py

What is pinned (all at 3840x2160, the size of dataset/d1 and dataset/d3):

  notebook_call   the literal call of main.ipynb:781 -> compute_disparity_map(imgL, imgR, 16, 0)
                  (main.ipynb:655-668: blockSize 11, P1 2904, P2 11616, ...), on the raw grayscale pairs
                  dataset/d3 (what the notebook loads, main.ipynb:358-363) and dataset/d1, in
                  MODE_SGBM (the notebook's), MODE_HH and MODE_SGBM_3WAY.  This parameter set selects the
                  saturating S accumulation (5 * (189*121 + 11616) > 65535).
  d3_cloud        the rest of the notebook cell on d3: /16 + mask (main.ipynb:668-670),
                  reprojectImageTo3D with the recorded Q (main.ipynb:598-607, 697), mask + gather of
                  points and colours (main.ipynb:726-737).
  rectified_d1    dataset/d1 rectified by the notebook recipe (main.ipynb:401-444, 474-500; K of
                  main.ipynb:24-26) -- SIFT/FLANN/F/E/recoverPose/stereoRectify run HERE, the resulting
                  R1, R2, P1, P2, Q are stored so that the GPU test runs initUndistortRectifyMap + remap
                  through the repo's own warp (chain n1 -> a1-a7 -> a9) -- then numDisparities = 128 with
                  the notebook's parameters, /16 + mask, reprojectImageTo3D with that Q.
  cfg5_cloud      BASELINE.json configs[4]: synthetic 3840x2160 D=256 MODE_SGBM_3WAY, then the notebook's
                  tail with the notebook's Q: XYZ digest, point count, compacted points digest.

The JPEG pairs are copied to tests/golden/dataset/ (the GPU box has no /root/reference); the digests of
the decoded grayscale images are stored so that a different JPEG decoder FAILS the test loudly.
"""
import hashlib
import json
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
import cv2  # noqa: E402

from synth import make_pair  # noqa: E402

REF = "/root/reference/dataset"
NOTEBOOK_Q = np.array([[1, 0, 0, -1909.9754], [0, 1, 0, -1057.74529], [0, 0, 0, 2045.48384],
                       [0, 0, -1, 0]], np.float64)                      # main.ipynb:598-607
NOTEBOOK_K = np.array([[2.25370759e+03, 0, 1.92969309e+03], [0, 2.24471892e+03, 1.05763445e+03],
                       [0, 0, 1]], np.float64)                          # main.ipynb:24-26


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def notebook_sgbm(ndisp, mindis, mode):                                  # main.ipynb:655-666 (+ mode)
    return cv2.StereoSGBM_create(minDisparity=mindis, numDisparities=ndisp, blockSize=11, P1=8 * 3 * 11 ** 2,
                                 P2=32 * 3 * 11 ** 2, disp12MaxDiff=1, preFilterCap=63, uniquenessRatio=10,
                                 speckleWindowSize=100, speckleRange=32, mode=mode)


def notebook_tail(disp_i16, Q, color_bgr=None):
    """main.ipynb:668-670, 697, 726-737."""
    f = disp_i16.astype(np.float32) / 16.0
    f = f * (f > 0).astype(np.float32)
    xyz = cv2.reprojectImageTo3D(f, Q)
    mask = ~np.isnan(xyz[:, :, 0]) & ~np.isinf(xyz[:, :, 0]) & (f > 0)
    out = {"xyz_sha256": sha(xyz), "n_points": int(mask.sum()), "points_sha256": sha(xyz[mask])}
    if color_bgr is not None:
        rgb = cv2.cvtColor(color_bgr, cv2.COLOR_BGR2RGB)                 # main.ipynb:792
        out["rgb_sha256"] = sha(rgb[mask])
    return out


def notebook_rectify_recipe(imgL, imgR, K0):
    """main.ipynb:401-444 (SIFT + FLANN + F + E + recoverPose) and 491-493 (stereoRectify, alpha = 1)."""
    sift = cv2.SIFT_create()
    kl, dl = sift.detectAndCompute(imgL, None)
    kr, dr = sift.detectAndCompute(imgR, None)
    flann = cv2.FlannBasedMatcher(dict(algorithm=1, trees=5), dict(checks=50))
    pl, pr = [], []
    for m, n in flann.knnMatch(dl, dr, k=2):
        if m.distance < 0.7 * n.distance:
            pl.append(kl[m.queryIdx].pt)
            pr.append(kr[m.trainIdx].pt)
    pl, pr = np.int32(pl), np.int32(pr)
    F, mask = cv2.findFundamentalMat(pl, pr, cv2.FM_LMEDS)
    pl, pr = pl[mask.ravel() == 1], pr[mask.ravel() == 1]
    E, mask = cv2.findEssentialMat(pl, pr, K0, method=cv2.RANSAC, prob=0.999, threshold=1.0)
    _, R, T, _ = cv2.recoverPose(E, pl, pr, K0)
    size = (imgL.shape[1], imgL.shape[0])
    R1, R2, P1, P2, Q, _, _ = cv2.stereoRectify(K0, None, K0, None, size, R, T, alpha=1.0)
    return R1, R2, P1, P2, Q


def main():
    cv2.setNumThreads(0)
    out = {"cv2_version": cv2.__version__, "images": {}, "notebook_call": {}}
    os.makedirs(os.path.join(HERE, "dataset"), exist_ok=True)
    imgs = {}
    for ds in ("d1", "d3"):
        for k, fn in (("left", "img1.jpg"), ("right", "img2.jpg")):
            dst = os.path.join(HERE, "dataset", "%s_%s" % (ds, fn))
            shutil.copyfile(os.path.join(REF, ds, fn), dst)
            os.chmod(dst, 0o644)
            imgs[ds, k] = cv2.imread(dst, cv2.IMREAD_GRAYSCALE)
        out["images"][ds] = {"left_sha256": sha(imgs[ds, "left"]), "right_sha256": sha(imgs[ds, "right"]),
                             "shape": list(imgs[ds, "left"].shape)}
    for ds in ("d3", "d1"):
        for mode in (0, 1, 2):
            disp = notebook_sgbm(16, 0, mode).compute(imgs[ds, "left"], imgs[ds, "right"])
            out["notebook_call"]["%s_m%d" % (ds, mode)] = {"dataset": ds, "mode": mode, "numDisparities": 16,
                                                          "disp_sha256": sha(disp),
                                                          "valid_fraction": float((disp >= 0).mean())}
            if ds == "d3" and mode == 0:
                col = cv2.imread(os.path.join(HERE, "dataset", "d3_img1.jpg"))
                out["d3_cloud"] = dict(notebook_tail(disp, NOTEBOOK_Q, col), color_sha256=sha(col))
            print(ds, mode, out["notebook_call"]["%s_m%d" % (ds, mode)]["valid_fraction"], flush=True)
    # ---- rectified d1 ---------------------------------------------------------------------------------
    R1, R2, P1, P2, Q = notebook_rectify_recipe(imgs["d1", "left"], imgs["d1", "right"], NOTEBOOK_K)
    size = (3840, 2160)
    mL1, mL2 = cv2.initUndistortRectifyMap(NOTEBOOK_K, None, R1, P1, size, cv2.CV_32F)     # main.ipynb:496-497
    mR1, mR2 = cv2.initUndistortRectifyMap(NOTEBOOK_K, None, R2, P2, size, cv2.CV_32F)
    Lr = cv2.remap(imgs["d1", "left"], mL1, mL2, interpolation=cv2.INTER_LINEAR)           # main.ipynb:499-500
    Rr = cv2.remap(imgs["d1", "right"], mR1, mR2, interpolation=cv2.INTER_LINEAR)
    rect = {"K": NOTEBOOK_K.tolist(), "R1": R1.tolist(), "R2": R2.tolist(), "P1": P1.tolist(), "P2": P2.tolist(),
            "Q": Q.tolist(), "left_rect_sha256": sha(Lr), "right_rect_sha256": sha(Rr), "numDisparities": 128, "modes": {}}
    for mode in (0, 1, 2):
        disp = notebook_sgbm(128, 0, mode).compute(Lr, Rr)
        rect["modes"]["m%d" % mode] = {"disp_sha256": sha(disp), "valid_fraction": float((disp >= 0).mean())}
        if mode == 0:
            rect["cloud"] = notebook_tail(disp, Q)
        print("rect", mode, rect["modes"]["m%d" % mode]["valid_fraction"], flush=True)
    out["rectified_d1"] = rect
    # ---- cfg5 -----------------------------------------------------------------------------------------
    W, H, D = 3840, 2160, 256
    l, r, _ = make_pair(W, H, D, seed=0)
    disp = cv2.StereoSGBM_create(minDisparity=0, numDisparities=D, blockSize=5, P1=200, P2=800, disp12MaxDiff=1,
                                 preFilterCap=63, uniquenessRatio=10, speckleWindowSize=100, speckleRange=32,
                                 mode=2).compute(l, r)
    out["cfg5_cloud"] = dict(notebook_tail(disp, NOTEBOOK_Q, np.stack([l, r, l], -1)), W=W, H=H, D=D, mode=2, seed=0,
                             left_sha256=sha(l), right_sha256=sha(r), disp_sha256=sha(disp))
    json.dump(out, open(os.path.join(HERE, "golden_full.json"), "w"), indent=1, sort_keys=True)
    print("wrote golden_full.json")


if __name__ == "__main__":
    main()
