"""Full-size parity at the reference's OWN calls (tests/golden/golden_full.json, made by cv2 with
tests/golden/make_golden_full.py):

  * main.ipynb:781 -> compute_disparity_map(imgL, imgR, 16, 0) on the raw 3840x2160 pairs dataset/d3
    (what the notebook loads) and dataset/d1: blockSize 11, P1 2904, P2 11616 -- the one BASELINE
    configuration (configs[0]) that selects the saturating sweep kernels and the D = 16 lane mapping;
  * the rest of that notebook cell on d3 (float conversion, reprojectImageTo3D with the recorded Q,
    mask + gather of the cloud);
  * dataset/d1 rectified by the notebook recipe THROUGH THE REPO'S OWN WARP, then numDisparities = 128
    with the notebook's parameters and the cloud (chain n1 -> a1-a7 -> a9);
  * BASELINE configs[4]: 3WAY at 3840x2160 D=256 + the notebook's tail with the notebook's Q;
  * the saturating sweep instantiations forced (SGBM_SWEEP_SAT=1) on the cfg3 / cfg4 digests.

Everything is SHA-256 equality with what cv2 produced; a missing digest, an image that decodes
differently or a synthetic generator that drifted FAILS (nothing here skips).
The CPU part pins the C oracle at the notebook's literal call at full size.
"""
import hashlib
import json
import os

import numpy as np
import pytest

import oracle
from oracle import OracleParams
from synth import make_pair

HERE = os.path.dirname(os.path.abspath(__file__))
FULL = json.load(open(os.path.join(HERE, "golden", "golden_full.json")))
NOTEBOOK_Q = np.array([[1, 0, 0, -1909.9754], [0, 1, 0, -1057.74529], [0, 0, 0, 2045.48384], [0, 0, -1, 0]],
                      np.float64)                                                    # main.ipynb:598-607
NOTEBOOK = dict(blockSize=11, P1=8 * 3 * 11 ** 2, P2=32 * 3 * 11 ** 2, disp12MaxDiff=1, preFilterCap=63,
                uniquenessRatio=10, speckleWindowSize=100, speckleRange=32)         # main.ipynb:655-666


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def load_pair(ds, color=False):
    """The committed copies of /root/reference/dataset/<ds>/img{1,2}.jpg, decoded like main.ipynb:362-364."""
    import cv2
    l = cv2.imread(os.path.join(HERE, "golden", "dataset", "%s_img1.jpg" % ds), cv2.IMREAD_GRAYSCALE)
    r = cv2.imread(os.path.join(HERE, "golden", "dataset", "%s_img2.jpg" % ds), cv2.IMREAD_GRAYSCALE)
    g = FULL["images"][ds]
    assert sha(l) == g["left_sha256"] and sha(r) == g["right_sha256"], \
        "the JPEG decoder here does not reproduce the pixels the digests were made from (cv2 %s there)" % FULL["cv2_version"]
    if color:
        return l, r, cv2.imread(os.path.join(HERE, "golden", "dataset", "%s_img1.jpg" % ds))
    return l, r


def notebook_tail_numpy(disp_i16):
    f = disp_i16.astype(np.float32) / 16.0                                           # main.ipynb:668
    return f * (f > 0).astype(np.float32)                                            # main.ipynb:669-670


# ------------------------------------------------------------------------------------------------
# CPU: the oracle at the notebook's literal call, full size
# ------------------------------------------------------------------------------------------------
def test_oracle_notebook_call_full_size_d3():
    l, r = load_pair("d3")
    p = OracleParams(minDisparity=0, numDisparities=16, mode=0, **NOTEBOOK)
    assert sha(oracle.compute(p, l, r)) == FULL["notebook_call"]["d3_m0"]["disp_sha256"]


def test_full_size_fixture_is_complete():
    assert sorted(FULL["notebook_call"]) == ["d1_m0", "d1_m1", "d1_m2", "d3_m0", "d3_m1", "d3_m2"]
    assert sorted(FULL["rectified_d1"]["modes"]) == ["m0", "m1", "m2"]
    for k in ("d3_cloud", "cfg5_cloud"):
        assert {"xyz_sha256", "n_points", "points_sha256", "rgb_sha256"} <= set(FULL[k])
    # the notebook's parameter set is outside the "sum of paths fits 16 bits" domain: it needs the saturating kernels
    assert 5 * (189 * 121 + NOTEBOOK["P2"]) > 65535


# ------------------------------------------------------------------------------------------------
# GPU
# ------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def sg():
    import torch
    assert torch.cuda.is_available()
    import stereo_reconstruction_cv_b200 as sg
    return sg


@pytest.mark.gpu
@pytest.mark.parametrize("ds", ["d3", "d1"])
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_notebook_call_full_size(sg, ds, mode):
    """stereo.compute(imgL, imgR) of main.ipynb:668 with the parameters of main.ipynb:655-666 (16, 0)."""
    l, r = load_pair(ds)
    g = FULL["notebook_call"]["%s_m%d" % (ds, mode)]
    st = sg.StereoSGBM_create(minDisparity=0, numDisparities=16, mode=mode, **NOTEBOOK)
    got = st.compute(l, r)
    assert got.shape == (2160, 3840) and got.dtype == np.int16
    assert abs(float((got >= 0).mean()) - g["valid_fraction"]) < 1e-12
    assert sha(got) == g["disp_sha256"]
    if mode == 0:                                          # and again through the batched device path, twice (hand-offs are timing dependent)
        import torch
        lt, rt = torch.from_numpy(np.stack([l, l])).cuda(), torch.from_numpy(np.stack([r, r])).cuda()
        for _ in range(2):
            out = st.compute(lt, rt).cpu().numpy()
            assert sha(out[0]) == g["disp_sha256"] and sha(out[1]) == g["disp_sha256"]


@pytest.mark.gpu
def test_notebook_cell_on_d3(sg):
    """main.ipynb:781-792 end to end: disparity, /16 + mask, reprojectImageTo3D(Q), mask + gather."""
    import torch
    from stereo_reconstruction_cv_b200 import disparity_tab as dt
    l, r, col = load_pair("d3", color=True)
    g = FULL["d3_cloud"]
    assert sha(col) == g["color_sha256"]
    d = dt.compute_disparity_map(l, r, 16, 0)                                        # main.ipynb:781
    pts = dt.reconstruct_3D(d, NOTEBOOK_Q)                                           # main.ipynb:790
    assert pts.shape == (2160, 3840, 3) and pts.dtype == np.float32
    assert sha(pts) == g["xyz_sha256"]
    vp, vc = dt.point_cloud_arrays(pts, col[:, :, ::-1], d)                          # main.ipynb:726-737 (RGB of imgL_color)
    assert len(vp) == g["n_points"] and sha(vp) == g["points_sha256"] and sha(vc) == g["rgb_sha256"]
    # the fused device tail gives the same cloud from the int16 map
    st = sg.StereoSGBM_create(minDisparity=0, numDisparities=16, mode=0, **NOTEBOOK)
    di = st.compute(torch.from_numpy(l).cuda(), torch.from_numpy(r).cuda())
    xyz, rgb = sg.reprojectCompact(di, NOTEBOOK_Q, torch.from_numpy(col).cuda())
    assert xyz.shape[0] == g["n_points"]
    assert sha(xyz.cpu().numpy()) == g["points_sha256"] and sha(rgb.cpu().numpy()) == g["rgb_sha256"]


@pytest.mark.gpu
def test_rectified_d1_chain(sg):
    """dataset/d1 -> the repo's initUndistortRectifyMap + remap with the recipe's R1/R2/P1/P2 (main.ipynb:496-500)
    -> SGBM numDisparities 128 at the notebook's parameters -> cloud with the recipe's Q."""
    from stereo_reconstruction_cv_b200 import disparity_tab as dt
    l, r = load_pair("d1")
    g = FULL["rectified_d1"]
    K = np.array(g["K"])
    Lr, Rr = dt.rectify_pair(l, r, K, K, np.array(g["R1"]), np.array(g["R2"]), np.array(g["P1"]), np.array(g["P2"]))
    assert sha(Lr) == g["left_rect_sha256"] and sha(Rr) == g["right_rect_sha256"]
    disp0 = None
    for mode in (0, 1, 2):
        st = sg.StereoSGBM_create(minDisparity=0, numDisparities=g["numDisparities"], mode=mode, **NOTEBOOK)
        got = st.compute(Lr, Rr)
        assert sha(got) == g["modes"]["m%d" % mode]["disp_sha256"], mode
        if mode == 0:
            disp0 = got
    f = notebook_tail_numpy(disp0)
    Q = np.array(g["Q"])
    pts = sg.reprojectImageTo3D(f, Q)
    assert sha(pts) == g["cloud"]["xyz_sha256"]
    mask = ~np.isnan(pts[:, :, 0]) & ~np.isinf(pts[:, :, 0]) & (f > 0)
    assert int(mask.sum()) == g["cloud"]["n_points"] and sha(pts[mask]) == g["cloud"]["points_sha256"]


@pytest.mark.gpu
def test_cfg5_cloud_full_size(sg):
    """BASELINE configs[4]: 3WAY 3840x2160 D=256, then XYZ / point count / compacted cloud with the notebook's Q."""
    import torch
    g = FULL["cfg5_cloud"]
    l, r, _ = make_pair(g["W"], g["H"], g["D"], seed=g["seed"])
    assert sha(l) == g["left_sha256"] and sha(r) == g["right_sha256"], "synthetic generator drifted: regenerate the digests"
    st = sg.StereoSGBM_create(minDisparity=0, numDisparities=g["D"], blockSize=5, P1=200, P2=800, disp12MaxDiff=1,
                              preFilterCap=63, uniquenessRatio=10, speckleWindowSize=100, speckleRange=32, mode=2)
    di = st.compute(torch.from_numpy(l).cuda(), torch.from_numpy(r).cuda())
    assert sha(di.cpu().numpy()) == g["disp_sha256"]
    f = sg.disparityToFloat(di)
    xyz = sg.reprojectImageTo3D(f, NOTEBOOK_Q)
    assert sha(xyz.cpu().numpy()) == g["xyz_sha256"]
    col = torch.from_numpy(np.stack([l, r, l], -1)).cuda()
    pts, rgb = sg.reprojectCompact(di, NOTEBOOK_Q, col)
    assert pts.shape[0] == g["n_points"]
    assert sha(pts.cpu().numpy()) == g["points_sha256"] and sha(rgb.cpu().numpy()) == g["rgb_sha256"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["cfg3_3840x2160_D256_HH", "cfg4_1920x1080_D192_SGBM_seed0", "cfg2_1280x720_D128_SGBM"])
def test_forced_saturating_sweep_full_size(sg, golden_meta, monkeypatch, name):
    """k_sweep<.,.,SAT=1,.> (saturating S accumulation, normally selected only by large blockSize / P2) must
    give the same bits as the plain-add instantiation wherever the latter is exact: forced on the digests."""
    g = golden_meta["digests"][name]
    l, r, _ = make_pair(g["W"], g["H"], g["D"], seed=g["seed"])
    assert sha(l) == g["left_sha256"] and sha(r) == g["right_sha256"], "synthetic generator drifted: regenerate the digests"
    monkeypatch.setenv("SGBM_SWEEP_SAT", "1")
    st = sg.StereoSGBM_create(minDisparity=0, numDisparities=g["D"], blockSize=5, P1=200, P2=800, disp12MaxDiff=1,
                              preFilterCap=63, uniquenessRatio=10, speckleWindowSize=100, speckleRange=32, mode=g["mode"])
    assert sha(st.compute(l, r)) == g["disp_sha256"]


@pytest.mark.gpu
def test_reproject_missing_values_and_ddepth(sg):
    """cv2.reprojectImageTo3D(handleMissingValues=True / ddepth=CV_16S, CV_32S) (A.8) against live cv2."""
    import cv2
    rng = np.random.default_rng(0)
    Qg = rng.normal(size=(4, 4))
    big = np.array([[1e9, 0, 0, 0.5], [0, 1, 0, 0.5], [0, 0, 3e9, 2.5], [0, 0, 0, 1]], np.float64)
    H, W = 37, 53
    d = (rng.integers(-16, 600, (H, W)) / 16).astype(np.float32)
    d[rng.random((H, W)) < 0.2] = d.min()
    d[0, :8] = [d.min(), d.min() + np.float32(1.0e-7), d.min() + np.float32(1.19e-7), d.min() + np.float32(2.4e-7), 1, 2, 3, 4]
    for Q in (NOTEBOOK_Q, Qg, big):
        for dd in (d, (d * 16).astype(np.int16), np.clip(d, 0, 255).astype(np.uint8), (d * 16).astype(np.int32)):
            for hmv in (False, True):
                for ddepth in (-1, cv2.CV_16S, cv2.CV_32S, cv2.CV_32F):
                    ref = cv2.reprojectImageTo3D(dd, Q, handleMissingValues=hmv, ddepth=ddepth)
                    got = sg.reprojectImageTo3D(dd, Q, handleMissingValues=hmv, ddepth=ddepth)
                    assert got.dtype == ref.dtype and got.shape == ref.shape
                    assert np.array_equal(got.view(np.uint8), ref.view(np.uint8)), (dd.dtype, hmv, ddepth)
    with pytest.raises(sg.error):
        sg.reprojectImageTo3D(d, NOTEBOOK_Q, ddepth=cv2.CV_64F)
