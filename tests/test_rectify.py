"""Rectification warp: oracle/rectify.py against the cv2-generated golden vectors
(CPU), and the CUDA kernels through the C ABI against both (GPU).  Bit-exact: the maps as fp32 bit
patterns, the warped images as uint8."""
import os

import numpy as np
import pytest

from oracle import rectify as orc

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_rectify.npz"))
NAMES = [str(n) for n in GOLD["names"]]


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _R(name):
    R = GOLD[name + "_R"]
    return None if R.size == 0 else R


@pytest.mark.parametrize("name", NAMES)
def test_oracle_maps_vs_golden(name):
    size = tuple(int(v) for v in GOLD[name + "_size"])
    m1, m2 = orc.init_undistort_rectify_map(GOLD[name + "_K"], _R(name), GOLD[name + "_P"], size)
    assert np.array_equal(_bits(m1), _bits(GOLD[name + "_map1"]))
    assert np.array_equal(_bits(m2), _bits(GOLD[name + "_map2"]))


@pytest.mark.parametrize("name", NAMES + ["rand"])
def test_oracle_remap_vs_golden(name):
    mx, my = GOLD[name + "_mx"], GOLD[name + "_my"]
    assert np.array_equal(orc.remap_linear(GOLD["src_gray"], mx, my), GOLD[name + "_remap_gray"])
    assert np.array_equal(orc.remap_linear(GOLD["src_bgr"], mx, my), GOLD[name + "_remap_bgr"])


def test_oracle_vs_live_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    for it in range(4):
        K = np.array([[rng.uniform(800, 3000), 0, rng.uniform(300, 2000)], [0, rng.uniform(800, 3000), rng.uniform(300, 1200)],
                      [0, 0, 1]])
        R, _ = cv2.Rodrigues(rng.normal(0, 0.03, 3))
        P = np.array([[rng.uniform(800, 3000), 0, rng.uniform(300, 2000), 0], [0, rng.uniform(800, 3000), rng.uniform(300, 1200), 0],
                      [0, 0, 1, 0]])
        size = (int(rng.integers(100, 700)), int(rng.integers(60, 300)))
        m1, m2 = cv2.initUndistortRectifyMap(K, None, R, P, size, cv2.CV_32F)
        o1, o2 = orc.init_undistort_rectify_map(K, R, P, size)
        bad = int((_bits(o1) != _bits(m1)).sum() + (_bits(o2) != _bits(m2)).sum())
        # bit-identical where cv2 dispatches to its 8-lane (AVX-512) row loop; at most 1 ulp in a few entries elsewhere
        assert bad == 0 or (bad < 1e-4 * m1.size and float(np.abs(o1 - m1).max()) < 1e-3), bad
        src = rng.integers(0, 256, (int(rng.integers(80, 300)), int(rng.integers(80, 300))), dtype=np.uint8)
        mx = rng.uniform(-6, src.shape[1] + 6, (50, 70)).astype(np.float32)
        my = rng.uniform(-6, src.shape[0] + 6, (50, 70)).astype(np.float32)
        assert np.array_equal(orc.remap_linear(src, mx, my), cv2.remap(src, mx, my, interpolation=cv2.INTER_LINEAR))


# ------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def sg():
    import torch
    assert torch.cuda.is_available()
    import stereo_reconstruction_cv_b200 as sg
    return sg


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_maps(sg, name):
    size = tuple(int(v) for v in GOLD[name + "_size"])
    m1, m2 = sg.initUndistortRectifyMap(GOLD[name + "_K"], None, _R(name), GOLD[name + "_P"], size, sg.CV_32F)
    assert np.array_equal(_bits(m1.cpu().numpy()), _bits(GOLD[name + "_map1"]))
    assert np.array_equal(_bits(m2.cpu().numpy()), _bits(GOLD[name + "_map2"]))


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES + ["rand"])
def test_gpu_remap(sg, name):
    mx, my = GOLD[name + "_mx"], GOLD[name + "_my"]
    assert np.array_equal(sg.remap(GOLD["src_gray"], mx, my, interpolation=sg.INTER_LINEAR), GOLD[name + "_remap_gray"])
    assert np.array_equal(sg.remap(GOLD["src_bgr"], mx, my), GOLD[name + "_remap_bgr"])


@pytest.mark.gpu
def test_gpu_rectify_full_size_vs_oracle(sg):
    """3840x2160 (the dataset's size): maps and warped image against the oracle; device-resident chain."""
    import torch
    K = np.array([[1733.74, 0, 792.27], [0, 1733.74, 541.89], [0, 0, 1]])
    th = 0.01
    R = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1.0]])
    P = np.array([[1700.0, 0, 800.0, 0], [0, 1700.0, 540.0, 0], [0, 0, 1, 0]])
    size = (3840, 2160)
    m1, m2 = sg.initUndistortRectifyMap(K, None, R, P, size)
    o1, o2 = orc.init_undistort_rectify_map(K, R, P, size)
    assert np.array_equal(_bits(m1.cpu().numpy()), _bits(o1)) and np.array_equal(_bits(m2.cpu().numpy()), _bits(o2))
    img = np.random.default_rng(5).integers(0, 256, (2160, 3840), dtype=np.uint8)
    out = sg.remap(torch.from_numpy(img).cuda(), m1, m2)
    assert out.is_cuda and np.array_equal(out.cpu().numpy(), orc.remap_linear(img, o1, o2))
    with pytest.raises(sg.error):
        sg.initUndistortRectifyMap(K, np.array([0.1, 0, 0, 0, 0]), R, P, size)
