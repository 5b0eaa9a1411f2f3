"""the notebook's dense-reconstruction functions / the README's "Tab 6" on top of the engine."""
import numpy as np
import pytest

import oracle
from oracle import OracleParams
from stereo_reconstruction_cv_b200.synth import make_pair


def test_point_cloud_arrays_is_the_notebook_mask():
    from stereo_reconstruction_cv_b200 import disparity_tab as dt
    rng = np.random.default_rng(3)
    pts = rng.normal(size=(6, 7, 3)).astype(np.float32)
    pts[1, 2, 0] = np.inf; pts[3, 3, 0] = np.nan; pts[4, 1, 1] = np.inf          # only X is tested (main.ipynb:727-731)
    disp = rng.uniform(-1, 3, size=(6, 7)).astype(np.float32)
    col = rng.integers(0, 256, size=(6, 7, 3), dtype=np.uint8)
    vp, vc = dt.point_cloud_arrays(pts, col, disp)
    mask = ~np.isnan(pts[:, :, 0]) & ~np.isinf(pts[:, :, 0]) & (disp > 0)
    assert np.array_equal(vp, pts[mask], equal_nan=True) and np.array_equal(vc, col[mask])
    assert dt.NOTEBOOK_PARAMS["blockSize"] == 11 and dt.NOTEBOOK_PARAMS["P2"] == 32 * 3 * 121
    assert callable(dt.create_disparity_tab) and callable(dt.rectify_pair)


@pytest.mark.gpu
def test_notebook_functions_on_gpu(tmp_path):
    from stereo_reconstruction_cv_b200 import disparity_tab as dt
    W, H, D = 360, 120, 16
    l, r, _ = make_pair(W, H, D, seed=4)
    d = dt.compute_disparity_map(l, r, D, 0)                                   # main.ipynb:781 calls it with (16, 0)
    p = OracleParams(0, D, 11, 8 * 3 * 121, 32 * 3 * 121, 1, 63, 10, 100, 32, 0)
    ref = oracle.compute(p, l, r).astype(np.float32) / 16.0
    ref = ref * (ref > 0).astype(np.float32)
    assert d.dtype == np.float32 and np.array_equal(d, ref)
    Q = np.array([[1, 0, 0, -W / 2], [0, 1, 0, -H / 2], [0, 0, 0, 300.0], [0, 0, -1, 0]], np.float64)
    pts = dt.reconstruct_3D(d, Q)
    assert pts.shape == (H, W, 3) and np.array_equal(pts.view(np.uint32), oracle.reproject_f32(d, Q).view(np.uint32))
    col = np.repeat(l[:, :, None], 3, 2)
    vp, vc = dt.point_cloud_arrays(pts, col, d)
    assert len(vp) == int((np.isfinite(pts[:, :, 0]) & (d > 0)).sum()) > 0
    n = dt.save_point_cloud(str(tmp_path / "c.ply"), pts)
    assert n == W * H                                                          # the notebook writes every pixel
    assert dt.reconstruct_3D(d, np.eye(3)) is None                             # Q must be 4x4: error -> None
