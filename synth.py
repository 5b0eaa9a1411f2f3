"""Synthetic rectified stereo pairs (SURVEY.md section 8(d) generator) -- input generator of bench.py, the tests, smoke() and
the tools; NOT part of the product package (it uses cv2 for the survey's exact recipe when cv2 is importable).

The reference has no generator of its own (it only ships three JPEG pairs, dataset/d1..d3), so
the survey defines one: a blurred-noise texture warped by a smooth ground-truth disparity field
with 8 constant rectangles, plus N(0,2) noise on the left image.  cv2 is used for the blur /
resize / remap steps when it is importable (the survey's exact recipe); otherwise an equivalent
scipy.ndimage recipe is used (same statistics, different bits) so that bench.py never depends on
cv2 being present.
"""
import numpy as np


def _have_cv2():
    try:
        import cv2  # noqa: F401
        return True
    except Exception:
        return False


def make_pair(W, H, D, seed=0, use_cv2=None):
    """Returns (left, right, gt) -- uint8 HxW, uint8 HxW, float32 HxW ground-truth disparity."""
    if use_cv2 is None:
        use_cv2 = _have_cv2()
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, (H, W + D), dtype=np.uint8)
    low = rng.uniform(1, D - 2, (max(H // 120, 2), max(W // 120, 2))).astype(np.float32)
    if use_cv2:
        import cv2
        tex = cv2.GaussianBlur(base, (0, 0), 1.5)
        tex = cv2.normalize(tex, None, 0, 255, cv2.NORM_MINMAX).astype(np.uint8)
        gt = cv2.resize(low, (W, H), interpolation=cv2.INTER_CUBIC)
    else:
        from scipy import ndimage
        t = ndimage.gaussian_filter(base.astype(np.float32), 1.5, mode="reflect")
        tex = np.clip(np.rint((t - t.min()) * (255.0 / max(float(t.max() - t.min()), 1e-6))), 0, 255).astype(np.uint8)
        gt = ndimage.zoom(low, (H / low.shape[0], W / low.shape[1]), order=3, mode="nearest")[:H, :W]
        gt = np.ascontiguousarray(gt, np.float32)
    gt = np.clip(gt, 1, D - 2).astype(np.float32)
    for _ in range(8):
        x0 = int(rng.integers(0, max(W - W // 8, 1)))
        y0 = int(rng.integers(0, max(H - H // 8, 1)))
        val = float(rng.uniform(1, D - 2))
        gt[y0:y0 + H // 8, x0:x0 + W // 8] = val
    right = np.ascontiguousarray(tex[:, D:D + W])
    xs = np.arange(W, dtype=np.float32)[None, :] + np.float32(D) - gt
    ys = np.broadcast_to(np.arange(H, dtype=np.float32)[:, None], (H, W)).copy()
    if use_cv2:
        import cv2
        warped = cv2.remap(tex, xs, ys, cv2.INTER_LINEAR).astype(np.float32)
    else:
        from scipy import ndimage
        warped = ndimage.map_coordinates(tex.astype(np.float32), [ys, xs], order=1, mode="nearest")
    noise = rng.normal(0.0, 2.0, (H, W)).astype(np.float32)
    left = np.clip(np.rint(warped + noise), 0, 255).astype(np.uint8)
    return left, right, gt


def make_noise_pair(W, H, seed=0):
    """Adversarial pair: independent uniform noise (exercises saturation / rejection / speckles)."""
    rng = np.random.default_rng(seed)
    return (rng.integers(0, 256, (H, W), dtype=np.uint8), rng.integers(0, 256, (H, W), dtype=np.uint8))
