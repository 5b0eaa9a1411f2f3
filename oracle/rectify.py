"""CPU restatement (numpy) of the rectification warp in front of the dense-stereo path --
TEST INFRASTRUCTURE, not a fallback.

    cv2.initUndistortRectifyMap(K, None, R, P, size, cv2.CV_32F)   main.ipynb:496-497, gui.py:160-161
    cv2.remap(img, map1, map2, interpolation=cv2.INTER_LINEAR)      main.ipynb:499-500, gui.py:163-164

The arithmetic lives in third-party OpenCV (modules/calib3d undistort + modules/imgproc remap; pinned
at opencv-python==4.11.0.86 by environment.yml:89, 4.13.0.92 installed here), not under
/root/reference.  Both functions were pinned against the installed cv2 binary (0 mismatches, see
tests/golden/make_golden_rectify.py and tests/test_rectify.py::test_oracle_vs_live_cv2).
"""
import numpy as np


def _inv3(S):
    """cv::invert for 3x3 doubles: determinant + adjugate in OpenCV's association order."""
    d = S[0, 0] * (S[1, 1] * S[2, 2] - S[1, 2] * S[2, 1]) - S[0, 1] * (S[1, 0] * S[2, 2] - S[1, 2] * S[2, 0]) \
        + S[0, 2] * (S[1, 0] * S[2, 1] - S[1, 1] * S[2, 0])
    d = 1.0 / d
    t = np.empty(9)
    t[0] = (S[1, 1] * S[2, 2] - S[1, 2] * S[2, 1]) * d
    t[1] = (S[0, 2] * S[2, 1] - S[0, 1] * S[2, 2]) * d
    t[2] = (S[0, 1] * S[1, 2] - S[0, 2] * S[1, 1]) * d
    t[3] = (S[1, 2] * S[2, 0] - S[1, 0] * S[2, 2]) * d
    t[4] = (S[0, 0] * S[2, 2] - S[0, 2] * S[2, 0]) * d
    t[5] = (S[0, 2] * S[1, 0] - S[0, 0] * S[1, 2]) * d
    t[6] = (S[1, 0] * S[2, 1] - S[1, 1] * S[2, 0]) * d
    t[7] = (S[0, 1] * S[2, 0] - S[0, 0] * S[2, 1]) * d
    t[8] = (S[0, 0] * S[1, 1] - S[0, 1] * S[1, 0]) * d
    return t


def init_undistort_rectify_map(K, R, P, size, lanes=8):
    """(map1, map2) float32 HxW for zero distortion.  `lanes` is the SIMD width of the reference's
    row loop: 8 doubles (AVX-512 dispatch) on the build and GPU-box hosts."""
    K = np.asarray(K, np.float64)
    P = np.asarray(P, np.float64)[:3, :3]
    W, H = size
    if R is None:
        A = P.copy()
    else:
        R = np.asarray(R, np.float64)
        A = np.zeros((3, 3))
        for i in range(3):
            for j in range(3):
                s = 0.0
                for k in range(3):
                    s += P[i, k] * R[k, j]
                A[i, j] = s
    ir = _inv3(A)
    fx, fy, u0, v0 = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    i = np.arange(H, dtype=np.float64)
    nb = W // lanes

    def walk(a0, step):
        M = np.empty((H, nb + 1))
        M[:, 0] = a0
        M[:, 1:] = lanes * step
        B = np.add.accumulate(M, axis=1)                         # block bases, accumulated sequentially
        full = (B[:, :nb, None] + (step * np.arange(lanes, dtype=np.float64))[None, None, :]).reshape(H, nb * lanes)
        tail = W - nb * lanes
        if tail:
            T = np.empty((H, tail))
            T[:, 0] = B[:, nb]
            T[:, 1:] = step
            full = np.concatenate([full, np.add.accumulate(T, axis=1)], axis=1)
        return full

    xs = walk(i * ir[1] + ir[2], ir[0])
    ys = walk(i * ir[4] + ir[5], ir[3])
    ws = walk(i * ir[7] + ir[8], ir[6])
    w = 1.0 / ws
    return (fx * (xs * w) + u0).astype(np.float32), (fy * (ys * w) + v0).astype(np.float32)


def remap_linear(src, map1, map2):
    """cv2.remap(src, map1, map2, INTER_LINEAR) for uint8 HxW or HxWxC, BORDER_CONSTANT 0."""
    src = np.asarray(src)
    sh, sw = src.shape[:2]
    sx = np.rint(map1.astype(np.float32) * np.float32(32)).astype(np.int64)      # cvRound: ties to even
    sy = np.rint(map2.astype(np.float32) * np.float32(32)).astype(np.int64)
    fx, fy = sx & 31, sy & 31
    ix = np.clip(sx >> 5, -32768, 32767)
    iy = np.clip(sy >> 5, -32768, 32767)
    w00 = (32 - fx) * (32 - fy) * 32
    w01 = fx * (32 - fy) * 32
    w10 = (32 - fx) * fy * 32
    w11 = fx * fy * 32
    q = (fx == 0) & (fy == 0)                                     # saturate_cast<short>(32768) + sum fix-up
    w00 = np.where(q, 32767, w00)
    w11 = np.where(q, 1, w11)

    def tap(yy, xx):
        ok = (xx >= 0) & (xx < sw) & (yy >= 0) & (yy < sh)
        v = src[np.clip(yy, 0, sh - 1), np.clip(xx, 0, sw - 1)].astype(np.int64)
        return np.where(ok[..., None] if src.ndim == 3 else ok, v, 0)

    def wx(w):
        return w[..., None] if src.ndim == 3 else w

    acc = tap(iy, ix) * wx(w00) + tap(iy, ix + 1) * wx(w01) + tap(iy + 1, ix) * wx(w10) + tap(iy + 1, ix + 1) * wx(w11)
    return np.clip((acc + 16384) >> 15, 0, 255).astype(np.uint8)
