"""Host-side mirror of the cv2 API the reference calls on its dense-reconstruction path.

    cv2.StereoSGBM_create(...)            main.ipynb:655-666  ->  StereoSGBM_create(...)
    stereo.compute(imgL, imgR)            main.ipynb:668      ->  StereoSGBM.compute(left, right)
    cv2.reprojectImageTo3D(disparity, Q)  main.ipynb:697      ->  reprojectImageTo3D(disparity, Q)
    cv2.filterSpeckles / cv2.medianBlur   (stages of compute) ->  filterSpeckles / medianBlur3

Same names, argument order, defaults and result dtypes as the cv2 binding, so the notebook's
compute_disparity_map / reconstruct_3D run unchanged with `cv2` replaced by this module.  numpy
arrays go through the C ABI's host entry point (pinned staging + H2D/D2H); torch CUDA tensors are
passed as device pointers on the current stream and the result stays on the device.
All arithmetic runs in libsgbm_b200.so (CUDA, sm_100a); there is no CPU fallback.
"""
import ctypes as C

import numpy as np

from . import _hostpool, _lib
from ._lib import SgbmParams, check, error  # noqa: F401

MODE_SGBM = 0
MODE_HH = 1
MODE_SGBM_3WAY = 2
MODE_HH4 = 3
DISP_SHIFT = 4
DISP_SCALE = 16
INTER_LINEAR = 1          # cv2.INTER_LINEAR
CV_32F = 5                # cv2.CV_32F / CV_32FC1 (m1type of initUndistortRectifyMap)
CV_32FC1 = 5

_PARAM_NAMES = ("minDisparity", "numDisparities", "blockSize", "P1", "P2", "disp12MaxDiff",
                "preFilterCap", "uniquenessRatio", "speckleWindowSize", "speckleRange", "mode")


def _torch():
    import torch
    return torch


def _is_tensor(x):
    return type(x).__module__.startswith("torch")


def _stream_ptr(device):
    torch = _torch()
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class StereoSGBM:
    """Drop-in for the object returned by cv2.StereoSGBM_create (main.ipynb:655)."""

    def __init__(self, **kw):
        self._p = SgbmParams(*[int(kw.get(n, d)) for n, d in zip(_PARAM_NAMES, (0, 16, 3, 0, 0, 0, 0, 0, 0, 0, 0))])
        self._h = C.c_void_p()
        check(_lib.lib().sgbm_create(C.byref(self._p), C.byref(self._h)))

    def __del__(self):
        try:
            if getattr(self, "_h", None) is not None and self._h.value:
                _lib.lib().sgbm_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    # -- cv2-style accessors ----------------------------------------------------------------------
    def _set(self, name, v):
        setattr(self._p, name, int(v))
        check(_lib.lib().sgbm_set_params(self._h, C.byref(self._p)))

    def getMinDisparity(self): return self._p.minDisparity
    def setMinDisparity(self, v): self._set("minDisparity", v)
    def getNumDisparities(self): return self._p.numDisparities
    def setNumDisparities(self, v): self._set("numDisparities", v)
    def getBlockSize(self): return self._p.blockSize
    def setBlockSize(self, v): self._set("blockSize", v)
    def getP1(self): return self._p.P1
    def setP1(self, v): self._set("P1", v)
    def getP2(self): return self._p.P2
    def setP2(self, v): self._set("P2", v)
    def getDisp12MaxDiff(self): return self._p.disp12MaxDiff
    def setDisp12MaxDiff(self, v): self._set("disp12MaxDiff", v)
    def getPreFilterCap(self): return self._p.preFilterCap
    def setPreFilterCap(self, v): self._set("preFilterCap", v)
    def getUniquenessRatio(self): return self._p.uniquenessRatio
    def setUniquenessRatio(self, v): self._set("uniquenessRatio", v)
    def getSpeckleWindowSize(self): return self._p.speckleWindowSize
    def setSpeckleWindowSize(self, v): self._set("speckleWindowSize", v)
    def getSpeckleRange(self): return self._p.speckleRange
    def setSpeckleRange(self, v): self._set("speckleRange", v)
    def getMode(self): return self._p.mode
    def setMode(self, v): self._set("mode", v)

    def workspaceBytes(self, W, H, channels=1, batch=1):
        out = C.c_size_t()
        check(_lib.lib().sgbm_workspace_bytes(self._h, W, H, channels, batch, C.byref(out)))
        return out.value

    # -- compute -----------------------------------------------------------------------------------
    def compute(self, left, right, disparity=None):
        """int16 disparity x16 (invalid = (minDisparity-1)*16), like cv2.StereoSGBM.compute.

        numpy uint8 HxW or HxWx3 -> new C-contiguous numpy int16 HxW.
        torch.uint8 CUDA tensors (H,W), (H,W,3), (B,H,W) or (B,H,W,3 with channels_last=True)
        -> torch.int16 CUDA tensor of shape (H,W) / (B,H,W)."""
        if _is_tensor(left):
            return self._compute_torch(left, right, disparity)
        left = np.asarray(left)
        right = np.asarray(right)
        if left.dtype != np.uint8 or right.dtype != np.uint8:
            raise error(-1, "left/right must be uint8 (cv2: stereosgbm.cpp:2213 assertion)")
        if left.shape != right.shape or left.ndim not in (2, 3):
            raise error(-1, "left and right must have the same HxW[xC] shape")
        cn = 1 if left.ndim == 2 else left.shape[2]
        if cn == 1 and left.ndim == 3:
            left, right = left[:, :, 0], right[:, :, 0]
        # accept non-contiguous views the way cv2 does: rows must be dense, pitch may be anything
        if left.strides[-1] != 1 or (left.ndim == 3 and left.strides[1] != cn):
            left = np.ascontiguousarray(left)
        if right.strides[-1] != 1 or (right.ndim == 3 and right.strides[1] != cn):
            right = np.ascontiguousarray(right)
        if left.strides[0] != right.strides[0] or left.strides[0] < left.shape[1] * cn:
            left, right = np.ascontiguousarray(left), np.ascontiguousarray(right)
        H, W = left.shape[:2]
        if (disparity is not None and isinstance(disparity, np.ndarray) and disparity.shape == (H, W)
                and disparity.dtype == np.int16 and disparity.strides[1] == 2 and disparity.strides[0] >= 2 * W):
            out = disparity                              # the caller's array is written directly
        else:
            out = _hostpool.empty((H, W), np.int16)        # page-locked: the D2H copy lands in the result itself
        check(_lib.lib().sgbm_compute_host(self._h, left.ctypes.data, right.ctypes.data, W, H, cn,
                                           left.strides[0], 1, out.ctypes.data, out.strides[0]))
        if disparity is not None and out is not disparity:
            disparity[...] = out
            return disparity
        return out

    def compute_batch(self, lefts, rights, disparity=None):
        """Disparity of B independent pairs: numpy uint8 (B,H,W) or (B,H,W,3) -> numpy int16 (B,H,W).

        One call into the C ABI's host entry point with batch = B: frames are double-buffered through
        pinned memory, so the host copies and the H2D / D2H transfers of neighbouring frames overlap the
        kernels (video / batch use; cv2 has no counterpart, the reference loops over compute())."""
        if _is_tensor(lefts):
            return self._compute_torch(lefts, rights, disparity)
        lefts = np.ascontiguousarray(lefts)
        rights = np.ascontiguousarray(rights)
        if lefts.dtype != np.uint8 or rights.dtype != np.uint8 or lefts.shape != rights.shape or lefts.ndim not in (3, 4):
            raise error(-1, "lefts/rights must be uint8 arrays of the same (B,H,W) or (B,H,W,3) shape")
        B, H, W = lefts.shape[:3]
        cn = 1 if lefts.ndim == 3 else lefts.shape[3]
        out = disparity if disparity is not None else _hostpool.empty((B, H, W), np.int16)
        if out.shape != (B, H, W) or out.dtype != np.int16 or not out.flags.c_contiguous:
            raise error(-1, "bad output array")
        check(_lib.lib().sgbm_compute_host(self._h, lefts.ctypes.data, rights.ctypes.data, W, H, cn, W * cn, B,
                                           out.ctypes.data, W * 2))
        return out

    def _compute_torch(self, left, right, disparity=None):
        torch = _torch()
        if not (left.is_cuda and right.is_cuda):
            raise error(-1, "torch inputs must be CUDA tensors (use numpy arrays for host data)")
        if left.dtype != torch.uint8 or right.dtype != torch.uint8 or left.shape != right.shape:
            raise error(-1, "left/right must be uint8 tensors of the same shape")
        left, right = left.contiguous(), right.contiguous()
        shp = tuple(left.shape)
        if left.dim() == 2:
            B, H, W, cn = 1, shp[0], shp[1], 1
        elif left.dim() == 3 and shp[2] == 3:                 # H x W x 3 colour image
            B, H, W, cn = 1, shp[0], shp[1], 3
        elif left.dim() == 3:
            B, H, W, cn = shp[0], shp[1], shp[2], 1
        elif left.dim() == 4 and shp[3] == 3:
            B, H, W, cn = shp[0], shp[1], shp[2], 3
        else:
            raise error(-1, "unsupported tensor shape %s" % (shp,))
        out_shape = (H, W) if (left.dim() == 2 or (left.dim() == 3 and cn == 3)) else (B, H, W)
        if disparity is None:
            disparity = torch.empty(out_shape, dtype=torch.int16, device=left.device)
        elif tuple(disparity.shape) != out_shape or disparity.dtype != torch.int16 or not disparity.is_contiguous():
            raise error(-1, "bad output tensor")
        with torch.cuda.device(left.device):
            check(_lib.lib().sgbm_compute(self._h, left.data_ptr(), right.data_ptr(), W, H, cn, W * cn, B,
                                          disparity.data_ptr(), W * 2, _stream_ptr(left.device)))
        return disparity

    # -- test hooks ----------------------------------------------------------------------------------
    def status(self):
        """Raise if a sweep of an already completed asynchronous compute() gave up (sgbm_status)."""
        check(_lib.lib().sgbm_status(self._h))

    def _debug_keep(self, on=True):
        check(_lib.lib().sgbm_debug_keep(self._h, 1 if on else 0))

    def _debug_fetch(self, which, shape):
        out = np.empty(shape, np.int16)
        check(_lib.lib().sgbm_debug_fetch(self._h, which, out.ctypes.data, out.nbytes))
        return out


def StereoSGBM_create(minDisparity=0, numDisparities=16, blockSize=3, P1=0, P2=0, disp12MaxDiff=0,
                      preFilterCap=0, uniquenessRatio=0, speckleWindowSize=0, speckleRange=0,
                      mode=MODE_SGBM):
    """cv2.StereoSGBM_create with the binding's defaults ([P15])."""
    return StereoSGBM(minDisparity=minDisparity, numDisparities=numDisparities, blockSize=blockSize, P1=P1,
                      P2=P2, disp12MaxDiff=disp12MaxDiff, preFilterCap=preFilterCap,
                      uniquenessRatio=uniquenessRatio, speckleWindowSize=speckleWindowSize,
                      speckleRange=speckleRange, mode=mode)


def _q16(Q):
    Q = np.ascontiguousarray(np.asarray(Q, dtype=np.float64))
    if Q.shape != (4, 4):
        raise error(-1, "Q must be 4x4 (cv2: stereo_geom.cpp:19)")
    return Q


CV_8U, CV_16S, CV_32S = 0, 3, 4    # cv2 depth codes (CV_32F = 5 above)


def reprojectImageTo3D(disparity, Q, _3dImage=None, handleMissingValues=False, ddepth=-1):
    """cv2.reprojectImageTo3D (main.ipynb:697): HxWx3, float32 unless ddepth asks for CV_16S / CV_32S.

    disparity: uint8 / int16 / int32 / float32, HxW; integer disparities are used as they are (no /16).
    handleMissingValues=True sets Z = 10000 where the disparity equals its minimum (A.8).  numpy in ->
    numpy out, CUDA tensor in -> CUDA tensor out."""
    Q = _q16(Q)
    L = _lib.lib()
    torch = _torch()
    if ddepth not in (-1, CV_16S, CV_32S, CV_32F):
        raise error(-1, "ddepth must be -1, CV_16S, CV_32S or CV_32F (cv2: stereo_geom.cpp)")
    if _is_tensor(disparity):
        d = disparity.contiguous()
        if not d.is_cuda or d.dim() != 2:
            raise error(-1, "disparity tensor must be a 2-D CUDA tensor")
        depth = {torch.uint8: CV_8U, torch.int16: CV_16S, torch.int32: CV_32S, torch.float32: CV_32F}.get(d.dtype)
        if depth is None:
            raise error(-1, "disparity must be uint8, int16, int32 or float32 (cv2: stereo_geom.cpp:17)")
        H, W = d.shape
        odt = {CV_16S: torch.int16, CV_32S: torch.int32}.get(ddepth, torch.float32)
        out = torch.empty((H, W, 3), dtype=odt, device=d.device)
        with torch.cuda.device(d.device):
            if not handleMissingValues and odt is torch.float32 and depth in (CV_16S, CV_32F):
                fn = L.sgbm_reproject_f32 if depth == CV_32F else L.sgbm_reproject_i16
                check(fn(d.data_ptr(), Q.ctypes.data, W, H, out.data_ptr(), None, _stream_ptr(d.device)))
            else:
                scratch = torch.empty((16,), dtype=torch.uint8, device=d.device)
                check(L.sgbm_reproject_ex(d.data_ptr(), depth, Q.ctypes.data, W, H, 1 if handleMissingValues else 0,
                                          ddepth, out.data_ptr(), scratch.data_ptr(), _stream_ptr(d.device)))
        return out
    d = np.asarray(disparity)
    if d.ndim != 2:
        raise error(-1, "disparity must be HxW")
    if d.dtype == np.float64:
        raise error(-1, "float64 disparity is not accepted (cv2: stereo_geom.cpp:17)")
    if d.dtype not in (np.float32, np.int16, np.uint8, np.int32):
        raise error(-1, "unsupported disparity dtype %s" % d.dtype)
    dt = torch.from_numpy(np.ascontiguousarray(d)).cuda()
    dev = reprojectImageTo3D(dt, Q, None, handleMissingValues, ddepth)
    # the result (99.5 MB at 3840x2160) is read into a page-locked array from the recycling pool
    res = _hostpool.empty(tuple(dev.shape), {torch.int16: np.int16, torch.int32: np.int32}.get(dev.dtype, np.float32))
    torch.from_numpy(res).copy_(dev, non_blocking=True)
    torch.cuda.current_stream(dev.device).synchronize()
    if _3dImage is not None:
        _3dImage[...] = res
        return _3dImage
    return res


def disparityToFloat(disp_x16):
    """disp.astype(float32)/16 followed by the positivity mask of main.ipynb:668-670 (CUDA tensor in/out)."""
    torch = _torch()
    d = disp_x16.contiguous()
    out = torch.empty(d.shape, dtype=torch.float32, device=d.device)
    H = int(np.prod(d.shape[:-1]))
    with torch.cuda.device(d.device):
        check(_lib.lib().sgbm_disp_to_float(d.data_ptr(), d.shape[-1], H, out.data_ptr(), _stream_ptr(d.device)))
    return out


_PINNED = {}


def _pinned(shape, dtype):
    """Cached pinned host buffer (one per shape/dtype) for device -> host reads of point clouds."""
    torch = _torch()
    key = (tuple(shape), dtype)
    buf = _PINNED.get(key)
    if buf is None:
        buf = torch.empty(shape, dtype=dtype).pin_memory()
        _PINNED[key] = buf
    return buf


def reprojectCompact(disp_x16, Q, colors_bgr=None, to_host=False):
    """Fused tail of the notebook (main.ipynb:668-670, 697, 726-737) on the device.

    disp_x16: int16 CUDA tensor HxW (output of compute).  Returns (xyz float32 Nx3, rgb uint8 Nx3 or None):
    the points with finite X and disparity > 0 in row-major pixel order, colours swapped BGR->RGB.
    to_host=True returns numpy views of cached pinned buffers instead (valid until the next call)."""
    torch = _torch()
    L = _lib.lib()
    Q = _q16(Q)
    d = disp_x16.contiguous()
    H, W = d.shape
    dev = d.device
    xyz = torch.empty((H * W, 3), dtype=torch.float32, device=dev)
    rgb = None
    bgr_ptr, bgr_cn, bgr_pitch = None, 0, 0
    if colors_bgr is not None:
        cb = colors_bgr.contiguous()
        bgr_cn = 1 if cb.dim() == 2 else int(cb.shape[2])
        bgr_ptr, bgr_pitch = cb.data_ptr(), W * bgr_cn
        rgb = torch.empty((H * W, 3), dtype=torch.uint8, device=dev)
    nbytes = C.c_size_t()
    check(L.sgbm_reproject_compact_scratch_bytes(W, H, C.byref(nbytes)))
    scratch = torch.empty((nbytes.value,), dtype=torch.uint8, device=dev)
    n = torch.zeros((1,), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        check(L.sgbm_reproject_compact(d.data_ptr(), Q.ctypes.data, W, H, bgr_ptr, bgr_cn, bgr_pitch, xyz.data_ptr(),
                                       rgb.data_ptr() if rgb is not None else None, n.data_ptr(), scratch.data_ptr(),
                                       nbytes.value, _stream_ptr(dev)))
    cnt = int(n.item())
    if to_host:
        hx = _pinned((H * W, 3), torch.float32)
        hx[:cnt].copy_(xyz[:cnt], non_blocking=True)
        hc = None
        if rgb is not None:
            hc = _pinned((H * W, 3), torch.uint8)
            hc[:cnt].copy_(rgb[:cnt], non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        return hx[:cnt].numpy(), (hc[:cnt].numpy() if hc is not None else None)
    return xyz[:cnt], (rgb[:cnt] if rgb is not None else None)


def filterSpeckles(img, newVal, maxSpeckleSize, maxDiff, buf=None):
    """cv2.filterSpeckles: in place on an int16 image; returns (img, buf) like the cv2 binding."""
    torch = _torch()
    L = _lib.lib()
    if _is_tensor(img):
        if img.dtype != torch.int16 or not img.is_cuda or not img.is_contiguous() or img.dim() != 2:
            raise error(-1, "img must be a contiguous 2-D int16 CUDA tensor")
        H, W = img.shape
        scratch = torch.empty((H * W * 8,), dtype=torch.uint8, device=img.device)
        with torch.cuda.device(img.device):
            check(L.sgbm_filter_speckles(img.data_ptr(), W, H, int(newVal), int(maxSpeckleSize), int(maxDiff),
                                         scratch.data_ptr(), H * W * 8, _stream_ptr(img.device)))
        return img, buf
    a = np.asarray(img)
    if a.dtype != np.int16 or a.ndim != 2:
        raise error(-1, "img must be a 2-D int16 array")
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    filterSpeckles(t, newVal, maxSpeckleSize, maxDiff)
    img[...] = t.cpu().numpy()
    return img, buf


def medianBlur3(img):
    """cv2.medianBlur(img, 3) for int16 images (the always-on post filter of compute, A.7)."""
    torch = _torch()
    host = not _is_tensor(img)
    t = torch.from_numpy(np.ascontiguousarray(np.asarray(img, np.int16))).cuda() if host else img.contiguous()
    H, W = t.shape
    out = torch.empty_like(t)
    with torch.cuda.device(t.device):
        check(_lib.lib().sgbm_median3x3(t.data_ptr(), out.data_ptr(), W, H, _stream_ptr(t.device)))
    return out.cpu().numpy() if host else out


def initUndistortRectifyMap(cameraMatrix, distCoeffs, R, newCameraMatrix, size, m1type=CV_32F, device=None):
    """cv2.initUndistortRectifyMap(K, None, R, P, (W, H), cv2.CV_32F) (main.ipynb:496-497, gui.py:160-161).

    Returns (map1, map2): float32 HxW x- and y-coordinate maps as torch CUDA tensors (they feed remap on
    the device; call .cpu().numpy() for the arrays cv2 returns).  distCoeffs must be None or all zero,
    m1type CV_32F / CV_32FC1 -- the forms the reference uses."""
    torch = _torch()
    if m1type not in (CV_32F,):
        raise error(-3, "only m1type=CV_32F (CV_32FC1) is implemented")
    K = np.ascontiguousarray(np.asarray(cameraMatrix, np.float64))
    P = np.ascontiguousarray(np.asarray(newCameraMatrix, np.float64))
    if K.shape != (3, 3) or P.shape not in ((3, 3), (3, 4)):
        raise error(-1, "cameraMatrix must be 3x3 and newCameraMatrix 3x3 or 3x4")
    Rm = None if R is None or np.size(R) == 0 else np.ascontiguousarray(np.asarray(R, np.float64))
    if Rm is not None and Rm.shape != (3, 3):
        raise error(-1, "R must be 3x3")
    dist = None if distCoeffs is None else np.ascontiguousarray(np.asarray(distCoeffs, np.float64).ravel())
    W, H = int(size[0]), int(size[1])
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    m1 = torch.empty((H, W), dtype=torch.float32, device=dev)
    m2 = torch.empty((H, W), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.lib().sgbm_init_rectify_map(K.ctypes.data, dist.ctypes.data if dist is not None else None,
                                               0 if dist is None else int(dist.size), Rm.ctypes.data if Rm is not None else None,
                                               P.ctypes.data, int(P.shape[1]), W, H, m1.data_ptr(), m2.data_ptr(),
                                               _stream_ptr(dev)))
    return m1, m2


def remap(src, map1, map2, interpolation=INTER_LINEAR, dst=None):
    """cv2.remap(src, map1, map2, interpolation=cv2.INTER_LINEAR) for uint8 images (main.ipynb:499-500):
    bilinear in 1/32-pixel fixed point, constant border 0, bit-identical to cv2.  numpy in -> numpy out,
    CUDA tensors in -> CUDA tensor out; the maps may be numpy float32 arrays or CUDA tensors."""
    torch = _torch()
    if interpolation != INTER_LINEAR:
        raise error(-3, "only INTER_LINEAR is implemented")
    host = not _is_tensor(src)
    dev = (map1.device if _is_tensor(map1) and map1.is_cuda else torch.device("cuda", torch.cuda.current_device())) if host else src.device
    s = torch.from_numpy(np.ascontiguousarray(src)).to(dev) if host else src.contiguous()
    if s.dtype != torch.uint8 or s.dim() not in (2, 3):
        raise error(-1, "src must be a uint8 HxW or HxWxC image")
    cn = 1 if s.dim() == 2 else int(s.shape[2])
    m1 = (map1 if _is_tensor(map1) else torch.from_numpy(np.ascontiguousarray(map1, np.float32))).to(dev).contiguous()
    m2 = (map2 if _is_tensor(map2) else torch.from_numpy(np.ascontiguousarray(map2, np.float32))).to(dev).contiguous()
    if m1.dtype != torch.float32 or m2.dtype != torch.float32 or m1.shape != m2.shape or m1.dim() != 2:
        raise error(-1, "map1/map2 must be float32 HxW arrays of the same shape")
    H, W = m1.shape
    sh, sw = s.shape[:2]
    out = torch.empty((H, W) if cn == 1 and s.dim() == 2 else (H, W, cn), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(_lib.lib().sgbm_remap_linear_u8(s.data_ptr(), sw, sh, cn, sw * cn, m1.data_ptr(), m2.data_ptr(), W, H,
                                              out.data_ptr(), W * cn, _stream_ptr(dev)))
    if host:
        res = out.cpu().numpy()
        if dst is not None:
            dst[...] = res
            return dst
        return res
    return out


def microbench_int16(which):
    """Measured issue rate (G lane-ops/s) of the packed 16-bit integer instructions (roofline denominator)."""
    v = C.c_double()
    check(_lib.lib().sgbm_microbench_int16(int(which), C.byref(v)))
    return v.value


def device_info():
    L = _lib.lib()
    sm, ma, mi = C.c_int(), C.c_int(), C.c_int()
    name = C.create_string_buffer(128)
    check(L.sgbm_device_info(C.byref(sm), C.byref(ma), C.byref(mi), name, 128))
    return {"sm_count": sm.value, "cc": "%d.%d" % (ma.value, mi.value), "name": name.value.decode()}
