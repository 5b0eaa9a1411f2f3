"""ctypes binding of oracle/liboracle.so (the C port of SURVEY.md Appendix A).  Test infra only."""
import ctypes as C
import os
import subprocess
from dataclasses import dataclass, fields

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    """Compile liboracle.so with gcc (seconds)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "sgbm_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "liboracle.so"])
    return so


@dataclass
class OracleParams:
    minDisparity: int = 0
    numDisparities: int = 16
    blockSize: int = 3
    P1: int = 0
    P2: int = 0
    disp12MaxDiff: int = 0
    preFilterCap: int = 0
    uniquenessRatio: int = 0
    speckleWindowSize: int = 0
    speckleRange: int = 0
    mode: int = 0


class _CParams(C.Structure):
    _fields_ = [(f.name, C.c_int) for f in fields(OracleParams)]


def _lib():
    global _LIB
    if _LIB is None:
        lib = C.CDLL(build())
        lib.oracle_sgbm_compute.restype = C.c_int
        lib.oracle_sgbm_compute.argtypes = [C.POINTER(_CParams), C.c_void_p, C.c_void_p, C.c_int,
                                            C.c_int, C.c_int, C.c_ssize_t, C.c_ssize_t, C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.c_void_p]
        lib.oracle_median3x3_i16.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        lib.oracle_filter_speckles_i16.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        lib.oracle_reproject_f32.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        _LIB = lib
    return _LIB


def _cparams(p):
    return _CParams(*[int(getattr(p, f.name)) for f in fields(OracleParams)])


def _prep(img):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    if img.ndim == 2:
        return img, 1
    return img, img.shape[2]


def valid_width(p, W):
    maxD = p.minDisparity + p.numDisparities
    return W + min(p.minDisparity, 0) - max(maxD, 0)


def compute_debug(p, left, right, want_C=False, want_S=False, want_raw=False):
    """Returns dict(disp=int16 HxW, [C], [S], [raw])."""
    left, cn = _prep(left)
    right, cn2 = _prep(right)
    if left.shape != right.shape:
        raise ValueError("size mismatch")
    H, W = left.shape[:2]
    disp = np.empty((H, W), np.int16)
    W1 = valid_width(p, W)
    D = p.numDisparities
    outs = {}
    ptrs = []
    for name, want, shape in (("C", want_C, (H, max(W1, 0), D)), ("S", want_S, (H, max(W1, 0), D)),
                              ("raw", want_raw, (H, W))):
        if want:
            outs[name] = np.zeros(shape, np.int16)
            ptrs.append(outs[name].ctypes.data)
        else:
            ptrs.append(None)
    cp = _cparams(p)
    rc = _lib().oracle_sgbm_compute(C.byref(cp), left.ctypes.data, right.ctypes.data, W, H, cn,
                                    left.strides[0], right.strides[0], disp.ctypes.data, *ptrs)
    if rc != 0:
        raise ValueError("oracle: invalid parameters/size (code %d)" % rc)
    outs["disp"] = disp
    return outs


def compute(p, left, right):
    return compute_debug(p, left, right)["disp"]


def median3x3(img):
    img = np.ascontiguousarray(img, np.int16)
    out = np.empty_like(img)
    _lib().oracle_median3x3_i16(img.ctypes.data, out.ctypes.data, img.shape[1], img.shape[0])
    return out


def filter_speckles(img, newVal, maxSpeckleSize, maxDiff):
    out = np.array(img, dtype=np.int16, order="C", copy=True)
    _lib().oracle_filter_speckles_i16(out.ctypes.data, out.shape[1], out.shape[0], int(newVal),
                                      int(maxSpeckleSize), int(maxDiff))
    return out


def reproject_f32(disp, Q):
    disp = np.ascontiguousarray(disp, np.float32)
    Q = np.ascontiguousarray(Q, np.float64)
    out = np.empty(disp.shape + (3,), np.float32)
    _lib().oracle_reproject_f32(disp.ctypes.data, disp.shape[1], disp.shape[0], Q.ctypes.data,
                                out.ctypes.data)
    return out
