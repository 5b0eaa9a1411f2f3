"""ctypes binding of libsgbm_b200.so (C ABI: include/sgbm_b200.h).  There is no CPU fallback:
if the library is missing or no CUDA device is present, the calls raise."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsgbm_b200.so")
# Same sources built with -DSGBM_DEBUG_HOOKS (hand-off fault injection): loaded by one test only, never by the package.
DEBUG_LIB_PATH = os.path.join(_HERE, "libsgbm_b200_dbg.so")
_LIB = None

# every symbol include/sgbm_b200.h declares
SYMBOLS = [
    "sgbm_last_error", "sgbm_version", "sgbm_device_info", "sgbm_create", "sgbm_destroy",
    "sgbm_set_params", "sgbm_get_params", "sgbm_workspace_bytes", "sgbm_compute", "sgbm_compute_host",
    "sgbm_disp_to_float", "sgbm_reproject_f32", "sgbm_reproject_i16", "sgbm_reproject_compact",
    "sgbm_reproject_compact_scratch_bytes", "sgbm_filter_speckles", "sgbm_median3x3",
    "sgbm_debug_keep", "sgbm_debug_fetch", "sgbm_microbench_int16", "sgbm_kernel_launches",
    "sgbm_profile_enable", "sgbm_profile_read", "sgbm_init_rectify_map", "sgbm_remap_linear_u8",
    "sgbm_status", "sgbm_reproject_ex", "sgbm_host_alloc", "sgbm_host_free", "sgbm_debug_sweep_plan",
]


class SgbmParams(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("minDisparity", "numDisparities", "blockSize", "P1", "P2",
                                       "disp12MaxDiff", "preFilterCap", "uniquenessRatio",
                                       "speckleWindowSize", "speckleRange", "mode")]


class error(Exception):
    """Raised where cv2 would raise cv2.error (bad size / type / parameter) and on CUDA errors."""

    def __init__(self, code, msg):
        super().__init__("sgbm_b200 error %d: %s" % (code, msg))
        self.code = code


def load(path):
    """dlopen a build of the library and declare the prototypes of include/sgbm_b200.h."""
    if not os.path.exists(path):
        raise ImportError("%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a). This engine has no CPU fallback." % path)
    L = C.CDLL(path)
    vp, i, sz, pd = C.c_void_p, C.c_int, C.c_size_t, C.c_ssize_t
    L.sgbm_last_error.restype = C.c_char_p
    L.sgbm_version.restype = C.c_char_p
    L.sgbm_device_info.argtypes = [C.POINTER(i), C.POINTER(i), C.POINTER(i), C.c_char_p, i]
    L.sgbm_create.argtypes = [C.POINTER(SgbmParams), C.POINTER(vp)]
    L.sgbm_destroy.argtypes = [vp]
    L.sgbm_set_params.argtypes = [vp, C.POINTER(SgbmParams)]
    L.sgbm_get_params.argtypes = [vp, C.POINTER(SgbmParams)]
    L.sgbm_workspace_bytes.argtypes = [vp, i, i, i, i, C.POINTER(sz)]
    L.sgbm_compute.argtypes = [vp, vp, vp, i, i, i, pd, i, vp, pd, vp]
    L.sgbm_compute_host.argtypes = [vp, vp, vp, i, i, i, pd, i, vp, pd]
    L.sgbm_disp_to_float.argtypes = [vp, i, i, vp, vp]
    L.sgbm_reproject_f32.argtypes = [vp, vp, i, i, vp, vp, vp]
    L.sgbm_reproject_i16.argtypes = [vp, vp, i, i, vp, vp, vp]
    L.sgbm_reproject_compact.argtypes = [vp, vp, i, i, vp, i, pd, vp, vp, vp, vp, sz, vp]
    L.sgbm_reproject_compact_scratch_bytes.argtypes = [i, i, C.POINTER(sz)]
    L.sgbm_filter_speckles.argtypes = [vp, i, i, i, i, i, vp, sz, vp]
    L.sgbm_median3x3.argtypes = [vp, vp, i, i, vp]
    L.sgbm_init_rectify_map.argtypes = [vp, vp, i, vp, vp, i, i, i, vp, vp, vp]
    L.sgbm_remap_linear_u8.argtypes = [vp, i, i, i, pd, vp, vp, i, i, vp, pd, vp]
    L.sgbm_status.argtypes = [vp]
    L.sgbm_reproject_ex.argtypes = [vp, i, vp, i, i, i, i, vp, vp, vp]
    L.sgbm_host_alloc.argtypes = [sz, C.POINTER(C.c_void_p)]
    L.sgbm_host_free.argtypes = [vp]
    L.sgbm_debug_sweep_plan.argtypes = [vp, i, i, i, i, i, i, i, vp]
    L.sgbm_kernel_launches.argtypes = []
    L.sgbm_kernel_launches.restype = C.c_ulonglong
    L.sgbm_debug_keep.argtypes = [vp, i]
    L.sgbm_debug_fetch.argtypes = [vp, i, vp, sz]
    L.sgbm_microbench_int16.argtypes = [i, C.POINTER(C.c_double)]
    for name in SYMBOLS:
        fn = getattr(L, name)
        if name not in ("sgbm_last_error", "sgbm_version", "sgbm_kernel_launches"):
            fn.restype = C.c_int
    return L


def lib():
    global _LIB
    if _LIB is None:
        # SGBM_B200_LIB: development override (an experimental build of the SAME library, e.g. build/exp/...); there is
        # still no fallback of any kind -- a missing file raises
        _LIB = load(os.environ.get("SGBM_B200_LIB") or LIB_PATH)
    return _LIB


def check(rc):
    if rc != 0:
        raise error(rc, lib().sgbm_last_error().decode("utf-8", "replace"))
