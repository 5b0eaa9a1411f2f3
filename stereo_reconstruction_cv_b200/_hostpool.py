"""Result arrays in page-locked host memory.

`stereo.compute(imgL, imgR)` (main.ipynb:668) returns a NEW numpy array per call, as cv2 does.  A freshly allocated
pageable array costs a staging copy plus one page fault per 4 KB on first touch -- at 3840x2160 that was about 3 ms of
the call, more than the kernels of the notebook's own parameters take.  Results therefore come from a recycling pool
of page-locked blocks (C ABI: sgbm_host_alloc / sgbm_host_free): the device -> host copy lands in the array directly,
and a block goes back to the pool when the last array (or view) that uses it is garbage collected.  The arrays behave
like any other ndarray (writeable, C-contiguous; `owndata` is False).  Page-locking costs about a millisecond per megabyte,
so only results between MIN_BYTES and MAX_BYTES use the pool, and the pool is bounded: beyond MAX_OUTSTANDING
bytes held by live results -- a caller that keeps every frame -- new results are ordinary pageable arrays again."""
import atexit
import ctypes as C
import threading

import numpy as np

from . import _lib

MIN_BYTES = 1 << 20            # small results are not worth a page-locked block
MAX_BYTES = 256 << 20          # ... and page-locking a huge one (a whole batch's result) costs more than the staged copy it saves
MAX_OUTSTANDING = 2 << 30      # page-locked bytes in live result arrays
MAX_CACHED = 512 << 20         # page-locked bytes kept for reuse
_GRAIN = 1 << 20

_lock = threading.Lock()
_free = {}                     # capacity -> [pointers]
_cached = 0
_outstanding = 0
_closed = False


def _take(nbytes):
    global _cached, _outstanding
    cap = (nbytes + _GRAIN - 1) // _GRAIN * _GRAIN
    with _lock:
        if _closed or _outstanding + cap > MAX_OUTSTANDING:
            return None
        lst = _free.get(cap)
        if lst:
            _cached -= cap
            _outstanding += cap
            return lst.pop(), cap
    p = C.c_void_p()
    if _lib.lib().sgbm_host_alloc(cap, C.byref(p)) != 0 or not p.value:
        return None                                  # no page-locked memory left: the caller falls back to a pageable array
    with _lock:
        _outstanding += cap
    return p.value, cap


def _give(ptr, cap):
    global _cached, _outstanding
    with _lock:
        _outstanding -= cap
        if _closed:
            return                                   # interpreter shutdown: the CUDA context may be gone already
        if _cached + cap <= MAX_CACHED:
            _free.setdefault(cap, []).append(ptr)
            _cached += cap
            return
    _lib.lib().sgbm_host_free(ptr)


class _Block:
    """Owner of one page-locked block: numpy keeps it as the `base` of the result array and of every view of it."""
    __slots__ = ("ptr", "cap", "__array_interface__", "__weakref__")

    def __init__(self, ptr, cap, shape, dtype):
        self.ptr, self.cap = ptr, cap
        self.__array_interface__ = {"data": (ptr, False), "shape": tuple(shape), "typestr": np.dtype(dtype).str, "version": 3}

    def __del__(self):
        try:
            _give(self.ptr, self.cap)
        except Exception:                            # never raise from a finaliser (interpreter teardown)
            pass


def empty(shape, dtype):
    """np.empty(shape, dtype) in page-locked memory when the pool allows it, an ordinary array otherwise."""
    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape)) * dtype.itemsize
    if MIN_BYTES <= nbytes <= MAX_BYTES:
        blk = _take(nbytes)
        if blk is not None:
            return np.asarray(_Block(blk[0], blk[1], shape, dtype))
    return np.empty(shape, dtype)


def stats():
    with _lock:
        return {"outstanding_bytes": _outstanding, "cached_bytes": _cached}


def trim():
    """Free the cached (unused) blocks."""
    global _cached
    with _lock:
        ptrs = [p for lst in _free.values() for p in lst]
        _free.clear()
        _cached = 0
    for p in ptrs:
        _lib.lib().sgbm_host_free(p)


@atexit.register
def _close():
    global _closed
    with _lock:
        _closed = True
