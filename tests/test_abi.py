"""CPU tests of the boundary: the C-ABI library loads and exports every symbol that
include/sgbm_b200.h declares; parameter objects behave like the cv2 binding's."""
import ctypes
import os
import re

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as ge
    ge.build()
    from stereo_reconstruction_cv_b200 import _lib
    return _lib


def test_header_symbols_exported(built):
    hdr = open(os.path.join(ROOT, "include", "sgbm_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(sgbm_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 18
    L = ctypes.CDLL(built.LIB_PATH)
    for n in sorted(names):
        assert hasattr(L, n), "symbol %s declared in the header is not exported" % n
    assert names == set(built.SYMBOLS)


def test_no_cpu_fallback(built):
    """Without a CUDA device the product must fail loudly, not fall back to the oracle."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import stereo_reconstruction_cv_b200 as sg
    with pytest.raises(sg.error):
        sg.StereoSGBM_create(numDisparities=16)


def test_result_pool_without_page_locked_memory(built):
    """The recycling pool of page-locked result blocks (_hostpool.py) only decides WHERE a result array lives: when no
    page-locked memory can be had (here: no CUDA device) it hands out ordinary arrays, and small results never use it."""
    import numpy as np
    import torch
    from stereo_reconstruction_cv_b200 import _hostpool
    small = _hostpool.empty((64, 64), np.int16)
    assert small.flags.owndata and small.shape == (64, 64) and small.dtype == np.int16
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    big = _hostpool.empty((1200, 1000), np.int16)
    assert big.flags.owndata and big.shape == (1200, 1000)
    assert _hostpool.stats() == {"outstanding_bytes": 0, "cached_bytes": 0}


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "stereo_reconstruction_cv_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "liboracle" not in txt, f
                # ... nor the reference's implementation: the package computes everything itself
                assert "import cv2" not in txt and "from cv2" not in txt, f
