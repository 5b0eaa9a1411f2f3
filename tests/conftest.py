import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(GOLDEN_DIR, "golden_small.npz"))


@pytest.fixture(scope="session")
def golden_meta():
    return json.load(open(os.path.join(GOLDEN_DIR, "golden_digests.json")))


@pytest.fixture(scope="session")
def numdisp_cases():
    """(name, params, left, right, cv2 disparity) for numDisparities that are not a multiple of 8: committed cv2 outputs
    (tests/golden/make_golden_numdisp.py), inputs regenerated from the recorded recipe."""
    from oracle import OracleParams
    from synth import make_noise_pair, make_pair
    meta = json.load(open(os.path.join(GOLDEN_DIR, "golden_numdisp.json")))
    arrs = np.load(os.path.join(GOLDEN_DIR, "golden_numdisp.npz"))
    out = []
    for name, c in sorted(meta["cases"].items()):
        c = dict(c)
        W, H, kind, seed = c.pop("W"), c.pop("H"), c.pop("kind"), c.pop("seed")
        l, r = make_pair(W, H, max(c["numDisparities"], 8), seed=seed)[:2] if kind == "synth" else make_noise_pair(W, H, seed=seed)
        out.append((name, OracleParams(**c), l, r, arrs[name + "__disp"]))
    return out
