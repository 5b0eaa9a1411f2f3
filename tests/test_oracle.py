"""CPU tests: the oracle (C port of SURVEY.md Appendix A) against the golden vectors that were
produced by the reference's own implementation (cv2, via tests/golden/make_golden.py) and, when
cv2 is importable, against the live binary on a randomised sweep."""
import hashlib

import numpy as np
import pytest

import oracle
from oracle import OracleParams, cv2_ref
from synth import make_noise_pair, make_pair


def _params(meta, name):
    return OracleParams(**meta["cases"][name])


def test_golden_small_disparity(golden, golden_meta):
    names = sorted(golden_meta["cases"])
    assert len(names) >= 20
    for name in names:
        p = _params(golden_meta, name)
        got = oracle.compute(p, golden[name + "__left"], golden[name + "__right"])
        ref = golden[name + "__disp"]
        assert got.dtype == np.int16 and got.shape == ref.shape
        assert int((got != ref).sum()) == 0, name


def test_golden_post_filters(golden):
    raw = golden["post__in"]
    assert np.array_equal(oracle.median3x3(raw), golden["post__median"])
    assert np.array_equal(oracle.filter_speckles(raw, -16, 60, 32), golden["post__speckle_60_32"])


def test_golden_reproject(golden):
    d = golden["reproj__disp_f32"]
    for q, x in (("reproj__Q", "reproj__xyz"), ("reproj__Q_general", "reproj__xyz_general")):
        got = oracle.reproject_f32(d, golden[q])
        assert np.array_equal(got.view(np.uint32), golden[x].view(np.uint32))      # bit exact
    got = oracle.reproject_f32(golden["post__in"].astype(np.float32), golden["reproj__Q"])
    assert np.array_equal(got.view(np.uint32), golden["reproj__xyz_i16"].view(np.uint32))


def test_golden_digest_mid(golden_meta):
    for name in ("mid_640x360_D64_SGBM", "mid_640x360_D64_HH", "mid_640x360_D64_3WAY"):
        g = golden_meta["digests"][name]
        l, r, _ = make_pair(g["W"], g["H"], g["D"], seed=g["seed"])
        if hashlib.sha256(l.tobytes()).hexdigest() != g["left_sha256"]:
            pytest.skip("synthetic generator differs from the one that made the digests (cv2 version?)")
        p = OracleParams(0, g["D"], 5, 200, 800, 1, 63, 10, 100, 32, g["mode"])
        got = oracle.compute(p, l, r)
        assert hashlib.sha256(got.tobytes()).hexdigest() == g["disp_sha256"], name


def test_golden_num_disparities_not_a_multiple_of_8(numdisp_cases):
    """cv2 accepts any positive numDisparities ([P16]); the oracle equals the committed cv2 vectors for
    4 ... 100, odd values included (tests/golden/make_golden_numdisp.py)."""
    n = 0
    for name, p, l, r, ref in numdisp_cases:
        assert int((oracle.compute(p, l, r) != ref).sum()) == 0, name
        n += 1
    assert n == 36


def test_invalid_sizes_raise():
    l, r = make_noise_pair(40, 20, seed=0)
    with pytest.raises(ValueError):
        oracle.compute(OracleParams(0, 48, 5, 0, 0, 0, 0, 0, 0, 0, 0), l, r)     # W - D <= bs/2
    with pytest.raises(ValueError):
        oracle.compute(OracleParams(-14, 48, 3, 0, 0, 0, 0, 0, 0, 0, 2), np.zeros((20, 47), np.uint8),
                       np.zeros((20, 47), np.uint8))                              # empty valid range [P18]


@pytest.mark.skipif(not cv2_ref.available(), reason="cv2 not importable")
def test_live_cv2_random_sweep():
    """Randomised parameter sweep in the style of SURVEY probe P17 (parity domain: W1 > r)."""
    rng = np.random.default_rng(2024)
    n = 0
    for it in range(150):
        mode = it % 3
        W = int(rng.integers(40, 160)); H = int(rng.integers(28, 56))
        D = int(rng.choice([16, 32, 48])); minD = int(rng.integers(-20, 21))
        bs = int(rng.choice([1, 3, 5, 7, 9, 11]))
        if rng.random() < 0.5:
            P1, P2 = int(rng.integers(0, 400)), int(rng.integers(0, 3000))
        else:
            P1, P2 = 8 * 3 * bs * bs, 32 * 3 * bs * bs
        p = OracleParams(minD, D, bs, P1, P2, int(rng.integers(-1, 4)), int(rng.integers(0, 80)),
                         int(rng.integers(-1, 30)), int(rng.choice([0, 20, 100])),
                         int(rng.choice([1, 2, 32])), mode)
        W1 = W + min(minD, 0) - max(minD + D, 0)
        if W1 <= bs // 2 or not (W - (minD + D) > bs // 2):
            continue
        if mode == 2:
            ss = (H + 3) // 4
            if ss < bs // 2 + 1 + int(np.ceil(0.1 * ss)):
                continue
        if it % 2:
            l, r, _ = make_pair(W, H, max(D + max(minD, 0), 8), seed=it)
        else:
            l, r = make_noise_pair(W, H, seed=it)
        assert int((cv2_ref.compute(p, l, r) != oracle.compute(p, l, r)).sum()) == 0, (it, p, W, H)
        n += 1
    assert n > 80
