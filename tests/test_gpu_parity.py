"""GPU parity tests (run on the B200 box): the CUDA path, called through the C ABI, against the
oracle (C port of SURVEY.md Appendix A), the cv2-generated golden vectors and -- when importable --
the live cv2 binary.  Bit-exact for all integer outputs; reprojection bit-exact in fp32 (the
north star's tolerance is 1e-5 relative; the test states both)."""
import hashlib

import numpy as np
import pytest

import oracle
from oracle import OracleParams, cv2_ref
from synth import make_noise_pair, make_pair

pytestmark = pytest.mark.gpu

MODES = {0: "SGBM", 1: "HH", 2: "3WAY", 3: "HH4"}


@pytest.fixture(scope="module")
def sg():
    import torch
    assert torch.cuda.is_available()
    import stereo_reconstruction_cv_b200 as sg
    return sg


def _kw(p):
    return dict(p.__dict__)


def _mismatch(a, b):
    return int((np.asarray(a) != np.asarray(b)).sum())


# ------------------------------------------------------------------------------------------------
# stage-wise: cost volume, aggregated S, raw disparity (before median) vs the oracle's stages
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", [0, 1, 2, 3])
@pytest.mark.parametrize("D,bs,minD", [(16, 3, 0), (64, 5, 0), (32, 7, -5), (48, 1, 3), (128, 5, 0)])
def test_stages_vs_oracle(sg, mode, D, bs, minD):
    W, H = 200 + D, 60
    l, r, _ = make_pair(W, H, D + max(minD, 0), seed=D + mode)
    p = OracleParams(minD, D, bs, 8 * bs * bs, 32 * bs * bs, 1, 63, 10, 0, 0, mode)
    o = oracle.compute_debug(p, l, r, want_C=True, want_S=True, want_raw=True)
    st = sg.StereoSGBM_create(**_kw(p))
    st._debug_keep(True)
    disp = st.compute(l, r)
    W1 = oracle.port.valid_width(p, W)
    if mode != 2:      # the oracle's C dump for 3WAY is stripe 0 only
        C = st._debug_fetch(0, (H, W1, D))
        assert _mismatch(C, o["C"]) == 0, "cost volume"
    S = st._debug_fetch(1, (H, W1, D))
    assert _mismatch(S, o["S"]) == 0, "aggregated cost S"
    raw = st._debug_fetch(2, (H, W))
    assert _mismatch(raw, o["raw"]) == 0, "raw disparity (WTA / uniqueness / subpixel / LR check)"
    assert _mismatch(disp, o["disp"]) == 0, "final disparity"


# ------------------------------------------------------------------------------------------------
# golden vectors produced by the reference's implementation (cv2), committed under tests/golden
# ------------------------------------------------------------------------------------------------
def test_golden_small(sg, golden, golden_meta):
    for name in sorted(golden_meta["cases"]):
        p = OracleParams(**golden_meta["cases"][name])
        st = sg.StereoSGBM_create(**_kw(p))
        got = st.compute(golden[name + "__left"], golden[name + "__right"])
        assert got.dtype == np.int16
        assert _mismatch(got, golden[name + "__disp"]) == 0, name


@pytest.mark.parametrize("name", ["mid_640x360_D64_SGBM", "mid_640x360_D64_HH", "mid_640x360_D64_3WAY",
                                  "cfg2_1280x720_D128_SGBM", "cfg2_1280x720_D128_HH", "cfg2_1280x720_D128_3WAY",
                                  "cfg4_1920x1080_D192_SGBM_seed0", "cfg3_3840x2160_D256_HH",
                                  "cfg5_3840x2160_D256_3WAY"])
def test_golden_digests(sg, golden_meta, name):
    """BASELINE.json configs at full size: SHA-256 of the int16 disparity must equal cv2's."""
    assert name in golden_meta["digests"], "digest missing: run tests/golden/make_golden.py --full"
    g = golden_meta["digests"][name]
    l, r, _ = make_pair(g["W"], g["H"], g["D"], seed=g["seed"])
    # a different generator (e.g. another cv2 build) must FAIL here, not silently skip the full-size gates
    assert hashlib.sha256(l.tobytes()).hexdigest() == g["left_sha256"] and \
        hashlib.sha256(r.tobytes()).hexdigest() == g["right_sha256"], \
        "the synthetic generator no longer reproduces the inputs the digests were made from: regenerate them"
    st = sg.StereoSGBM_create(minDisparity=0, numDisparities=g["D"], blockSize=5, P1=200, P2=800, disp12MaxDiff=1,
                              preFilterCap=63, uniquenessRatio=10, speckleWindowSize=100, speckleRange=32,
                              mode=g["mode"])
    got = st.compute(l, r)
    assert abs(float((got >= 0).mean()) - g["valid_fraction"]) < 1e-12
    assert hashlib.sha256(got.tobytes()).hexdigest() == g["disp_sha256"]


# ------------------------------------------------------------------------------------------------
# randomised sweep against the oracle (and live cv2 when importable)
# ------------------------------------------------------------------------------------------------
def test_random_sweep(sg):
    rng = np.random.default_rng(77)
    n = 0
    use_cv2 = cv2_ref.available()
    for it in range(120):
        mode = it % 4
        W = int(rng.integers(40, 200)); H = int(rng.integers(28, 70))
        D = int(rng.choice([8, 16, 32, 48, 80, 96])); minD = int(rng.integers(-20, 21))
        bs = int(rng.choice([0, 1, 3, 4, 5, 7, 9, 11]))
        if rng.random() < 0.5:
            P1, P2 = int(rng.integers(0, 400)), int(rng.integers(0, 3000))
        else:
            P1, P2 = 8 * 3 * bs * bs, 32 * 3 * bs * bs
        p = OracleParams(minD, D, bs, P1, P2, int(rng.integers(-1, 4)), int(rng.integers(0, 80)),
                         int(rng.integers(-1, 30)), int(rng.choice([0, 20, 100])), int(rng.choice([1, 2, 32])), mode)
        r_eff = (bs if bs > 0 else (3 if mode == 2 else 5)) // 2
        W1 = W + min(minD, 0) - max(minD + D, 0)
        if W1 <= r_eff or not (W - (minD + D) > bs // 2):
            continue
        if mode == 2:
            ss = (H + 3) // 4
            if ss < bs // 2 + 1 + int(np.ceil(0.1 * ss)):
                continue
        if it % 3 == 0:
            l, r = make_noise_pair(W, H, seed=it)
        else:
            l, r, _ = make_pair(W, H, max(D + max(minD, 0), 8), seed=it)
        got = sg.StereoSGBM_create(**_kw(p)).compute(l, r)
        assert _mismatch(got, oracle.compute(p, l, r)) == 0, (it, p, W, H)
        if use_cv2 and mode != 3 and D % 16 == 0:
            assert _mismatch(got, cv2_ref.compute(p, l, r)) == 0, ("cv2", it, p, W, H)
        n += 1
    assert n > 60


def test_three_channel_and_views(sg):
    l, r, _ = make_pair(160, 48, 16, seed=5)
    l3 = np.stack([l, np.roll(l, 1, 1), l[::-1].copy()], -1)
    r3 = np.stack([r, np.roll(r, 1, 1), r[::-1].copy()], -1)
    p = OracleParams(0, 16, 5, 200, 800, 1, 63, 10, 0, 0, 0)
    st = sg.StereoSGBM_create(**_kw(p))
    assert _mismatch(st.compute(l3, r3), oracle.compute(p, l3, r3)) == 0
    # non-contiguous views are accepted like cv2 does
    big_l = np.zeros((48, 200), np.uint8); big_r = np.zeros((48, 200), np.uint8)
    big_l[:, 20:180] = l; big_r[:, 20:180] = r
    assert _mismatch(st.compute(big_l[:, 20:180], big_r[:, 20:180]), oracle.compute(p, l, r)) == 0


def test_torch_batch_path(sg):
    import torch
    frames = [make_pair(256, 64, 32, seed=s) for s in range(3)]
    L = torch.from_numpy(np.stack([f[0] for f in frames])).cuda()
    R = torch.from_numpy(np.stack([f[1] for f in frames])).cuda()
    p = OracleParams(0, 32, 5, 200, 800, 1, 63, 10, 50, 2, 1)
    st = sg.StereoSGBM_create(**_kw(p))
    out = st.compute(L, R)
    assert out.shape == (3, 64, 256) and out.dtype == torch.int16 and out.is_cuda
    for i, f in enumerate(frames):
        assert _mismatch(out[i].cpu().numpy(), oracle.compute(p, f[0], f[1])) == 0


def test_errors(sg):
    l, r = make_noise_pair(40, 20, seed=0)
    with pytest.raises(sg.error):
        sg.StereoSGBM_create(numDisparities=48, blockSize=5).compute(l, r)            # cv2: stereosgbm.cpp:511
    with pytest.raises(sg.error):
        sg.StereoSGBM_create(minDisparity=-14, numDisparities=48, mode=2).compute(np.zeros((20, 47), np.uint8),
                                                                                   np.zeros((20, 47), np.uint8))
    with pytest.raises(sg.error):
        sg.StereoSGBM_create(numDisparities=16).compute(l, r[:, :30])                 # size mismatch
    with pytest.raises(sg.error):
        sg.StereoSGBM_create(numDisparities=16).compute(l.astype(np.float32), r.astype(np.float32))
    with pytest.raises(sg.error):
        sg.reprojectImageTo3D(np.zeros((4, 4), np.float32), np.eye(3))                # cv2: stereo_geom.cpp:19
    with pytest.raises(sg.error):
        sg.reprojectImageTo3D(np.zeros((4, 4), np.float64), np.eye(4))                # cv2: stereo_geom.cpp:17


# ------------------------------------------------------------------------------------------------
# post filters and reprojection
# ------------------------------------------------------------------------------------------------
def test_post_filters_golden(sg, golden):
    raw = golden["post__in"]
    assert _mismatch(sg.medianBlur3(raw), golden["post__median"]) == 0
    img = raw.copy()
    sg.filterSpeckles(img, -16, 60, 32)
    assert _mismatch(img, golden["post__speckle_60_32"]) == 0
    rng = np.random.default_rng(3)
    for (w, h, win, rg) in [(333, 97, 25, 1), (64, 64, 400, 3), (500, 300, 100, 32)]:
        a = (rng.integers(-1, 40, (h, w)) * 16).astype(np.int16)
        a[rng.random((h, w)) < 0.3] = -16
        b = a.copy()
        sg.filterSpeckles(b, -16, win, 16 * rg)
        assert _mismatch(b, oracle.filter_speckles(a, -16, win, 16 * rg)) == 0


def test_reproject_golden(sg, golden):
    d = golden["reproj__disp_f32"]
    for q, x in (("reproj__Q", "reproj__xyz"), ("reproj__Q_general", "reproj__xyz_general")):
        got = sg.reprojectImageTo3D(d, golden[q])
        ref = golden[x]
        assert got.dtype == np.float32 and got.shape == ref.shape
        fin = np.isfinite(ref)
        assert np.array_equal(np.isfinite(got), fin)                                    # same non-finite pattern
        assert np.array_equal(got[~fin], ref[~fin]) or np.array_equal(np.sign(got[~fin]), np.sign(ref[~fin]))
        rel = np.abs(got[fin] - ref[fin]) / np.maximum(np.abs(ref[fin]), 1e-30)
        assert rel.max() <= 1e-5                                                        # north-star tolerance
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))                 # and in fact bit exact
    got = sg.reprojectImageTo3D(golden["post__in"], golden["reproj__Q"])                # int16 used as is
    assert np.array_equal(got.view(np.uint32), golden["reproj__xyz_i16"].view(np.uint32))


def test_fused_tail_matches_notebook_recipe(sg):
    """int16 disparity -> /16, mask, reproject, finite/positive mask, gather (main.ipynb:668-670,697,726-737)."""
    import torch
    l, r, _ = make_pair(320, 120, 32, seed=9)
    p = OracleParams(0, 32, 5, 200, 800, 1, 63, 10, 100, 32, 2)
    st = sg.StereoSGBM_create(**_kw(p))
    dt = st.compute(torch.from_numpy(l).cuda(), torch.from_numpy(r).cuda())
    Q = np.array([[1, 0, 0, -160.0], [0, 1, 0, -60.0], [0, 0, 0, 400.0], [0, 0, -1, 0]], np.float64)
    col = np.stack([l, r, (l // 2 + r // 2)], -1).astype(np.uint8)                       # BGR
    xyz, rgb = sg.reprojectCompact(dt, Q, torch.from_numpy(col).cuda())
    disp = dt.cpu().numpy()
    f = disp.astype(np.float32) / 16.0
    f = f * (f > 0).astype(np.float32)
    assert np.array_equal(sg.disparityToFloat(dt).cpu().numpy().view(np.uint32), f.view(np.uint32))
    pts = oracle.reproject_f32(f, Q)
    mask = ~np.isnan(pts[:, :, 0]) & ~np.isinf(pts[:, :, 0]) & (f > 0)
    assert xyz.shape[0] == int(mask.sum()) > 1000
    assert np.array_equal(xyz.cpu().numpy().view(np.uint32), pts[mask].view(np.uint32))
    assert np.array_equal(rgb.cpu().numpy(), col[:, :, ::-1][mask])


# ------------------------------------------------------------------------------------------------
# size-independent properties at BASELINE.json's full sizes
# ------------------------------------------------------------------------------------------------
def test_full_size_properties(sg):
    import torch
    W, H, D = 3840, 2160, 256
    l, r, gt = make_pair(W, H, D, seed=0)
    st = sg.StereoSGBM_create(minDisparity=0, numDisparities=D, blockSize=5, P1=200, P2=800, disp12MaxDiff=1,
                              preFilterCap=63, uniquenessRatio=10, speckleWindowSize=100, speckleRange=32, mode=1)
    lt, rt = torch.from_numpy(l).cuda(), torch.from_numpy(r).cuda()
    d1 = st.compute(lt, rt)
    d2 = st.compute(lt, rt)
    assert torch.equal(d1, d2)                                             # deterministic
    d = d1.cpu().numpy()
    assert d.shape == (H, W) and (d[:, :D] == -16).all()                   # columns [0, minX1) invalid (A.0)
    valid = d >= 0
    assert 0.3 < valid.mean() < 0.95
    assert d[valid].max() < D * 16
    err = np.abs(d[valid] / 16.0 - gt[valid])
    assert np.median(err) < 0.5                                            # recovers the synthetic ground truth
    # speckle filter is idempotent; running it again on the output changes nothing
    again = d1.clone()
    sg.filterSpeckles(again, -16, 100, 16 * 32)
    assert torch.equal(again, d1)
    # a horizontally cropped band of rows reproduces the 3WAY stripe-independence: rows of stripe 0
    st3 = sg.StereoSGBM_create(minDisparity=0, numDisparities=D, blockSize=5, P1=200, P2=800, disp12MaxDiff=1,
                               preFilterCap=63, uniquenessRatio=10, speckleWindowSize=0, speckleRange=0, mode=2)
    full = st3.compute(lt, rt).cpu().numpy()
    ss = (H + 3) // 4
    top = st3.compute(lt[:ss + 64].contiguous(), rt[:ss + 64].contiguous()).cpu().numpy()
    # stripe 0 of the full image only sees rows < ss + r + 1 => identical away from the crop's own stripes
    assert np.array_equal(full[: (ss + 64 + 3) // 4 - 8], top[: (ss + 64 + 3) // 4 - 8])


def test_host_batch_pipeline(sg):
    """compute_batch (double-buffered pinned staging inside sgbm_compute_host) == per-frame oracle results."""
    W, H, D = 272, 70, 48
    pairs = [make_pair(W, H, D, seed=40 + i)[:2] for i in range(5)]
    p = OracleParams(0, D, 5, 200, 800, 1, 63, 10, 50, 2, 0)
    st = sg.StereoSGBM_create(**_kw(p))
    lefts = np.stack([a for a, _ in pairs])
    rights = np.stack([b for _, b in pairs])
    for B in (1, 2, 5):
        out = st.compute_batch(lefts[:B], rights[:B])
        assert out.shape == (B, H, W) and out.dtype == np.int16
        for i in range(B):
            assert _mismatch(out[i], oracle.compute(p, pairs[i][0], pairs[i][1])) == 0, "frame %d of batch %d" % (i, B)


def test_results_live_in_recycled_page_locked_memory(sg, monkeypatch):
    """compute(numpy, numpy) returns a new array per call (cv2's contract) out of a recycling pool of page-locked
    blocks: results stay intact while they (or views of them) are alive, blocks are reused once they are dropped,
    the pool is bounded, a caller-supplied `disparity` array is written in place, and reprojectImageTo3D's host
    result takes the same route."""
    import gc
    from stereo_reconstruction_cv_b200 import _hostpool as _pinned
    W, H, D = 1100, 520, 32                          # 1.1 MB of int16: above the pool's minimum
    pairs = [make_pair(W, H, D, seed=70 + i)[:2] for i in range(3)]
    p = OracleParams(0, D, 5, 200, 800, 1, 63, 10, 100, 32, 0)
    refs = [oracle.compute(p, a, b) for a, b in pairs]
    st = sg.StereoSGBM_create(**_kw(p))
    gc.collect()
    base = _pinned.stats()["outstanding_bytes"]
    outs = [st.compute(a, b) for a, b in pairs]       # three live results: three different blocks
    assert len({o.ctypes.data for o in outs}) == 3
    assert _pinned.stats()["outstanding_bytes"] - base >= 3 * W * H * 2
    for o, ref in zip(outs, refs):
        assert o.dtype == np.int16 and o.flags.c_contiguous and o.flags.writeable and _mismatch(o, ref) == 0
    view = outs[0][100:200, 50:]                      # a view keeps its block out of the pool
    addr0 = outs[0].ctypes.data
    del outs, o
    gc.collect()
    again = [st.compute(a, b) for a, b in pairs]
    assert addr0 not in {a.ctypes.data for a in again}
    assert _mismatch(view, refs[0][100:200, 50:]) == 0
    addrs = {a.ctypes.data for a in again}
    del again, view
    gc.collect()
    assert _pinned.stats()["outstanding_bytes"] == base
    o = st.compute(*pairs[1])                         # a dropped block comes back
    assert o.ctypes.data in addrs | {addr0} and _mismatch(o, refs[1]) == 0
    # caller-supplied output, dense and strided
    dst = np.zeros((H, W), np.int16)
    assert st.compute(pairs[2][0], pairs[2][1], dst) is dst and _mismatch(dst, refs[2]) == 0
    wide = np.zeros((H, W + 6), np.int16)
    assert _mismatch(st.compute(pairs[2][0], pairs[2][1], wide[:, 3:W + 3]), refs[2]) == 0 and not wide[:, :3].any()
    # a bounded pool: beyond the cap results are ordinary arrays, and still right
    monkeypatch.setattr(_pinned, "MAX_OUTSTANDING", _pinned.stats()["outstanding_bytes"] + W * H * 2)
    o2 = st.compute(*pairs[0])
    assert o2.flags.owndata and _mismatch(o2, refs[0]) == 0
    monkeypatch.undo()
    # the point-cloud side
    Q = np.array([[1, 0, 0, -W / 2], [0, 1, 0, -H / 2], [0, 0, 0, 800.0], [0, 0, 1 / 60.0, 0]])
    f = (refs[0].astype(np.float32) / 16)
    xyz = sg.reprojectImageTo3D(f, Q)
    assert xyz.shape == (H, W, 3) and xyz.dtype == np.float32
    assert np.array_equal(xyz.view(np.uint32), oracle.reproject_f32(f, Q).view(np.uint32))
    _pinned.trim()


# ------------------------------------------------------------------------------------------------
# the sweep's strip hand-off: every legal number of rows per super-step, wide enough for many strips
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("R", [1, 2, 3, 4, 5, 6, 7, 8])
@pytest.mark.parametrize("mode", [0, 1])
def test_sweep_rows_per_superstep(sg, monkeypatch, R, mode):
    W, H, D = 1500, 96, 32
    l, r, _ = make_pair(W, H, D, seed=90 + R)
    p = OracleParams(0, D, 5, 200, 800, 1, 63, 10, 0, 0, mode)
    ref = oracle.compute(p, l, r)
    monkeypatch.setenv("SGBM_VR", str(R))
    st = sg.StereoSGBM_create(**_kw(p))
    for rep in range(3):                       # hand-off races are timing dependent: repeat
        assert _mismatch(st.compute(l, r), ref) == 0, (R, mode, rep)


@pytest.mark.parametrize("env", [{}, {"SGBM_SWEEP_W": "0"}, {"SGBM_SWEEP": "0"}, {"SGBM_ROWSTEP": "1"}, {"SGBM_COST3": "0"},
                                 {"SGBM_COST2": "0"}])
def test_kernel_generations_agree(sg, monkeypatch, env):
    """The fallback kernels (sweep without the WTA role, lock-step vertical kernel, row-at-a-time kernel,
    cost generations 1/2) stay bit-exact: they serve the geometries the newest kernels do not hold."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    W, H, D = 700, 80, 64
    l, r, _ = make_pair(W, H, D, seed=11)
    for mode in (0, 1, 2):
        p = OracleParams(0, D, 5, 200, 800, 1, 63, 10, 100, 32, mode)
        assert _mismatch(sg.StereoSGBM_create(**_kw(p)).compute(l, r), oracle.compute(p, l, r)) == 0, (env, mode)


@pytest.mark.parametrize("env", [{"SGBM_SWEEP_WRG": "1"}, {"SGBM_SWEEP_WRG": "2"}, {"SGBM_SWEEP_NWW": "1"}, {"SGBM_HH_SPLIT": "0"},
                                 {"SGBM_SWEEP_PF": "0"}, {"SGBM_SWEEP_PF": "3"}, {"SGBM_SWEEP_K": "3", "SGBM_SWEEP_NSC": "2"},
                                 {"SGBM_SWEEP_K": "8", "SGBM_SWEEP_NSI": "2"}, {"SGBM_VR": "16"}, {"SGBM_VR": "12", "SGBM_SWEEP_WRG": "3"},
                                 {"SGBM_SWEEP_RPS": "1"}, {"SGBM_SWEEP_RPS": "2"}, {"SGBM_SWEEP_RPS": "2", "SGBM_VR": "2"},
                                 {"SGBM_SWEEP_RPS": "2", "SGBM_SWEEP_K": "3", "SGBM_SWEEP_NSC": "2"},
                                 {"SGBM_SWEEP_RPS": "2", "SGBM_SWEEP_K": "12", "SGBM_SWEEP_NSC": "9", "SGBM_SWEEP_NSI": "4"}])
def test_sweep_schedule_knobs(sg, monkeypatch, env):
    """The round-2 schedule options of the sweeps -- row groups of the winner-take-all warps, warps per row, where MODE_HH
    reads L_hB, the producer's L2 prefetch distance, ring depths, 12 / 16 rows per super-step, one or two rows per ring
    stage -- never change a bit: narrow
    strips (many row groups) and wide ones, few and many disparities, saturating and plain accumulation, three frames each."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    for (W, H, D, bs, P1, P2) in [(1200, 90, 64, 5, 200, 800), (2600, 70, 16, 11, 2904, 11616), (500, 100, 128, 3, 72, 288)]:
        l, r, _ = make_pair(W, H, D, seed=W)
        for mode in (0, 1):
            p = OracleParams(0, D, bs, P1, P2, 1, 63, 10, 100, 32, mode)
            ref = oracle.compute(p, l, r)
            st = sg.StereoSGBM_create(**_kw(p))
            for rep in range(3):
                assert _mismatch(st.compute(l, r), ref) == 0, (env, W, D, mode, rep)


@pytest.mark.parametrize("H", [1, 2, 3, 9, 95, 97])
@pytest.mark.parametrize("rps", ["1", "2"])
def test_sweep_rows_per_stage_odd_heights(sg, monkeypatch, H, rps):
    """Two image rows per ring stage: the last stage of an odd-height image holds one row, super-steps end on stage
    boundaries; every mode that sweeps, forward and backward, against the same frames with one row per stage."""
    monkeypatch.setenv("SGBM_SWEEP_RPS", rps)
    for (W, D, bs) in [(900, 64, 5), (1400, 16, 9), (800, 192, 3)]:
        l, r, _ = make_pair(W, H, D, seed=H + D)
        for mode in (0, 1):
            p = OracleParams(0, D, bs, 8 * bs * bs, 32 * bs * bs, 1, 63, 10, 0, 0, mode)
            ref = oracle.compute(p, l, r)
            st = sg.StereoSGBM_create(**_kw(p))
            for rep in range(2):
                assert _mismatch(st.compute(l, r), ref) == 0, (H, rps, W, D, mode, rep)


@pytest.mark.parametrize("D", [4, 5, 6, 7, 9, 10, 12, 20, 21, 27, 36, 44, 99, 100, 250, 255])
def test_num_disparities_not_a_multiple_of_8(sg, D):
    """cv2 documents numDisparities % 16 == 0 but accepts anything ([P16]): values >= 4, odd ones too, run on volumes
    padded to the next multiple of 8 whose padding disparities are inert in the path step and masked in every
    winner-take-all.  MODE_SGBM, MODE_HH and MODE_HH4 against the oracle (which equals cv2 there) and against live cv2;
    the notebook's penalties (saturating accumulation), a 3-channel pair, minDisparity != 0, several strips."""
    import torch
    cases = [(900, 70, 5, 200, 800, 0, 1), (1300, 40, 11, 2904, 11616, 0, 1), (420, 60, 3, 216, 864, -3, 3), (2100, 33, 7, 392, 1568, 5, 1)]
    for (W, H, bs, P1, P2, minD, cn) in cases:
        if W - D - abs(minD) < 40:
            continue
        l, r, _ = make_pair(W, H, max(D, 8), seed=D + W)
        if cn == 3:
            l = np.stack([l, np.roll(l, 1, 0), l // 2 + 7], -1).astype(np.uint8)
            r = np.stack([r, np.roll(r, 1, 0), r // 2 + 7], -1).astype(np.uint8)
        for mode in (0, 1, 3):
            p = OracleParams(minD, D, bs, P1 * cn, P2 * cn, 1, 63, 10, 60, 16, mode)
            ref = oracle.compute(p, l, r)
            st = sg.StereoSGBM_create(**_kw(p))
            assert _mismatch(st.compute(l, r), ref) == 0, (D, W, bs, mode, "host")
            got = st.compute(torch.from_numpy(l).cuda(), torch.from_numpy(r).cuda()).cpu().numpy()
            assert _mismatch(got, ref) == 0, (D, W, bs, mode, "device")
            if cv2_ref.available() and mode != 3:             # (cv2's MODE_HH4 is not repeatable: DESIGN.md section 2)
                assert _mismatch(got, cv2_ref.compute(p, l, r)) == 0, (D, W, bs, mode, "cv2")
    # batches side by side and the fallback kernels
    l, r, _ = make_pair(700, 64, max(D, 8), seed=D)
    p = OracleParams(0, D, 5, 200, 800, 1, 63, 10, 0, 0, 0)
    ref = oracle.compute(p, l, r)
    out = sg.StereoSGBM_create(**_kw(p)).compute_batch(np.stack([l] * 5), np.stack([r] * 5))
    assert all(_mismatch(out[i], ref) == 0 for i in range(5))


@pytest.mark.parametrize("env", [{"SGBM_SWEEP_W": "0"}, {"SGBM_SWEEP": "0"}, {"SGBM_ROWSTEP": "1"}, {"SGBM_COST3": "0"}, {"SGBM_COST2": "0"},
                                 {"SGBM_SWEEP_RPS": "1"}])
def test_num_disparities_not_a_multiple_of_8_fallback_kernels(sg, monkeypatch, env):
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    for D in (12, 21, 36, 100):
        l, r, _ = make_pair(800, 50, D, seed=D)
        for mode in (0, 1, 3):
            p = OracleParams(0, D, 5, 200, 800, 1, 63, 10, 100, 32, mode)
            assert _mismatch(sg.StereoSGBM_create(**_kw(p)).compute(l, r), oracle.compute(p, l, r)) == 0, (env, D, mode)


def test_num_disparities_golden_vectors(sg, numdisp_cases):
    """The committed cv2 vectors for numDisparities 4 ... 100 (tests/golden/make_golden_numdisp.py)."""
    n = 0
    for name, p, l, r, ref in numdisp_cases:
        assert _mismatch(sg.StereoSGBM_create(**_kw(p)).compute(l, r), ref) == 0, name
        n += 1
    assert n == 36


def test_num_disparities_unsupported_values(sg):
    l, r, _ = make_pair(300, 40, 16, seed=1)
    for kw in (dict(numDisparities=3), dict(numDisparities=2), dict(numDisparities=20, mode=2), dict(numDisparities=21, mode=2), dict(numDisparities=1032),
               dict(numDisparities=20, blockSize=11, P1=2904, P2=22000)):
        with pytest.raises(sg.error):
            sg.StereoSGBM_create(**kw).compute(l, r)


@pytest.mark.parametrize("R", [8, 2])
def test_repeatability_under_load(sg, monkeypatch, R):
    """Strip hand-offs and role hand-offs are timing dependent: 150 back-to-back frames of a
    many-strip geometry must give the same disparity every time (and the oracle's).  R = 2 rows per
    super-step makes the warps of a role publish out of order (the case per-entry flags exist for)."""
    import torch
    monkeypatch.setenv("SGBM_VR", str(R))
    W, H, D = 1280, 360, 128
    frames = [make_pair(W, H, D, seed=s)[:2] for s in range(2)]
    for mode in (0, 1):
        p = OracleParams(0, D, 5, 200, 800, 1, 63, 10, 100, 32, mode)
        st = sg.StereoSGBM_create(**_kw(p))
        lt = [torch.from_numpy(f[0]).cuda() for f in frames]
        rt = [torch.from_numpy(f[1]).cuda() for f in frames]
        ref = [torch.from_numpy(oracle.compute(p, f[0], f[1])).cuda() for f in frames]
        for it in range(150):
            d = st.compute(lt[it % 2], rt[it % 2])
            assert bool((d == ref[it % 2]).all()), (mode, it)


def test_very_wide_image(sg):
    """7680 columns at numDisparities = 256: the strips do not fit the persistent sweeps, the row-at-a-time
    fallback takes over (cv2 has no width limit, neither has the drop-in)."""
    W, H, D = 7680, 40, 256
    l, r, _ = make_pair(W, H, D, seed=3)
    for mode in (0, 1):
        p = OracleParams(0, D, 5, 200, 800, 1, 63, 10, 100, 32, mode)
        assert _mismatch(sg.StereoSGBM_create(**_kw(p)).compute(l, r), oracle.compute(p, l, r)) == 0, mode


def test_handoff_watchdog_reports_and_recovers(sg, monkeypatch):
    """A sweep hand-off that never happens must not hang the GPU: the kernel drains after ~2 s, the call
    reports an error, the next call is fine.  The fault (role V of strip 0 withholds one arrival) exists only
    in the debug-hook build libsgbm_b200_dbg.so (-DSGBM_DEBUG_HOOKS), which this test loads by itself through
    the same C ABI; the product library ignores SGBM_DBG_* (second half of the test)."""
    import ctypes as C
    import time
    from stereo_reconstruction_cv_b200 import _lib
    W, H, D = 900, 64, 64
    l, r, _ = make_pair(W, H, D, seed=21)
    p = OracleParams(0, D, 5, 200, 800, 1, 63, 10, 0, 0, 0)
    ref = oracle.compute(p, l, r)
    dbg = _lib.load(_lib.DEBUG_LIB_PATH)

    def create(L):
        h = C.c_void_p()
        prm = _lib.SgbmParams(*[int(getattr(p, n)) for n, _ in _lib.SgbmParams._fields_])
        assert L.sgbm_create(C.byref(prm), C.byref(h)) == 0, L.sgbm_last_error()
        return h

    def compute(L, h):
        out = np.empty((H, W), np.int16)
        rc = L.sgbm_compute_host(h, l.ctypes.data, r.ctypes.data, W, H, 1, W, 1, out.ctypes.data, 2 * W)
        return rc, out, L.sgbm_last_error().decode()

    good = create(dbg)                                     # environment read at create: no fault for this handle
    rc, out, _ = compute(dbg, good)
    assert rc == 0 and _mismatch(out, ref) == 0
    monkeypatch.setenv("SGBM_DBG_STALL", "1")
    bad = create(dbg)
    t0 = time.time()
    rc, _, msg = compute(dbg, bad)
    assert rc != 0 and "hand-off timed out" in msg, (rc, msg)
    assert time.time() - t0 < 30
    rc, out, _ = compute(dbg, good)                        # the device and the other handle are fine
    assert rc == 0 and _mismatch(out, ref) == 0
    # the product library has no such hook: same environment, correct result
    assert _mismatch(sg.StereoSGBM_create(**_kw(p)).compute(l, r), ref) == 0
    monkeypatch.delenv("SGBM_DBG_STALL")
    dbg.sgbm_destroy(bad)
    dbg.sgbm_destroy(good)


@pytest.mark.parametrize("mode", [0, 1])
def test_batch_four_lanes_for_two_lane_columns(sg, mode):
    """numDisparities <= 32 (a column is two lanes): batches run FOUR frames side by side by default.  Nine frames (odd
    remainder), device tensors and host arrays, repeated calls; every frame equals the oracle."""
    import torch
    W, H, D, B = 1500, 64, 16, 9
    pairs = [make_pair(W, H, D, seed=500 + i)[:2] for i in range(B)]
    p = OracleParams(0, D, 11, 8 * 3 * 121, 32 * 3 * 121, 1, 63, 10, 100, 32, mode)       # the notebook's parameters (saturating)
    ref = np.stack([oracle.compute(p, l, r) for l, r in pairs])
    ls = np.stack([l for l, _ in pairs]); rs = np.stack([r for _, r in pairs])
    st = sg.StereoSGBM_create(**_kw(p))
    for rep in range(3):
        got = st.compute(torch.from_numpy(ls).cuda(), torch.from_numpy(rs).cuda())
        assert _mismatch(got.cpu().numpy(), ref) == 0, (mode, rep, "device")
    assert _mismatch(st.compute_batch(ls, rs), ref) == 0, (mode, "host")
    assert st.workspaceBytes(W, H, 1, batch=B) == 4 * st.workspaceBytes(W, H, 1, batch=1)


@pytest.mark.parametrize("mode", [0, 1])
def test_batch_side_by_side_schedule(sg, monkeypatch, mode):
    """Batched calls run two or three frames side by side, each on its share of the SMs (when the narrower
    sweeps hold the geometry): every frame must equal the oracle, for device tensors and host arrays, odd
    batch sizes, repeated calls, and with the schedule restricted (SGBM_LANES=2) or off (SGBM_LANES=1)."""
    import torch
    W, H, D, B = 1100, 72, 64, 5
    pairs = [make_pair(W, H, D, seed=40 + i)[:2] for i in range(B)]
    p = OracleParams(0, D, 5, 200, 800, 1, 63, 10, 100, 32, mode)
    ref = np.stack([oracle.compute(p, l, r) for l, r in pairs])
    ls = np.stack([l for l, _ in pairs]); rs = np.stack([r for _, r in pairs])
    for lanes in ("3", "2", "1"):
        monkeypatch.setenv("SGBM_LANES", lanes)
        st = sg.StereoSGBM_create(**_kw(p))
        for rep in range(4):
            got = st.compute(torch.from_numpy(ls).cuda(), torch.from_numpy(rs).cuda())
            assert got.shape == (B, H, W) and _mismatch(got.cpu().numpy(), ref) == 0, (lanes, rep, "device")
        assert _mismatch(st.compute_batch(ls, rs), ref) == 0, (lanes, "host")
        assert _mismatch(st.compute(ls[0], rs[0]), ref[0]) == 0


@pytest.mark.parametrize("bands", ["2", "3", "8"])
def test_row_band_schedule(sg, monkeypatch, bands):
    """Large frames are cut into row bands so that a band's horizontal kernel runs beside the next band's
    cost kernel (SGBM_BANDS forces the band count on a small frame): heights that are not a multiple of
    the band granularity, every mode (MODE_HH4 keeps one band), repeated frames and batches."""
    import torch
    monkeypatch.setenv("SGBM_BANDS", bands)
    W, D = 600, 64
    for H in (83, 48, 17):
        pairs = [make_pair(W, H, D, seed=60 + i)[:2] for i in range(3)]
        ls = np.stack([l for l, _ in pairs]); rs = np.stack([r for _, r in pairs])
        for mode in (0, 1, 2, 3):
            p = OracleParams(0, D, 5, 200, 800, 1, 63, 10, 100, 32, mode)
            ref = np.stack([oracle.compute(p, l, r) for l, r in pairs])
            st = sg.StereoSGBM_create(**_kw(p))
            for rep in range(3):
                assert _mismatch(st.compute(ls[rep], rs[rep]), ref[rep]) == 0, (bands, H, mode, rep)
            got = st.compute(torch.from_numpy(ls).cuda(), torch.from_numpy(rs).cuda())
            assert _mismatch(got.cpu().numpy(), ref) == 0, (bands, H, mode, "batch")


# ------------------------------------------------------------------------------------------------
# corners of the parameter / shape domain
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("bs,cap", [(13, 63), (15, 31), (17, 15)])
def test_large_block_sizes(sg, bs, cap):
    """blockSize > 11 leaves the register-resident cost kernel (r <= 5) for the older generations.  The block sum itself must
    fit int16 -- (2r+1)^2 * (min(2*ftzero, 255) + 63) <= 32767; beyond that cv2's `short` cost volume wraps
    (blockSize 21 with preFilterCap 63 can reach 83349) and no unsigned 16-bit engine follows it -- so preFilterCap shrinks
    as the block grows: 13 / 63, 15 / 31, 17 / 15 are the largest pairs inside the domain.  All four modes against the
    oracle (and live cv2)."""
    W, H, D = 420, 90, 32
    l, r, _ = make_pair(W, H, D, seed=bs)
    ftzero = max(cap, 15) | 1
    assert bs * bs * (min(2 * ftzero, 255) + 63) <= 32767
    for mode in (0, 1, 2, 3):
        p = OracleParams(0, D, bs, 600, 2400, 1, cap, 10, 100, 32, mode)
        ref = oracle.compute(p, l, r)
        assert _mismatch(sg.StereoSGBM_create(**_kw(p)).compute(l, r), ref) == 0, (bs, mode)
        if cv2_ref.available() and mode != 3:
            assert _mismatch(cv2_ref.compute(p, l, r), ref) == 0, ("cv2", bs, mode)


@pytest.mark.parametrize("D", [320, 512, 1024])
def test_many_disparities(sg, D):
    """numDisparities up to the supported maximum (1024): lane mappings with 16 or 32 lanes per column."""
    W, H = D + 180, 36
    l, r, _ = make_pair(W, H, D, seed=D)
    for mode in (0, 1, 2):
        if mode == 2 and H < 4 * (5 // 2 + 2):
            continue
        p = OracleParams(0, D, 5, 200, 800, 1, 63, 10, 50, 2, mode)
        assert _mismatch(sg.StereoSGBM_create(**_kw(p)).compute(l, r), oracle.compute(p, l, r)) == 0, (D, mode)


@pytest.mark.parametrize("H,W", [(1, 120), (2, 97), (3, 64), (5, 41), (37, 26)])
def test_tiny_and_odd_shapes(sg, H, W):
    """One to five rows, widths barely above numDisparities + blockSize / 2: every row is first and last row of its paths,
    strips are one or two columns wide.  (3WAY needs enough rows for its four stripes, so it only runs on the tallest.)"""
    D = 16
    l, r = make_noise_pair(W, H, seed=H * 100 + W)
    for mode in (0, 1, 3) + ((2,) if H >= 37 else ()):
        for bs in (1, 3, 5):
            p = OracleParams(0, D, bs, 8 * bs * bs, 32 * bs * bs, 1, 31, 5, 20, 2, mode)
            if oracle.port.valid_width(p, W) <= bs // 2:
                continue
            assert _mismatch(sg.StereoSGBM_create(**_kw(p)).compute(l, r), oracle.compute(p, l, r)) == 0, (H, W, mode, bs)


def test_one_handle_from_two_streams_and_two_threads(sg):
    """The workspace belongs to the handle, not to the stream: calls on one handle from different streams (no host
    synchronisation in between) and from different host threads must not overlap on it -- the library serialises the
    enqueue and orders each call behind the handle's previous one on the device.  Every result equals the oracle."""
    import threading
    import torch
    W, H, D = 1300, 200, 64
    pairs = [make_pair(W, H, D, seed=700 + i)[:2] for i in range(4)]
    p = OracleParams(0, D, 5, 200, 800, 1, 63, 10, 100, 32, 1)
    ref = [oracle.compute(p, l, r) for l, r in pairs]
    lt = [torch.from_numpy(l).cuda() for l, _ in pairs]
    rt = [torch.from_numpy(r).cuda() for _, r in pairs]
    st = sg.StereoSGBM_create(**_kw(p))
    streams = [torch.cuda.Stream() for _ in range(4)]
    torch.cuda.synchronize()
    for rep in range(5):
        outs = [None] * 4
        for i in range(4):                                  # back to back, each on its own stream, nothing waits in between
            with torch.cuda.stream(streams[i]):
                outs[i] = st.compute(lt[i], rt[i])
        torch.cuda.synchronize()
        for i in range(4):
            assert _mismatch(outs[i].cpu().numpy(), ref[i]) == 0, ("streams", rep, i)
    res, errs = [None] * 4, []

    def worker(i):
        try:
            for _ in range(6):
                with torch.cuda.stream(streams[i]):
                    res[i] = st.compute(lt[i], rt[i])
                streams[i].synchronize()
                assert _mismatch(res[i].cpu().numpy(), ref[i]) == 0
        except Exception as e:                              # pragma: no cover
            errs.append((i, e))

    th = [threading.Thread(target=worker, args=(i,)) for i in range(4)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs
