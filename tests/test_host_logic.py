"""CPU tests of the host-side planning logic of the C-ABI library (no device needed): the strip / ring schedule of the
persistent sweeps (sgbm_sweep.cu: sweep_plan) must respect the limits the kernels rely on for EVERY geometry, not only
the ones the GPU tests happen to visit -- shared memory and thread budgets, super-steps that are multiples of the ring
stage height, the depth of the halo ring against the lead a publishing strip can build up, strips that cover the image."""
import ctypes as C
import itertools

import pytest


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    from stereo_reconstruction_cv_b200 import _lib
    return _lib


NAMES = ("found", "nstrips", "SW", "R", "NB", "rps", "K", "NSC", "NSI", "nwV", "nwA", "nwW", "wRG", "wPR", "threads", "smem")
MAX_SMEM = 232448          # 227 KB: sm_100 opt-in maximum per CTA


def plan(lib, W, H, D, mode, sms, wrole, nab, bs=5, smem=MAX_SMEM):
    p = lib.SgbmParams(0, D, bs, 8 * bs * bs, 32 * bs * bs, 1, 63, 10, 100, 32, mode)
    out = (C.c_int * 16)()
    rc = lib.lib().sgbm_debug_sweep_plan(C.byref(p), W, H, 1, sms, smem, wrole, nab, out)
    assert rc == 0, lib.lib().sgbm_last_error()
    return dict(zip(NAMES, out))


def check(pl, W, D, wrole, sms):
    W1 = W - D
    assert 1 <= pl["nstrips"] <= sms
    assert pl["SW"] * pl["nstrips"] >= W1, "strips cover the valid columns"
    assert pl["smem"] <= MAX_SMEM and pl["threads"] <= 1024 and pl["threads"] % 32 == 0
    assert pl["rps"] in (1, 2) and pl["R"] >= 1 and pl["R"] % pl["rps"] == 0, "a ring stage never straddles two super-steps"
    assert pl["NB"] * pl["R"] >= pl["SW"] + pl["R"] - 1, "every chain leaves the strip before its batch restarts"
    assert pl["NSC"] >= 2 and pl["NSI"] >= 2 and pl["K"] >= (3 if wrole else 2)
    if pl["rps"] == 2:
        assert pl["NSC"] >= 4, "two rows per stage are only taken with a deep cost ring"
    # halo ring: 64 column entries per strip; a publishing role can lead the neighbour's consuming role by the S ring
    # (in rows) plus one super-step (plus one for granularity) and must not lap it
    assert pl["K"] * pl["rps"] + 2 * pl["R"] < 64 or pl["nstrips"] == 1
    if wrole:
        assert 1 <= pl["wPR"] <= 7 and pl["wRG"] >= 1 and pl["nwW"] == pl["wPR"] * pl["wRG"] <= 7
        assert pl["wRG"] <= max(pl["K"] - 2, 1), "the path roles keep two S slots to themselves"
        r4 = lambda v: (v + 3) & ~3                      # every role a whole number of warpgroups (setmaxnreg)
        assert pl["threads"] == (r4(pl["nwV"]) + 2 * r4(pl["nwA"]) + 8) * 32
    else:
        assert pl["threads"] == (pl["nwV"] + 2 * pl["nwA"] + 1) * 32


@pytest.mark.parametrize("wrole,nab", [(1, 2), (1, 1), (0, 2), (0, 1)])
def test_sweep_plan_invariants(lib, wrole, nab):
    found = 0
    for (W, H), D, sms in itertools.product([(640, 480), (1280, 720), (1920, 1080), (2560, 1440), (3840, 2160), (5000, 300), (333, 77)],
                                            [16, 20, 32, 64, 100, 128, 192, 256], [148, 74, 49, 37]):
        if W - D < 64:
            continue
        pl = plan(lib, W, H, D, 0, sms, wrole, nab)
        if pl["found"]:
            found += 1
            check(pl, W, D, wrole, sms)
    assert found > 100


def test_sweep_plan_known_geometries(lib):
    """The schedules of the BASELINE configurations (what DESIGN.md describes)."""
    cfg3 = plan(lib, 3840, 2160, 256, 1, 148, 1, 1)
    assert (cfg3["nstrips"], cfg3["SW"], cfg3["R"], cfg3["rps"], cfg3["threads"]) == (148, 25, 8, 1, 1024)
    cfg2 = plan(lib, 1280, 720, 128, 0, 148, 1, 2)
    assert (cfg2["nstrips"], cfg2["SW"], cfg2["R"], cfg2["rps"]) == (144, 8, 8, 2) and cfg2["NSC"] >= 5
    cfg1 = plan(lib, 3840, 2160, 16, 0, 148, 1, 2, bs=11)
    assert (cfg1["R"], cfg1["rps"]) == (16, 2)
    cfg4 = plan(lib, 1920, 1080, 192, 0, 148, 1, 2)
    assert cfg4["rps"] == 2 and cfg4["NSC"] >= 4
    # too wide for the persistent sweep at 8 lanes per column: the caller falls back (k_vertical / k_rowstep)
    assert plan(lib, 7680, 400, 256, 0, 148, 1, 2)["found"] == 0
    # little shared memory: one row per stage, shallower rings, or nothing
    small = plan(lib, 1280, 720, 128, 0, 148, 1, 2, smem=48 * 1024)
    assert small["found"] == 0 or small["smem"] <= 48 * 1024


def test_parameter_validation_without_a_device(lib):
    """make_geo (the same function sgbm_compute runs first) through the planner hook: what is rejected, and with which code."""
    def rc(W=640, H=48, cn=1, **kw):
        base = dict(minDisparity=0, numDisparities=64, blockSize=5, P1=200, P2=800, disp12MaxDiff=1, preFilterCap=63,
                    uniquenessRatio=10, speckleWindowSize=0, speckleRange=0, mode=0)
        base.update(kw)
        p = lib.SgbmParams(*[base[n] for n in ("minDisparity", "numDisparities", "blockSize", "P1", "P2", "disp12MaxDiff",
                                               "preFilterCap", "uniquenessRatio", "speckleWindowSize", "speckleRange", "mode")])
        out = (C.c_int * 16)()
        return lib.lib().sgbm_debug_sweep_plan(C.byref(p), W, H, cn, 148, MAX_SMEM, 1, 2, out)
    assert rc() == 0
    for D in (4, 5, 20, 21, 250, 255, 1024):                      # any numDisparities >= 4 outside MODE_SGBM_3WAY
        assert rc(W=1400, numDisparities=D) == 0, D
    assert rc(numDisparities=20, mode=2) != 0 and rc(numDisparities=24, mode=2) == 0
    for bad in (dict(numDisparities=0), dict(numDisparities=-16), dict(numDisparities=3), dict(numDisparities=1032, W=2000),
                dict(W=60), dict(W=0), dict(H=0), dict(cn=2), dict(mode=7), dict(P2=40000), dict(uniquenessRatio=101),
                dict(uniquenessRatio=100, mode=2), dict(minDisparity=3000), dict(numDisparities=20, blockSize=11, P1=2904, P2=22000)):
        assert rc(**bad) != 0, bad
    assert b"numDisparities" in lib.lib().sgbm_last_error()
