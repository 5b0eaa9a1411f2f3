"""the notebook's dense-reconstruction functions / the README's "Tab 6" on top of the engine."""
import numpy as np
import pytest

import oracle
from oracle import OracleParams
from synth import make_pair


def test_point_cloud_arrays_is_the_notebook_mask():
    from stereo_reconstruction_cv_b200 import disparity_tab as dt
    rng = np.random.default_rng(3)
    pts = rng.normal(size=(6, 7, 3)).astype(np.float32)
    pts[1, 2, 0] = np.inf; pts[3, 3, 0] = np.nan; pts[4, 1, 1] = np.inf          # only X is tested (main.ipynb:727-731)
    disp = rng.uniform(-1, 3, size=(6, 7)).astype(np.float32)
    col = rng.integers(0, 256, size=(6, 7, 3), dtype=np.uint8)
    vp, vc = dt.point_cloud_arrays(pts, col, disp)
    mask = ~np.isnan(pts[:, :, 0]) & ~np.isinf(pts[:, :, 0]) & (disp > 0)
    assert np.array_equal(vp, pts[mask], equal_nan=True) and np.array_equal(vc, col[mask])
    assert dt.NOTEBOOK_PARAMS["blockSize"] == 11 and dt.NOTEBOOK_PARAMS["P2"] == 32 * 3 * 121
    assert callable(dt.create_disparity_tab) and callable(dt.rectify_pair)


@pytest.mark.gpu
def test_notebook_functions_on_gpu(tmp_path):
    from stereo_reconstruction_cv_b200 import disparity_tab as dt
    W, H, D = 360, 120, 16
    l, r, _ = make_pair(W, H, D, seed=4)
    d = dt.compute_disparity_map(l, r, D, 0)                                   # main.ipynb:781 calls it with (16, 0)
    p = OracleParams(0, D, 11, 8 * 3 * 121, 32 * 3 * 121, 1, 63, 10, 100, 32, 0)
    ref = oracle.compute(p, l, r).astype(np.float32) / 16.0
    ref = ref * (ref > 0).astype(np.float32)
    assert d.dtype == np.float32 and np.array_equal(d, ref)
    Q = np.array([[1, 0, 0, -W / 2], [0, 1, 0, -H / 2], [0, 0, 0, 300.0], [0, 0, -1, 0]], np.float64)
    pts = dt.reconstruct_3D(d, Q)
    assert pts.shape == (H, W, 3) and np.array_equal(pts.view(np.uint32), oracle.reproject_f32(d, Q).view(np.uint32))
    col = np.repeat(l[:, :, None], 3, 2)
    vp, vc = dt.point_cloud_arrays(pts, col, d)
    assert len(vp) == int((np.isfinite(pts[:, :, 0]) & (d > 0)).sum()) > 0
    n = dt.save_point_cloud(str(tmp_path / "c.ply"), pts)
    assert n == W * H                                                          # the notebook writes every pixel
    assert dt.reconstruct_3D(d, np.eye(3)) is None                             # Q must be 4x4: error -> None


# ------------------------------------------------------------------------------------------------
# create_disparity_tab executed against a fake Tk (no display, tkinter itself is not even installed here)
# ------------------------------------------------------------------------------------------------
class _FakeTk:
    """Just enough of tkinter / ttk / messagebox for create_disparity_tab: widgets record what they are
    given, buttons keep their command so the test can press them."""

    def __init__(self):
        import types
        self.buttons, self.errors, self.tabs = {}, [], []
        fk = self

        class Widget:
            def __init__(self, parent=None, **kw):
                self.parent, self.kw, self.packed = parent, kw, False
                if "command" in kw:
                    fk.buttons[kw.get("text")] = kw["command"]

            def pack(self, **kw):
                self.packed = True

        class Text(Widget):
            def __init__(self, parent=None, **kw):
                super().__init__(parent, **kw)
                self.content = ""

            def insert(self, where, s):
                self.content += s

            def delete(self, a, b):
                self.content = ""

        class IntVar:
            def __init__(self, value=0):
                self.v = value

            def get(self):
                return self.v

            def set(self, v):
                self.v = v

        class Notebook(Widget):
            def add(self, tab, text=""):
                fk.tabs.append((tab, text))

        self.Notebook, self.TextCls, self.IntVarCls = Notebook, Text, IntVar
        self.tk = types.ModuleType("tkinter")
        self.tk.Text, self.tk.IntVar = Text, IntVar
        self.ttk = types.ModuleType("tkinter.ttk")
        for n in ("Frame", "Label", "Entry", "Button"):
            setattr(self.ttk, n, type(n, (Widget,), {}))
        self.mb = types.ModuleType("tkinter.messagebox")
        self.mb.showerror = lambda title, msg: fk.errors.append(msg)
        self.tk.ttk, self.tk.messagebox = self.ttk, self.mb

    def install(self, monkeypatch):
        import sys
        monkeypatch.setitem(sys.modules, "tkinter", self.tk)
        monkeypatch.setitem(sys.modules, "tkinter.ttk", self.ttk)
        monkeypatch.setitem(sys.modules, "tkinter.messagebox", self.mb)


class _FakeGui:
    """Shaped like gui.NotebookGUI (gui.py:326-377): .notebook and the per-tab result slots."""

    def __init__(self, notebook):
        self.notebook = notebook
        self.stereo_rect_results = None


def test_create_disparity_tab_wiring_without_gpu(monkeypatch):
    """Tab 6 of README.md:81-83: both buttons exist, refuse to run before Tab 2's results are there."""
    from stereo_reconstruction_cv_b200 import disparity_tab as dt
    fk = _FakeTk()
    fk.install(monkeypatch)
    gui = _FakeGui(fk.Notebook())
    tab = dt.create_disparity_tab(gui)
    assert fk.tabs == [(tab, "Disparity / Dense 3D")]
    assert set(fk.buttons) == {"Run Disparity", "Visualize 3D Point Cloud"}
    fk.buttons["Run Disparity"]()
    assert gui.disparity_results is None and "Run Stereo Rectification first" in fk.errors[-1]
    fk.buttons["Visualize 3D Point Cloud"]()
    assert len(fk.errors) == 2
    gui.stereo_rect_results = {"Rectified Left": np.zeros((4, 4), np.uint8)}
    fk.buttons["Run Disparity"]()
    assert "Rectified Right, Q" in fk.errors[-1]


@pytest.mark.gpu
def test_create_disparity_tab_runs_the_dense_path(monkeypatch):
    """Press "Run Disparity" and "Visualize 3D Point Cloud" (README.md:81-83, 103-104) on Tab 2's output:
    the disparity equals the oracle's at the notebook's parameters, the cloud is the notebook's mask + gather."""
    import sys
    from stereo_reconstruction_cv_b200 import disparity_tab as dt
    fk = _FakeTk()
    fk.install(monkeypatch)
    monkeypatch.setitem(sys.modules, "open3d", None)                           # "not installed": the tab falls back to a message
    W, H, D = 400, 128, 32
    l, r, _ = make_pair(W, H, D, seed=8)
    Q = np.array([[1, 0, 0, -W / 2], [0, 1, 0, -H / 2], [0, 0, 0, 350.0], [0, 0, -1, 0]], np.float64)
    col = np.stack([l, r, l // 2 + r // 2], -1).astype(np.uint8)               # BGR
    gui = _FakeGui(fk.Notebook())
    gui.stereo_rect_results = {"Rectified Left": l, "Rectified Right": r, "Q": Q, "Color Left": col}
    dt.create_disparity_tab(gui, ndisp=D, mindis=0)
    fk.buttons["Visualize 3D Point Cloud"]()
    assert "Run Disparity first" in fk.errors[-1]
    fk.buttons["Run Disparity"]()
    assert not fk.errors[1:], fk.errors
    d = gui.disparity_results["Disparity"]
    p = OracleParams(0, D, 11, 8 * 3 * 121, 32 * 3 * 121, 1, 63, 10, 100, 32, 0)
    ref = oracle.compute(p, l, r).astype(np.float32) / 16.0
    ref = ref * (ref > 0).astype(np.float32)
    assert np.array_equal(d, ref)
    fk.buttons["Visualize 3D Point Cloud"]()
    pts = oracle.reproject_f32(ref, Q)
    mask = np.isfinite(pts[:, :, 0]) & (ref > 0)
    assert np.array_equal(gui.disparity_results["Points"].view(np.uint32), pts[mask].view(np.uint32))
    assert np.array_equal(gui.disparity_results["Colors"], col[:, :, ::-1][mask])
    # a failing compute (width too small for numDisparities) ends in a message box, not an exception
    gui.stereo_rect_results = {"Rectified Left": l[:, :20], "Rectified Right": r[:, :20], "Q": Q}
    fk.buttons["Run Disparity"]()
    assert "too small" in fk.errors[-1]
