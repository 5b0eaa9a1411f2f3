"""The reference's dense-reconstruction call sites as functions, plus the Tk tab the README advertises
but gui.py never got (README.md:81-83, 103-104; gui.py:367-377 stops at Tab 4).

    compute_disparity_map(imgL, imgR, ndisp, mindis)      main.ipynb:655-670  (same name, arguments, result)
    reconstruct_3D(disparity_map, Q)                      main.ipynb:697      (returns None on error, like the notebook)
    point_cloud_arrays(points_3D, colors, disparity_map)  main.ipynb:726-747  (the mask + gather in front of Open3D)
    rectify_pair(imgL, imgR, K1, K2, R1, R2, P1, P2)      main.ipynb:496-500 / gui.py:160-164
    create_disparity_tab(gui)                             "Tab 6": Run Disparity / Visualize 3D Point Cloud

All numerics run on the GPU through the C ABI (stereo.py); nothing here imports cv2.  Tk and Open3D are
imported lazily and only by the functions that need a display.
"""
import numpy as np

from . import pointcloud
from .stereo import (MODE_SGBM, StereoSGBM_create, initUndistortRectifyMap, remap, reprojectCompact,
                     reprojectImageTo3D)

# parameters of the notebook's call (main.ipynb:655-666); blockSize 11 with the 3-channel P1/P2 recipe
NOTEBOOK_PARAMS = dict(blockSize=11, P1=8 * 3 * 11 ** 2, P2=32 * 3 * 11 ** 2, disp12MaxDiff=1, preFilterCap=63,
                       uniquenessRatio=10, speckleWindowSize=100, speckleRange=32, mode=MODE_SGBM)


def compute_disparity_map(imgL, imgR, ndisp, mindis, **overrides):
    """float32 H x W disparity in pixels, non-positive values zeroed (main.ipynb:655-670)."""
    kw = dict(NOTEBOOK_PARAMS, minDisparity=mindis, numDisparities=ndisp)
    kw.update(overrides)
    stereo = StereoSGBM_create(**kw)
    disparity_map = stereo.compute(np.ascontiguousarray(imgL), np.ascontiguousarray(imgR)).astype(np.float32) / 16.0
    mask = disparity_map > 0
    return disparity_map * mask.astype(np.float32)


def reconstruct_3D(disparity_map, Q):
    """H x W x 3 float32 points (main.ipynb:697); None on error, as the notebook's wrapper returns."""
    try:
        return reprojectImageTo3D(disparity_map, Q)
    except Exception as e:  # the notebook catches everything here (main.ipynb:696-701)
        print("Error in reconstruct_3D: %s" % e)
        return None


def point_cloud_arrays(points_3D, colors, disparity_map):
    """(valid_points N x 3 float32, valid_colors N x 3 uint8) exactly as main.ipynb:726-737 selects them."""
    points_3D = np.asarray(points_3D)
    mask = ~np.isnan(points_3D[:, :, 0]) & ~np.isinf(points_3D[:, :, 0]) & (np.asarray(disparity_map) > 0)
    return points_3D[mask], np.asarray(colors)[mask]


def dense_cloud(disparity_x16, Q, colors_bgr=None):
    """The same cloud straight from the int16 (x16) disparity on the GPU: /16, > 0 mask, reprojection,
    finite test and ordered compaction fused in one pass (sgbm_reproject_compact).  Colours come back RGB."""
    return reprojectCompact(disparity_x16, Q, colors_bgr)


def rectify_pair(imgL, imgR, K1, K2, R1, R2, P1, P2, size=None):
    """Both images through initUndistortRectifyMap + remap(INTER_LINEAR) (main.ipynb:496-500, distortion None)."""
    h, w = imgL.shape[:2]
    size = size or (w, h)
    m1x, m1y = initUndistortRectifyMap(K1, None, R1, P1, size)
    m2x, m2y = initUndistortRectifyMap(K2, None, R2, P2, size)
    return remap(imgL, m1x, m1y), remap(imgR, m2x, m2y)


def save_point_cloud(path, points_3D, colors=None, all_points=True):
    """The notebook's optional save (main.ipynb:795-797 writes ALL H*W points, non-finite ones included)."""
    return pointcloud.write_ply(path, points_3D, colors, keep_nonfinite=all_points)


def create_disparity_tab(gui, ndisp=16, mindis=0):
    """Add "Tab 6" to a gui.NotebookGUI-like object (needs .notebook and .stereo_rect_results).

    stereo_rect_results must carry the rectified pair and Q -- the reference's stereo_rect computes them
    and drops them (gui.py:204-209); keep them under the keys "Rectified Left", "Rectified Right", "Q"
    (optionally "Color Left").  Returns the tab frame; results land in gui.disparity_results.
    """
    import tkinter as tk
    from tkinter import messagebox, ttk

    tab = ttk.Frame(gui.notebook)
    gui.notebook.add(tab, text="Disparity / Dense 3D")
    nd_var, md_var = tk.IntVar(value=ndisp), tk.IntVar(value=mindis)
    row = ttk.Frame(tab)
    row.pack(pady=10)
    ttk.Label(row, text="numDisparities").pack(side="left")
    ttk.Entry(row, textvariable=nd_var, width=6).pack(side="left", padx=5)
    ttk.Label(row, text="minDisparity").pack(side="left")
    ttk.Entry(row, textvariable=md_var, width=6).pack(side="left", padx=5)
    out = tk.Text(tab, height=8, width=100)
    gui.disparity_results = None

    def rectified():
        r = getattr(gui, "stereo_rect_results", None) or {}
        missing = [k for k in ("Rectified Left", "Rectified Right", "Q") if k not in r]
        if missing:
            messagebox.showerror("Error", "Run Stereo Rectification first (missing: %s)" % ", ".join(missing))
            return None
        return r

    def run_disparity():
        r = rectified()
        if r is None:
            return
        try:
            d = compute_disparity_map(r["Rectified Left"], r["Rectified Right"], nd_var.get(), md_var.get())
        except Exception as e:
            messagebox.showerror("Error", str(e))
            return
        gui.disparity_results = {"Disparity": d}
        out.delete("1.0", "end")
        out.insert("end", "disparity %dx%d, valid %.1f %%, range %.2f .. %.2f px\n"
                   % (d.shape[1], d.shape[0], 100.0 * float((d > 0).mean()), float(d[d > 0].min()) if (d > 0).any() else 0.0,
                      float(d.max())))

    def visualize_3d():
        r = rectified()
        if r is None:
            return
        if not gui.disparity_results:
            messagebox.showerror("Error", "Run Disparity first")
            return
        d = gui.disparity_results["Disparity"]
        pts = reconstruct_3D(d, np.asarray(r["Q"], np.float64))
        if pts is None:
            return
        col = r.get("Color Left")
        col = np.repeat(np.asarray(r["Rectified Left"])[:, :, None], 3, 2) if col is None else np.asarray(col)[:, :, ::-1]
        vp, vc = point_cloud_arrays(pts, col, d)
        gui.disparity_results.update({"Points": vp, "Colors": vc})
        out.insert("end", "%d valid 3-D points\n" % len(vp))
        try:
            import open3d as o3d
            pc = o3d.geometry.PointCloud()
            p64, c64 = pointcloud.open3d_arrays(vp, vc)
            pc.points = o3d.utility.Vector3dVector(p64)
            pc.colors = o3d.utility.Vector3dVector(c64)
            o3d.visualization.draw_geometries([pc])
        except ImportError:
            out.insert("end", "open3d is not installed: use save_point_cloud() / write_ply() instead\n")

    ttk.Button(tab, text="Run Disparity", command=run_disparity).pack(pady=5)
    ttk.Button(tab, text="Visualize 3D Point Cloud", command=visualize_3d).pack(pady=5)
    out.pack(pady=10)
    return tab
