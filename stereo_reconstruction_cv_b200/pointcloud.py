"""Point-cloud output formats of the reference's 3-D tab.

    o3d.geometry.PointCloud + Vector3dVector(points) / colors / 255.0   main.ipynb:739-747
    o3d.io.write_point_cloud(..., points_3D.reshape(-1, 3))              main.ipynb:795-797

Open3D is not a dependency here: `open3d_arrays` returns exactly the arrays the reference hands to
Vector3dVector (float64 N x 3 points, float64 N x 3 colours in [0, 1]) and `write_ply` writes the
file Open3D would write for them (binary little-endian PLY, double x/y/z, uchar red/green/blue).
The reference's second call site writes ALL H*W points including non-finite ones
(main.ipynb:796); `write_ply(..., keep_nonfinite=True)` mirrors that, the default drops them the
way the first call site's mask does (main.ipynb:726-737).
"""
import numpy as np


def _host(a):
    if a is None:
        return None
    if type(a).__module__.startswith("torch"):
        a = a.detach().cpu().numpy()
    return np.asarray(a)


def open3d_arrays(xyz, rgb=None):
    """(points float64 N x 3, colors float64 N x 3 in [0,1] or None) as passed to o3d.utility.Vector3dVector."""
    pts = _host(xyz).reshape(-1, 3).astype(np.float64)
    col = None
    if rgb is not None:
        col = _host(rgb).reshape(-1, 3).astype(np.float64) / 255.0          # main.ipynb:744
    return pts, col


def write_ply(path, xyz, rgb=None, keep_nonfinite=False, dtype=np.float64):
    """Binary little-endian PLY of an N x 3 (or H x W x 3) cloud, optional uint8 colours.  Returns N written."""
    pts = _host(xyz).reshape(-1, 3)
    col = _host(rgb).reshape(-1, 3).astype(np.uint8) if rgb is not None else None
    if col is not None and col.shape[0] != pts.shape[0]:
        raise ValueError("xyz and rgb disagree on the number of points")
    if not keep_nonfinite:
        ok = np.isfinite(pts).all(axis=1)
        pts = pts[ok]
        col = col[ok] if col is not None else None
    dtype = np.dtype(dtype)
    if dtype not in (np.dtype(np.float32), np.dtype(np.float64)):
        raise ValueError("dtype must be float32 or float64")
    tname = "double" if dtype == np.dtype(np.float64) else "float"
    fields = [("x", "<" + dtype.str[1:]), ("y", "<" + dtype.str[1:]), ("z", "<" + dtype.str[1:])]
    header = ["ply", "format binary_little_endian 1.0", "comment stereo_reconstruction_cv_b200",
              "element vertex %d" % pts.shape[0], "property %s x" % tname, "property %s y" % tname,
              "property %s z" % tname]
    if col is not None:
        fields += [("red", "u1"), ("green", "u1"), ("blue", "u1")]
        header += ["property uchar red", "property uchar green", "property uchar blue"]
    header.append("end_header")
    rec = np.empty(pts.shape[0], dtype=np.dtype(fields))
    rec["x"], rec["y"], rec["z"] = pts[:, 0], pts[:, 1], pts[:, 2]
    if col is not None:
        rec["red"], rec["green"], rec["blue"] = col[:, 0], col[:, 1], col[:, 2]
    with open(path, "wb") as f:
        f.write(("\n".join(header) + "\n").encode("ascii"))
        f.write(rec.tobytes())
    return int(pts.shape[0])


def read_ply(path):
    """Reader for the files write_ply produces (tests and round trips): (xyz, rgb or None)."""
    with open(path, "rb") as f:
        lines = []
        while True:
            ln = f.readline().decode("ascii").strip()
            lines.append(ln)
            if ln == "end_header":
                break
        n = int([l for l in lines if l.startswith("element vertex")][0].split()[-1])
        props = [l.split()[1:] for l in lines if l.startswith("property")]
        m = {"double": "<f8", "float": "<f4", "uchar": "u1"}
        dt = np.dtype([(name, m[t]) for t, name in props])
        rec = np.frombuffer(f.read(n * dt.itemsize), dtype=dt, count=n)
    xyz = np.stack([rec["x"], rec["y"], rec["z"]], 1)
    rgb = np.stack([rec["red"], rec["green"], rec["blue"]], 1) if "red" in rec.dtype.names else None
    return xyz, rgb
