"""The reference's own implementation of the path: the installed cv2 binary.  Test infra only.

The reference calls cv2.StereoSGBM_create(...).compute (main.ipynb:655-668) and
cv2.reprojectImageTo3D (main.ipynb:697); OpenCV is a pip dependency (environment.yml:89-90), its
sources are not under /root/reference, so the 'real reference' arm is this binary, imported from
site-packages (it is part of the image, so it is also present on the GPU box).
"""
import numpy as np


def available():
    try:
        import cv2  # noqa: F401
        return True
    except Exception:
        return False


def compute(p, left, right, threads=None):
    import cv2
    if threads is not None:
        cv2.setNumThreads(int(threads))
    st = cv2.StereoSGBM_create(minDisparity=p.minDisparity, numDisparities=p.numDisparities,
                               blockSize=p.blockSize, P1=p.P1, P2=p.P2, disp12MaxDiff=p.disp12MaxDiff,
                               preFilterCap=p.preFilterCap, uniquenessRatio=p.uniquenessRatio,
                               speckleWindowSize=p.speckleWindowSize, speckleRange=p.speckleRange,
                               mode=p.mode)
    return st.compute(np.ascontiguousarray(left), np.ascontiguousarray(right))


def reproject(disp, Q):
    import cv2
    return cv2.reprojectImageTo3D(disp, np.asarray(Q))


def version():
    import cv2
    return cv2.__version__
