"""B200-native dense-stereo engine: drop-in for the cv2 calls on the reference's dense
reconstruction path (main.ipynb:655-668, 697).  See stereo.py and include/sgbm_b200.h."""
from .stereo import (DISP_SCALE, DISP_SHIFT, MODE_HH, MODE_HH4, MODE_SGBM, MODE_SGBM_3WAY, StereoSGBM,
                     StereoSGBM_create, device_info, disparityToFloat, error, filterSpeckles, medianBlur3,
                     microbench_int16, reprojectCompact, reprojectImageTo3D, initUndistortRectifyMap, remap,
                     INTER_LINEAR, CV_32F, CV_32FC1)

from . import disparity_tab, pointcloud, sharding  # noqa: E402,F401
from .pointcloud import open3d_arrays, write_ply  # noqa: E402,F401
from .sharding import compute_shard, gather_point_cloud, shard_range  # noqa: E402,F401

__all__ = ["initUndistortRectifyMap", "remap", "INTER_LINEAR", "CV_32F", "CV_32FC1", "pointcloud", "sharding", "write_ply", "open3d_arrays", "shard_range", "compute_shard", "gather_point_cloud",
           "StereoSGBM", "StereoSGBM_create", "reprojectImageTo3D", "reprojectCompact", "disparityToFloat",
           "filterSpeckles", "medianBlur3", "microbench_int16", "device_info", "error", "MODE_SGBM", "MODE_HH",
           "MODE_SGBM_3WAY", "MODE_HH4", "DISP_SHIFT", "DISP_SCALE"]
