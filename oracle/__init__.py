"""CPU oracle for the dense-stereo hot path -- TEST INFRASTRUCTURE, not a fallback.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  It wraps oracle/liboracle.so (plain C restatement of SURVEY.md Appendix A,
see sgbm_oracle.c) and, when available, the installed cv2 binary that the reference itself calls
(main.ipynb:655-668, 697) -- see cv2_ref.py.
"""
from .port import (OracleParams, build, compute, compute_debug, median3x3, filter_speckles,
                   reproject_f32)

__all__ = ["OracleParams", "build", "compute", "compute_debug", "median3x3", "filter_speckles",
           "reproject_f32"]
