"""ctypes binding of libgslift.so (include/gslift.h).  No fallback: a missing library or a
non-zero return raises."""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgslift.so")
ABI_VERSION = 4

# Mirrors `GslView` in include/gslift.h.
VIEW_DTYPE = np.dtype(
    [
        ("R", np.float64, (9,)),
        ("t", np.float64, (3,)),
        ("fx", np.float64),
        ("fy", np.float64),
        ("half_w", np.float64),
        ("half_h", np.float64),
        ("width", np.float64),
        ("height", np.float64),
        ("scale_x", np.float64),
        ("scale_y", np.float64),
        ("seg_w", np.int32),
        ("seg_h", np.int32),
        ("map_offset", np.int64),
    ],
    align=True,
)
assert VIEW_DTYPE.itemsize == 176

# Every symbol include/gslift.h declares, with its ctypes signature.
_i64, _i32, _vp, _dbl, _sz = ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_double, ctypes.c_size_t
SIGNATURES = {
    "gsl_version": (_i32, []),
    "gsl_last_error": (ctypes.c_char_p, []),
    "gsl_device_count": (_i32, []),
    "gsl_launch_count": (ctypes.c_ulonglong, []),
    "gsl_packed_map_bytes": (_i64, [_i32, _i32]),
    "gsl_pack_labels": (_i32, [_vp, _i32, _i32, _i32, _vp, _i32, _i32, _vp, _vp]),
    "gsl_host_pack_labels": (_i32, [_vp, _vp, _i32, _i32, _i32, _vp, _i32, _vp, _vp]),
    "gsl_tile_codes": (_i32, [_vp, _i32, _i32, _i32, _vp, _vp]),
    "gsl_label_range": (_i32, [_vp, _i64, _vp, _vp]),
    "gsl_lift_workspace_bytes": (_sz, [_i64, _i32]),
    "gsl_lift_votes": (_i32, [_vp, _i64, _vp, _i32, _vp, _i32, _i32, _vp, _vp, _dbl, _vp, _sz, _vp]),
    "gsl_lift_prepare": (_i32, [_vp, _i64, _vp, _i32, _vp, _sz, _vp]),
    "gsl_lift_sweep": (_i32, [_vp, _i64, _vp, _i32, _vp, _i32, _i32, _vp, _vp, _vp, _sz, _vp]),
    "gsl_lift_gather": (_i32, [_vp, _i64, _vp, _i32, _vp, _vp, _sz, _vp]),
    "gsl_lift_gather_range": (_i32, [_vp, _i64, _vp, _i32, _i32, _i32, _vp, _vp, _sz, _vp]),
    "gsl_lift_majority": (_i32, [_i64, _i32, _i32, _i32, _vp, _vp, _vp, _sz, _vp]),
    "gsl_lift_merge": (_i32, [_vp, _vp, _vp, _vp, _i64, _vp]),
    "gsl_lift_near": (_i32, [_vp, _i64, _vp, _i32, _vp, _dbl, _vp, _sz, _vp]),
    "gsl_div_selftest": (_i32, [_vp, _vp, _vp, _i64, _vp, _vp]),
    "gsl_kmeans_workspace_bytes": (_sz, [_i64, _i32, _i32]),
    "gsl_kmeans_assign": (_i32, [_vp, _i64, _i32, _vp, _i32, _vp, _vp, _sz, _vp]),
    "gsl_kmeans_step": (_i32, [_vp, _i64, _i32, _vp, _i32, _vp, _vp, _vp, _sz, _vp]),
    "gsl_kmeans_finalize": (_i32, [_vp, _vp, _i32, _i32, _vp, _vp, _vp]),
    "gsl_kmeans_exchange_bytes": (_sz, [_i32, _i32, _i32]),
    "gsl_kmeans_step_exchange": (_i32, [_vp, _i64, _i32, _vp, _i32, _vp, _i32, _i32, _vp, ctypes.c_uint64, _vp, _vp, _vp, _vp, _sz, _vp]),
    "gsl_kmeans_update_ordered": (_i32, [_vp, _vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _sz, _vp]),
    "gsl_kmeans_screen_selftest": (_i32, [_vp, _i64, _i32, _vp, _i32, _vp, _vp, _vp]),
    "gsl_ply_format_ascii": (_i64, [_vp, _i64, _i32, _i32, _vp, _vp, _vp, _i64, _i32]),
    "gsl_recolor": (_i32, [_vp, _i64, _vp, _vp, _vp]),
    "gsl_viewer_sort_workspace_bytes": (_sz, [_i64]),
    "gsl_viewer_depth_sort": (_i32, [_vp, _i64, _i32, _vp, _vp, _vp, _sz, _vp]),
    "gsl_region_workspace_bytes": (_sz, [_i64]),
    "gsl_region_knn_pca": (_i32, [_vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "gsl_region_grow": (_i64, [_vp, _i32, _vp, _vp, _i64, _dbl, _dbl, _vp, _vp]),
    "gsl_viewer_hit_workspace_bytes": (_sz, []),
    "gsl_viewer_hit_test": (_i32, [_vp, _vp, _i64, _i32, _vp, _dbl, _dbl, _dbl, _dbl, _i32, _vp, _vp, _vp, _sz, _vp]),
}


class GslError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libgslift error {code}: {msg}")
        self.code = code


_lib = None


def lib() -> ctypes.CDLL:
    """Load libgslift.so (built in-tree by build.py).  Raises if it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  There is no CPU fallback."
            )
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        if L.gsl_version() != ABI_VERSION:
            raise ImportError(f"libgslift ABI {L.gsl_version()} != {ABI_VERSION}; rebuild")
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise GslError(rc, lib().gsl_last_error().decode("utf-8", "replace"))
