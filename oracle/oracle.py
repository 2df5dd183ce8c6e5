"""CPU oracle for the hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product package never does.

It wraps ``oracle/_build/libgsl_oracle.so`` (built from ``gsl_oracle.c`` by ``oracle/Makefile``)
and adds two slow, independent pure-Python restatements used to cross-check the C code on
small cases.  All ``file:line`` citations are into ``/root/reference``.

Parity pin: the reference ships no tests or golden vectors for this path.
The pins are the fixtures under ``tests/golden/`` produced by ``oracle/make_golden.py``,
which runs the reference's own functions verbatim in the build container.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from fractions import Fraction

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libgsl_oracle.so")

# Mirrors `OrcView` in gsl_oracle.c.
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


def build(force: bool = False) -> str:
    """Compile the C oracle if it is missing (or `force`)."""
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith(".c")] + [os.path.join(_HERE, "Makefile")]
    if force or not os.path.exists(_LIB_PATH) or max(os.path.getmtime(f) for f in srcs) > os.path.getmtime(_LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        i64, i32, vp, dbl = ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_double
        L.orc_version.restype = i32
        L.orc_max_threads.restype = i32
        L.orc_set_threads.argtypes = [i32]
        L.orc_set_threads.restype = None
        L.orc_translation.argtypes = [vp, vp, vp]
        L.orc_translation.restype = None
        L.orc_project.argtypes = [vp, vp, vp, vp]
        L.orc_project.restype = i32
        L.orc_lift_votes.argtypes = [vp, i64, vp, i32, vp, i32, i32, vp, vp, dbl, vp]
        L.orc_lift_votes.restype = i32
        L.orc_sqdist.argtypes = [vp, vp, i32]
        L.orc_sqdist.restype = dbl
        L.orc_kmeans_assign.argtypes = [vp, i64, i32, vp, i32, vp, vp]
        L.orc_kmeans_assign.restype = None
        L.orc_kmeans_update.argtypes = [vp, vp, i64, i32, i32, vp, vp, vp]
        L.orc_kmeans_update.restype = None
        L.orc_kmeans_update_f64.argtypes = [vp, vp, i64, i32, i32, vp, vp]
        L.orc_kmeans_update_f64.restype = None
        L.orc_viewer_depth_sort.argtypes = [vp, i64, i32, vp, vp]
        L.orc_viewer_depth_sort.restype = None
        L.orc_viewer_buckets.argtypes = [vp, i64, i32, vp, vp]
        L.orc_viewer_buckets.restype = None
        L.orc_multiply4.argtypes = [vp, vp, vp]
        L.orc_multiply4.restype = None
        L.orc_viewer_hit_test.argtypes = [vp, vp, i64, i32, vp, dbl, dbl, dbl, dbl, i32, vp]
        L.orc_viewer_hit_test.restype = i32
        L.orc_region_knn_pca.argtypes = [vp, i64, i32, vp, vp, vp, vp, vp, vp]
        L.orc_region_knn_pca.restype = None
        L.orc_region_knn_pca_range.argtypes = [vp, i64, i32, vp, vp, vp, vp, vp, vp, i64, i64]
        L.orc_region_knn_pca_range.restype = None
        L.orc_region_grow.argtypes = [vp, i32, vp, vp, i64, dbl, dbl, vp]
        L.orc_region_grow.restype = i64
        _lib = L
    return _lib


def max_threads() -> int:
    return int(lib().orc_max_threads())


def set_threads(n: int | None = None) -> int:
    """Use `n` OpenMP threads (default: every core this process may run on, whatever
    OMP_NUM_THREADS says -- torchrun sets it to 1).  Returns the count in effect."""
    if n is None:
        try:
            n = len(os.sched_getaffinity(0))
        except (AttributeError, OSError):
            n = os.cpu_count() or 1
    lib().orc_set_threads(int(n))
    return max_threads()


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


# --------------------------------------------------------------------------------------
# lifting
# --------------------------------------------------------------------------------------
def translation(R, p) -> np.ndarray:
    """t = -R @ p with the probed dgemv rounding (deep_learning_segmentation.py:66)."""
    R = np.ascontiguousarray(R, np.float64).reshape(9)
    p = np.ascontiguousarray(p, np.float64).reshape(3)
    t = np.empty(3, np.float64)
    lib().orc_translation(_ptr(R), _ptr(p), _ptr(t))
    return t


def make_views(cameras, map_shapes, image_sizes=None) -> np.ndarray:
    """Build the oracle's view table from reference camera dicts.

    cameras      list of dicts with the cameras.json fields (deep_learning_segmentation.py:54-63)
    map_shapes   per view (seg_h, seg_w)                          (:267)
    image_sizes  per view (orig_w, orig_h) of the opened image    (:263); default = map size
    Maps are assumed concatenated in view order (map_offset = running sum of seg_h*seg_w).
    """
    V = len(cameras)
    views = np.zeros(V, VIEW_DTYPE)
    off = 0
    for v, cam in enumerate(cameras):
        R = np.array(cam["rotation"], np.float64)
        p = np.array(cam["position"], np.float64)
        seg_h, seg_w = (int(s) for s in map_shapes[v])
        ow, oh = (seg_w, seg_h) if image_sizes is None else (int(s) for s in image_sizes[v])
        views[v]["R"] = R.reshape(9)
        views[v]["t"] = translation(R, p)
        views[v]["fx"] = cam["fx"]
        views[v]["fy"] = cam["fy"]
        views[v]["half_w"] = cam["width"] / 2
        views[v]["half_h"] = cam["height"] / 2
        views[v]["width"] = cam["width"]
        views[v]["height"] = cam["height"]
        views[v]["scale_x"] = seg_w / ow
        views[v]["scale_y"] = seg_h / oh
        views[v]["seg_w"] = seg_w
        views[v]["seg_h"] = seg_h
        views[v]["map_offset"] = off
        off += seg_h * seg_w
    return views


def lift_votes(pos, views, maps, label_min=-1, n_classes=255, eps=0.0, want_near=False):
    """assign_labels (deep_learning_segmentation.py:241-308) given the segmentation maps.

    Returns (labels int32[N], near uint8[N] or None, visible_pairs int).
    """
    pos = np.ascontiguousarray(pos, np.float32).reshape(-1, 3)
    maps = np.ascontiguousarray(maps, np.int32).reshape(-1)
    views = np.ascontiguousarray(views)
    assert views.dtype == VIEW_DTYPE
    N = pos.shape[0]
    labels = np.empty(N, np.int32)
    near = np.zeros(N, np.uint8) if want_near else None
    vis = np.zeros(1, np.int64)
    rc = lib().orc_lift_votes(
        _ptr(pos), N, _ptr(views), len(views), _ptr(maps), int(label_min), int(n_classes),
        _ptr(labels), _ptr(near) if want_near else None, float(eps), _ptr(vis),
    )
    if rc != 0:
        raise ValueError("label map value outside [label_min, label_min + n_classes)")
    return labels, near, int(vis[0])


def _fma(a, b, c) -> float:
    a, b, c = float(a), float(b), float(c)
    if not (np.isfinite(a) and np.isfinite(b) and np.isfinite(c)):
        with np.errstate(all="ignore"):
            return float(np.float64(a) * np.float64(b) + np.float64(c))  # nan/inf: rounding is moot
    return float(Fraction(a) * Fraction(b) + Fraction(c))


def project_py(position, camera):
    """Pure-Python project_gaussian (deep_learning_segmentation.py:43-82), dgemv rounding
    made explicit with an exact-rational fma.  Independent of the C file."""
    fx, fy = float(camera["fx"]), float(camera["fy"])
    width, height = camera["width"], camera["height"]
    R = [[float(v) for v in row] for row in camera["rotation"]]
    p = [float(v) for v in camera["position"]]
    x0, y0, z0 = (float(np.float32(v)) for v in position)

    def dot(row, a, b, c):
        return _fma(row[2], c, _fma(row[0], a, row[1] * b))

    t = [dot([-e for e in row], *p) for row in R]
    cam = [dot(R[r], x0, y0, z0) + t[r] for r in range(3)]
    if cam[2] <= 0:
        return None
    x = (fx * cam[0] / cam[2]) + width / 2
    y = (fy * cam[1] / cam[2]) + height / 2
    if 0 <= x < width and 0 <= y < height:
        return (int(x), int(y))
    return None


def lift_votes_py(pos, cameras, seg_maps, image_sizes=None):
    """Pure-Python assign_labels (deep_learning_segmentation.py:251-306), dict-of-dicts and
    all, for tiny cases.  `seg_maps` is a list of 2-D int arrays, one per camera."""
    votes = {}
    for v, cam in enumerate(cameras):
        seg = seg_maps[v]
        seg_h, seg_w = seg.shape
        ow, oh = (seg_w, seg_h) if image_sizes is None else image_sizes[v]
        hs, ws = seg_h / oh, seg_w / ow
        for i in range(len(pos)):
            pr = project_py(pos[i], cam)
            if pr is None:
                continue
            xs = min(max(0, int(pr[0] * ws)), seg_w - 1)
            ys = min(max(0, int(pr[1] * hs)), seg_h - 1)
            lab = int(seg[ys, xs])
            d = votes.setdefault(i, {})
            d[lab] = d.get(lab, 0) + 1
    out = np.full(len(pos), -1, np.int32)
    for i, d in votes.items():
        out[i] = max(d.items(), key=lambda kv: kv[1])[0]
    return out


# --------------------------------------------------------------------------------------
# K-means
# --------------------------------------------------------------------------------------
def kmeans_assign(data, centroids, want_gap=False):
    """Nearest centroid, scipy-order float64 distance (k_means.py:116-122)."""
    data = np.ascontiguousarray(data, np.float32)
    centroids = np.ascontiguousarray(centroids, np.float32)
    N, D = data.shape
    K = centroids.shape[0]
    labels = np.empty(N, np.int64)
    gap = np.empty(N, np.float64) if want_gap else None
    lib().orc_kmeans_assign(_ptr(data), N, D, _ptr(centroids), K, _ptr(labels),
                            _ptr(gap) if want_gap else None)
    return (labels, gap) if want_gap else labels


def kmeans_update(data, labels, centroids):
    """New centroids, float32 sequential mean (k_means.py:125-128).  Returns (f32[K,D], counts)."""
    data = np.ascontiguousarray(data, np.float32)
    labels = np.ascontiguousarray(labels, np.int64)
    centroids = np.ascontiguousarray(centroids, np.float32)
    N, D = data.shape
    K = centroids.shape[0]
    new = np.empty((K, D), np.float32)
    counts = np.empty(K, np.int64)
    lib().orc_kmeans_update(_ptr(data), _ptr(labels), N, D, K, _ptr(centroids), _ptr(new), _ptr(counts))
    return new, counts


def kmeans_update_f64(data, labels, centroids):
    """Exact-mean yardstick (float64 accumulation); not reference arithmetic."""
    data = np.ascontiguousarray(data, np.float32)
    labels = np.ascontiguousarray(labels, np.int64)
    centroids = np.ascontiguousarray(centroids, np.float32)
    N, D = data.shape
    K = centroids.shape[0]
    new = np.empty((K, D), np.float64)
    lib().orc_kmeans_update_f64(_ptr(data), _ptr(labels), N, D, K, _ptr(centroids), _ptr(new))
    return new


def sqdist(u, v) -> float:
    u = np.ascontiguousarray(u, np.float32)
    v = np.ascontiguousarray(v, np.float32)
    return float(lib().orc_sqdist(_ptr(u), _ptr(v), u.shape[0]))


def kmeans_run(data, k, max_iter=100, tol=1e-4, init_idx=None, trace=None):
    """Lloyd loop of k_means_kd_tree / k_means_with_color (k_means.py:46-96 / 107-144)
    on the concatenated feature block, without the prints and the recolouring.

    init: `np.random.choice(N, k, replace=False)` from the global stream (:63/:111) unless
    `init_idx` is given.  Returns (centroids f32[K,D], labels int64[N], iterations_run).
    `trace`, if a list, receives (centroids_used, labels, new_centroids, shift) per iteration.
    """
    data = np.ascontiguousarray(data, np.float32)
    N = data.shape[0]
    if init_idx is None:
        init_idx = np.random.choice(N, k, replace=False)
    centroids = data[init_idx]
    it = 0
    for it in range(max_iter):
        labels = kmeans_assign(data, centroids)
        new, _ = kmeans_update(data, labels, centroids)
        shift = np.linalg.norm(new - centroids)            # :83/:131
        if trace is not None:
            trace.append((centroids.copy(), labels.copy(), new.copy(), float(shift)))
        if shift < tol:                                    # :84-86 break BEFORE the swap
            break
        centroids = new                                    # :88/:136
    labels = kmeans_assign(data, centroids)                # :92-96/:140-144
    return centroids, labels, it + 1


# --------------------------------------------------------------------------------------
# viewer-side label consumers (gaussians_selection.js, `gs`); see viewer_oracle.c
# --------------------------------------------------------------------------------------
NO_SELECTION = -999999          # gs:6


def _rows(pos):
    pos = np.ascontiguousarray(pos, np.float32)
    assert pos.ndim == 2 and pos.shape[1] >= 3
    return pos, pos.shape[0], pos.shape[1]


def viewer_depth_sort(pos, view_proj) -> np.ndarray:
    """runSort (gs:427-457).  pos f32[N][stride >= 3]; view_proj 16 doubles.  Returns depthIndex uint32[N]."""
    pos, n, stride = _rows(pos)
    vp_ = np.ascontiguousarray(view_proj, np.float64).reshape(16)
    out = np.empty(n, np.uint32)
    lib().orc_viewer_depth_sort(_ptr(pos), n, stride, _ptr(vp_), _ptr(out))
    return out


def viewer_buckets(pos, view_proj) -> np.ndarray:
    pos, n, stride = _rows(pos)
    vp_ = np.ascontiguousarray(view_proj, np.float64).reshape(16)
    out = np.empty(n, np.int32)
    lib().orc_viewer_buckets(_ptr(pos), n, stride, _ptr(vp_), _ptr(out))
    return out


def multiply4(a, b) -> np.ndarray:
    """multiply4 (gs:110-123)."""
    a = np.ascontiguousarray(a, np.float64).reshape(16)
    b = np.ascontiguousarray(b, np.float64).reshape(16)
    out = np.empty(16, np.float64)
    lib().orc_multiply4(_ptr(a), _ptr(b), _ptr(out))
    return out


def viewer_hit_test(pos, labels, matrix, x, y, viewport, no_selection=NO_SELECTION):
    """performHitTesting (gs:361-395) with the combined matrix given.  Returns (label, index)."""
    pos, n, stride = _rows(pos)
    labels = np.ascontiguousarray(labels, np.int32)
    m = np.ascontiguousarray(matrix, np.float64).reshape(16)
    idx = np.zeros(1, np.int64)
    lab = lib().orc_viewer_hit_test(_ptr(pos), _ptr(labels), n, stride, _ptr(m), float(x), float(y),
                                    float(viewport[0]), float(viewport[1]), int(no_selection), _ptr(idx))
    return int(lab), int(idx[0])


def js_to_int32_py(d: float) -> int:
    """ECMA-262 ToInt32 on a Python float, written from the specification text."""
    import math
    if math.isnan(d) or math.isinf(d):
        return 0
    t = int(d)                      # truncation towards zero, exact
    t %= 1 << 32
    return t - (1 << 32) if t >= 1 << 31 else t


def viewer_depth_sort_py(pos, view_proj):
    """Independent pure-Python runSort (gs:427-457): Python floats are binary64 like JS numbers;
    the typed arrays' out-of-range behaviour is written out with dictionaries."""
    pos = np.asarray(pos, np.float32)
    n = pos.shape[0]
    vp_ = [float(v) for v in np.asarray(view_proj, np.float64).reshape(16)]
    size_list = [0] * n
    max_depth, min_depth = float("-inf"), float("inf")
    for i in range(n):
        x, y, z = (float(v) for v in pos[i, :3])
        depth = js_to_int32_py((vp_[2] * x + vp_[6] * y + vp_[10] * z) * 4096)
        size_list[i] = depth
        max_depth = max(max_depth, depth)
        min_depth = min(min_depth, depth)
    with np.errstate(all="ignore"):
        depth_inv = float(np.float64(65536.0) / np.float64(max_depth - min_depth)) if n else 0.0
    counts0 = [0] * 65536
    for i in range(n):
        with np.errstate(all="ignore"):
            size_list[i] = js_to_int32_py(float(np.float64(size_list[i] - min_depth) * np.float64(depth_inv)))
        if 0 <= size_list[i] < 65536:
            counts0[size_list[i]] += 1
    starts0 = [0] * 65536
    for i in range(1, 65536):
        starts0[i] = starts0[i - 1] + counts0[i - 1]
    depth_index = [0] * n
    for i in range(n):
        k = size_list[i]
        if 0 <= k < 65536:          # otherwise: NaN index, nothing stored
            depth_index[starts0[k]] = i
            starts0[k] += 1
    return np.array(depth_index, np.uint32)


def viewer_hit_test_py(pos, labels, matrix, x, y, viewport, no_selection=NO_SELECTION):
    """Independent pure-Python performHitTesting (gs:361-395)."""
    import math
    pos = np.asarray(pos, np.float32)
    m = [float(v) for v in np.asarray(matrix, np.float64).reshape(16)]
    closest_dist = closest_depth = float("inf")
    selected, sel = no_selection, -1
    for i in range(pos.shape[0]):
        p = [float(pos[i, 0]), float(pos[i, 1]), float(pos[i, 2]), 1.0]
        r = [p[0] * m[k] + p[1] * m[k + 4] + p[2] * m[k + 8] + p[3] * m[k + 12] for k in range(4)]
        if r[3] <= 0:
            continue
        with np.errstate(all="ignore"):
            w = np.float64(r[3])
            sx = float((np.float64(r[0]) / w + 1) * 0.5 * viewport[0])
            sy = float((np.float64(r[1]) / w + 1) * 0.5 * viewport[1])
            depth = float(np.float64(r[2]) / w)
        dx, dy = abs(sx - x), abs(sy - y)
        if math.isinf(dx) or math.isinf(dy):
            dist = float("inf")
        elif math.isnan(dx) or math.isnan(dy):
            dist = float("nan")
        else:
            mx = max(dx, dy)
            if mx == 0:
                dist = 0.0
            else:                   # V8 MathHypot, Kahan loop written out
                s = c = 0.0
                for v in (dx, dy):
                    q = v / mx
                    summand = q * q - c
                    prelim = s + summand
                    c = (prelim - s) - summand
                    s = prelim
                dist = math.sqrt(s) * mx
        if dist < 10 and (dist < closest_dist or (dist == closest_dist and depth < closest_depth)):
            closest_dist, closest_depth, selected, sel = dist, depth, int(labels[i]), i
    return selected, sel


# --------------------------------------------------------------------------------------
# region growing (3D_clustering/region_growing.py, `rg`); see region_oracle.c
# --------------------------------------------------------------------------------------
def region_knn_pca(pos, k, normals_in=None, want_knn=False, queries=None):
    """compute_normals (rg:78-127) + compute_residuals (rg:130-163) for one k; `queries` = (q0, q1)
    restricts the work to those points (the other rows of the outputs are left uninitialised).
    Returns dict(normals f64[N,3], residuals f64[N], centroids f32[N,3], gap f64[N], knn int32[N,k] or None)."""
    pos = np.ascontiguousarray(pos, np.float32).reshape(-1, 3)
    n = pos.shape[0]
    normals = np.empty((n, 3), np.float64)
    residuals = np.empty(n, np.float64)
    centroids = np.empty((n, 3), np.float32)
    gap = np.empty(n, np.float64)
    knn = np.empty((n, k), np.int32) if want_knn else None
    nin = None if normals_in is None else np.ascontiguousarray(normals_in, np.float64)
    q0, q1 = (0, n) if queries is None else (int(queries[0]), int(queries[1]))
    lib().orc_region_knn_pca_range(_ptr(pos), n, int(k), _ptr(nin) if nin is not None else None, _ptr(normals), _ptr(residuals),
                                   _ptr(centroids), _ptr(knn) if want_knn else None, _ptr(gap), q0, q1)
    return dict(normals=normals, residuals=residuals, centroids=centroids, gap=gap, knn=knn)


def region_grow(knn, normals, residuals, residual_threshold, angle_threshold):
    """segmentation_3D (rg:166-221) given neighbour lists.  Returns region_of int32[N] (creation order)."""
    knn = np.ascontiguousarray(knn, np.int32)
    normals = np.ascontiguousarray(normals, np.float64)
    residuals = np.ascontiguousarray(residuals, np.float64)
    n, k = knn.shape
    region_of = np.empty(n, np.int32)
    r = lib().orc_region_grow(_ptr(knn), k, _ptr(normals), _ptr(residuals), n, float(residual_threshold),
                              float(angle_threshold), _ptr(region_of))
    return region_of, int(r)
