"""Deterministic synthetic scenes for the configurations of BASELINE.json.

Nothing here is reference behaviour: the bundled ``data/point_cloud.ply`` is missing from the
mount and there is no network, so benchmarks and tests use these stand-ins.  Camera dicts
follow the ``cameras.json`` schema read at ``deep_learning_segmentation.py:54-63``.
"""
from __future__ import annotations

import math

import numpy as np

# Property order of a standard 3DGS point_cloud.ply (62 float32 columns).
PLY_3DGS_PROPS = (
    ["x", "y", "z", "nx", "ny", "nz"]
    + [f"f_dc_{i}" for i in range(3)]
    + [f"f_rest_{i}" for i in range(45)]
    + ["opacity"]
    + [f"scale_{i}" for i in range(3)]
    + [f"rot_{i}" for i in range(4)]
)


def lookat_cameras(n_views, radius=6.0, width=1920, height=1080, hfov_deg=60.0, seed=3):
    """Cameras on a Fibonacci sphere looking at the origin.

    ``rotation`` rows are the camera x, y, z axes in world coordinates, so that the
    reference's un-transposed ``R @ (X - p)`` (deep_learning_segmentation.py:60-69) puts the
    cloud in front of the camera (z > 0)."""
    rng = np.random.default_rng(seed)
    fx = (width / 2) / math.tan(math.radians(hfov_deg) / 2)
    cams = []
    golden = math.pi * (3.0 - math.sqrt(5.0))
    for i in range(n_views):
        zc = 1.0 - 2.0 * (i + 0.5) / n_views
        r = math.sqrt(max(0.0, 1.0 - zc * zc))
        th = golden * i
        p = radius * np.array([r * math.cos(th), r * math.sin(th), zc])
        p = p + rng.normal(0.0, 0.05, 3)
        fwd = -p / np.linalg.norm(p)
        up = np.array([0.0, 0.0, 1.0]) if abs(fwd[2]) < 0.95 else np.array([0.0, 1.0, 0.0])
        right = np.cross(fwd, up)
        right /= np.linalg.norm(right)
        down = np.cross(fwd, right)
        R = np.stack([right, down, fwd])
        cams.append(
            {
                "id": i,
                "img_name": f"view_{i:04d}",
                "width": int(width),
                "height": int(height),
                "position": [float(v) for v in p],
                "rotation": [[float(v) for v in row] for row in R],
                "fy": float(fx),
                "fx": float(fx),
            }
        )
    return cams


def gaussian_cloud(n, sigma=1.5, seed=3):
    """positions ~ N(0, sigma^2)^3 as float32 [n, 3]."""
    rng = np.random.default_rng(seed)
    return (rng.standard_normal((n, 3), dtype=np.float32) * np.float32(sigma)).astype(np.float32)


def block_label_map(height, width, block=32, lo=-1, hi=149, seed=0, dtype=np.int32):
    """Piecewise-constant label map: `block` x `block` px cells, labels U{lo..hi}."""
    rng = np.random.default_rng(seed)
    gh, gw = -(-height // block), -(-width // block)
    cells = rng.integers(lo, hi + 1, size=(gh, gw)).astype(dtype)
    return np.repeat(np.repeat(cells, block, axis=0), block, axis=1)[:height, :width].copy()


def block_label_maps(n_views, height, width, block=32, lo=-1, hi=149, seed=1000, dtype=np.int32, out=None):
    """Stack of per-view block maps [n_views, height, width] (seed + view)."""
    if out is None:
        out = np.empty((n_views, height, width), dtype)
    for v in range(n_views):
        out[v] = block_label_map(height, width, block, lo, hi, seed + v, dtype)
    return out


def blob_features(n, d, n_blobs=64, seed=5):
    """C5 features: mixture of `n_blobs` blobs in the 3 position dims (sigma 0.3, centres
    U(-2,2)^3), remaining dims N(0,1).  float32 [n, d]."""
    rng = np.random.default_rng(seed)
    centres = rng.uniform(-2.0, 2.0, size=(n_blobs, 3)).astype(np.float32)
    which = rng.integers(0, n_blobs, size=n)
    x = rng.standard_normal((n, d), dtype=np.float32)
    x[:, :3] = x[:, :3] * np.float32(0.3) + centres[which]
    return x


def standin_3dgs_vertices(n, seed=1, n_blobs=10):
    """Stand-in for the missing data/point_cloud.ply: structured array with the 62 float32
    3DGS properties; positions = `n_blobs` blobs (sigma 0.3), f_dc ~ N(0,1)."""
    rng = np.random.default_rng(seed)
    dt = np.dtype([(name, "<f4") for name in PLY_3DGS_PROPS])
    v = np.zeros(n, dt)
    centres = rng.uniform(-2.0, 2.0, size=(n_blobs, 3))
    which = rng.integers(0, n_blobs, size=n)
    pos = centres[which] + rng.normal(0.0, 0.3, size=(n, 3))
    for i, name in enumerate(("x", "y", "z")):
        v[name] = pos[:, i].astype(np.float32)
    for name in PLY_3DGS_PROPS[6:]:
        v[name] = rng.standard_normal(n).astype(np.float32)
    return v
