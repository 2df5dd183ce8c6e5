"""Generate tests/golden/*.npz by running the REFERENCE's own functions verbatim.

Runs only in the build container (needs /root/reference, which does not exist on the GPU
box).  The reference modules import `plyfile`, `matplotlib` and `ultralytics`, none of which
is installed and none of which the hot path touches, so empty stub modules are injected.
`assign_labels` is then run unmodified with its segmenter monkeypatched to hand back
precomputed maps (the injection point SURVEY.md 8b names).

    python oracle/make_golden.py            # regenerates every fixture

The fixtures hold inputs AND the reference's outputs, so the tests need neither the
reference nor this script.  Also re-runs the arithmetic probes (dgemv rounding, scipy
distance order, NumPy float32 mean) and stores known-answer vectors for them.
"""
from __future__ import annotations

import contextlib
import importlib
import io
import json
import os
import sys
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)


def _stub_modules():
    ply = types.ModuleType("plyfile")
    ply.PlyData = type("PlyData", (), {})
    ply.PlyElement = type("PlyElement", (), {})
    sys.modules.setdefault("plyfile", ply)
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules.setdefault("matplotlib", mpl)
    sys.modules.setdefault("matplotlib.pyplot", plt)
    ul = types.ModuleType("ultralytics")
    ul.YOLO = type("YOLO", (), {})
    sys.modules.setdefault("ultralytics", ul)


def load_reference():
    _stub_modules()
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(REF, "3D_clustering"))
    dls = importlib.import_module("deep_learning_segmentation")
    km = importlib.import_module("k_means")
    return dls, km


def run_assign_labels(dls, pos, cameras, seg_maps, image_sizes):
    """Call the reference's assign_labels unmodified; only its segmenter / file probes are
    replaced.  Returns (labels, captured stdout)."""
    gaussians = np.zeros(len(pos), dtype=[("position", np.float32, 3), ("scale", np.float32, 3),
                                          ("rotation", np.float32, 4)])
    gaussians["position"] = pos
    by_name = {c["img_name"]: i for i, c in enumerate(cameras)}

    class _Model:
        def to(self, _):
            return self

    class _Img:
        def __init__(self, size):
            self.size = size

    def fake_segment(image_path, output_dir, processor, model, device, model_type):
        name = os.path.splitext(os.path.basename(image_path))[0]
        return seg_maps[by_name[name]]

    def fake_open(path):
        name = os.path.splitext(os.path.basename(path))[0]
        return _Img(tuple(image_sizes[by_name[name]]))

    saved = (dls.initialize_model, dls.segment_image, dls.Image.open, os.path.exists)
    dls.initialize_model = lambda model_type, device: (None, _Model())
    dls.segment_image = fake_segment
    dls.Image.open = fake_open
    os.path.exists = lambda p: True
    buf = io.StringIO()
    try:
        with contextlib.redirect_stdout(buf):
            labels = dls.assign_labels(gaussians, cameras, "in", "out", model_type="mask2former")
    finally:
        dls.initialize_model, dls.segment_image, dls.Image.open, os.path.exists = saved
    return labels, buf.getvalue()


def region_label_map(height, width, seed, n_regions=9, lo=-1, hi=12):
    """Smooth-ish regions (nearest of a few random sites), so views agree more often than
    independent random blocks and the majority is not just 'first seen'."""
    rng = np.random.default_rng(seed)
    sites = rng.uniform(0, 1, size=(n_regions, 2))
    labs = rng.integers(lo, hi + 1, size=n_regions)
    yy, xx = np.mgrid[0:height, 0:width]
    d = (yy[..., None] / height - sites[:, 0]) ** 2 + (xx[..., None] / width - sites[:, 1]) ** 2
    return labs[np.argmin(d, axis=-1)].astype(np.int32)


def save_lift_case(name, pos, cameras, seg_maps, image_sizes, labels, note):
    maps16 = [m.astype(np.int16) for m in seg_maps]
    shapes = np.array([m.shape for m in seg_maps], np.int32)
    flat = np.concatenate([m.reshape(-1) for m in maps16])
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        pos=pos.astype(np.float32),
        cameras=np.array(json.dumps(cameras)),
        map_shapes=shapes,
        maps_flat=flat,
        image_sizes=np.array(image_sizes, np.int32),
        labels=labels.astype(np.int32),
        note=np.array(note),
    )
    vis = int((labels != -1).sum())
    print(f"[golden] {name}: N={len(pos)} V={len(cameras)} labelled={vis} "
          f"distinct={len(np.unique(labels))}")


def make_lifting(dls):
    from importlib import import_module
    scene = import_module("3d_gaussian_splatting_project_b200.scene")
    bundled = json.load(open(os.path.join(REF, "Web_Viewer_Gaussians_Selection", "cameras.json")))
    rng = np.random.default_rng(11)

    # L1: bundled cameras (3114x2075), half-resolution image AND seg map -> scale 1.0, the
    # clamp at :285-286 fires for x >= 1557.
    cams = [bundled[i] for i in range(0, 311, 40)]
    pos = (rng.standard_normal((2500, 3)) * 2.0).astype(np.float32)
    maps = [scene.block_label_map(1038, 1557, 64, -1, 149, 2 + v) for v in range(len(cams))]
    sizes = [(1557, 1038)] * len(cams)
    lab, out = run_assign_labels(dls, pos, cams, maps, sizes)
    assert out.count("Processing image") == len(cams)
    save_lift_case("lift_bundled_halfres", pos, cams, maps, sizes, lab,
                   "bundled cameras.json[::40], image+seg 1557x1038, camera 3114x2075")

    # L2: synthetic look-at cameras, map == image == camera size (scale 1.0 exactly).
    cams = scene.lookat_cameras(16, radius=6.0, width=640, height=360, seed=7)
    pos = scene.gaussian_cloud(4000, 1.5, seed=8)
    maps = [scene.block_label_map(360, 640, 16, -1, 149, 1000 + v) for v in range(len(cams))]
    sizes = [(640, 360)] * len(cams)
    lab, _ = run_assign_labels(dls, pos, cams, maps, sizes)
    save_lift_case("lift_lookat_fullres", pos, cams, maps, sizes, lab,
                   "16 look-at cameras 640x360, 16px random blocks, 151 labels")

    # L3: seg map at a non-integer ratio of the image (scale_x = 300/640, scale_y = 200/360).
    maps = [scene.block_label_map(200, 300, 8, -1, 149, 2000 + v) for v in range(len(cams))]
    lab, _ = run_assign_labels(dls, pos, cams, maps, sizes)
    save_lift_case("lift_lookat_rescaled", pos, cams, maps, sizes, lab,
                   "same cameras, seg 300x200 vs image 640x360 (non-integer scale)")

    # L4: coherent regions with -1 background and few classes: real majorities and ties.
    maps = [region_label_map(360, 640, 3000 + (v // 2)) for v in range(len(cams))]
    lab, _ = run_assign_labels(dls, pos, cams, maps, sizes)
    save_lift_case("lift_lookat_regions", pos, cams, maps, sizes, lab,
                   "same cameras, region maps shared by view pairs (majorities + ties)")

    # L5: a camera list with one view 'missing' is the caller's business (:257-259); here we
    # pin the degenerate cases instead: points behind every camera / exactly at a camera.
    cams2 = cams[:3]
    pos2 = np.array([[100, 100, 100], cams2[0]["position"], [0, 0, 0], [np.nan, 0, 0],
                     [np.inf, 0, 0], [1e-30, -1e-30, 1e-30]], np.float32)
    maps2 = maps[:3]
    with np.errstate(all="ignore"):
        lab, _ = run_assign_labels(dls, pos2, cams2, maps2, sizes[:3])
    save_lift_case("lift_degenerate", pos2, cams2, maps2, sizes[:3], lab,
                   "far / at-camera / origin / nan / inf / denormal-ish points")


def make_probes():
    """Known-answer vectors for the third-party arithmetic the oracle restates."""
    from scipy.spatial import KDTree
    bundled = json.load(open(os.path.join(REF, "Web_Viewer_Gaussians_Selection", "cameras.json")))
    rng = np.random.default_rng(5)
    # (1) dgemv: t = -R @ p for all 311 cameras, and R @ pos(float32) + t samples.
    R = np.array([c["rotation"] for c in bundled])
    p = np.array([c["position"] for c in bundled])
    t = np.stack([-R[i] @ p[i] for i in range(len(bundled))])
    pos = (rng.standard_normal((len(bundled) * 4, 3)) * 2).astype(np.float32)
    cam_idx = np.arange(len(pos)) % len(bundled)
    pos_cam = np.stack([R[cam_idx[i]] @ pos[i] + t[cam_idx[i]] for i in range(len(pos))])
    # (2) scipy distance order at several D (tail lengths 0..3).
    dist = {}
    for D in (3, 6, 7, 8, 59, 61):
        c = rng.standard_normal((40, D)).astype(np.float32)
        x = rng.standard_normal((200, D)).astype(np.float32)
        tree = KDTree(c)
        dd = np.empty(len(x)); ii = np.empty(len(x), np.int64)
        for j, pt in enumerate(x):               # per-point query, as k_means.py:120-122
            dd[j], ii[j] = tree.query(pt)
        dist[f"c{D}"] = c; dist[f"x{D}"] = x; dist[f"d{D}"] = dd; dist[f"i{D}"] = ii
    np.savez_compressed(os.path.join(OUT, "probes.npz"), R=R, p=p, t=t, pos=pos, cam_idx=cam_idx,
                        pos_cam=pos_cam, **dist)
    print("[golden] probes: dgemv", t.shape, pos_cam.shape, "scipy D", [3, 6, 7, 8, 59, 61])


def make_kmeans(km):
    from importlib import import_module
    scene = import_module("3d_gaussian_splatting_project_b200.scene")
    rng = np.random.default_rng(21)

    def run(fn, seed, *args, **kw):
        buf = io.StringIO()
        np.random.seed(seed)
        with contextlib.redirect_stdout(buf):
            res = fn(*args, **kw)
        return res, buf.getvalue()

    # K1: k_means_with_color, the __main__ configuration (K=10, max_iter=10), D=3+3.
    v = scene.standin_3dgs_vertices(4000, seed=1)
    points = np.column_stack((v["x"], v["y"], v["z"]))
    colors = np.column_stack((v["f_dc_0"], v["f_dc_1"], v["f_dc_2"]))
    (cen, lab, col), out = run(km.k_means_with_color, 0, points, 10, colors.copy(), max_iter=10)
    np.savez_compressed(os.path.join(OUT, "kmeans_color_k10.npz"), points=points, colors=colors,
                        seed=0, k=10, max_iter=10, centroids=cen, labels=lab, colors_out=col,
                        stdout=np.array(out))
    print("[golden] kmeans_color_k10:", cen.shape, cen.dtype, lab.dtype, np.bincount(lab).tolist())

    # K2: generic k_means_kd_tree, K=64, D=59 (the C5 shape, small N), 6 iterations.
    data = scene.blob_features(6000, 59, n_blobs=64, seed=5)
    colors = rng.standard_normal((6000, 3)).astype(np.float32)
    (cen, lab, col), out = run(km.k_means_kd_tree, 0, data, 64, colors.copy(), max_iter=6)
    np.savez_compressed(os.path.join(OUT, "kmeans_kdtree_k64_d59.npz"), data=data, colors=colors,
                        seed=0, k=64, max_iter=6, centroids=cen, labels=lab, colors_out=col,
                        stdout=np.array(out))
    print("[golden] kmeans_kdtree_k64_d59:", cen.shape, lab.dtype, "clusters used", len(np.unique(lab)))

    # K3: convergence path (:132-134): well separated blobs converge before max_iter, and the
    # break leaves `centroids` at the PRE-update value.
    centres = np.array([[0, 0, 0], [10, 0, 0], [0, 10, 0], [0, 0, 10]], np.float32)
    which = rng.integers(0, 4, 800)
    pts = (centres[which] + rng.normal(0, 0.05, (800, 3))).astype(np.float32)
    cols = rng.standard_normal((800, 3)).astype(np.float32) * np.float32(0.01)
    (cen, lab, col), out = run(km.k_means_with_color, 3, pts, 4, cols.copy(), max_iter=50)
    assert "Converged" in out
    np.savez_compressed(os.path.join(OUT, "kmeans_converged_k4.npz"), points=pts, colors=cols,
                        seed=3, k=4, max_iter=50, centroids=cen, labels=lab, colors_out=col,
                        stdout=np.array(out))
    print("[golden] kmeans_converged_k4:", out.strip().splitlines()[-1])

    # K4: an empty cluster (:126 keeps the old centroid).  Two of the initial centroids are
    # copies of the same point, so every member ties exactly; with K <= leafsize (10) the
    # cKDTree is one leaf scanned in index order and the lowest index wins, which leaves the
    # other copy empty in the first iteration (it re-captures its own point afterwards).  (For K > 10 the winner of an exact tie depends on
    # the tree layout -- the documented exemption -- so this case stays at K = 8.)
    pts = np.repeat(rng.standard_normal((30, 3)).astype(np.float32), 20, axis=0)
    cols = np.zeros((600, 3), np.float32)
    for seed in range(100):
        np.random.seed(seed)
        idx = np.random.choice(600, 8, replace=False)
        if len(np.unique(idx // 20)) < 8:
            break
    (cen, lab, col), out = run(km.k_means_with_color, seed, pts, 8, cols.copy(), max_iter=4)
    np.savez_compressed(os.path.join(OUT, "kmeans_empty_k8.npz"), points=pts, colors=cols,
                        seed=seed, k=8, max_iter=4, centroids=cen, labels=lab, colors_out=col,
                        stdout=np.array(out))
    print("[golden] kmeans_empty_k8: seed", seed, "clusters used", len(np.unique(lab)), "of 8")


def make_bundled_cameras():
    """The reference's bundled camera file (311 views of 3114 x 2075, BASELINE config C2) as a test
    input: the GPU box has no /root/reference.  Parsed and re-serialised (repr round-trips floats)."""
    import gzip
    bundled = json.load(open(os.path.join(REF, "Web_Viewer_Gaussians_Selection", "cameras.json")))
    with gzip.GzipFile(os.path.join(OUT, "bundled_cameras.json.gz"), "wb", mtime=0) as fh:
        fh.write(json.dumps(bundled).encode())
    print(f"[golden] bundled_cameras.json.gz: {len(bundled)} cameras")


def region_cloud(seed, n):
    """Three noisy planar patches and a blob: surfaces for the normals, a volume where they are ill defined."""
    rng = np.random.default_rng(seed)
    parts = []
    m = n // 4
    for axis, off in ((2, 0.0), (0, 1.5), (1, -1.0)):
        p = rng.uniform(-1, 1, size=(m, 3))
        p[:, axis] = off + rng.normal(0, 0.01, m)
        parts.append(p)
    parts.append(rng.normal(0, 0.3, size=(n - 3 * m, 3)) + [3.0, 0, 0])
    pts = np.concatenate(parts).astype(np.float32)
    return pts[rng.permutation(n)]


def make_region():
    """compute_normals / compute_residuals / segmentation_3D of region_growing.py run verbatim
    (rg:78-221); only their progress prints are swallowed."""
    sys.path.insert(0, os.path.join(REF, "3D_clustering"))
    rgm = importlib.import_module("region_growing")
    from scipy.spatial import KDTree
    for name, seed, n, k, kseg in (("region_planes_k40", 21, 1500, 40, 10), ("region_planes_k400", 22, 1200, 400, 10),
                                ("region_planes_k2000", 23, 2600, 2000, 10)):
        pts = region_cloud(seed, n)
        with contextlib.redirect_stdout(io.StringIO()):
            normals = rgm.compute_normals(pts, k)
            residuals = rgm.compute_residuals(pts, normals, k)
            regions = rgm.segmentation_3D(pts, normals, residuals, residual_threshold=0.1, angle_threshold=0.05, k=kseg)
        region_of = np.full(n, -1, np.int32)
        for r, members in enumerate(regions):          # regions come back sorted by size, largest first (rg:219)
            region_of[np.array(members, np.int64)] = r
        knn_seg = KDTree(pts).query(pts, kseg)[1].astype(np.int32)      # the lists rg:203 sees
        np.savez_compressed(os.path.join(OUT, name + ".npz"), pos=pts, k=k, kseg=kseg, normals=normals, residuals=residuals,
                            region_of=region_of, n_regions=len(regions), knn_seg=knn_seg,
                            note="region_growing.py compute_normals/compute_residuals/segmentation_3D verbatim")
        print(f"[golden] {name}: {n} points, k={k}, {len(regions)} regions")


def main():
    os.makedirs(OUT, exist_ok=True)
    make_bundled_cameras()
    dls, km = load_reference()
    make_probes()
    make_lifting(dls)
    make_kmeans(km)
    make_region()


if __name__ == "__main__":
    main()
