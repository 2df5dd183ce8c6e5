"""The two reference entry points, run as scripts with the reference's flags on a stand-in scene
(config C1 / C2 shape, small): outputs are labelled PLY files the reference's viewer can read."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from util import load_lift_case, pkg

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _standin_ply(path, n, seed):
    scene, plyio = pkg("scene"), pkg("plyio")
    v = scene.standin_3dgs_vertices(n, seed=seed)
    plyio.write_ply(path, [("vertex", v)], text=False)
    return v


def test_k_means_script_c1(oracle, tmp_path):
    """python 3D_clustering/k_means.py --file_path .. --save_path .. [--k 10]  (k_means.py:198-215)"""
    plyio = pkg("plyio")
    src, dst = tmp_path / "point_cloud.ply", tmp_path / "clustered.ply"
    v = _standin_ply(src, 20000, seed=1)
    code = ("import numpy as np, runpy, sys; np.random.seed(0); sys.argv = sys.argv[1:]; "
            "runpy.run_path(sys.argv[0], run_name='__main__')")
    out = subprocess.run([sys.executable, "-c", code, os.path.join(ROOT, "3D_clustering", "k_means.py"),
                          "--file_path", str(src), "--save_path", str(dst)], capture_output=True, text=True, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = out.stdout.splitlines()
    assert f"New PLY file with label added saved to {dst}" in lines[-1]
    assert sum(1 for ln in lines if ln.strip().isdigit()) == 10          # max_iter=10, one index per iteration
    back = plyio.read_ply(dst)
    assert back.text and back["vertex"].data.dtype.names[-1] == "label"
    got = back["vertex"]["label"]
    # the oracle's run of the same algorithm from the same seeded start
    data = np.column_stack((v["x"], v["y"], v["z"], v["f_dc_0"], v["f_dc_1"], v["f_dc_2"])).astype(np.float32)
    np.random.seed(0)
    _, want, _ = oracle.kmeans_run(data, 10, max_iter=10)
    assert np.array_equal(got.astype(np.int64), want)
    for name in ("x", "f_dc_2", "rot_3"):
        assert np.array_equal(back["vertex"][name], v[name])              # %.18g round-trips float32


def test_lifting_script_c2(oracle, tmp_path):
    """python deep_learning_segmentation.py --ply_file .. --camera_file .. --input_dir .. --output_dir ..
    --output_file ..  with the bundled cameras and precomputed <img>_segmap.npy files (:165, :335-375)."""
    from PIL import Image
    plyio = pkg("plyio")
    c = load_lift_case("lift_bundled_halfres")
    src, dst = tmp_path / "scene.ply", tmp_path / "labelled.ply"
    v = _standin_ply(src, 30000, seed=2)
    v["x"] *= 2; v["y"] *= 2; v["z"] *= 2
    plyio.write_ply(src, [("vertex", v)], text=False)
    img_dir, seg_dir = tmp_path / "images", tmp_path / "seg"
    os.makedirs(img_dir); os.makedirs(seg_dir)
    json.dump(c["cameras"], open(tmp_path / "cameras.json", "w"))
    for cam, m, (w, h) in zip(c["cameras"], c["maps"], c["sizes"]):
        Image.new("L", (w, h)).save(img_dir / (cam["img_name"] + ".png"))
        np.save(seg_dir / f"{cam['img_name']}_segmap.npy", m)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "deep_learning_segmentation.py"), "--ply_file", str(src),
                          "--camera_file", str(tmp_path / "cameras.json"), "--input_dir", str(img_dir),
                          "--output_dir", str(seg_dir), "--output_file", str(dst)], capture_output=True, text=True, cwd=ROOT,
                         env=dict(os.environ, GSLIFT_REUSE_SEGMAPS="1"))        # reuse of segment_image's files is opt-in
    assert out.returncode == 0, out.stderr[-2000:]
    assert "Loading cameras..." in out.stdout and "Label statistics:" in out.stdout
    assert out.stdout.count("Using precomputed segmentation map") == len(c["cameras"])
    assert out.stdout.count("Processing image") == len(c["cameras"])
    back = plyio.read_ply(dst)
    assert not back.text and back["vertex"].data.dtype.names[-1] == "label"
    pos = np.column_stack((v["x"], v["y"], v["z"])).astype(np.float32)
    want, _, vis = oracle.lift_votes(pos, oracle.make_views(c["cameras"], c["shapes"], c["sizes"]), c["flat"])
    assert np.array_equal(back["vertex"]["label"], want)
    assert f"Total gaussians: {len(pos)}" in out.stdout
    n_unseen = int((want == -1).sum())
    assert f"Label -1: {n_unseen} gaussians" in out.stdout


def test_region_growing_script(oracle, tmp_path):
    """python 3D_clustering/region_growing.py in.ply out.ply: the script's __main__ flow (rg:263-285:
    normals and residuals with k = 2000, growth with k = 10) on a 6000-vertex stand-in; the regions
    written as colours equal the ones the oracle grows from the oracle's own normals / residuals."""
    plyio = pkg("plyio")
    src, dst = tmp_path / "point_cloud.ply", tmp_path / "clustering.ply"
    v = _standin_ply(src, 6000, seed=3)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "3D_clustering", "region_growing.py"), str(src), str(dst)],
                         capture_output=True, text=True, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = out.stdout.splitlines()
    assert lines[0] == "Calculating normals..." and "Calculating residuals..." in lines and lines[-1] == "writing new data"
    n_segments = int([ln for ln in lines if ln.startswith("number of segments: ")][0].split(": ")[1])
    back = plyio.read_ply(dst)
    assert not back.text and back["vertex"].data.dtype == v.dtype
    for name in ("x", "y", "z", "opacity", "rot_3"):
        assert np.array_equal(back["vertex"][name], v[name])
    # every region got one random colour triple: vertices with equal triples form the regions
    colour = np.stack([back["vertex"][f"f_dc_{i}"] for i in range(3)], 1)
    _, region_written = np.unique(colour, axis=0, return_inverse=True)
    pos = np.column_stack((v["x"], v["y"], v["z"])).astype(np.float32)
    ref = oracle.region_knn_pca(pos, 2000)
    knn = oracle.region_knn_pca(pos, 10, want_knn=True)["knn"]
    region_ref, n_ref = oracle.region_grow(knn, ref["normals"], ref["residuals"], 0.1, 0.05)
    # the GPU's normals differ from the oracle's in the last bits, which can move a borderline smoothness
    # test (|cos| vs cos 0.05): allow a handful of points to change region
    assert abs(n_segments - n_ref) <= max(3, n_ref // 100)
    pairs = {}
    for a, b in zip(region_written.reshape(-1).tolist(), region_ref.tolist()):
        pairs.setdefault(a, {}).setdefault(b, 0)
        pairs[a][b] += 1
    agree = sum(max(d.values()) for d in pairs.values())
    assert agree >= 0.98 * len(pos)
