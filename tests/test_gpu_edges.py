"""Error behaviour of the C ABI on a device, and randomised small cases against the independent
pure-Python restatement of the reference (oracle.lift_votes_py) and the C oracle."""
import numpy as np
import pytest
import torch

from util import pkg

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_abi_error_codes_on_device():
    native, ops, scene = pkg("_native"), pkg("ops"), pkg("scene")
    L = native.lib()
    cams = scene.lookat_cameras(3, width=64, height=48, seed=1)
    views = ops.make_views(cams, [(48, 64)] * 3)
    pos = torch.zeros(10, 3, device=DEV)
    packed = torch.zeros(3 * ops.packed_map_bytes(48, 64), dtype=torch.uint8, device=DEV)
    labels = torch.empty(10, dtype=torch.int32, device=DEV)
    ws = torch.empty(64, dtype=torch.uint8, device=DEV)
    rc = L.gsl_lift_votes(pos.data_ptr(), 10, views.ctypes.data, 3, packed.data_ptr(), -1, 254, labels.data_ptr(),
                          None, 0.0, ws.data_ptr(), ws.numel(), None)
    assert rc == -2 and b"workspace" in L.gsl_last_error()                      # GSL_EWORKSPACE
    big = torch.empty(L.gsl_lift_workspace_bytes(10, 3), dtype=torch.uint8, device=DEV)
    rc = L.gsl_lift_votes(pos.data_ptr(), 10, views.ctypes.data, 3, packed.data_ptr(), -1, 255, labels.data_ptr(),
                          None, 0.0, big.data_ptr(), big.numel(), None)
    assert rc == -1 and b"n_classes" in L.gsl_last_error()                      # GSL_EINVAL
    bad = views.copy()
    bad["seg_w"][1] = 0
    rc = L.gsl_lift_votes(pos.data_ptr(), 10, bad.ctypes.data, 3, packed.data_ptr(), -1, 254, labels.data_ptr(),
                          None, 0.0, big.data_ptr(), big.numel(), None)
    assert rc == -1 and b"view 1" in L.gsl_last_error()
    maps = torch.zeros(1024 + 1, dtype=torch.int32, device=DEV)
    err = torch.zeros(1, dtype=torch.int32, device=DEV)
    rc = L.gsl_pack_labels(maps.data_ptr() + 2, 1, 32, 32, packed.data_ptr(), -1, 254, err.data_ptr(), None)
    assert rc == -1 and b"aligned" in L.gsl_last_error()
    rc = L.gsl_pack_labels(maps.data_ptr(), 1, 0, 32, packed.data_ptr(), -1, 254, err.data_ptr(), None)
    assert rc == -1 and b"shape" in L.gsl_last_error()
    rc = L.gsl_lift_sweep(pos.data_ptr(), 10, views.ctypes.data, 3, None, -1, 254, labels.data_ptr(), None,
                          big.data_ptr(), big.numel(), None)
    assert rc == -1 and b"null packed" in L.gsl_last_error()
    with pytest.raises(native.GslError, match="shared memory"):
        ops.kmeans_assign(torch.zeros(100, 200, device=DEV), torch.zeros(900, 200, device=DEV))
    with pytest.raises(TypeError):
        ops.kmeans_assign(torch.zeros(10, 3, device=DEV, dtype=torch.float64), torch.zeros(2, 3, device=DEV))
    torch.cuda.synchronize()                                                   # nothing above left a sticky CUDA error
    assert ops.kmeans_assign(torch.zeros(4, 3, device=DEV), torch.zeros(2, 3, device=DEV)).tolist() == [0, 0, 0, 0]


def _random_camera(rng, i, width, height, integral):
    A = rng.standard_normal((3, 3))
    Q, _ = np.linalg.qr(A)
    if rng.random() < 0.3:
        Q = Q * rng.uniform(0.5, 2.0)                                          # not even orthonormal: the reference does not care
    w = int(width) if integral else float(width) + 0.37
    return {"id": i, "img_name": f"v{i}", "width": w, "height": int(height) if integral else float(height) + 0.61,
            "position": [float(v) for v in rng.standard_normal(3) * 3], "rotation": [[float(v) for v in r] for r in Q],
            "fx": float(rng.uniform(20, 400)), "fy": float(rng.uniform(20, 400))}


@pytest.mark.parametrize("seed", range(8))
def test_random_small_scenes_against_pure_python(oracle, seed):
    """Arbitrary rotations (incl. non-orthonormal), odd intrinsics, non-integer camera sizes, ragged
    map sizes, image size != map size: three implementations must agree label for label."""
    ops = pkg("ops")
    rng = np.random.default_rng(100 + seed)
    V = int(rng.integers(1, 7))
    N = int(rng.integers(1, 120))
    integral = seed % 3 != 2
    cams = [_random_camera(rng, i, rng.integers(8, 200), rng.integers(8, 200), integral) for i in range(V)]
    shapes = [(int(rng.integers(1, 90)), int(rng.integers(1, 90))) for _ in range(V)]
    sizes = [(int(rng.integers(1, 300)), int(rng.integers(1, 300))) for _ in range(V)]
    lo, hi = (-1, 5) if seed % 2 else (0, 252)             # the widest window: 254 codes from -1
    maps = [rng.integers(lo, hi + 1, size=s).astype(np.int32) for s in shapes]
    pos = (rng.standard_normal((N, 3)) * rng.uniform(0.5, 6)).astype(np.float32)
    want_py = oracle.lift_votes_py(pos, cams, maps, sizes)
    flat = np.concatenate([m.reshape(-1) for m in maps])
    want_c, _, _ = oracle.lift_votes(pos, oracle.make_views(cams, shapes, sizes), flat)
    assert np.array_equal(want_c, want_py), "C oracle and pure-Python restatement disagree"
    views = ops.make_views(cams, shapes, sizes)
    packed = ops.pack_labels(torch.from_numpy(flat).to(DEV), shapes)
    got = ops.lift_votes(torch.from_numpy(pos).to(DEV), views, packed).cpu().numpy()
    assert np.array_equal(got, want_c)


@pytest.mark.parametrize("n,d,k", [(300, 1, 4), (300, 8, 16), (1000, 64, 64), (1000, 64, 65), (700, 7, 64),
                                   (5, 12, 5), (513, 33, 17), (2000, 59, 15), (2000, 59, 16)])
def test_kmeans_path_boundaries(oracle, n, d, k):
    """Shapes on both sides of every dispatch boundary (tensor-core path needs 16 <= K <= 64 and
    8 <= D <= 64; float32 screening needs D <= 64) must all give the oracle's labels and sums."""
    ops = pkg("ops")
    rng = np.random.default_rng(n + 7 * d + k)
    data = rng.standard_normal((n, d)).astype(np.float32)
    cen = data[rng.choice(n, k, replace=False)] + (rng.standard_normal((k, d)) * 0.1).astype(np.float32)
    want = oracle.kmeans_assign(data, cen)
    lab, sums = ops.kmeans_step(torch.from_numpy(data).to(DEV), torch.from_numpy(cen).to(DEV))
    assert np.array_equal(lab.cpu().numpy().astype(np.int64), want)
    s = sums.cpu().numpy()
    for c in range(k):
        m = want == c
        assert s[c, d] == m.sum()
        assert np.allclose(s[c, :d], data[m].astype(np.float64).sum(0), rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("seed", range(6))
def test_screened_sweep_equals_float64_sweep_on_random_configs(oracle, seed, monkeypatch):
    """Differential test of the two sweeps (float32-screened default, GSLIFT_LIFT_F64=1) and the
    oracle over random scenes that mix, inside one 16-view window, integer frames with and
    without rescaling, maps larger and smaller than the frame (clamp active), one non-integer
    frame (sent to the float64 kernel), and intrinsics from wide-angle to telephoto."""
    ops, scene = pkg("ops"), pkg("scene")
    rng = np.random.default_rng(500 + seed)
    V = int(rng.integers(17, 40))
    cams = scene.lookat_cameras(V, width=200, height=120, seed=600 + seed)
    shapes, sizes = [], []
    for i, c in enumerate(cams):
        w, h = int(rng.integers(40, 400)), int(rng.integers(30, 300))
        c["width"], c["height"] = w, h
        c["fx"], c["fy"] = float(rng.uniform(0.3, 6.0) * w), float(rng.uniform(0.3, 6.0) * h)
        mode = int(rng.integers(0, 4))
        if mode == 0:                                   # map == frame, unit scale: zero-ring variant
            shapes.append((h, w)); sizes.append((w, h))
        elif mode == 1:                                 # map != frame, unit scale (clamp may fire)
            shapes.append((int(rng.integers(10, 300)), int(rng.integers(10, 400)))); sizes.append((shapes[-1][1], shapes[-1][0]))
        else:                                           # rescaled
            shapes.append((int(rng.integers(10, 300)), int(rng.integers(10, 400)))); sizes.append((int(rng.integers(20, 500)), int(rng.integers(20, 500))))
    if seed % 2:
        cams[3]["width"] = cams[3]["width"] + 0.5       # non-integer frame: the view takes the float64 path
    maps = [rng.integers(-1, 150, size=s).astype(np.int32) for s in shapes]
    pos = (rng.standard_normal((30_000, 3)) * rng.uniform(0.5, 3.0)).astype(np.float32)
    flat = np.concatenate([m.reshape(-1) for m in maps])
    want, _, _ = oracle.lift_votes(pos, oracle.make_views(cams, shapes, sizes), flat)
    views = ops.make_views(cams, shapes, sizes)
    packed = ops.pack_labels(torch.from_numpy(flat).to(DEV), shapes)
    d_pos = torch.from_numpy(pos).to(DEV)
    got = ops.lift_votes(d_pos, views, packed).cpu().numpy()
    monkeypatch.setenv("GSLIFT_LIFT_F64", "1")
    got64 = ops.lift_votes(d_pos, views, packed).cpu().numpy()
    monkeypatch.delenv("GSLIFT_LIFT_F64")
    monkeypatch.setenv("GSLIFT_LIFT_ORDER", "0")
    got_plain = ops.lift_votes(d_pos, views, packed).cpu().numpy()
    monkeypatch.delenv("GSLIFT_LIFT_ORDER")
    assert np.array_equal(got, got64) and np.array_equal(got, got_plain)
    assert np.array_equal(got, want), f"{(got != want).sum()} labels differ from the oracle"
