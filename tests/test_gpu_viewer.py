"""GPU parity of the viewer-side label consumers (csrc/viewer.cu through the C ABI) with the
oracle (oracle/viewer_oracle.c): depthIndex and the hit-test selection must be identical.
Citations: gs = Web_Viewer_Gaussians_Selection/gaussians_selection.js (reference)."""
import numpy as np
import pytest
import torch

from util import pkg

pytestmark = pytest.mark.gpu


def _cloud(rng, n, stride, scale):
    return (rng.normal(size=(n, stride)) * scale).astype(np.float32)


@pytest.mark.parametrize("n, stride, scale", [(1, 3, 1.0), (2, 8, 1.0), (31, 8, 5.0), (8191, 3, 1.0), (8192, 8, 1.0),
                                              (8193, 8, 0.001), (100_003, 8, 2.0), (1_000_000, 3, 40.0)])
def test_run_sort_equals_oracle(oracle, n, stride, scale):
    viewer = pkg("viewer")
    rng = np.random.default_rng(n)
    pos = _cloud(rng, n, stride, scale)
    vp = rng.normal(size=16)
    want = oracle.viewer_depth_sort(pos, vp)
    got = viewer.run_sort(torch.from_numpy(pos).cuda(), vp).cpu().numpy()
    assert got.dtype == np.uint32 and np.array_equal(got, want)
    # the deepest Gaussians fall off the 65536-entry tables (bucket 65536) and leave zeros at the tail
    dropped = int((oracle.viewer_buckets(pos, vp) == 65536).sum())
    if dropped:
        assert (got[n - dropped:] == 0).all()


def test_run_sort_edge_cases(oracle):
    viewer = pkg("viewer")
    vp = np.zeros(16); vp[10] = 1.0
    for zs in ([0, 1, 2, 3], [3, 1, 2, 0, 3], [5, 5, 5], [2, 0, 1, 0, 2, 1, 2], [0, 7],
               [1, np.nan, np.inf, 3e9, -2, -np.inf, 1e-30]):
        pos = np.zeros((len(zs), 3), np.float32); pos[:, 2] = zs
        got = viewer.run_sort(pos, vp).cpu().numpy()
        assert np.array_equal(got, oracle.viewer_depth_sort(pos, vp)), zs
    assert viewer.run_sort(np.zeros((0, 3), np.float32), vp).numel() == 0
    # heavy ties: 200 000 Gaussians on 7 depth planes -> the order inside a bucket is the input order
    rng = np.random.default_rng(3)
    pos = np.zeros((200_000, 8), np.float32); pos[:, 2] = rng.integers(0, 7, 200_000)
    got = viewer.run_sort(pos, vp).cpu().numpy()
    assert np.array_equal(got, oracle.viewer_depth_sort(pos, vp))


def _camera(oracle, dist=6.0, f=1.2):
    view = np.eye(4).reshape(16).copy(); view[14] = dist
    proj = np.array([f, 0, 0, 0, 0, f, 0, 0, 0, 0, 1.01, 1, 0, 0, -0.2, 0], np.float64)
    return view, proj


@pytest.mark.parametrize("n, stride", [(1, 3), (300, 8), (50_000, 8), (2_000_000, 3)])
def test_hit_test_equals_oracle(oracle, n, stride):
    viewer = pkg("viewer")
    rng = np.random.default_rng(n + 1)
    pos = _cloud(rng, n, stride, 1.0)
    labels = rng.integers(-1, 150, n).astype(np.int32)
    view, proj = _camera(oracle)
    m = oracle.multiply4(proj, view)
    assert viewer.multiply4(proj, view) == m.tolist()
    dpos, dlab = torch.from_numpy(pos).cuda(), torch.from_numpy(labels).cuda()
    hits = 0
    for k in range(12):
        x, y = rng.uniform(250, 550), rng.uniform(150, 450)
        want = oracle.viewer_hit_test(pos, labels, m, x, y, (800, 600))
        got = viewer.perform_hit_testing(x, y, view, proj, (800, 600), dpos, dlab, return_index=True)
        assert got == want, (k, got, want)
        hits += want[1] >= 0
    if n >= 50_000:
        assert hits >= 6


def test_hit_test_ties_and_skips(oracle):
    viewer = pkg("viewer")
    vw, vh = 512.0, 256.0
    proj = np.zeros(16)
    proj[0], proj[12], proj[5], proj[13], proj[10], proj[15] = 2.0 / vw, -1.0, 2.0 / vh, -1.0, 1.0, 1.0
    view = np.eye(4).reshape(16)
    pos = np.array([[100, 100, 5], [103, 104, 9], [97, 96, 1], [103, 104, 2], [300, 200, 0], [110, 100, 0]], np.float32)
    labels = np.array([10, 11, 12, 13, 14, 15], np.int32)
    hit = lambda p, l, x=100, y=100: viewer.perform_hit_testing(x, y, view, proj, (vw, vh), p, l, return_index=True)
    assert hit(pos, labels) == (10, 0)
    assert hit(pos[1:], labels[1:]) == (12, 1)                       # distance tie -> smallest depth
    p2 = pos[[1, 3, 3]].copy(); p2[:, 2] = 7
    assert hit(p2, np.array([1, 2, 3], np.int32)) == (1, 0)         # full tie -> first index
    assert hit(pos[4:], labels[4:]) == (viewer.NO_SELECTION, -1)     # dist == 10 is not < 10
    # a tie-heavy crowd: 100 000 Gaussians on a 3 x 3 px lattice around the click with 4 depth values
    rng = np.random.default_rng(9)
    crowd = np.zeros((100_000, 8), np.float32)
    crowd[:, 0] = 100 + rng.integers(-1, 2, 100_000)
    crowd[:, 1] = 100 + rng.integers(-1, 2, 100_000)
    crowd[:, 2] = rng.integers(0, 4, 100_000)
    lab = rng.integers(0, 1000, 100_000).astype(np.int32)
    m = oracle.multiply4(proj, view)
    for x, y in ((100, 100), (100.5, 100.5), (99, 101.25)):
        assert hit(crowd, lab, x, y) == oracle.viewer_hit_test(crowd, lab, m, x, y, (vw, vh))
    # NaN depth on a distance tie: reachable only through overflow (r2 = w = Infinity -> depth = NaN while
    # x / w = 0 keeps the distance finite).  The sequential scan keeps the FIRST Gaussian at the smallest
    # distance when its depth is NaN, and never lets a later NaN displace the current one (gs:387).
    pn = np.zeros(16)
    pn[6], pn[10], pn[11], pn[15] = 1.0, 1e300, 1e300, 1.0       # r2 = y + 1e300 z, w = 1e300 z + 1, x = y = 0
    mn = oracle.multiply4(pn, view)
    nan_pt, near, far = [0, 0, 3e38], [0, 0.5, 0], [0, 1.0, 0]
    for order, want in (([nan_pt, far, near], 0), ([far, nan_pt, near], 2), ([near, nan_pt, far], 0),
                        ([nan_pt, nan_pt, near], 0), ([far, near, nan_pt], 1)):
        p3 = np.array(order, np.float32)
        l3 = np.array([7, 8, 9], np.int32)
        ref = oracle.viewer_hit_test(p3, l3, mn, 256, 128, (vw, vh))
        assert ref == (int(l3[want]), want), (order, ref)
        got = viewer.perform_hit_testing(256, 128, view, pn, (vw, vh), p3, l3, return_index=True)
        assert got == ref, (order, got, ref)
