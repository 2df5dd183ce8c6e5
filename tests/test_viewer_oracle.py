"""CPU: the viewer-side oracle (oracle/viewer_oracle.c) against hand-computed known answers and an
independent pure-Python restatement.  No JavaScript engine exists in the build container, so the
pin is the language semantics (ECMA-262) these cases spell out; see the header of viewer_oracle.c.
Citations: gs = Web_Viewer_Gaussians_Selection/gaussians_selection.js (reference)."""
import math

import numpy as np
import pytest


def _z_cloud(zs):
    pos = np.zeros((len(zs), 3), np.float32)
    pos[:, 2] = zs
    vp = np.zeros(16)
    vp[10] = 1.0
    return pos, vp


def test_to_int32_known_answers(oracle):
    # ECMA-262 7.1.6: truncate, modulo 2^32, upper half wraps negative; NaN and infinities give 0
    cases = [(0.0, 0), (-0.0, 0), (1.9, 1), (-1.9, -1), (2147483647.0, 2147483647), (2147483648.0, -2147483648),
             (4294967296.0 + 5.5, 5), (-2147483649.0, 2147483647), (float("nan"), 0), (float("inf"), 0),
             (-float("inf"), 0), (1e20, 1661992960)]          # 1e20 mod 2^32 = 1661992960
    for d, want in cases:
        assert oracle.js_to_int32_py(d) == want, d
    assert 10 ** 20 % 2 ** 32 == 1661992960


# depth = z * 4096 exactly; buckets by hand: ((depth - min) * (65536 / (max - min))) | 0
@pytest.mark.parametrize("zs, buckets, index", [
    # 65536 / 12288 * 12288 == 65536.0 in binary64, so the deepest Gaussian lands in bucket 65536: outside
    # the 65536-entry typed arrays (gs:444, :450) -> never stored, the tail of depthIndex stays 0 (gs:453)
    ([0, 1, 2, 3], [0, 21845, 43690, 65536], [0, 1, 2, 0]),
    ([3, 1, 2, 0, 3], [65536, 21845, 43690, 0, 65536], [3, 1, 2, 0, 0]),
    # equal depths: max - min = 0, depthInv = Infinity, 0 * Infinity = NaN, NaN | 0 = 0: input order kept
    ([5, 5, 5], [0, 0, 0], [0, 1, 2]),
    # stability inside a bucket (counting sort scatters in index order, gs:454-457)
    ([2, 0, 1, 0, 2, 1, 2], [65536, 0, 32768, 0, 65536, 32768, 65536], [1, 3, 2, 5, 0, 0, 0]),
    ([0, 7], [0, 65536], [0, 0]),
])
def test_depth_sort_known_answers(oracle, zs, buckets, index):
    pos, vp = _z_cloud(zs)
    assert oracle.viewer_buckets(pos, vp).tolist() == buckets
    assert oracle.viewer_depth_sort(pos, vp).tolist() == index
    assert oracle.viewer_depth_sort_py(pos, vp).tolist() == index


def test_depth_sort_c_equals_python_on_random_clouds(oracle):
    rng = np.random.default_rng(11)
    for n, stride in ((1, 3), (2, 8), (257, 8), (4000, 3)):
        pos = (rng.normal(size=(n, stride)) * rng.choice([0.01, 1.0, 300.0])).astype(np.float32)
        vp = rng.normal(size=16)
        a, b = oracle.viewer_depth_sort(pos, vp), oracle.viewer_depth_sort_py(pos, vp)
        assert a.dtype == np.uint32 and (a == b).all()
    # non-finite positions: NaN | 0 == 0 (gs:437); a huge coordinate wraps modulo 2^32
    pos = np.array([[0, 0, 1], [0, 0, np.nan], [0, 0, np.inf], [0, 0, 3e9], [0, 0, -2]], np.float32)
    vp = np.zeros(16); vp[10] = 1.0
    assert (oracle.viewer_depth_sort(pos, vp) == oracle.viewer_depth_sort_py(pos, vp)).all()


def test_multiply4_known_answer(oracle):
    a = np.arange(1, 17, dtype=np.float64)
    ident = np.eye(4).reshape(16)
    assert (oracle.multiply4(a, ident) == a).all() and (oracle.multiply4(ident, a) == a).all()
    # gs:110-123 is column-major: result[4r + c] = sum_k b[4r + k] * a[c + 4k]
    b = np.array([2, 0, 0, 0, 0, 3, 0, 0, 0, 0, 4, 0, 1, 1, 1, 1], np.float64)
    want = [sum(b[4 * r + k] * a[c + 4 * k] for k in range(4)) for r in range(4) for c in range(4)]
    assert oracle.multiply4(a, b).tolist() == want


def _ortho(vw, vh):
    """Matrix that maps (x, y, z, 1) to clip = (2x/vw - 1, 2y/vh - 1, z, 1): screen = (x, y), depth = z."""
    m = np.zeros(16)
    m[0], m[12] = 2.0 / vw, -1.0
    m[5], m[13] = 2.0 / vh, -1.0
    m[10] = 1.0
    m[15] = 1.0
    return m


def test_hit_test_known_answers(oracle):
    vw, vh = 512.0, 256.0          # powers of two: the screen coordinates below are exact
    m = _ortho(vw, vh)
    pos = np.array([[100, 100, 5], [103, 104, 9], [97, 96, 1], [103, 104, 2], [300, 200, 0], [110, 100, 0]], np.float32)
    labels = np.array([10, 11, 12, 13, 14, 15], np.int32)
    # click on Gaussian 0: distance 0 wins
    assert oracle.viewer_hit_test(pos, labels, m, 100, 100, (vw, vh)) == (10, 0)
    # remove it: 1, 2, 3 all at distance exactly 5 (3-4-5); the smallest depth wins -> Gaussian 2 (gs:387)
    assert oracle.viewer_hit_test(pos[1:], labels[1:], m, 100, 100, (vw, vh)) == (12, 1)
    # equal distance and equal depth: the first one in index order stays
    p2 = pos[[1, 3, 3]]; p2[:, 2] = 7
    assert oracle.viewer_hit_test(p2, np.array([1, 2, 3], np.int32), m, 100, 100, (vw, vh)) == (1, 0)
    # Gaussian 5 at distance exactly 10 is NOT selected (dist < 10, gs:387); nothing else near
    assert oracle.viewer_hit_test(pos[4:], labels[4:], m, 100, 100, (vw, vh)) == (oracle.NO_SELECTION, -1)
    # w <= 0 is skipped (gs:403)
    mw = m.copy(); mw[15] = 0.0; mw[11] = 1.0            # w = z
    behind = np.array([[100, 100, -1], [100, 100, 0]], np.float32)
    assert oracle.viewer_hit_test(behind, np.array([1, 2], np.int32), mw, 100, 100, (vw, vh)) == (oracle.NO_SELECTION, -1)
    for args in ((pos, labels, m, 100, 100, (vw, vh)), (pos[1:], labels[1:], m, 101.5, 99.25, (vw, vh))):
        assert oracle.viewer_hit_test(*args) == oracle.viewer_hit_test_py(*args)


def test_hit_test_c_equals_python_and_hypot_form(oracle):
    rng = np.random.default_rng(5)
    pos = rng.normal(size=(3000, 8)).astype(np.float32)
    labels = rng.integers(-1, 150, 3000).astype(np.int32)
    view = np.eye(4).reshape(16).copy(); view[14] = 6.0                       # push the cloud to z = 6
    f = 1.2
    proj = np.array([f, 0, 0, 0, 0, f, 0, 0, 0, 0, 1.01, 1, 0, 0, -0.2, 0], np.float64)   # w = z
    m = oracle.multiply4(proj, view)
    hits = 0
    for k in range(40):
        x, y = rng.uniform(0, 800), rng.uniform(0, 600)
        a = oracle.viewer_hit_test(pos, labels, m, x, y, (800, 600))
        assert a == oracle.viewer_hit_test_py(pos, labels, m, x, y, (800, 600))
        hits += a[1] >= 0
    assert hits > 5
    # the two-term Kahan loop of V8's Math.hypot collapses to sqrt(n0^2 + n1^2) * max: 3-4-5 is exact
    assert math.sqrt((3 / 4) ** 2 + 1.0) * 4 == 5.0
