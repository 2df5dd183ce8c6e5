"""The oracle against the reference's own outputs (tests/golden, made by oracle/make_golden.py)
and against the third-party arithmetic it restates (NumPy mean, scipy cKDTree)."""
import numpy as np
import pytest

from util import GOLDEN, KMEANS_CASES, LIFT_CASES, load_kmeans_case, load_lift_case


def test_probe_dgemv_rounding(oracle):
    z = np.load(f"{GOLDEN}/probes.npz")
    for i in range(len(z["R"])):
        assert np.array_equal(oracle.translation(z["R"][i], z["p"][i]), z["t"][i])


def test_probe_scipy_distance_order(oracle):
    z = np.load(f"{GOLDEN}/probes.npz")
    for D in (3, 6, 7, 8, 59, 61):
        c, x, d, idx = z[f"c{D}"], z[f"x{D}"], z[f"d{D}"], z[f"i{D}"]
        lab = oracle.kmeans_assign(x, c)
        assert np.array_equal(lab, idx)
        mine = np.array([np.sqrt(oracle.sqdist(c[idx[j]], x[j])) for j in range(len(x))])
        assert np.array_equal(mine, d), f"D={D}: scipy summation order not reproduced"


@pytest.mark.parametrize("name", LIFT_CASES)
def test_lift_matches_verbatim_reference(oracle, name):
    c = load_lift_case(name)
    views = oracle.make_views(c["cameras"], c["shapes"], c["sizes"])
    labels, _, _ = oracle.lift_votes(c["pos"], views, c["flat"], label_min=-1, n_classes=255)
    assert np.array_equal(labels, c["labels"])


@pytest.mark.parametrize("name", ["lift_lookat_regions", "lift_degenerate"])
def test_lift_pure_python_restatement(oracle, name):
    c = load_lift_case(name)
    n = min(len(c["pos"]), 300)
    with np.errstate(all="ignore"):
        lab = oracle.lift_votes_py(c["pos"][:n], c["cameras"], c["maps"], c["sizes"])
    assert np.array_equal(lab, c["labels"][:n])


@pytest.mark.parametrize("name", KMEANS_CASES)
def test_kmeans_matches_verbatim_reference(oracle, name):
    c = load_kmeans_case(name)
    np.random.seed(c["seed"])
    cen, lab, iters = oracle.kmeans_run(c["data"], c["k"], max_iter=c["max_iter"])
    assert np.array_equal(lab, c["labels"])
    assert np.array_equal(cen, c["centroids"]), "float32 sequential mean not reproduced"
    printed_iters = sum(1 for ln in c["stdout"].splitlines() if ln.strip().isdigit())
    assert iters == printed_iters


def test_kmeans_assign_vs_scipy_and_update_vs_numpy(oracle):
    from scipy.spatial import KDTree
    rng = np.random.default_rng(0)
    for (N, D, K) in [(20000, 6, 10), (8000, 59, 64), (5000, 5, 3)]:
        data = rng.standard_normal((N, D)).astype(np.float32)
        cen = data[rng.choice(N, K, replace=False)]
        lab, gap = oracle.kmeans_assign(data, cen, want_gap=True)
        _, ref = KDTree(cen).query(data)
        bad = lab != ref
        assert not np.any(bad & (gap > 0)), "mismatch away from an exact tie"
        new, counts = oracle.kmeans_update(data, lab, cen)
        for k in range(K):
            m = lab == k
            want = data[m].mean(axis=0) if m.any() else cen[k]
            assert np.array_equal(new[k], want)
            assert counts[k] == m.sum()


def test_update_large_cluster_is_sequential_f32(oracle):
    rng = np.random.default_rng(1)
    data = (rng.standard_normal((300000, 4)) + 3).astype(np.float32)
    lab = np.zeros(len(data), np.int64)
    lab[::7] = 1
    cen = np.zeros((3, 4), np.float32)
    new, _ = oracle.kmeans_update(data, lab, cen)
    for k in range(2):
        assert np.array_equal(new[k], data[lab == k].mean(axis=0))
    assert np.array_equal(new[2], cen[2])
