"""The C-ABI library: loads, exports every symbol include/gslift.h declares, fails loudly
without a device.  No compute calls here (no GPU on the CPU runner)."""
import ctypes
import os
import re

import pytest

from util import pkg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "gslift.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gsl_[a-z_0-9]+)\s*\(", text)))


def test_library_is_built_and_exports_every_declared_symbol():
    native = pkg("_native")
    assert os.path.exists(native.LIB_PATH), "run __graft_entry__.build() first"
    raw = ctypes.CDLL(native.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 12
    for name in names:
        assert hasattr(raw, name), f"{name} declared in gslift.h but not exported"
    assert sorted(native.SIGNATURES) == names, "ctypes SIGNATURES out of step with gslift.h"


def test_version_and_view_struct_size():
    native = pkg("_native")
    assert native.lib().gsl_version() == native.ABI_VERSION == 4
    hdr = open(os.path.join(ROOT, "include", "gslift.h")).read()
    assert "176 bytes" in hdr and native.VIEW_DTYPE.itemsize == 176


def test_workspace_queries_are_host_only():
    L = pkg("_native").lib()
    assert L.gsl_lift_workspace_bytes(0, 0) >= 0
    n = L.gsl_lift_workspace_bytes(1000, 10)
    assert n >= 1024 * (12 + 4)                  # sorted positions + permutation
    assert L.gsl_lift_workspace_bytes(1000, 300) > n
    assert L.gsl_kmeans_workspace_bytes(100000, 59, 64) >= 64 * 60 * 8


def test_argument_validation_needs_no_device():
    native = pkg("_native")
    L = native.lib()
    assert L.gsl_pack_labels(None, 1, 8, 8, None, -1, 254, None, None) == -1
    assert L.gsl_packed_map_bytes(1920, 1080) == 122 * 137 * 128 + 33440 and L.gsl_packed_map_bytes(0, 5) == 0
    assert b"null" in L.gsl_last_error()
    assert L.gsl_lift_votes(None, -1, None, 0, None, -1, 254, None, None, 0.0, None, 0, None) == -1
    assert L.gsl_kmeans_assign(None, 10, 0, None, 4, None, None, 0, None) == -1
    assert L.gsl_kmeans_assign(None, 10, 3, None, 4000, None, None, 0, None) == -1
    with pytest.raises(native.GslError):
        native.check(L.gsl_kmeans_finalize(None, None, 4, 3, None, None, None))


def test_next_row_entry_points_validate_arguments_without_a_device():
    """Viewer / region-growing entry points: bad arguments are refused before any
    CUDA call, workspace queries are host arithmetic."""
    import numpy as np
    native = pkg("_native")
    L = native.lib()
    vp = np.zeros(16)
    assert L.gsl_viewer_depth_sort(None, 10, 8, vp.ctypes.data, None, None, 0, None) == -1
    assert L.gsl_viewer_depth_sort(None, 0, 8, vp.ctypes.data, None, None, 0, None) == 0          # empty cloud: nothing to do
    assert L.gsl_viewer_depth_sort(None, 10, 2, vp.ctypes.data, None, None, 0, None) == -1        # stride < 3
    assert L.gsl_viewer_hit_test(None, None, 10, 8, vp.ctypes.data, 0.0, 0.0, 1.0, 1.0, -999999, None, None, None, 0, None) == -1
    assert L.gsl_viewer_sort_workspace_bytes(1_000_000) >= 3 * 4 * 1_000_000 and L.gsl_viewer_hit_workspace_bytes() > 0
    assert L.gsl_region_knn_pca(None, 10, 0, None, None, None, None, None, None, None, 0, None) == -1    # k < 1
    assert L.gsl_region_knn_pca(None, 10, 3, None, None, None, None, None, None, None, 0, None) == -1    # null points
    assert L.gsl_region_knn_pca(None, 0, 3, None, None, None, None, None, None, None, 0, None) == 0
    assert L.gsl_region_workspace_bytes(200_000) >= 200_000 * (16 + 16) + 8 ** 7 * 4


def test_region_grow_host_helper_equals_oracle(oracle):
    """gsl_region_grow is host code inside the product library (the growth loop of segmentation_3D, rg:188-215,
    is serial): same partition as the oracle's restatement and as the reference run verbatim."""
    import numpy as np
    from util import GOLDEN
    L = pkg("_native").lib()
    for name in ("region_planes_k40", "region_planes_k400", "region_planes_k2000"):
        g = np.load(os.path.join(GOLDEN, name + ".npz"))
        knn = np.ascontiguousarray(g["knn_seg"], np.int32)
        normals = np.ascontiguousarray(g["normals"], np.float64)
        residuals = np.ascontiguousarray(g["residuals"], np.float64)
        n, k = knn.shape
        region_of = np.empty(n, np.int32)
        sizes = np.empty(n, np.int64)
        r = L.gsl_region_grow(knn.ctypes.data, k, normals.ctypes.data, residuals.ctypes.data, n, 0.1, 0.05,
                              region_of.ctypes.data, sizes.ctypes.data)
        want, n_want = oracle.region_grow(knn, normals, residuals, 0.1, 0.05)
        assert r == n_want == int(g["n_regions"])
        assert np.array_equal(region_of, want)                        # both number regions in creation order
        assert np.array_equal(np.bincount(region_of, minlength=r), sizes[:r])
        assert len(set(zip(region_of.tolist(), g["region_of"].tolist()))) == r


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    ops = pkg("ops")
    L = pkg("_native").lib()
    assert L.gsl_device_count() == -3            # GSL_ECUDA, with a message
    assert L.gsl_last_error()
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.lift_votes(torch.zeros(4, 3), ops.make_views([], []), torch.zeros(0, dtype=torch.uint8))
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.kmeans_assign(torch.zeros(4, 3), torch.zeros(2, 3))


def test_product_never_imports_the_oracle():
    pkg_dir = os.path.join(ROOT, "3d_gaussian_splatting_project_b200")
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "libgsl_oracle" not in text, f


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """include/gslift.h must compile as C (no C++, no torch types) and a C program must be able to
    call the library: the boundary a non-Python host would bind."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    native = pkg("_native")
    src = tmp_path / "abi_probe.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "gslift.h"
int main(void)
{
    GslView v;
    if (sizeof(v) != 176) return 2;
    if (gsl_version() != GSL_ABI_VERSION) return 3;
    if (gsl_packed_map_bytes(1920, 1080) != 122LL * 137 * 128 + 33440) return 4;
    if (gsl_kmeans_exchange_bytes(8, 59, 64) != 256 + 2u * 8 * 64 * 60 * 8) return 5;
    if (gsl_lift_votes(NULL, -1, NULL, 0, NULL, -1, 254, NULL, NULL, 0.0, NULL, 0, NULL) != GSL_EINVAL) return 6;
    if (strlen(gsl_last_error()) == 0) return 7;
    printf("abi ok %d\n", gsl_version());
    return 0;
}
''')
    exe = tmp_path / "abi_probe"
    lib_dir = os.path.dirname(native.LIB_PATH)
    cc = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                         "-L", lib_dir, "-lgslift", f"-Wl,-rpath,{lib_dir}"], capture_output=True, text=True)
    assert cc.returncode == 0, cc.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True)
    assert run.returncode == 0 and "abi ok 4" in run.stdout, (run.returncode, run.stdout, run.stderr)
