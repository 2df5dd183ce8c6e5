"""In-tree build of libgslift.so (nvcc, sm_100a only).  `python -m` friendly:

    python 3d_gaussian_splatting_project_b200/build.py [--force]

The .so is written next to this file so it travels with the repository snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libgslift.so")
SOURCES = ["api.cu", "lift.cu", "lift_order.cu", "lift_sort.cu", "kmeans.cu", "kmeans_tc.cu", "kmeans_ordered.cu", "ply_format.cu", "host_stage.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",            # no implicit contraction: fused ops are spelled fma() in the source
    "-Xcompiler", "-fPIC", "-Xcompiler", "-pthread", "-shared",
    "-I", os.path.join(ROOT, "include"),
]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "gslift.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    env = dict(os.environ)
    env.pop("CC", None)       # the image exports a CC wrapper that nvcc must not pick up
    env.pop("CXX", None)
    subprocess.check_call(cmd, env=env)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
