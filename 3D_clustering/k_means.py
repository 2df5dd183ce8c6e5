"""Entry point with the reference's path, name and flags:

    python 3D_clustering/k_means.py --file_path in.ply --save_path out.ply [--k 10]

The implementation is 3d_gaussian_splatting_project_b200/k_means.py.
"""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
_impl = importlib.import_module("3d_gaussian_splatting_project_b200.k_means")
globals().update({k: getattr(_impl, k) for k in dir(_impl) if not k.startswith("__")})

if __name__ == "__main__":
    _impl.main()
