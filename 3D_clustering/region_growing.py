"""Entry point with the reference's path and name:

    python 3D_clustering/region_growing.py [input.ply [output.ply]]

(the reference hard-codes data\\point_cloud.ply and 3D_clustering\\clustering.ply, its defaults here).
The implementation is 3d_gaussian_splatting_project_b200/region_growing.py.
"""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
_impl = importlib.import_module("3d_gaussian_splatting_project_b200.region_growing")
globals().update({k: getattr(_impl, k) for k in dir(_impl) if not k.startswith("__")})

if __name__ == "__main__":
    _impl.main(*sys.argv[1:3])
