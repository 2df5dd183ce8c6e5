"""gslift: B200-native (sm_100a) label lifting and K-means labelling for 3D Gaussian splats.

The package name starts with a digit, so import it with importlib:

    import importlib
    gs = importlib.import_module("3d_gaussian_splatting_project_b200")
    labels = gs.deep_learning_segmentation.assign_labels(gaussians, cameras, in_dir, out_dir)

Modules
    deep_learning_segmentation   drop-in for the reference script of the same name
    k_means                      drop-in for 3D_clustering/k_means.py
    ops                          device-tensor operators over the C ABI (include/gslift.h)
    plyio                        PLY reader / writer (plyfile stand-in)
    scene                        synthetic scenes for tests and benchmarks
    sharding                     one-process-per-GPU helpers (Gaussian slices, NCCL)
    viewer                       the viewer worker's depth sort and click hit test (gaussians_selection.js)
"""
from importlib import import_module as _imp

__all__ = ["deep_learning_segmentation", "k_means", "ops", "plyio", "scene", "sharding", "viewer"]


def __getattr__(name):
    if name in __all__:
        return _imp(f"{__name__}.{name}")
    raise AttributeError(name)
