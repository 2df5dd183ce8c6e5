"""Entry point with the reference's name and flags:

    python deep_learning_segmentation.py --ply_file ... --camera_file ... --input_dir ... \
        --output_dir ... --output_file ... [--model segformer|mask2former|yolo]

The implementation is 3d_gaussian_splatting_project_b200/deep_learning_segmentation.py.
"""
import importlib

_impl = importlib.import_module("3d_gaussian_splatting_project_b200.deep_learning_segmentation")
globals().update({k: getattr(_impl, k) for k in dir(_impl) if not k.startswith("__")})

if __name__ == "__main__":
    _impl.main()
