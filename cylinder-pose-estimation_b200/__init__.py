"""lgx: B200-native laser-grid point extractor (stages 1-2 of the reference's detect_grid).

Directory name `cylinder-pose-estimation_b200` is not an identifier; import it as
`cylinder_pose_estimation_b200` (repo-root shim) or put this directory on sys.path and import the
drop-in modules by the reference's names (python_grid_detection_cylinder / _plane).
"""
from . import _lib, synth, frontend, iotool    # noqa: F401
from .frontend import (Frontend, FrontendResult, load_and_preprocess_image, extract_joints,   # noqa: F401
                       detect_points_batch, get_frontend, stage12_batch, unpack_mask)

__all__ = ["Frontend", "FrontendResult", "load_and_preprocess_image", "extract_joints",
           "detect_points_batch", "get_frontend", "stage12_batch", "unpack_mask", "synth", "iotool"]
