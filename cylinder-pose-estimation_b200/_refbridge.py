"""Loads the reference's own utils/util_cylinder.py / utils/util_plane.py (stages 3-6 stay the reference's
unchanged code) and swaps the two hot-path functions for the lgx ones.

The reference checkout is found through $LGX_REFERENCE_ROOT (the directory that holds
python_grid_detection_cylinder.py and utils/).  Nothing of the reference is vendored in this repository.
"""
import importlib.util
import os
import sys

from . import frontend

_loaded = {}


def reference_root():
    root = os.environ.get("LGX_REFERENCE_ROOT")
    if not root or not os.path.isfile(os.path.join(root, "utils", "util_cylinder.py")):
        raise ImportError("set LGX_REFERENCE_ROOT to a checkout of cv3vpl-lab/cylinder-pose-estimation "
                          "(stages 3-6 of detect_grid are the reference's own code)")
    return root


def load_reference_utils(name):
    """name: 'util_cylinder' or 'util_plane'.  Returns a private copy of the reference module whose
    load_and_preprocess_image / extract_joints are the B200 implementations."""
    if name in _loaded:
        return _loaded[name]
    root = reference_root()
    if root not in sys.path:
        sys.path.append(root)      # the reference modules import `utils.*` siblings by package name
    path = os.path.join(root, "utils", name + ".py")
    spec = importlib.util.spec_from_file_location("lgx_ref_" + name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.reference_load_and_preprocess_image = mod.load_and_preprocess_image
    mod.reference_extract_joints = mod.extract_joints
    mod.load_and_preprocess_image = frontend.load_and_preprocess_image
    mod.extract_joints = frontend.extract_joints
    _loaded[name] = mod
    return mod
