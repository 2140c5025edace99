"""Import the UNMODIFIED reference (read-only checkout) in this container.

Build-container only: /root/reference does not exist on the GPU box, so nothing
that runs there may call this.  Used by oracle/make_golden.py and by the CPU
tests that pin oracle/ref_port.py against the reference's own functions.

Shim: np.RankWarning alias (utils/util_cylinder.py:17 uses a
name NumPy 2 removed), stub matplotlib.pyplot, stub skimage.
"""
import importlib
import os
import sys

REFERENCE_ROOT = os.environ.get("LGX_REFERENCE_ROOT", "/root/reference")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "refshim")


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "utils", "util_cylinder.py"))


def load():
    """Returns (python_grid_detection_cylinder, python_grid_detection_plane)."""
    if not available():
        raise RuntimeError(f"reference checkout not found at {REFERENCE_ROOT}")
    import numpy as np
    if not hasattr(np, "RankWarning"):
        np.RankWarning = np.exceptions.RankWarning
    for p in (_SHIM, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    cyl = importlib.import_module("python_grid_detection_cylinder")
    pla = importlib.import_module("python_grid_detection_plane")
    return cyl, pla
