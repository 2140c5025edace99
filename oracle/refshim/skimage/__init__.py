"""Stub of scikit-image (pinned 0.19.3 in the reference's requirements.txt:5;
not installable here: no network).  Restates the three functions the hot path
reaches (utils/util_cylinder.py:10-11, :1736-1737) from the published
scikit-image 0.19 source: see feature.py."""
import numpy as np


def img_as_float(a):
    a = np.asarray(a)
    if a.dtype == np.uint8:
        return a / 255.0
    if a.dtype == np.uint16:
        return a / 65535.0
    return a.astype(np.float64)
