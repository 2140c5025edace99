"""Stub of scikit-image (pinned 0.19.3 in the reference's requirements.txt:5;
not installable here: no network).  Restates the three functions the hot path
reaches (utils/util_cylinder.py:10-11, :1736-1737) from the published
scikit-image 0.19 source: see feature.py."""
import numpy as np


# scikit-image 0.19 `util.dtype._convert`, unsigned -> float branch:
#     image = np.multiply(image, 1. / imax_in, dtype=computed_float_type)
# i.e. a multiplication by the rounded reciprocal, NOT a division (the two differ by one ulp for 24 of the 256
# 8-bit levels and 88 of the 65536 16-bit levels).  MULTIPLY_BY_RECIPROCAL = False restores the division
# (what round 1 of this repository assumed; kernels: LGX_OPT_FLOAT_DIV).
MULTIPLY_BY_RECIPROCAL = True


def img_as_float(a):
    a = np.asarray(a)
    if a.dtype == np.uint8 or a.dtype == np.uint16:
        imax = float(np.iinfo(a.dtype).max)
        if MULTIPLY_BY_RECIPROCAL:
            return np.multiply(a, 1. / imax, dtype=np.float64)
        return a / imax
    return a.astype(np.float64)
