"""skimage.feature.hessian_matrix / hessian_matrix_eigvals, restated (0.19.3
semantics: gaussian_filter(mode, cval, truncate=4) -> np.gradient -> second
np.gradient per axis pair; closed-form 2x2 eigenvalues, descending)."""
from itertools import combinations_with_replacement

import numpy as np
import scipy.ndimage as ndi

from . import img_as_float

# Which mixed derivative 0.19.3 forms for order='rc' cannot be checked offline
# (SURVEY.md §8c).  The published 0.19 source reads `if order == 'rc': axes = reversed(axes)`
# (0.20 changed the test to 'xy'), which makes the element list [Hcc, d(g_c)/dr, Hrr]: the
# eigenvalue formula is symmetric in M00 <-> M11 bit for bit, so the only observable effect is
# the mixed term d(g_c)/dr (True, the default since round 2; kernels: LGX_OPT_MIXED_FROM_COLS = 1).
# False: d(g_r)/dc (scikit-image >= 0.20).  The two differ by an ulp of Hrc.
REVERSED_AXES_FOR_RC = True


def hessian_matrix(image, sigma=1, mode='constant', cval=0, order='rc'):
    image = img_as_float(image)
    g = ndi.gaussian_filter(image, sigma=sigma, mode=mode, cval=cval)
    gradients = np.gradient(g)
    axes = range(image.ndim)
    if (order == 'xy') != REVERSED_AXES_FOR_RC:
        axes = reversed(axes)
    return [np.gradient(gradients[a0], axis=a1)
            for a0, a1 in combinations_with_replacement(axes, 2)]


def hessian_matrix_eigvals(H_elems):
    M00, M01, M11 = H_elems
    l1 = (M00 + M11) / 2 + np.sqrt(4 * M01 ** 2 + (M00 - M11) ** 2) / 2
    l2 = (M00 + M11) / 2 - np.sqrt(4 * M01 ** 2 + (M00 - M11) ** 2) / 2
    return np.stack([l1, l2])
