"""TEST INFRASTRUCTURE — CPU port of the reference's stages 1-2, same library calls.

This is the oracle of record on the GPU box (where /root/reference does not
exist) and the timed CPU baseline (`cpu_baseline.kind == "port"`).  It makes
the *same third-party calls in the same order* as the reference, so it runs at
the reference's speed and inherits the libraries' exact arithmetic:

    stage 1  utils/util_cylinder.py:1769-1802  (= utils/util_plane.py:2459-2492)
        ridge       :1734-1738  -> skimage hessian_matrix / hessian_matrix_eigvals
                                   (restated here: scikit-image is not installed)
        sauvola     :1740-1766
    stage 2  utils/util_cylinder.py:1805-1827  (= utils/util_plane.py:2494-2516)

Unlike the reference functions it also returns every intermediate (g, b, T,
float centroids, contour sums) so kernels can be checked stage by stage.
Pinned against the unmodified reference by tests/test_oracle.py and by the
committed vectors in tests/golden/.

Parity unpinned by the reference's own tests: it has none.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import cv2
import numpy as np
import scipy.ndimage as ndi

SIGMA = 3.0           # util_cylinder.py:1793
SAUVOLA_WINDOW = 15   # util_cylinder.py:1797
SAUVOLA_K = 0.5
SAUVOLA_R = 128
OPEN_LEN = 20         # util_cylinder.py:1810-1811


@dataclass
class Stage1:
    original: np.ndarray
    gray: np.ndarray
    blurred: np.ndarray
    g: np.ndarray          # gaussian-filtered float image
    b: np.ndarray          # smaller Hessian eigenvalue ("minima_ridges")
    T: np.ndarray          # Sauvola threshold
    binary: np.ndarray     # u8 {0,255}


@dataclass
class Stage2:
    hmask: np.ndarray
    vmask: np.ndarray
    joints: np.ndarray
    centroids: list                       # [(cX, cY)] in findContours order
    centroids_f: np.ndarray = field(default=None)   # [N,2] f64 (m10/m00, m01/m00)
    first_pixels: np.ndarray = field(default=None)  # [Ncontours,2] contour[0] of every contour
    n_contours: int = 0


def to_float(img, float_div=False):
    """skimage.img_as_float for the two integer depths the front-end accepts: scikit-image 0.19 multiplies by
    the rounded reciprocal, `np.multiply(image, 1. / imax_in, dtype=float64)` (util/dtype.py, `_convert`);
    float_div=True is the division round 1 assumed (kernels: LGX_OPT_FLOAT_DIV)."""
    if img.dtype == np.uint8 or img.dtype == np.uint16:
        imax = float(np.iinfo(img.dtype).max)
        return img / imax if float_div else np.multiply(img, 1. / imax, dtype=np.float64)
    return img.astype(np.float64)


def ridge_min_eigenvalue(blurred, sigma=SIGMA, mixed_from_cols=True, float_div=False, both=False):
    """detect_ridges(...)  (util_cylinder.py:1734-1738); returns (g, b) with b = minima_ridges, or
    (g, l1, b) with both=True (the reference computes both eigenvalues and drops l1: the timed CPU baseline
    does the same work).

    mixed_from_cols=True (default, scikit-image 0.19.x for order='rc'): Hrc = d(g_c)/dr;
    False (scikit-image >= 0.20): Hrc = d(g_r)/dc.  SURVEY.md §8c; oracle/refshim/skimage/feature.py.
    """
    f = to_float(blurred, float_div)
    g = ndi.gaussian_filter(f, sigma=sigma, mode="constant", cval=0)
    g_r, g_c = np.gradient(g)
    Hrr = np.gradient(g_r, axis=0)
    Hcc = np.gradient(g_c, axis=1)
    Hrc = np.gradient(g_c, axis=0) if mixed_from_cols else np.gradient(g_r, axis=1)
    if both:
        l1 = (Hrr + Hcc) / 2 + np.sqrt(4 * Hrc ** 2 + (Hrr - Hcc) ** 2) / 2
        l2 = (Hrr + Hcc) / 2 - np.sqrt(4 * Hrc ** 2 + (Hrr - Hcc) ** 2) / 2
        return g, l1, l2
    root = np.sqrt(4 * Hrc ** 2 + (Hrr - Hcc) ** 2)
    b = (Hrr + Hcc) / 2 - root / 2
    return g, b


def sauvola_threshold(b, window=SAUVOLA_WINDOW, k=SAUVOLA_K, R=SAUVOLA_R):
    """sauvola_threshold_fast (util_cylinder.py:1740-1766)."""
    b = b.astype(np.float64)
    ks = (window, window)
    mean = cv2.boxFilter(b, ddepth=-1, ksize=ks, borderType=cv2.BORDER_REPLICATE)
    mean_sq = cv2.boxFilter(b * b, ddepth=-1, ksize=ks, borderType=cv2.BORDER_REPLICATE)
    var = mean_sq - mean * mean
    var[var < 0] = 0
    std = np.sqrt(var)
    return mean * (1 + k * ((std / R) - 1))


def stage1(img, mixed_from_cols=True, float_div=False, as_reference=False) -> Stage1:
    """load_and_preprocess_image (util_cylinder.py:1769-1802)."""
    if img.ndim == 2:
        original = cv2.cvtColor(img, cv2.COLOR_GRAY2BGR)
    elif img.ndim == 3:
        original = img.copy()
    else:
        raise ValueError(f"Unexpected input dimensions: {img.ndim}")
    gray = cv2.cvtColor(original, cv2.COLOR_BGR2GRAY)
    blurred = cv2.GaussianBlur(gray, (5, 5), 0)
    if as_reference:     # every array pass the reference makes, incl. the eigenvalue it throws away (timed baseline)
        g, _l1, b = ridge_min_eigenvalue(blurred, SIGMA, mixed_from_cols, float_div, both=True)
    else:
        g, b = ridge_min_eigenvalue(blurred, SIGMA, mixed_from_cols, float_div)
    T = sauvola_threshold(b)
    binary = (255 - (b > T).astype(np.uint8) * 255).astype(np.uint8)
    return Stage1(original, gray, blurred, g, b, T, binary)


def contours(mask):
    """The contour part of extract_joints (util_cylinder.py:1817-1825) on any u8 mask: findContours(RETR_EXTERNAL,
    CHAIN_APPROX_SIMPLE), moments, int() centroids in contour order.  Returns (list of (int, int), f64 [n, 2],
    first contour points i32 [m, 2], number of contours)."""
    found, _ = cv2.findContours(mask, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
    cents, cents_f, firsts = [], [], []
    for cnt in found:
        firsts.append(cnt[0, 0])
        M = cv2.moments(cnt)
        if M["m00"] != 0:
            fx = M["m10"] / M["m00"]
            fy = M["m01"] / M["m00"]
            cents.append((int(fx), int(fy)))
            cents_f.append((fx, fy))
    return (cents, np.array(cents_f, dtype=np.float64).reshape(-1, 2),
            np.array(firsts, dtype=np.int32).reshape(-1, 2), len(found))


def stage2(binary) -> Stage2:
    """extract_joints (util_cylinder.py:1805-1827)."""
    hk = cv2.getStructuringElement(cv2.MORPH_RECT, (OPEN_LEN, 1))
    vk = cv2.getStructuringElement(cv2.MORPH_RECT, (1, OPEN_LEN))
    hmask = cv2.morphologyEx(binary, cv2.MORPH_OPEN, hk)
    vmask = cv2.morphologyEx(binary, cv2.MORPH_OPEN, vk)
    joints = cv2.bitwise_and(hmask, vmask)
    cents, cents_f, firsts, ncontours = contours(joints)
    return Stage2(hmask, vmask, joints, cents, cents_f, firsts, ncontours)


def frontend(img, mixed_from_cols=True, float_div=False, as_reference=False):
    s1 = stage1(img, mixed_from_cols, float_div, as_reference)
    return s1, stage2(s1.binary)


# ---- input side --------------------------------------------------------------------------------
def undistort_image(image, camera_params):
    """utils/iotool.py:22-39: cv2.undistort with the camera matrix np.array(IntrinsicMatrix) and the coefficient vector
    np.hstack((RadialDistortion, TangentialDistortion)) — radial first, as the reference forms it.  Oracle of record for
    lgx_undistort on the GPU box."""
    K = np.array(camera_params["IntrinsicMatrix"])
    dist = np.hstack((camera_params["RadialDistortion"], camera_params["TangentialDistortion"]))
    return cv2.undistort(image, K, dist)
