"""TEST INFRASTRUCTURE — operation-order restatement of stages 1-2 (the kernel spec).

Where oracle/ref_port.py calls OpenCV / SciPy / NumPy like the reference does,
this file restates what those library routines compute *in their operation
order*, with plain NumPy element-wise arithmetic (every f64 operation rounded
individually, no FMA), so that each CUDA kernel has a bit-exact CPU twin:

    step                        library routine restated              reference call site
    gray (BGR input)            cv2.cvtColor BGR2GRAY                 utils/util_cylinder.py:1789
    blur5                       cv2.GaussianBlur((5,5),0) u8/u16      :1790
    to_float                    skimage.img_as_float                  :1736
    gauss25                     scipy.ndimage.gaussian_filter s=3     :1736 (correlate1d, axis 0 then 1)
    hessian / min eigenvalue    np.gradient x4 + skimage eigvals      :1736-1737
    box15                       cv2.boxFilter f64 RowSum/ColumnSum    :1755-1757
    sauvola + compare           NumPy expression                      :1760-1765, :1798-1800
    open_h / open_v             cv2.morphologyEx OPEN 20x1 / 1x20     :1813-1814
    contour_sums                cv2.findContours(EXTERNAL) + moments  :1817-1825

Recipes: SURVEY.md App. A (each verified there against cv2 4.13 / scipy 1.18 /
numpy 2.3; re-verified by tests/test_oracle.py against ref_port.py).
scipy.ndimage.label / binary_fill_holes are used only as a CPU connected-
component labeller in `contour_sums` (integer work, no arithmetic order).
"""
from __future__ import annotations

import numpy as np
import scipy.ndimage as ndi

BLUR_K = np.array([1, 4, 6, 4, 1], dtype=np.int64)


def gray_from_bgr(img):
    """cv2 BGR2GRAY fixed point (15-bit coefficients, cv2 >= 4.x)."""
    b = img[..., 0].astype(np.int64)
    g = img[..., 1].astype(np.int64)
    r = img[..., 2].astype(np.int64)
    return ((3735 * b + 19235 * g + 9798 * r + 16384) >> 15).astype(img.dtype)


def blur5(gray):
    """Separable [1 4 6 4 1]/16, exact integer, BORDER_REFLECT_101, rounding (acc+128)>>8."""
    p = np.pad(gray.astype(np.int64), 2, mode="reflect")
    H, W = gray.shape
    hor = sum(BLUR_K[j] * p[:, j:j + W] for j in range(5))
    acc = sum(BLUR_K[i] * hor[i:i + H, :] for i in range(5))
    return ((acc + 128) >> 8).astype(gray.dtype)


def float_lut(dtype, float_div=False):
    """img_as_float as a table: v * (1/255) (u8) or v * (1/65535) (u16), one f64 multiplication by the rounded
    reciprocal (scikit-image 0.19 `_convert`); float_div=True: the f64 division v / imax."""
    if dtype == np.uint8 or dtype == np.uint16:
        n = int(np.iinfo(dtype).max)
        v = np.arange(n + 1, dtype=np.float64)
        return v / float(n) if float_div else v * (1.0 / float(n))
    raise TypeError(dtype)


def gauss_weights(sigma=3.0, truncate=4.0):
    r = int(truncate * sigma + 0.5)
    x = np.arange(-r, r + 1)
    phi = np.exp(-0.5 / (sigma * sigma) * x ** 2)
    return phi / phi.sum()


def _corr1d_sym(f, w, axis):
    """NI_Correlate1D symmetric branch, zero padded:
    tmp = x[l]*w[c]; for j=-r..-1: tmp += (x[l+j] + x[l-j]) * w[c+j]."""
    r = len(w) // 2
    pad = [(0, 0), (0, 0)]
    pad[axis] = (r, r)
    p = np.pad(f, pad)
    n = f.shape[axis]

    def sl(o):
        s = [slice(None), slice(None)]
        s[axis] = slice(r + o, r + o + n)
        return p[tuple(s)]

    tmp = sl(0) * w[r]
    for j in range(-r, 0):
        tmp = tmp + (sl(j) + sl(-j)) * w[r + j]
    return tmp


def gauss25(f, sigma=3.0):
    w = gauss_weights(sigma)
    return _corr1d_sym(_corr1d_sym(f, w, 0), w, 1)


def grad(f, axis):
    """np.gradient, unit spacing, edge_order=1."""
    f = np.moveaxis(f, axis, 0)
    out = np.empty_like(f)
    out[1:-1] = (f[2:] - f[:-2]) / 2.0
    out[0] = (f[1] - f[0]) / 1.0
    out[-1] = (f[-1] - f[-2]) / 1.0
    return np.moveaxis(out, 0, axis)


def min_eigenvalue(g, mixed_from_cols=True):
    g_r = grad(g, 0)
    g_c = grad(g, 1)
    Hrr = grad(g_r, 0)
    Hcc = grad(g_c, 1)
    Hrc = grad(g_c, 0) if mixed_from_cols else grad(g_r, 1)
    s = Hrr + Hcc
    d = Hrr - Hcc
    root = np.sqrt(4 * (Hrc * Hrc) + d * d)
    return s / 2 - root / 2


def row_sums15(b):
    """cv2 RowSum<double,double>, ksize 15, BORDER_REPLICATE pad 7 in x.  Returns [H, W]
    (rows are NOT padded here; the column pass replicates rows by index)."""
    H, W = b.shape
    p = np.pad(b, ((0, 0), (7, 7)), mode="edge")
    rs = np.empty((H, W), dtype=np.float64)
    s = np.zeros(H, dtype=np.float64)
    for i in range(15):
        s = s + p[:, i]
    rs[:, 0] = s
    for i in range(W - 1):
        s = s + (p[:, i + 15] - p[:, i])
        rs[:, i + 1] = s
    return rs


def col_sums15(rs):
    """cv2 ColumnSum<double,double>: SUM over the first 14 padded rows, then
    s0 = SUM + new; out = s0 * (1/225); SUM = s0 - old."""
    H, W = rs.shape
    idx = np.clip(np.arange(-7, H + 7), 0, H - 1)   # padded row -> image row
    SUM = np.zeros(W, dtype=np.float64)
    for pr in range(14):
        SUM = SUM + rs[idx[pr]]
    out = np.empty((H, W), dtype=np.float64)
    scale = 1.0 / 225
    for y in range(H):
        s0 = SUM + rs[idx[y + 14]]
        out[y] = s0 * scale
        SUM = s0 - rs[idx[y]]
    return out


def box15(b):
    return col_sums15(row_sums15(b))


def sauvola_T(b):
    m = box15(b)
    msq = box15(b * b)
    var = msq - m * m
    var = np.where(var < 0, 0.0, var)
    std = np.sqrt(var)
    return m * (1 + 0.5 * ((std / 128) - 1))


def binarize(b, T):
    return np.where(b > T, 0, 255).astype(np.uint8)


def _shifted(m, d, axis, fill):
    """m[x+d] along axis with `fill` outside."""
    out = np.full_like(m, fill)
    n = m.shape[axis]
    src = [slice(None)] * 2
    dst = [slice(None)] * 2
    if d >= 0:
        src[axis] = slice(d, n)
        dst[axis] = slice(0, max(n - d, 0))
    else:
        src[axis] = slice(0, max(n + d, 0))
        dst[axis] = slice(-d, n)
    out[tuple(dst)] = m[tuple(src)]
    return out


def open_line(mask, axis, length=20):
    """cv2.morphologyEx(OPEN) with a `length`-long line: erode then dilate, both over
    offsets [-length/2, length/2-1]; outside = white for erode, black for dilate."""
    m = mask > 0
    lo, hi = -(length // 2), length - length // 2 - 1
    er = np.ones_like(m)
    for d in range(lo, hi + 1):
        er &= _shifted(m, d, axis, True)
    out = np.zeros_like(m)
    for d in range(lo, hi + 1):
        out |= _shifted(er, d, axis, False)
    return out.astype(np.uint8) * 255


def fill_holes(mask):
    """Background components not 4-connected to the outside of the image."""
    return ndi.binary_fill_holes(mask)


def contour_sums(joints):
    """Tracing-free equivalent of findContours(EXTERNAL)+moments (App. A.13).

    Returns (first_pixel_raster_index, a00, a10, a01) per reported contour, in
    findContours order (descending first-pixel raster index)."""
    m = joints > 0
    H, W = m.shape
    F = fill_holes(m)
    lab, n = ndi.label(F, structure=np.ones((3, 3), dtype=int))
    if n == 0:
        z = np.zeros(0, dtype=np.int64)
        return z, z, z, z
    flat = lab.ravel()
    idx = np.flatnonzero(flat)
    first = np.full(n + 1, H * W, dtype=np.int64)
    np.minimum.at(first, flat[idx], idx)
    # quads: TL at (x,y), x in [-1, W-1], y in [-1, H-1]
    P = np.pad(F, 1).astype(np.int64)
    L = np.pad(lab, 1)
    tl, tr, bl, br = P[:-1, :-1], P[:-1, 1:], P[1:, :-1], P[1:, 1:]
    k = tl + tr + bl + br
    qy, qx = np.mgrid[-1:H, -1:W]
    qlab = np.maximum(np.maximum(L[:-1, :-1], L[:-1, 1:]), np.maximum(L[1:, :-1], L[1:, 1:]))
    a00 = np.zeros(n + 1, dtype=np.int64)
    a10 = np.zeros(n + 1, dtype=np.int64)
    a01 = np.zeros(n + 1, dtype=np.int64)
    full = k == 4
    np.add.at(a00, qlab[full], 2)
    np.add.at(a10, qlab[full], 6 * qx[full] + 3)
    np.add.at(a01, qlab[full], 6 * qy[full] + 3)
    three = k == 3
    sx = tl * qx + tr * (qx + 1) + bl * qx + br * (qx + 1)
    sy = tl * qy + tr * qy + bl * (qy + 1) + br * (qy + 1)
    np.add.at(a00, qlab[three], 1)
    np.add.at(a10, qlab[three], sx[three])
    np.add.at(a01, qlab[three], sy[three])
    order = np.argsort(-first[1:], kind="stable") + 1
    return first[order], a00[order], a10[order], a01[order]


def centroids_from_sums(a00, a10, a01):
    """cv2.moments scaling + the reference's int() truncation; a00 == 0 dropped."""
    keep = a00 != 0
    m00 = a00[keep] * 0.5
    m10 = a10[keep] * (1.0 / 6)   # 0.16666666666666666
    m01 = a01[keep] * (1.0 / 6)
    fx = m10 / m00
    fy = m01 / m00
    ints = np.stack([np.trunc(fx), np.trunc(fy)], axis=1).astype(np.int32)
    return ints, np.stack([fx, fy], axis=1)


def frontend(img, mixed_from_cols=True, float_div=False):
    """Full restated stages 1-2.  Returns a dict of every intermediate."""
    gray = gray_from_bgr(img) if img.ndim == 3 else img
    bl = blur5(gray)
    f = float_lut(bl.dtype, float_div)[bl]
    g = gauss25(f)
    b = min_eigenvalue(g, mixed_from_cols)
    rs_b = row_sums15(b)
    rs_b2 = row_sums15(b * b)
    m = col_sums15(rs_b)
    msq = col_sums15(rs_b2)
    var = msq - m * m
    var = np.where(var < 0, 0.0, var)
    T = m * (1 + 0.5 * ((np.sqrt(var) / 128) - 1))
    binary = binarize(b, T)
    hmask = open_line(binary, 1)
    vmask = open_line(binary, 0)
    joints = hmask & vmask
    first, a00, a10, a01 = contour_sums(joints)
    ints, flt = centroids_from_sums(a00, a10, a01)
    return dict(gray=gray, blurred=bl, g=g, b=b, rs_b=rs_b, rs_b2=rs_b2, T=T, binary=binary,
                hmask=hmask, vmask=vmask, joints=joints, first=first, a00=a00, a10=a10,
                a01=a01, centroids=ints, centroids_f=flt)


# ---- input side: what cv2.undistort computes per frame --------------------------------------------
def remap_bilinear_fixed(src, map_xy, map_frac):
    """cv2.remap(src, map_xy (CV_16SC2), map_frac (CV_16UC1), INTER_LINEAR, BORDER_CONSTANT 0) for 8-bit images with
    1 or 3 channels, in exact integer arithmetic: OpenCV's fixed-point bilinear table is
    w = [(32-fy)(32-fx), (32-fy)fx, fy(32-fx), fy*fx] * 32 (INTER_BITS 5, INTER_REMAP_COEF_BITS 15; the products
    are exact, so the table's normalisation fix-up never fires), samples outside the image are 0, and the result is
    (sum + 2^14) >> 15.  Spec of csrc/lgx_undistort.cu; tests/test_undistort.py pins it to cv2.remap / cv2.undistort."""
    H, W = src.shape[:2]
    sx = map_xy[..., 0].astype(np.int64)
    sy = map_xy[..., 1].astype(np.int64)
    fxy = map_frac.astype(np.int64) & 1023
    fx, fy = fxy & 31, fxy >> 5
    s = src.reshape(H, W, -1).astype(np.int64)

    def sample(y, x):
        inside = (x >= 0) & (x < W) & (y >= 0) & (y < H)
        return np.where(inside[..., None], s[np.clip(y, 0, H - 1), np.clip(x, 0, W - 1)], 0)

    w00, w01, w10, w11 = ((32 - fy) * (32 - fx) * 32, (32 - fy) * fx * 32, fy * (32 - fx) * 32, fy * fx * 32)
    acc = (sample(sy, sx) * w00[..., None] + sample(sy, sx + 1) * w01[..., None]
           + sample(sy + 1, sx) * w10[..., None] + sample(sy + 1, sx + 1) * w11[..., None])
    out = ((acc + (1 << 14)) >> 15).astype(np.uint8)
    return out.reshape(src.shape)
