"""CPU: the model of the strip-local first pass of the contour stage (oracle/strip_model.py: word-runs numbered per strip,
union with the left word and the three words above, per-run quad sums with the kernel's ownership rules, merge of the
components that cross strip boundaries) against an independent labelling (scipy.ndimage.label + per-quad sums), for strips of
every height, and - for hole-free components - against the restatement that tests/test_oracle.py ties to cv2.findContours."""
import numpy as np
import pytest

from oracle import restate, strip_model


def _mask(kind, w, h, seed):
    rng = np.random.default_rng(seed)
    if kind == "blobs":
        import cv2
        r = cv2.GaussianBlur(rng.random((h, w)), (0, 0), 2.0)
        return r > np.quantile(r, 0.55)
    if kind == "bars":
        m = np.zeros((h, w), bool)
        m[:, ::3] = True
        m[::11, :] = False
        return m
    if kind == "diag":
        yy, xx = np.mgrid[:h, :w]
        return ((yy + xx) % 5 == 0) | ((yy - xx) % 7 == 0)
    return rng.random((h, w)) < float(kind)


@pytest.mark.parametrize("kind,w,h", [("0.08", 70, 41), ("0.3", 97, 37), ("0.5", 66, 45), ("0.62", 40, 50), ("0.9", 65, 20),
                                      ("blobs", 130, 60), ("bars", 75, 40), ("diag", 64, 33), ("0.5", 31, 9), ("0.5", 33, 2)])
@pytest.mark.parametrize("rows", [2, 3, 7, 26, 64])
def test_strip_model_equals_independent_labelling(kind, w, h, rows):
    m = _mask(kind, w, h, seed=w * 100 + h)
    got = strip_model.first_pass(m, rows)
    want = strip_model.reference_first_pass(m)
    assert got.shape == want.shape
    assert np.array_equal(got, want)


def test_strip_model_hole_free_components_are_the_contour_sums():
    """where a component has no hole (Euler number 1) the first-pass sums are what cv2.moments of its outer contour gives
    (restate.contour_sums, pinned to cv2 by test_quad_sums_match_findcontours_moments)"""
    m = _mask("blobs", 160, 90, seed=5)
    got = strip_model.first_pass(m, 13)
    first, a00, a10, a01 = restate.contour_sums(m.astype(np.uint8) * 255)
    by_first = {int(f): (int(a), int(b), int(c)) for f, a, b, c in zip(first, a00, a10, a01)}
    checked = 0
    for f, a, b, c, e4 in got:
        if e4 == 4 and int(f) in by_first:
            assert by_first[int(f)] == (int(a), int(b), int(c))
            checked += 1
    assert checked >= 5
