"""Shared synthetic inputs for the CPU and GPU test-suites (seeded, small enough for the oracle)."""
import numpy as np

from cylinder_pose_estimation_b200 import synth


def grid_u8(w, h, seed=0, n=9, pitch=14.0, **kw):
    return synth.render_u8(w, h, seed=seed, n=n, pitch=pitch, **kw)


def grid_u16(w, h, seed=0, n=9, pitch=14.0, **kw):
    return synth.render_u16(w, h, seed=seed, n=n, pitch=pitch, **kw)


def noise_u8(w, h, seed=0):
    return np.random.default_rng(seed).integers(0, 256, (h, w), dtype=np.uint8)


def smooth_noise_u8(w, h, seed=0):
    """low-pass noise: gives large, irregular binary regions (many joints, some with holes)."""
    import cv2
    r = np.random.default_rng(seed).normal(0, 1, (h, w))
    r = cv2.GaussianBlur(r, (0, 0), 2.5)
    r = (r - r.min()) / (r.max() - r.min())
    return np.clip(np.rint(r * 255 + np.random.default_rng(seed + 1).normal(0, 1.0, (h, w))), 0, 255).astype(np.uint8)


def random_mask(w, h, fill, seed=0):
    return (np.random.default_rng(seed).random((h, w)) < fill).astype(np.uint8) * 255


def blob_mask(w, h, seed=0, sigma=2.0, thr=0.55):
    """smooth random blobs with holes and nested islands."""
    import cv2
    r = np.random.default_rng(seed).random((h, w))
    r = cv2.GaussianBlur(r, (0, 0), sigma)
    r = (r - r.min()) / (r.max() - r.min())
    return (r > thr).astype(np.uint8) * 255


SMALL_SIZES = [(2, 2), (3, 2), (2, 5), (7, 9), (24, 25), (25, 24), (31, 33), (64, 60), (61, 64), (97, 131),
               (200, 37), (37, 200), (320, 256)]
