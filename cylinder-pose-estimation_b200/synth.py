"""Synthetic laser-grid frames (host side).

The reference ships no sample images, so every test and bench
frame is rendered.  The recipe is SURVEY.md App. C: a lit square region, `n`
Gaussian-profile vertical lines, `n` horizontal lines bent like a grid seen on
a cylinder, a saturated zero-order spot (needed by the reference's stage 5,
utils/util_cylinder.py:1974-1980) and sensor noise everywhere (flat regions
are knife-edge for the Sauvola compare).

`render_base` is the noise-free float32 image; `add_noise_u8/u16` quantise it
with a per-frame seed.  The batched benches upload a few base images and add
noise on the device (csrc/lgx_synth.cu); the host functions here are used for
golden vectors and small parity cases.
"""
from __future__ import annotations

import numpy as np

CYLINDER_2448 = dict(width=2448, height=2048, n=31, pitch=28.0, lw=1.8, noise=1.0, curv=1e-5)
PLANE_1280 = dict(width=1280, height=1024, n=25, pitch=28.0, lw=1.8, noise=1.0, curv=0.0)
CYLINDER_4096 = dict(width=4096, height=3000, n=61, pitch=28.0, lw=1.8, noise=1.0, curv=6e-6)
DENSE_4096 = dict(width=4096, height=3000, n=200, pitch=14.0, lw=1.5, noise=1.0, curv=6e-6)


def render_base(width, height, n=31, pitch=28.0, lw=1.8, curv=1e-5, shift=0.0,
                spot=True, dtype=np.float64):
    """Noise-free frame, float, before rounding (App. C, everything but the rng draw).

    `shift` moves the whole pattern horizontally (stereo disparity for L/R pairs).
    """
    y, x = np.mgrid[0:height, 0:width].astype(np.float64)
    cx = width / 2 + 3.3 + shift
    cy = height / 2 - 2.7
    half = (n - 1) / 2
    ext = half * pitch + 0.6 * pitch
    lit = ((np.abs(x - cx) < ext) & (np.abs(y - cy) < ext)).astype(np.float64)
    img = 12.0 + 18.0 * lit
    dx = x - cx
    dy = y - cy
    dx2 = dx * dx
    for k in range(n):
        off = (k - half) * pitch
        d = x - (cx + off + 0.01 * dy)
        img += lit * 170.0 * np.exp(-0.5 * (d / lw) ** 2)
        d = y - (cy + off + curv * dx2 * np.sign(off) * abs(off) / ext * 3 + 0.008 * dx)
        img += lit * 170.0 * np.exp(-0.5 * (d / lw) ** 2)
    if spot:
        img += 400.0 * np.exp(-0.5 * (dx2 + dy * dy) / 81.0)
    return img.astype(dtype, copy=False)


def add_noise_u8(base, seed, noise=1.0):
    rng = np.random.default_rng(seed)
    img = base.astype(np.float64) + rng.normal(0.0, noise, base.shape)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def add_noise_u16(base, seed, noise=1.0):
    """16-bit variant: the 8-bit scene scaled by 257 with 16-bit-resolution noise."""
    rng = np.random.default_rng(seed)
    img = (base.astype(np.float64) + rng.normal(0.0, noise, base.shape)) * 257.0
    return np.clip(np.rint(img), 0, 65535).astype(np.uint16)


def render_u8(width, height, seed=0, noise=1.0, **kw):
    return add_noise_u8(render_base(width, height, **kw), seed, noise)


def render_u16(width, height, seed=0, noise=1.0, **kw):
    return add_noise_u16(render_base(width, height, **kw), seed, noise)


def render_multi_cylinder(width=4096, height=3000, seed=0, noise=1.0):
    """Config 5: three cylinders (different curvature / offsets), the nearer one
    occluding its neighbours, dense grids."""
    parts = [
        dict(n=120, pitch=9.0, lw=1.3, curv=9e-6, shift=-1100.0),
        dict(n=120, pitch=9.0, lw=1.3, curv=-7e-6, shift=1100.0),
        dict(n=200, pitch=7.0, lw=1.2, curv=5e-6, shift=0.0),
    ]
    img = np.full((height, width), 12.0)
    for p in parts:
        layer = render_base(width, height, spot=(p["shift"] == 0.0), **p)
        lit = layer > 12.0 + 9.0
        # nearer cylinder (later in the list) hides what is behind it
        img = np.where(lit, layer, img)
    return add_noise_u8(img, seed, noise)


def render_base_torch(width, height, n=31, pitch=28.0, lw=1.8, curv=1e-5, shift=0.0, spot=True, device="cuda"):
    """render_base evaluated with torch on the device (bench plumbing: a 5 MP scene takes ~10 s in NumPy).
    Same formula; last-bit differences in exp() are irrelevant because parity checks copy the rendered
    frames back to the host."""
    import torch
    y = torch.arange(height, dtype=torch.float64, device=device)[:, None]
    x = torch.arange(width, dtype=torch.float64, device=device)[None, :]
    cx = width / 2 + 3.3 + shift
    cy = height / 2 - 2.7
    half = (n - 1) / 2
    ext = half * pitch + 0.6 * pitch
    lit = ((torch.abs(x - cx) < ext) & (torch.abs(y - cy) < ext)).to(torch.float64)
    img = 12.0 + 18.0 * lit
    dx, dy = x - cx, y - cy
    dx2 = dx * dx
    for k in range(n):
        off = (k - half) * pitch
        d = x - (cx + off + 0.01 * dy)
        img = img + lit * 170.0 * torch.exp(-0.5 * (d / lw) ** 2)
        sgn = (off > 0) - (off < 0)
        d = y - (cy + off + curv * dx2 * sgn * abs(off) / ext * 3 + 0.008 * dx)
        img = img + lit * 170.0 * torch.exp(-0.5 * (d / lw) ** 2)
    if spot:
        img = img + 400.0 * torch.exp(-0.5 * (dx2 + dy * dy) / 81.0)
    return img.to(torch.float32)
