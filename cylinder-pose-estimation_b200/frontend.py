"""Host-side mirror of the reference's stage-1/2 interface on top of the lgx C ABI.

Same names, arguments, return values and error behaviour as the reference functions

    load_and_preprocess_image(input_img_array) -> (original_img, gray_img, blurred_img, binary_img)
        /root/reference/utils/util_cylinder.py:1769-1802  (= utils/util_plane.py:2459-2492)
    extract_joints(binary_img) -> (horizontal_mask, vertical_mask, centroids)
        /root/reference/utils/util_cylinder.py:1805-1827  (= utils/util_plane.py:2494-2516)

plus the additive batched / device-resident API (`Frontend.run`, `Frontend.run_host`,
`detect_points_batch`).  PyTorch is only the carrier of device memory and of the CUDA stream; all
arithmetic happens in liblgx.so.  There is no CPU path: without the library or a B200 every call raises.
"""
from __future__ import annotations

import ctypes as C
import weakref
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _lib
from ._lib import LgxError, check

_NP_BITS = {np.dtype(np.uint8): 8, np.dtype(np.uint16): 16}


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise LgxError("no CUDA device: lgx is a B200-only implementation and has no CPU fallback")
    return torch


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _np_ptr(a):
    return C.c_void_p(a.ctypes.data) if a is not None else None


def default_max_centroids(h, w):
    return max(1024, (h * w) // 64)


@dataclass
class FrontendResult:
    """Device-resident outputs of stages 1-2 for a batch (torch tensors on the frontend's device)."""
    binary: Optional[object]        # [B,H,W] u8 {0,255}
    hmask: Optional[object]         # [B,H,W] u8
    vmask: Optional[object]         # [B,H,W] u8
    blurred: Optional[object]       # [B,H,W] u8/u16
    centroids: object               # [B,maxN,2] i32 (cX,cY), reference list order
    centroids_f: Optional[object]   # [B,maxN,2] f64
    counts: object                  # [B] i32
    flags: object                   # [B] u32 as int32 tensor

    def centroid_lists(self):
        """The reference's `centroids` value per frame: list[tuple[int,int]] (one D2H of the used part)."""
        counts = self.counts.cpu().numpy()
        flags = self.flags.cpu().numpy()
        if (flags & (_lib.LGX_FLAG_COMP_OVERFLOW | _lib.LGX_FLAG_CENT_OVERFLOW)).any():
            raise LgxError("centroid / component capacity exceeded; raise max_centroids / max_components")
        nmax = int(counts.max()) if len(counts) else 0
        cent = self.centroids[:, :nmax].cpu().numpy()
        return [_tuples(cent[i, :counts[i]]) for i in range(len(counts))]


class Frontend:
    """One lgx handle (device, capacity) plus torch-side buffer management."""

    def __init__(self, max_w, max_h, chunk_frames=8, max_components=0, device=None):
        torch = _torch()
        self._lib = _lib.load()
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self.max_w, self.max_h, self.chunk_frames = int(max_w), int(max_h), int(chunk_frames)
        h = C.c_void_p()
        check(self._lib.lgx_create(self.device.index, self.max_w, self.max_h, self.chunk_frames,
                                   int(max_components), C.byref(h)), "lgx_create")
        self._h = h
        self._set_caller_gauss_weights()

    def _set_caller_gauss_weights(self):
        # scipy.ndimage._gaussian_kernel1d(3.0, 0, 12) as the caller's own numpy evaluates it
        x = np.arange(-12, 13)
        phi = np.exp(-0.5 / 9.0 * x ** 2)
        w = np.ascontiguousarray(phi / phi.sum(), dtype=np.float64)
        check(self._lib.lgx_set_gauss_weights(self._h, w.ctypes.data_as(C.POINTER(C.c_double))), "lgx_set_gauss_weights")

    def close(self):
        if getattr(self, "_h", None):
            self._lib.lgx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_mixed_from_cols(self, on):
        """Mixed second derivative: True (default) d(g_c)/dr as scikit-image 0.19.x forms it for order='rc',
        False d(g_r)/dc (scikit-image >= 0.20).  An ulp-level choice; SURVEY.md §8c."""
        check(self._lib.lgx_set_option(self._h, _lib.LGX_OPT_MIXED_FROM_COLS, int(bool(on))))

    def set_float_div(self, on):
        """img_as_float: False (default) v * (1/imax) as scikit-image 0.19 computes it, True the division v / imax."""
        check(self._lib.lgx_set_option(self._h, _lib.LGX_OPT_FLOAT_DIV, int(bool(on))))

    def set_ridge_warps(self, n):
        """Tuning knob: instantiation of the ridge kernel (16 = warp-specialised TMA pipeline, 8 / 4 = phase kernel with
        64- / 32-row bands, 0 = chosen by launch size).  Results are identical."""
        check(self._lib.lgx_set_option(self._h, _lib.LGX_OPT_RIDGE_WARPS, int(n)))

    def set_sauvola_variant(self, v):
        """Tuning / cross-check knob: 0 = column kernel with direct loads (default), 2 = TMA ring kernel when the
        planes allow it.  Results are identical."""
        check(self._lib.lgx_set_option(self._h, _lib.LGX_OPT_SAUVOLA, int(v)))

    def set_fused(self, mode):
        """Stage-1 kernel choice: 0 (default) = blur / ridge / sauvola as three kernels, 1 = the fused ridge + sauvola
        kernel when the batch fills every CTA group, 2 = whenever the geometry allows.  Results are identical; the fused
        kernel is the experimental no-f64-planes path (parity-green, slower: DESIGN.md section 6)."""
        check(self._lib.lgx_set_option(self._h, _lib.LGX_OPT_FUSED, int(mode)))

    def set_joints_global(self, on):
        """Cross-check knob: True = first pass of the contour stage as the whole-frame union-find (csrc/lgx_joints.cu),
        False (default) = strip-local labelling in shared memory (csrc/lgx_joints_local.cu).  Results are identical."""
        check(self._lib.lgx_set_option(self._h, _lib.LGX_OPT_JOINTS_GLOBAL, int(bool(on))))

    def last_joints_kernel(self):
        return self._lib.lgx_last_joints_kernel(self._h).decode()

    def last_ridge_kernel(self):
        return self._lib.lgx_last_ridge_kernel(self._h).decode()

    def set_timing(self, on):
        check(self._lib.lgx_set_option(self._h, _lib.LGX_OPT_TIMING, int(bool(on))))

    def stats(self, reset=True):
        """(ms per kernel group [blur5, ridge, sauvola, open_hv, joints], kernel groups timed, kernels launched)"""
        ms = (C.c_double * 5)()
        chunks, launches = C.c_longlong(0), C.c_longlong(0)
        check(self._lib.lgx_get_stats(self._h, ms, C.byref(chunks), C.byref(launches), int(reset)))
        return list(ms), chunks.value, launches.value

    def render_noisy(self, base, batch, sigma=1.0, seed0=0, bits=8):
        """Synthetic batch on the device: base [n_base,H,W] float32 CUDA tensor + per-frame noise."""
        torch = _torch()
        nb, H, W = base.shape
        out = torch.empty((batch, H, W), dtype=torch.uint8 if bits == 8 else torch.uint16, device=base.device)
        check(self._lib.lgx_render_noisy(_ptr(base), nb, batch, H, W, C.c_float(sigma), C.c_uint64(seed0), bits,
                                         _ptr(out), self._stream()), "lgx_render_noisy")
        return out

    # ---- device-resident batch API ------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(_torch().cuda.current_stream(self.device).cuda_stream)

    @staticmethod
    def _bits_of(t):
        import torch
        if t.dtype == torch.uint8:
            return 8
        if t.dtype in (torch.uint16, torch.int16):
            return 16
        raise TypeError(f"frames must be uint8 or uint16, got {t.dtype}")

    def run(self, frames, masks=True, blurred=False, floats=False, max_centroids=None) -> FrontendResult:
        """Stages 1+2 on a CUDA tensor [B,H,W] (u8 / u16, last dim contiguous).  Asynchronous on the
        current stream; outputs stay on the device."""
        torch = _torch()
        if frames.dim() == 2:
            frames = frames[None]
        if frames.dim() != 3 or not frames.is_cuda or frames.stride(2) != 1:
            raise ValueError("frames must be a CUDA tensor [B,H,W] with contiguous rows")
        bits = self._bits_of(frames)
        B, H, W = frames.shape
        es = bits // 8
        dev = frames.device
        n = int(max_centroids or default_max_centroids(H, W))
        u8 = dict(dtype=torch.uint8, device=dev)
        binary = torch.empty((B, H, W), **u8) if masks else None
        hmask = torch.empty((B, H, W), **u8) if masks else None
        vmask = torch.empty((B, H, W), **u8) if masks else None
        blur = torch.empty((B, H, W), dtype=frames.dtype, device=dev) if blurred else None
        cent = torch.empty((B, n, 2), dtype=torch.int32, device=dev)
        centf = torch.empty((B, n, 2), dtype=torch.float64, device=dev) if floats else None
        counts = torch.empty((B,), dtype=torch.int32, device=dev)
        flags = torch.empty((B,), dtype=torch.int32, device=dev)
        if B == 0:
            return FrontendResult(binary, hmask, vmask, blur, cent, centf, counts, flags)
        check(self._lib.lgx_frontend(self._h, _ptr(frames), bits, B, H, W, frames.stride(1) * es,
                                     (frames.stride(0) if B > 1 else H * frames.stride(1)) * es,
                                     _ptr(binary), _ptr(hmask), _ptr(vmask), _ptr(blur), _ptr(cent), _ptr(centf), n,
                                     _ptr(counts), _ptr(flags), self._stream()), "lgx_frontend")
        return FrontendResult(binary, hmask, vmask, blur, cent, centf, counts, flags)

    def extract_joints_device(self, binary, floats=False, max_centroids=None) -> FrontendResult:
        torch = _torch()
        if binary.dim() == 2:
            binary = binary[None]
        binary = binary.contiguous()
        B, H, W = binary.shape
        dev = binary.device
        n = int(max_centroids or default_max_centroids(H, W))
        hmask = torch.empty((B, H, W), dtype=torch.uint8, device=dev)
        vmask = torch.empty((B, H, W), dtype=torch.uint8, device=dev)
        cent = torch.empty((B, n, 2), dtype=torch.int32, device=dev)
        centf = torch.empty((B, n, 2), dtype=torch.float64, device=dev) if floats else None
        counts = torch.empty((B,), dtype=torch.int32, device=dev)
        flags = torch.empty((B,), dtype=torch.int32, device=dev)
        check(self._lib.lgx_extract_joints(self._h, _ptr(binary), B, H, W, _ptr(hmask), _ptr(vmask), _ptr(cent),
                                           _ptr(centf), n, _ptr(counts), _ptr(flags), self._stream()),
              "lgx_extract_joints")
        return FrontendResult(binary, hmask, vmask, None, cent, centf, counts, flags)

    def contour_centroids_device(self, mask, floats=False, max_centroids=None) -> FrontendResult:
        """The contour part of extract_joints alone on device-resident u8 masks [B,H,W] (lgx_contour_centroids)."""
        torch = _torch()
        if mask.dim() == 2:
            mask = mask[None]
        mask = mask.contiguous()
        B, H, W = mask.shape
        dev = mask.device
        n = int(max_centroids or default_max_centroids(H, W))
        cent = torch.empty((B, n, 2), dtype=torch.int32, device=dev)
        centf = torch.empty((B, n, 2), dtype=torch.float64, device=dev) if floats else None
        counts = torch.empty((B,), dtype=torch.int32, device=dev)
        flags = torch.empty((B,), dtype=torch.int32, device=dev)
        check(self._lib.lgx_contour_centroids(self._h, _ptr(mask), B, H, W, _ptr(cent), _ptr(centf), n, _ptr(counts),
                                              _ptr(flags), self._stream()), "lgx_contour_centroids")
        return FrontendResult(None, None, None, None, cent, centf, counts, flags)

    # ---- host-buffer API (what a reference-side caller holds) ----------------------------------
    def host_buffers(self, B, H, W, dtype=np.uint8, masks=True, blurred=False, floats=False, max_centroids=None, packed=False):
        """Page-locked output buffers for run_host(..., buffers=...).  With pinned inputs and outputs the
        host path overlaps copy-in, compute and copy-out; results then live in these buffers until the next
        call that uses them.  packed=True: the three masks as bit planes [B,H,lgx_bits_pitch(W)] uint32."""
        torch = _torch()
        n = int(max_centroids or default_max_centroids(H, W))

        def pinned(shape, dt):
            return torch.empty(shape, dtype=dt).pin_memory().numpy()
        tdt = torch.uint8 if np.dtype(dtype) == np.uint8 else torch.uint16
        mshape, mdt = ((B, H, self._lib.lgx_bits_pitch(W)), torch.int32) if packed else ((B, H, W), torch.uint8)
        mk = (lambda: pinned(mshape, mdt).view(np.uint32)) if packed else (lambda: pinned(mshape, mdt))
        return dict(packed=bool(packed), binary=mk() if masks else None,
                    hmask=mk() if masks else None,
                    vmask=mk() if masks else None,
                    blurred=pinned((B, H, W), tdt) if blurred else None,
                    cent=pinned((B, n, 2), torch.int32),
                    centf=pinned((B, n, 2), torch.float64) if floats else None,
                    counts=pinned((B,), torch.int32), flags=pinned((B,), torch.int32).view(np.uint32), n=n)

    def run_host(self, frames: np.ndarray, masks=True, blurred=False, floats=False, max_centroids=None, buffers=None, packed=False):
        """lgx_frontend_host: NumPy in, NumPy out, synchronous.  Returns a dict with `binary`, `hmask`,
        `vmask`, `blurred` ([B,H,W]) and `centroids` (list of [n_i,2] int32 arrays), `centroids_f`, `flags`.
        `buffers` (from host_buffers) makes the outputs page-locked and reused.  packed=True (or packed buffers): the three
        masks come back as bit planes [B,H,lgx_bits_pitch(W)] uint32 (an eighth of the PCIe bytes; unpack_mask() restores
        the reference's u8 planes)."""
        _torch()
        frames = np.ascontiguousarray(frames)
        if frames.ndim == 2:
            frames = frames[None]
        if frames.ndim != 3 or frames.dtype not in _NP_BITS:
            raise TypeError("frames must be [B,H,W] uint8 or uint16")
        bits = _NP_BITS[frames.dtype]
        B, H, W = frames.shape
        if buffers is not None:
            n = buffers["n"]
            binary, hmask, vmask, blur = buffers["binary"], buffers["hmask"], buffers["vmask"], buffers["blurred"]
            cent, centf, counts, flags = buffers["cent"], buffers["centf"], buffers["counts"], buffers["flags"]
            packed = bool(buffers.get("packed", False))
            _validate_host_buffers(buffers, B, H, W, frames.dtype, self._lib.lgx_bits_pitch(W) if packed else 0)
            floats = centf is not None
        else:
            n = int(max_centroids or default_max_centroids(H, W))
            mshape, mdt = ((B, H, self._lib.lgx_bits_pitch(W)), np.uint32) if packed else ((B, H, W), np.uint8)
            binary = np.empty(mshape, mdt) if masks else None
            hmask = np.empty(mshape, mdt) if masks else None
            vmask = np.empty(mshape, mdt) if masks else None
            blur = np.empty((B, H, W), frames.dtype) if blurred else None
            cent = np.empty((B, n, 2), np.int32)
            centf = np.empty((B, n, 2), np.float64) if floats else None
            counts = np.empty((B,), np.int32)
            flags = np.empty((B,), np.uint32)
        check(self._lib.lgx_set_option(self._h, _lib.LGX_OPT_PACKED_MASKS, int(bool(packed))))
        try:
            check(self._lib.lgx_frontend_host(self._h, _np_ptr(frames), bits, B, H, W, _np_ptr(binary), _np_ptr(hmask),
                                              _np_ptr(vmask), _np_ptr(blur), _np_ptr(cent), _np_ptr(centf), n,
                                              _np_ptr(counts), _np_ptr(flags), self._stream()), "lgx_frontend_host")
        finally:
            if packed:
                check(self._lib.lgx_set_option(self._h, _lib.LGX_OPT_PACKED_MASKS, 0))
        if (flags & (_lib.LGX_FLAG_COMP_OVERFLOW | _lib.LGX_FLAG_CENT_OVERFLOW)).any():
            raise LgxError("centroid / component capacity exceeded; raise max_centroids / max_components")
        return dict(binary=binary, hmask=hmask, vmask=vmask, blurred=blur,
                    centroids=[cent[i, :counts[i]] for i in range(B)],
                    centroids_f=[centf[i, :counts[i]] for i in range(B)] if floats else None,
                    counts=counts, flags=flags)

    def run_one(self, gray: np.ndarray):
        """One gray frame through lgx_frontend_host with page-locked staging buffers that live in this handle (the path the
        reference-named functions take: no pageable DMA, no per-call allocation of the big mirrors).  Returns fresh arrays
        (blurred, binary, hmask, vmask, centroids [n,2] int32) the caller owns.  (Handing the page-locked planes out directly,
        with reuse guarded by reference counts, was measured: no faster, the call is bound by the tuple list and the checksums.)"""
        torch = _torch()
        H, W = gray.shape
        key = (H, W, gray.dtype.str)
        st = getattr(self, "_one", None)
        if st is None or st[0] != key:
            tdt = torch.uint8 if gray.dtype == np.uint8 else torch.uint16
            pin = torch.empty((1, H, W), dtype=tdt).pin_memory().numpy()
            st = (key, pin, self.host_buffers(1, H, W, dtype=gray.dtype, masks=True, blurred=True))
            self._one = st
        _, pin, bufs = st
        np.copyto(pin[0], gray)
        out = self.run_host(pin, buffers=bufs)
        return (out["blurred"][0].copy(), out["binary"][0].copy(), out["hmask"][0].copy(), out["vmask"][0].copy(),
                out["centroids"][0].copy())

    def debug_contours(self, frame_in_chunk=0, capacity=1 << 20):
        """(first_pixel, a00, a10, a01) of every reported contour of a frame of the last chunk."""
        out = np.empty((capacity, 4), np.int64)
        n = C.c_int(0)
        check(self._lib.lgx_debug_contours(self._h, frame_in_chunk, _np_ptr(out), capacity, C.byref(n)))
        return out[:min(n.value, capacity)].copy()


def unpack_mask(bits, width):
    """bit planes [..., H, lgx_bits_pitch(W)] uint32 (LGX_OPT_PACKED_MASKS) -> the reference's u8 planes [..., H, W] {0, 255}"""
    b = np.unpackbits(np.ascontiguousarray(bits).view(np.uint8), axis=-1, bitorder="little")[..., :width]
    return b * np.uint8(255)


def _validate_host_buffers(buf, B, H, W, frame_dtype, packed_words=0):
    """every supplied output buffer is overrun-proof: exact plane shapes and dtypes, C-contiguous, lists large enough"""
    n = int(buf["n"])

    def need(name, shape_ok, dtype):
        a = buf.get(name)
        if a is None:
            return
        if not isinstance(a, np.ndarray) or a.dtype != np.dtype(dtype) or not a.flags["C_CONTIGUOUS"] or not shape_ok(a.shape):
            raise ValueError(f"buffers[{name!r}] does not match the batch: got "
                             f"{getattr(a, 'shape', None)} {getattr(a, 'dtype', None)}, frames are {(B, H, W)} {frame_dtype}")
    plane = lambda sh: tuple(sh) == (B, H, W)
    mplane = (lambda sh: tuple(sh) == (B, H, packed_words)) if packed_words else plane
    for name in ("binary", "hmask", "vmask"):
        need(name, mplane, np.uint32 if packed_words else np.uint8)
    need("blurred", plane, frame_dtype)
    if buf.get("cent") is None or buf.get("counts") is None or buf.get("flags") is None:
        raise ValueError("buffers must hold 'cent', 'counts' and 'flags'")
    need("cent", lambda sh: len(sh) == 3 and sh[0] >= B and sh[1] == n and sh[2] == 2, np.int32)
    need("centf", lambda sh: len(sh) == 3 and sh[0] >= B and sh[1] == n and sh[2] == 2, np.float64)
    need("counts", lambda sh: len(sh) == 1 and sh[0] >= B, np.int32)
    need("flags", lambda sh: len(sh) == 1 and sh[0] >= B, np.uint32)
    if n < 1:
        raise ValueError("buffers['n'] must be >= 1")


# ---- module-level functions with the reference's names ---------------------------------------------

_frontends = {}


def get_frontend(height, width, chunk_frames=1, device=None) -> Frontend:
    """Cached handle big enough for (height, width).  Lives in this module, which MATLAB's
    importlib.reload of the entry-point module (utils/makePyGridPts.m:16) does not touch."""
    torch = _torch()
    dev = torch.cuda.current_device() if device is None else device
    key = (dev, chunk_frames)
    fe = _frontends.get(key)
    if fe is None or fe.max_w < width or fe.max_h < height:
        if fe is not None:
            fe.close()
        fe = Frontend(max(width, fe.max_w if fe else 0), max(height, fe.max_h if fe else 0), chunk_frames, 0, dev)
        _frontends[key] = fe
    return fe


# Stage 2 is computed in the same device pass as stage 1; extract_joints(binary_img) answers from here when the
# caller passes the binary image stage 1 returned, UNCHANGED: the entry is keyed by object identity and verified
# by content (two 64-bit checksums of the bytes), so a caller that edits binary_img in place gets a fresh computation,
# as with the reference (util_cylinder.py:1805-1827 recomputes on every call).  A hit hands the stored arrays over and
# forgets them: nothing the module returns is ever aliased by a later answer.
_stage2_cache = []   # [(weakref(binary), checksums, hmask, vmask, centroids)], newest last


def _checksums(a):
    flat = a.reshape(-1)
    n8 = flat.size & ~7
    w = flat[:n8].view(np.uint64)
    tail = int(flat[n8:].astype(np.uint64).sum()) if n8 < flat.size else 0
    return (int(w.sum(dtype=np.uint64)), int(np.bitwise_xor.reduce(w)) if n8 else 0, tail, a.shape, a.dtype.str)


def _remember(binary, hmask, vmask, cents):
    _stage2_cache.append((weakref.ref(binary), _checksums(binary), hmask, vmask, cents))
    del _stage2_cache[:-2]


def _recall(binary):
    for k in range(len(_stage2_cache) - 1, -1, -1):
        ref, sums, hmask, vmask, cents = _stage2_cache[k]
        if ref() is binary and isinstance(binary, np.ndarray) and binary.flags["C_CONTIGUOUS"] and _checksums(binary) == sums:
            del _stage2_cache[k]
            return hmask, vmask, cents
    return None


# Stage 1 of a batch computed ahead (prime_stage12: the folder CLIs decode and undistort a batch of files, run stages 1-2 for
# all of them in one device pass and then call the unchanged detect_grid per file): load_and_preprocess_image(img) answers
# from here when `img` is one of the primed arrays, unchanged (identity + checksums, consumed on the hit).
_stage1_cache = []   # [(weakref(img), checksums, (original, gray, blurred, binary, hmask, vmask, centroids))]


def prime_stage12(images, chunk_frames=8):
    """images: list of same-shape uint8 / uint16 arrays [H,W] or [H,W,3] (BGR) that the caller is about to pass, one by one
    and unchanged, to load_and_preprocess_image / detect_grid.  Runs BGR2GRAY and stages 1-2 for all of them in one batched
    device pass (masks come back as bit planes) and remembers the results.  Returns the number of primed images."""
    torch = _torch()
    images = [im for im in images if isinstance(im, np.ndarray) and im.dtype in _NP_BITS and im.ndim in (2, 3)]
    if not images or any(im.shape != images[0].shape for im in images) or (images[0].ndim == 3 and images[0].shape[2] != 3):
        return 0
    H, W = images[0].shape[:2]
    stack = np.ascontiguousarray(np.stack(images))
    if stack.ndim == 4:
        lib = _lib.load()
        d = torch.from_numpy(stack).cuda()
        g = torch.empty(stack.shape[:3], dtype=d.dtype, device=d.device)
        check(lib.lgx_bgr2gray(_ptr(d), _NP_BITS[stack.dtype], len(images), H, W, _ptr(g),
                               C.c_void_p(torch.cuda.current_stream().cuda_stream)), "lgx_bgr2gray")
        grays = g.cpu().numpy()
    else:
        grays = stack
    out = get_frontend(H, W, chunk_frames).run_host(grays, masks=True, blurred=True, packed=True)
    del _stage1_cache[:]
    for i, im in enumerate(images):
        original = im.copy() if im.ndim == 3 else _gray2bgr(im)
        planes = [unpack_mask(out[k][i], W) for k in ("binary", "hmask", "vmask")]
        _stage1_cache.append((weakref.ref(im), _checksums(im),
                              (original, grays[i].copy(), out["blurred"][i].copy(), planes[0], planes[1], planes[2],
                               _tuples(out["centroids"][i]))))
    return len(images)


def _recall1(img):
    for k in range(len(_stage1_cache)):
        ref, sums, res = _stage1_cache[k]
        if ref() is img and img.flags["C_CONTIGUOUS"] and _checksums(img) == sums:
            del _stage1_cache[k]
            return res
    return None


def _tuples(arr):
    a = np.asarray(arr)
    return list(zip(a[:, 0].tolist(), a[:, 1].tolist())) if len(a) else []


def load_and_preprocess_image(input_img_array):
    """Reference stage 1 (util_cylinder.py:1769-1802).  Stage 2 is computed in the same device pass and
    remembered, so the extract_joints(binary_img) call that follows costs nothing."""
    arr = np.asarray(input_img_array)
    if arr.ndim not in (2, 3):
        raise ValueError(f"Unexpected input dimensions: {arr.ndim}")
    if arr.dtype not in _NP_BITS:
        raise TypeError(f"lgx front-end accepts uint8 / uint16 images, got {arr.dtype}")
    if _stage1_cache and isinstance(input_img_array, np.ndarray):
        hit = _recall1(input_img_array)
        if hit is not None:
            original, gray, blurred, binary, hmask, vmask, cents = hit
            _remember(binary, hmask, vmask, cents)
            return original, gray, blurred, binary
    arr = np.ascontiguousarray(arr)
    if arr.ndim == 2:
        gray = arr
        original = _gray2bgr(arr)                                # cv2.cvtColor(GRAY2BGR)
    else:
        if arr.shape[2] != 3:
            raise ValueError(f"Unexpected channel count: {arr.shape[2]}")
        original = arr.copy()
        gray = _bgr2gray(original)                               # cv2.cvtColor(BGR2GRAY)
    H, W = gray.shape
    blurred, binary, hmask, vmask, cents = get_frontend(H, W).run_one(gray)
    _remember(binary, hmask, vmask, _tuples(cents))
    return original, gray.copy() if gray is arr else gray, blurred, binary


def _gray2bgr(gray):
    """three identical channels; through cv2 when it is importable (the reference's later stages need it anyway: threaded
    copy, ~10x faster than NumPy's repeat at 5 MP), else NumPy"""
    try:
        import cv2
        return cv2.cvtColor(gray, cv2.COLOR_GRAY2BGR)
    except ImportError:
        out = np.empty(gray.shape + (3,), gray.dtype)
        out[...] = gray[:, :, None]
        return out


def _bgr2gray(bgr):
    torch = _torch()
    lib = _lib.load()
    H, W, _ = bgr.shape
    d = torch.from_numpy(bgr).cuda()
    g = torch.empty((H, W), dtype=d.dtype, device=d.device)
    check(lib.lgx_bgr2gray(_ptr(d), _NP_BITS[bgr.dtype], 1, H, W, _ptr(g),
                           C.c_void_p(torch.cuda.current_stream().cuda_stream)), "lgx_bgr2gray")
    return g.cpu().numpy()


def extract_joints(binary_img):
    """Reference stage 2 (util_cylinder.py:1805-1827)."""
    hit = _recall(binary_img)
    if hit is not None:
        return hit
    torch = _torch()
    b = np.ascontiguousarray(np.asarray(binary_img))
    if b.ndim != 2 or b.dtype != np.uint8:
        raise TypeError("binary_img must be a 2-D uint8 image")
    H, W = b.shape
    res = get_frontend(H, W).extract_joints_device(torch.from_numpy(b).cuda())
    cents = res.centroid_lists()[0]
    return res.hmask[0].cpu().numpy(), res.vmask[0].cpu().numpy(), cents


def detect_points_batch(frames, chunk_frames=8):
    """Additive API: stages 1-2 for a stack of frames [B,H,W] (NumPy).  Returns the per-frame centroid
    lists in the reference's order as [n_i,2] int32 arrays (no masks copied back)."""
    frames = np.asarray(frames)
    if frames.ndim != 3:
        raise ValueError("frames must be [B,H,W]")
    fe = get_frontend(frames.shape[1], frames.shape[2], chunk_frames)
    return fe.run_host(frames, masks=False)["centroids"]


def stage12_batch(frames, chunk_frames=8):
    """Stages 1-2 for a stack of gray frames [B,H,W] in one device pass.  Returns, per frame, exactly what the
    reference's two functions return: (original_img, gray_img, blurred_img, binary_img, horizontal_mask,
    vertical_mask, centroids) — the inputs of the reference's stages 3-6."""
    frames = np.ascontiguousarray(np.asarray(frames))
    if frames.ndim != 3:
        raise ValueError("frames must be [B,H,W] (gray)")
    if frames.dtype not in _NP_BITS:
        raise TypeError(f"lgx front-end accepts uint8 / uint16 images, got {frames.dtype}")
    B, H, W = frames.shape
    # the three masks cross PCIe as bit planes (an eighth of the bytes) and are unpacked next to their consumer
    out = get_frontend(H, W, chunk_frames).run_host(frames, masks=True, blurred=True, packed=True)
    res = []
    for i in range(B):
        gray = frames[i].copy()
        res.append((_gray2bgr(gray), gray, out["blurred"][i], unpack_mask(out["binary"][i], W), unpack_mask(out["hmask"][i], W),
                    unpack_mask(out["vmask"][i], W), _tuples(out["centroids"][i])))
    return res
