"""TEST INFRASTRUCTURE — CPU model of the strip-local first pass of the contour stage (csrc/lgx_joints_local.cu).

The kernels label the joints bit plane strip by strip in shared memory: word-runs (maximal runs of set bits inside
one 32-bit word) are the union-find elements, numbered in raster order inside a strip, linked to the run left of
the word boundary and to the runs they touch in the row above (8-connectivity); every run adds the sums of the 2x2
quads it owns to its component's record; components with a run that touches a set pixel of the neighbouring strip
are merged through a global union-find over their first pixels (jl_border, jl_merge) and the live records are
copied in raster order (jl_compact).  This file is that dataflow in plain Python, bit masks as Python ints, with
the kernel's own window / ownership formulas, so that the algorithm can be checked on the CPU against an
independent labelling (scipy.ndimage.label + oracle/restate.py's per-quad sums): tests/test_strip_model.py.

What it returns is the state the first pass leaves behind (before the hole logic): for every 8-connected component
of the mask, in ascending order of its first raster pixel: (first pixel, a00, a10, a01, e4) where a00, a10, a01 are
the Green sums of the component's outline INCLUDING its holes' outlines (cv2.moments of the outer contour equals
them only for hole-free components) and e4 = 4 * Euler number (4 = no hole).

Reference: /root/reference/utils/util_cylinder.py:1817-1825 (findContours + moments), SURVEY.md App. A.13.
"""
from __future__ import annotations

import numpy as np

M34 = (1 << 34) - 1


def _pack(mask):
    """[H, W] bool -> list of rows, each a list of 32-bit words (bit i of word w = pixel 32 w + i)"""
    H, W = mask.shape
    WW = (W + 31) // 32
    pad = np.zeros((H, WW * 32), bool)
    pad[:, :W] = mask
    words = np.packbits(pad, axis=1, bitorder="little").view(np.uint32)
    return [[int(v) for v in row] for row in words], WW


def _runs(word):
    """(start bit, end bit) of the runs of ones of a 32-bit word, ascending"""
    out, s = [], None
    for i in range(33):
        on = i < 32 and (word >> i) & 1
        if on and s is None:
            s = i
        if not on and s is not None:
            out.append((s, i - 1))
            s = None
    return out


def _window34(row, w, WW):
    """bit i <-> pixel 32 w - 1 + i of the row (lgx_joints.cuh window34)"""
    c = row[w]
    p = (row[w - 1] >> 31) if w > 0 else 0
    n = (row[w + 1] & 1) if w + 1 < WW else 0
    return p | (c << 1) | (n << 33)


def _sum_bits(m):
    s = 0
    while m:
        low = m & -m
        s += low.bit_length() - 1
        m ^= low
    return s


def _popc(m):
    return bin(m).count("1")


class _UF:
    def __init__(self):
        self.p = {}

    def add(self, a):
        self.p.setdefault(a, a)

    def find(self, a):
        while self.p[a] != a:
            a = self.p[a]
        return a

    def union(self, a, b):
        a, b = self.find(a), self.find(b)
        if a != b:
            self.p[max(a, b)] = min(a, b)       # the smaller id (earlier in raster order) becomes the root: atomicMin


def first_pass(mask, rows_per_strip):
    """mask: [H, W] bool / u8.  Returns an int64 array [n, 5]: first pixel, a00, a10, a01, e4 per component, ascending."""
    mask = np.asarray(mask) > 0
    H, W = mask.shape
    rows, WW = _pack(mask)
    zero = [0] * WW
    R = rows_per_strip
    records = []                       # per strip: list of [first pixel, a00, a10, a01, e4, border flag]
    gl = _UF()                         # global union-find over first pixels of border components
    boundary_run = {}                  # pixel of a boundary run start -> first pixel of its component (the sparse map L)
    for y0 in range(0, H, R):
        nrows = min(R, H - y0)
        sw = lambda ly: rows[y0 + ly] if 0 <= y0 + ly < H else zero          # incl. the halo rows
        # ---- runs in raster order
        run_of = {}                    # (ly, w, start) -> id
        runs = []
        for ly in range(nrows):
            for w in range(WW):
                for (s, e) in _runs(sw(ly)[w]):
                    run_of[(ly, w, s)] = len(runs)
                    runs.append((ly, w, s, e))
        uf = _UF()
        for r in range(len(runs)):
            uf.add(r)
        start_of = lambda ly, w, bit: max(s for (s, e) in _runs(sw(ly)[w]) if s <= bit)
        # ---- union: left word, three words above (inside the strip)
        for r, (ly, w, s, e) in enumerate(runs):
            if s == 0 and w > 0 and (sw(ly)[w - 1] >> 31):
                uf.union(r, run_of[(ly, w - 1, start_of(ly, w - 1, 31))])
            if ly == 0:
                continue
            U = _window34(sw(ly - 1), w, WW)
            mm = U & ((1 << (e + 3)) - 1) & ~((1 << s) - 1)
            i = 0
            while mm >> i:
                if (mm >> i) & 1:
                    if i == 0:
                        wu, bb = w - 1, 31
                    elif i <= 32:
                        wu, bb = w, i - 1
                    else:
                        wu, bb = w + 1, 0
                    uf.union(r, run_of[(ly - 1, wu, start_of(ly - 1, wu, bb))])
                i += 1
        # ---- roots ranked in raster order (run ids are raster ordered, the root is the smallest id of its component)
        root_rank = {}
        recs = []
        for r, (ly, w, s, e) in enumerate(runs):
            if uf.find(r) == r:
                root_rank[r] = len(recs)
                recs.append([(y0 + ly) * W + w * 32 + s, 0, 0, 0, 0, 0])
        # ---- per-run quad sums (ownership rules of jl_sums_word / jl_local), contact with the neighbouring strips
        for r, (ly, w, s, e) in enumerate(runs):
            rec = recs[root_rank[uf.find(r)]]
            y = y0 + ly
            A, Bn, Up = _window34(sw(ly), w, WW), _window34(sw(ly + 1), w, WW), _window34(sw(ly - 1), w, WW)
            own = ((1 << (e + 2)) - 1) & ~((1 << (s + 1)) - 1)
            if not (A >> s) & 1:
                own |= 1 << s
            ownb = (1 << s) | (1 << (e + 1))
            tl, tr, bl, br = A, A >> 1, Bn, Bn >> 1
            inv = lambda v: ~v & M34
            q4 = tl & tr & bl & br & own
            q3 = ((tl & tr & (bl ^ br)) | (bl & br & (tl ^ tr))) & own
            k1 = ((tl ^ tr) & inv(bl) & inv(br)) | ((bl ^ br) & inv(tl) & inv(tr))
            kd = (tl & br & inv(tr) & inv(bl)) | (tr & bl & inv(tl) & inv(br))
            kb = (A ^ (A >> 1)) & inv(Up) & inv(Up >> 1)
            xbase = w * 32 - 1
            n4, n3 = _popc(q4), _popc(q3)
            rec[4] += _popc(k1 & own) + _popc(kb & ownb) - n3 - 2 * _popc(kd & own)
            sx4 = n4 * xbase + _sum_bits(q4)
            sx3 = n3 * xbase + _sum_bits(q3)
            rec[1] += 2 * n4 + n3
            rec[2] += 6 * sx4 + 3 * n4 + 3 * sx3 + _popc(q3 & tr) + _popc(q3 & br)
            rec[3] += n4 * (6 * y + 3) + 3 * y * n3 + _popc(q3 & bl) + _popc(q3 & br)
            span = ((1 << (e + 3)) - 1) & ~((1 << s) - 1)
            if (ly == 0 and (Up & span)) or (ly == nrows - 1 and (Bn & span)):
                rec[5] = 1
                boundary_run[y * W + w * 32 + s] = rec[0]
                gl.add(rec[0])
        records.append(recs)
    # ---- links across strip boundaries (jl_border)
    for y in range(R, H, R):
        for w in range(WW):
            for (s, e) in _runs(rows[y][w]):
                U = _window34(rows[y - 1], w, WW)
                mm = U & ((1 << (e + 3)) - 1) & ~((1 << s) - 1)
                i = 0
                while mm >> i:
                    if (mm >> i) & 1:
                        if i == 0:
                            wu, bb = w - 1, 31
                        elif i <= 32:
                            wu, bb = w, i - 1
                        else:
                            wu, bb = w + 1, 0
                        st = max(s2 for (s2, e2) in _runs(rows[y - 1][wu]) if s2 <= bb)
                        gl.union(boundary_run[y * W + w * 32 + s], boundary_run[(y - 1) * W + wu * 32 + st])
                    i += 1
    # ---- merge (jl_merge) and compact (jl_compact)
    by_pixel = {rec[0]: rec for recs in records for rec in recs}
    dead = set()
    for recs in records:
        for rec in recs:
            if rec[5]:
                root = gl.find(rec[0])
                if root != rec[0]:
                    dst = by_pixel[root]
                    for k in (1, 2, 3, 4):
                        dst[k] += rec[k]
                    dead.add(rec[0])
    out = [rec[:5] for recs in records for rec in recs if rec[0] not in dead]
    return np.array(out, dtype=np.int64).reshape(-1, 5)


def reference_first_pass(mask):
    """The same state from an independent labelling: scipy.ndimage.label (8-connectivity) and per-quad sums over the padded
    mask (the arithmetic of oracle/restate.py contour_sums, without its hole filling), Euler number from the quad counts."""
    import scipy.ndimage as ndi
    m = np.asarray(mask) > 0
    H, W = m.shape
    lab, n = ndi.label(m, structure=np.ones((3, 3), dtype=int))
    if n == 0:
        return np.zeros((0, 5), np.int64)
    flat = lab.ravel()
    idx = np.flatnonzero(flat)
    first = np.full(n + 1, H * W, dtype=np.int64)
    np.minimum.at(first, flat[idx], idx)
    P = np.pad(m, 1).astype(np.int64)
    L = np.pad(lab, 1)
    tl, tr, bl, br = P[:-1, :-1], P[:-1, 1:], P[1:, :-1], P[1:, 1:]
    k = tl + tr + bl + br
    qy, qx = np.mgrid[-1:H, -1:W]
    qlab = np.maximum(np.maximum(L[:-1, :-1], L[:-1, 1:]), np.maximum(L[1:, :-1], L[1:, 1:]))
    a00, a10, a01, e4 = (np.zeros(n + 1, dtype=np.int64) for _ in range(4))
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
    # 4 * Euler number (8-connectivity) = Q1 - Q3 - 2 QD (Gray's bit-quad formula)
    diag = (k == 2) & (tl == br)
    np.add.at(e4, qlab[k == 1], 1)
    np.add.at(e4, qlab[three], -1)
    np.add.at(e4, qlab[diag], -2)
    order = np.argsort(first[1:], kind="stable") + 1
    return np.stack([first[order], a00[order], a10[order], a01[order], e4[order]], axis=1)
