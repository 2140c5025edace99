"""TEST INFRASTRUCTURE — lane-level model of the column role of csrc/lgx_fused.cu.

The fused ridge + Sauvola kernel replaces the K1 -> f64 planes -> K2 hand-over by a role ("EC") in which
lane = image row, lanes are skewed by one column each (lane l works on column x = t - l at slot t), the
cv2.boxFilter row chain stays in the lane's registers and the column chain (cv2 ColumnSum, serial from the top
row) travels down the lanes by shuffle, one lane per slot.  This file is that dataflow in NumPy, one warp
instruction = one vector operation over 32 lanes, with the same rings, hand-over blocks and index arithmetic
as the kernel, so that the schedule can be checked bit for bit against oracle/restate.py on the CPU
(tests/test_fused_model.py).  It models, per band of 124 rows and per warp w (rows 32w-4 .. 32w+27 of the band):

    Bs ring     [16 columns][32 lanes]  b of the last 16 columns of each lane (b(x-15) for the row chain, b(y-7, x-8)
                                        of lane l-7 for the compare)
    Rs ring     [16 columns][32 lanes]  row sums (b, b*b) of the last 16 columns (lane l reads lane l-14: rs(y-14))
    hand-over   rs rows of lanes 18..31, b rows of lanes 25..31, the running column sums of lane 31, per column:
                128-column rings between the warps of a CTA, full-width arrays between bands (items)
    bits        each lane shifts its compare results into a 32-bit word and stores it when the word is complete

Reference arithmetic: /root/reference/utils/util_cylinder.py:1755-1765, 1798-1800 (cv2.boxFilter x2, Sauvola
formula, compare), as restated in oracle/restate.py (row_sums15 / col_sums15 / sauvola_T / binarize).
"""
from __future__ import annotations

import numpy as np

BAND = 124          # rows of b per band (CTA work item)
LAG = 44            # slots warp w runs behind warp w-1 (>= 32: the column chain crosses 32 lanes; + one batch + slack)
RING = 128          # columns of the intra-CTA hand-over rings
SCALE = 1.0 / 225


class HandOver:
    """rs rows of the last 14 lanes, b rows of the last 7 lanes and the column sums below the last lane, per column."""

    def __init__(self, ncols, mask):
        self.mask = mask
        self.rs = np.full((14, ncols, 2), np.nan)
        self.b = np.full((7, ncols), np.nan)
        self.sum = np.full((ncols, 2), np.nan)

    def col(self, p):
        return (p + 32) & self.mask


def fused_column_model(b_plane, H, W):
    """b_plane: [H, W] f64 (oracle/restate.py min_eigenvalue).  Returns (binary u8 [H, W], T f64 [H, W])."""
    nbands = (H + 7 + BAND - 1) // BAND
    x_first, x_last = -16, W + 7                       # a lane's sweep: columns x_first .. x_last
    sweep = x_last - x_first + 1
    wfull = W + 64
    T = np.full((H, W), np.nan)
    words = np.zeros((H, (W + 31) // 32), dtype=np.uint32)
    lanes = np.arange(32)
    prev_item = None                                   # hand-over of the band above (full width)
    for band in range(nbands):
        y0 = band * BAND
        item_out = HandOver(wfull, 0xFFFFFFFF)
        rings = [None] + [HandOver(RING, RING - 1) for _ in range(3)]     # rings[w]: produced by w-1, consumed by w
        st = []
        for w in range(4):
            local = 32 * w - 4 + lanes
            y = y0 + local
            valid = (local >= 0) & (local < BAND) & (y < H + 7)
            st.append(dict(
                y=y, valid=valid, real=valid & (y < H), virt=valid & (y >= H),
                first=int(np.argmax(valid)) if valid.any() else 32,
                Bs=np.full((16, 32), np.nan), Rs=np.full((16, 32, 2), np.nan),
                chain=np.zeros((32, 2)), b0=np.zeros(32), blast=np.zeros(32),
                sum_out=np.zeros((32, 2)), new_prev=np.zeros((32, 2)), word=np.zeros(32, dtype=np.uint32)))
        for tau in range(sweep + 31 + 3 * LAG + 1):
            for w in range(4):
                s = st[w]
                if not s["valid"].any():
                    continue
                t = tau - LAG * w
                x = x_first + t - lanes                 # column of each lane at this slot
                act = s["valid"] & (x >= x_first) & (x <= x_last)
                if not act.any():
                    continue
                hin = (prev_item if w == 0 else rings[w])
                hout = (item_out if w == 3 else rings[w + 1])
                y, first = s["y"], s["first"]
                p = x - 8                                # column of the row sums / of the output
                # ---- E + row chain (lane = row, own registers): b at column x, rs_new = s(x - 8) -------------
                in_img = act & s["real"] & (x >= 0) & (x < W)
                bv = np.zeros(32)
                bv[in_img] = b_plane[y[in_img], x[in_img]]
                s["blast"] = np.where(in_img, bv, s["blast"])
                s["b0"] = np.where(in_img & (x == 0), bv, s["b0"])
                right = act & s["real"] & (x >= W)
                bv = np.where(right, s["blast"], bv)                       # b(min(c + 7, W - 1))
                # b(x - 15) from the lane's own ring (b(0) left of the image)
                bold = s["Bs"][(x - 15) & 15, lanes]
                bold = np.where(x - 15 <= 0, s["b0"], bold)
                rs_new = s["chain"].copy()                                 # s(x - 8)
                c = x - 7
                upd = act & s["real"] & (c > 0) & (c < W)
                s["chain"][upd, 0] = s["chain"][upd, 0] + (bv[upd] - bold[upd])
                s["chain"][upd, 1] = s["chain"][upd, 1] + (bv[upd] * bv[upd] - bold[upd] * bold[upd])
                init = act & s["real"] & (c == 0)
                for l in np.flatnonzero(init):                             # cv2 RowSum start: 15 replicated-border terms
                    sb = sq = 0.0
                    for k in range(15):
                        bi = k - 7 if k > 7 else 0
                        v = s["b0"][l] if bi == 0 else (bv[l] if bi == 7 else s["Bs"][bi & 15, l])
                        sb = sb + v
                        sq = sq + v * v
                    s["chain"][l] = (sb, sq)
                # ---- loads of this slot that come from rings (all written at least 7 slots ago) ---------------------
                pv = act & (p >= 0) & (p < W)                              # the column role is active on this lane
                i14 = lanes - first                                        # hand-over row index of rs(y - 14)
                top = (band == 0 and w == 0)
                old = np.zeros((32, 2))
                bcmp = np.zeros(32)
                for l in np.flatnonzero(pv):
                    yl = y[l]
                    if top and yl < 14:
                        if yl >= 7:
                            old[l] = s["Rs"][p[l] & 15, first]             # rs(0, p): the first lane of the image
                    elif i14[l] < 14:
                        old[l] = hin.rs[i14[l], hin.col(p[l])]
                    else:
                        old[l] = s["Rs"][p[l] & 15, l - 14]
                    if yl >= 7:
                        bcmp[l] = hin.b[i14[l], hin.col(p[l])] if i14[l] < 7 else s["Bs"][p[l] & 15, l - 7]
                # ---- new value of the column chain: this lane's row sum, or (virtual rows) the row above -------------
                new = rs_new.copy()
                up = np.roll(s["new_prev"], 1, axis=0)                     # shfl_up of the previous slot's `new`
                for l in np.flatnonzero(pv & s["virt"]):
                    new[l] = up[l] if l > first else hin.rs[13, hin.col(p[l])]
                # ---- column chain: SUM from the lane above (previous slot) or from the hand-over ---------------
                sum_in = np.roll(s["sum_out"], 1, axis=0)
                for l in np.flatnonzero(pv):
                    yl = y[l]
                    if l == first:
                        sum_in[l] = (0.0, 0.0) if top else hin.sum[hin.col(p[l])]
                    if yl == 0:                                            # ColumnSum start: row 0 counted 8 times
                        acc = np.zeros(2)
                        for _ in range(8):
                            acc = acc + new[l]
                        s["sum_out"][l] = acc
                        continue
                    s0 = sum_in[l] + new[l]
                    if yl < 7:
                        s["sum_out"][l] = s0
                        continue
                    s["sum_out"][l] = s0 - old[l]
                    m, msq = s0[0] * SCALE, s0[1] * SCALE
                    var = msq - m * m
                    if var < 0:
                        var = 0.0
                    thr = m * (1 + 0.5 * ((np.sqrt(var) / 128) - 1))
                    # the kernel forms the bracket as fma(sd, 2^-8, -0.5) + 1: the same value (power-of-two scalings commute
                    # with rounding), asserted here for every pixel the model visits
                    assert thr == m * (1.0 + (np.sqrt(var) / 256 - 0.5))
                    yo = yl - 7
                    T[yo, p[l]] = thr
                    white = not (bcmp[l] > thr)
                    s["word"][l] = (s["word"][l] >> np.uint32(1)) | (np.uint32(0x80000000) if white else np.uint32(0))
                    if (p[l] & 31) == 31 or p[l] == W - 1:
                        nb = (p[l] & 31) + 1
                        words[yo, p[l] >> 5] = s["word"][l] >> np.uint32(32 - nb)
                        s["word"][l] = 0
                s["new_prev"] = np.where(pv[:, None], new, s["new_prev"])
                # ---- stores of this slot: rings and hand-over ---------------------------------------------------
                for l in np.flatnonzero(in_img):
                    s["Bs"][x[l] & 15, l] = bv[l]
                for l in np.flatnonzero(pv):
                    s["Rs"][p[l] & 15, l] = new[l]
                    if l >= 18:
                        hout.rs[l - 18, hout.col(p[l])] = new[l]
                    if l == 31:
                        hout.sum[hout.col(p[l])] = s["sum_out"][l]
                for l in np.flatnonzero(in_img & (lanes >= 25)):
                    hout.b[l - 25, hout.col(x[l])] = bv[l]
        prev_item = item_out
    binary = np.where(np.unpackbits(words.view(np.uint8), axis=1, bitorder="little")[:, :W] > 0, 255, 0).astype(np.uint8)
    return binary, T
