"""Overlap stitching by a minimum-error boundary cut (oracle; test infra only).  PARITY UNPINNED (oracle/__init__.py).

Sequential restatement of train-gray-patch-batch-overlap.lua:457-694 (duplicated for the train image at :788-1025), in the
reference's patch order and with its arithmetic types: the script never sets a default tensor type, so `torch.Tensor` is a
DoubleTensor there -- the cost tables are float64 over the float32 generator output -- and every pixel written is a copy of a
generated pixel (no blending).

Geometry (:386-388): L = (fineSize - overlap) / (patchSize - overlap) patches per line, patch i (0-based) sits at
(x, y) = (i // L, i % L), top-left pixel (x * overlap, y * overlap); the strip it shares with a neighbour is `overlap` wide and
is read from the neighbour at offset patchSize - overlap (consistent for patchSize == 2 * overlap, the reference's 8 / 4).

Per patch, in index order (:467-694):
  x == 0, y == 0  plain copy (:484-490)
  x == 0, y  > 0  left seam (:492-551)
  x  > 0          top seam (:553-620), then, when y > 0, the left seam as well (:622-692) -- whose write covers the whole patch
                  again, so for interior patches the top seam never survives (kept: results must match the reference's)
Left seam: delta[a][b] = |left[a][p-ov+b] - cur[a][b]|, cumulative cost down the rows with the 3-neighbour minimum of the row
above (2 at the borders), start of the back-track = the LAST column holding the minimum of the bottom row (the reference's loop
keeps overwriting), back-track preference: same column, then b+1, then b-1 (:518-541); row a then takes its first index[a]
pixels from the left neighbour and the rest from the current patch (:543-550).  Top seam: the transposed programme -- but its
COST is taken against patch i-1 (:557; the previous patch in index order, not the patch above) while the pixels are copied from
patch i-L (:611-618).  `fix_top_cost` switches the cost to patch i-L (what the author presumably meant); default is the quirk.
"""
from __future__ import annotations

import numpy as np


def _seam(delta):
    """delta [n][ov] float64 -> index[n] (1-based count of pixels taken from the neighbour on each of the n lines)."""
    n, ov = delta.shape
    path = np.zeros_like(delta)
    path[0] = delta[0]
    for a in range(1, n):
        for b in range(ov):
            lo, hi = max(b - 1, 0), min(b + 1, ov - 1)
            path[a, b] = delta[a, b] + min(path[a - 1, lo:hi + 1])
    idx = np.zeros(n, np.int64)
    m = path[n - 1].min()
    for b in range(ov):
        if path[n - 1, b] == m:
            idx[n - 1] = b + 1
    for a in range(n - 2, -1, -1):
        nxt = idx[a + 1]
        if nxt == 1:
            idx[a] = 1 if path[a, 0] == min(path[a, 0], path[a, 1]) else 2
        elif nxt == ov:
            idx[a] = ov if path[a, ov - 1] == min(path[a, ov - 1], path[a, ov - 2]) else ov - 1
        else:
            b = nxt - 1
            m3 = min(path[a, b], path[a, b - 1], path[a, b + 1])
            if path[a, b] == m3:
                idx[a] = nxt
            elif path[a, b + 1] == m3:
                idx[a] = nxt + 1
            else:
                idx[a] = nxt - 1
    return idx


def line_count(fine, patch, overlap):
    return (fine - overlap) // (patch - overlap)


def stitch(patches, fine, patch, overlap, fix_top_cost=False):
    """patches [L*L][patch][patch] float32 of ONE image -> stitched image [fine][fine] float32 (uncovered pixels 0)."""
    P = np.asarray(patches, np.float32)
    p, ov = patch, overlap
    L = line_count(fine, p, ov)
    assert P.shape == (L * L, p, p) and ov >= 2 and p > ov
    out = np.zeros((fine, fine), np.float64)
    Pd = P.astype(np.float64)
    for i in range(L * L):
        x, y = i // L, i % L
        r0, c0 = x * ov, y * ov

        def left_seam():
            nb = Pd[i - 1]
            delta = np.abs(nb[:, p - ov:] - Pd[i][:, :ov])                    # [p][ov]
            idx = _seam(delta)
            for a in range(p):
                for b in range(idx[a]):
                    out[r0 + a, c0 + b] = nb[a, p - ov + b]
                for b in range(idx[a], p):
                    out[r0 + a, c0 + b] = Pd[i][a, b]

        if x == 0:
            if y == 0:
                out[r0:r0 + p, c0:c0 + p] = Pd[i]
            else:
                left_seam()
        else:
            cost_nb = Pd[i - L] if fix_top_cost else Pd[i - 1]
            delta = np.abs(cost_nb[p - ov:, :] - Pd[i][:ov, :])               # [ov][p]
            idx = _seam(delta.T.copy())                                       # programme along the columns
            top = Pd[i - L]
            for b in range(p):
                for a in range(idx[b]):
                    out[r0 + a, c0 + b] = top[p - ov + a, b]
                for a in range(idx[b], p):
                    out[r0 + a, c0 + b] = Pd[i][a, b]
            if y != 0:
                left_seam()
    return out.astype(np.float32)
