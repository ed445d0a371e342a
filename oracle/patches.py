"""Patch extraction / re-assembly of the patch scripts (oracle; test infra only).  PARITY UNPINNED (oracle/__init__.py).

Restates, loop for loop, train-gray-patch.lua:267-273 (extraction), :588-595 (re-assembly), the batched forms
train-gray-patch-batch.lua:258-264 / :434-442 and the overlapping extraction train-gray-patch-batch-overlap.lua:393-399.
"""
from __future__ import annotations

import numpy as np


def extract(images, patch, line, nper, stride):
    """images [K][H][W] -> patches [K*nper][patch][patch].

    reference (1-based): real_none[(k-1)*patchNumber + i][a][b] =
        img[floor((i-1)/L)*S + a][((i-1) - floor((i-1)/L)*L)*S + b]   with (L, S) = (patchSize, patchSize)
    (train-gray-patch-batch.lua:258-264; note L is patchSize, not fineSize/patchSize) or (overlapPatchLine, overlap)
    (train-gray-patch-batch-overlap.lua:393-399)."""
    images = np.asarray(images)
    K = images.shape[0]
    out = np.zeros((K * nper, patch, patch), images.dtype)
    for k in range(K):
        for i in range(nper):
            r0, c0 = (i // line) * stride, (i % line) * stride
            for a in range(patch):
                for b in range(patch):
                    out[k * nper + i, a, b] = images[k, r0 + a, c0 + b]
    return out


def assemble(patches, images, patch, line, nper, stride):
    """Inverse scatter in the reference's loop order (train-gray-patch-batch.lua:434-442): later patches overwrite earlier
    ones where they overlap; pixels no patch covers keep the value of `images`."""
    patches = np.asarray(patches)
    out = np.array(images, copy=True)
    K = out.shape[0]
    for k in range(K):
        for i in range(nper):
            r0, c0 = (i // line) * stride, (i % line) * stride
            for a in range(patch):
                for b in range(patch):
                    out[k, r0 + a, c0 + b] = patches[k * nper + i, a, b]
    return out
