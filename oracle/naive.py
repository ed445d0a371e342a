"""Independent naive numpy restatement of conv / transposed conv (oracle; test infra only).

Written from the defining sums, sharing no code with the
``torch.nn.functional`` based ``oracle.ops`` -- it exists to catch a bug shared between
that oracle and the library.  Pure loops over the kernel taps, float64, small cases only.
"""
from __future__ import annotations

import numpy as np


def conv2d_fwd(x, w, s, p):
    """y[b,o,i,j] = sum_{c,u,v} w[o,c,u,v] * x[b,c,i*s+u-p,j*s+v-p]   (train.lua:108)."""
    B, C, H, W = x.shape
    O, _, K, _ = w.shape
    Ho, Wo = (H + 2 * p - K) // s + 1, (W + 2 * p - K) // s + 1
    xp = np.zeros((B, C, H + 2 * p, W + 2 * p), dtype=np.float64)
    xp[:, :, p:p + H, p:p + W] = x
    y = np.zeros((B, O, Ho, Wo), dtype=np.float64)
    for u in range(K):
        for v in range(K):
            patch = xp[:, :, u:u + (Ho - 1) * s + 1:s, v:v + (Wo - 1) * s + 1:s]   # B,C,Ho,Wo
            y += np.einsum("bchw,oc->bohw", patch, w[:, :, u, v])
    return y


def fullconv2d_fwd(x, w, s, p, adj=0):
    """y[b,o,i*s-p+u,j*s-p+v] += w[c,o,u,v] * x[b,c,i,j]   (train.lua:99)."""
    B, C, H, W = x.shape
    _, O, K, _ = w.shape
    Ho, Wo = (H - 1) * s - 2 * p + K + adj, (W - 1) * s - 2 * p + K + adj
    full = np.zeros((B, O, (H - 1) * s + K + adj, (W - 1) * s + K + adj), dtype=np.float64)
    for u in range(K):
        for v in range(K):
            contrib = np.einsum("bchw,co->bohw", x, w[:, :, u, v])
            full[:, :, u:u + (H - 1) * s + 1:s, v:v + (W - 1) * s + 1:s] += contrib
    return full[:, :, p:p + Ho, p:p + Wo]


def conv2d_wgrad(x, dy, K, s, p):
    B, C, H, W = x.shape
    O = dy.shape[1]
    Ho, Wo = dy.shape[2], dy.shape[3]
    xp = np.zeros((B, C, H + 2 * p, W + 2 * p), dtype=np.float64)
    xp[:, :, p:p + H, p:p + W] = x
    dw = np.zeros((O, C, K, K), dtype=np.float64)
    for u in range(K):
        for v in range(K):
            patch = xp[:, :, u:u + (Ho - 1) * s + 1:s, v:v + (Wo - 1) * s + 1:s]
            dw[:, :, u, v] = np.einsum("bohw,bchw->oc", dy, patch)
    return dw


def conv2d_dgrad(dy, w, x_shape, s, p):
    B, C, H, W = x_shape
    O, _, K, _ = w.shape
    Ho, Wo = dy.shape[2], dy.shape[3]
    dxp = np.zeros((B, C, H + 2 * p + s, W + 2 * p + s), dtype=np.float64)
    for u in range(K):
        for v in range(K):
            contrib = np.einsum("bohw,oc->bchw", dy, w[:, :, u, v])
            dxp[:, :, u:u + (Ho - 1) * s + 1:s, v:v + (Wo - 1) * s + 1:s] += contrib
    return dxp[:, :, p:p + H, p:p + W]
