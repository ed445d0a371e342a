"""calPSNR / calSSIM of the eval sweeps (oracle; test infra only).  PARITY UNPINNED (oracle/__init__.py).

Restates train-gray-3.lua:143-151 (calPSNR) and :156-221 (calSSIM, after coupriec/VideoPredictionICLR2016) including the
upstream `image` package pieces they call (un-vendored torch/image): `image.gaussian(size, sigma, amplitude)` with sigma
relative to the size and centre 0.5*size + 0.5 (1-based), `image.convolve(x, k, 'full')` = zero-padded true convolution.
"""
from __future__ import annotations

import numpy as np


def psnr(a, b):
    """MSE = sum((a-b)^2) / (H*W); PSNR = 10*log(1/MSE)/log(10), 99 if MSE == 0   (train-gray-3.lua:143-151)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    mse = ((a - b) ** 2).sum() / (a.shape[0] * a.shape[1])
    return 10.0 * np.log(1.0 / mse) / np.log(10.0) if mse > 0 else 99.0


def gaussian_window(size=11, sigma=1.5 / 11):
    """image.gaussian(11, 1.5/11, 0.0708) normalised by its sum (train-gray-3.lua:199-201); the amplitude cancels."""
    i = np.arange(1, size + 1, dtype=np.float64)
    c = 0.5 * size + 0.5
    g = np.exp(-(((i - c) / (sigma * size)) ** 2) / 2.0)
    w = np.outer(g, g)
    return w / w.sum()


def convolve_full(x, k):
    """image.convolve(x, k, 'full'): out[i][j] = sum_{a,b} k[a][b] * x[i-a][j-b], zero outside."""
    H, W = x.shape
    kh, kw = k.shape
    out = np.zeros((H + kh - 1, W + kw - 1), np.float64)
    for a in range(kh):
        for b in range(kw):
            out[a:a + H, b:b + W] += k[a, b] * x
    return out


def ssim(a, b):
    """train-gray-3.lua:185-218."""
    x = (np.asarray(a, np.float64) + 1) / 2 * 255
    y = (np.asarray(b, np.float64) + 1) / 2 * 255
    C1, C2 = (0.01 * 255) ** 2, (0.03 * 255) ** 2
    w = gaussian_window()
    mu1, mu2 = convolve_full(x, w), convolve_full(y, w)
    mu1_sq, mu2_sq, mu1_mu2 = mu1 * mu1, mu2 * mu2, mu1 * mu2
    s1 = convolve_full(x * x, w) - mu1_sq
    s2 = convolve_full(y * y, w) - mu2_sq
    s12 = convolve_full(x * y, w) - mu1_mu2
    m = ((2 * mu1_mu2 + C1) * (2 * s12 + C2)) / ((mu1_sq + mu2_sq + C1) * (s1 + s2 + C2))
    return float(m.mean())
