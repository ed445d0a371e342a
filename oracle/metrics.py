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


def scale_bilinear(img, dst_h, dst_w):
    """image.scale(src, W, H) in its default 'bilinear' mode, the eval sweeps' baseline (train-gray-3.lua:399,
    train-gray-patch-batch-overlap.lua:424 -- the `bilinear` argument there is an unset global, i.e. the default).  Upstream
    torch/image (un-vendored) scales the rows then the columns, each a 1-D linear interpolation with the END POINTS ALIGNED:
    for dst_len > src_len, scale = (src_len-1)/(dst_len-1), dst[d] = (1-f)*src[i] + f*src[i+1] with i = floor(d*scale),
    f = d*scale - i, and the last sample copied; float32 arithmetic with a float32 intermediate image.  Only enlarging (or
    equal) sizes are restated: that is all the reference uses."""
    src = np.asarray(img, np.float32)
    assert src.ndim == 2 and dst_h >= src.shape[0] and dst_w >= src.shape[1]

    def lin(x, dst_len):                       # along the last axis
        src_len = x.shape[-1]
        if dst_len == src_len:
            return x.copy()
        out = np.empty(x.shape[:-1] + (dst_len,), np.float32)
        if src_len == 1:
            out[...] = x
            return out
        scale = np.float32(src_len - 1) / np.float32(dst_len - 1)
        d = np.arange(dst_len - 1, dtype=np.float32)
        sf = d * scale
        si = sf.astype(np.int64)
        f = sf - si.astype(np.float32)
        out[..., :-1] = (np.float32(1) - f) * x[..., si] + f * x[..., si + 1]
        out[..., -1] = x[..., -1]
        return out

    tmp = lin(src, dst_w)                      # width first (rows of the source)
    return lin(tmp.T.copy(), dst_h).T.copy()
