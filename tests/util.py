"""Shared helpers for the parity tests (oracle = checker, library = thing under test)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from oracle import nets as onets
from oracle import step as ostep

STRICT_TOL = 1e-5     # north_star: strict fp32 within 1e-5 relative (norm-wise, vs the float64 oracle)
FAST_TOL = 2e-3       # north_star: TF32 fast mode within 2e-3


def rel_err(a, b):
    """norm-wise relative error max|a-b| / max|b|."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    d = np.max(np.abs(a - b)) if a.size else 0.0
    n = np.max(np.abs(b)) if b.size else 0.0
    return d / n if n > 0 else d


def rng(seed):
    return np.random.Generator(np.random.Philox(seed))


def t64(a):
    return torch.from_numpy(np.asarray(a, np.float64))


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def oracle_net(specs, seed, dtype=torch.float64):
    net = onets.Sequential(specs, dtype)
    onets.weights_init(net, seed)
    return net


def smooth_images(r, shape, lo, hi):
    """low-pass filtered uniform noise (3x3 box blur) so GAN losses are non-degenerate."""
    x = r.uniform(lo, hi, size=shape)
    p = np.pad(x, ((0, 0), (0, 0), (1, 1), (1, 1)), mode="edge")
    acc = np.zeros_like(x)
    for dy in range(3):
        for dx in range(3):
            acc += p[:, :, dy:dy + shape[2], dx:dx + shape[3]]
    return (acc / 9.0).astype(np.float32)


def ostep_cfg(step: dict, **kw):
    d = dict(step)
    d.update(kw)
    return ostep.StepCfg(**d)
