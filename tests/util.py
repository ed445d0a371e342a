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


# ---------------------------------------------------------------------------------------------------------
# TF32 emulation of the ORACLE (test infrastructure).  FAST_TF32 feeds the tensor cores fp32 activations (the
# hardware ignores the low 13 mantissa bits) and TF32-rounded weights (cvt.rna at pack time).  On whole nets
# with BatchNorm and ReLU kinks a *correct* TF32 evaluation is 3-11 % (max-norm) away from the float64 oracle
# in the backward pass (scripts/tf32_conditioning.py), so net-level FAST tests compare against the float64
# oracle evaluated with the same operand rounding; what remains is summation order.
# ---------------------------------------------------------------------------------------------------------
import contextlib

from oracle import ops as _ops


def trunc_tf32(x):
    a = x.to(torch.float32).contiguous().view(torch.int32)
    return (a & ~0x1FFF).view(torch.float32).to(x.dtype)


def rna_tf32(x):
    a = x.to(torch.float32).contiguous().view(torch.int32)
    return ((a + 0x1000) & ~0x1FFF).view(torch.float32).to(x.dtype)


@contextlib.contextmanager
def tf32_oracle(act_round=trunc_tf32, thin_exact=True):
    """Inside the block every convolution of the oracle rounds its operands like FAST_TF32 does.  Layers whose
    contraction channel count is not a multiple of 8 (C in {1, 3}) run on the fp32 SIMT kernels in the library,
    so they stay exact when `thin_exact`."""
    names = ("conv2d_fwd", "conv2d_dgrad", "conv2d_wgrad", "fullconv2d_fwd", "fullconv2d_dgrad", "fullconv2d_wgrad")
    orig = {k: getattr(_ops, k) for k in names}

    def tc2(cin, cout, npix):
        """True when the library runs this contraction on the tensor cores: contraction channels a multiple of 8, and not
        one of the streaming fp32 kernels (thin input: <= 4 contraction channels; thin output: <= 4 output channels with
        few pixels -- kernels_thin.cu:thin_out_supported)."""
        if not thin_exact:
            return True
        if cin % 8 or cin <= 4:
            return False
        if cout <= 4 and cin % 4 == 0 and npix <= 148 * 32:
            return False
        return True

    def r(x, on):
        return act_round(x) if on else x

    def w_(w, on):
        return rna_tf32(w) if on else w

    def _wg(ws):
        return (not thin_exact) or (ws[0] % 4 == 0 and ws[1] % 4 == 0 and ws[0] >= 8 and ws[1] >= 8)

    # forward contracts over cin, dgrad over cout, wgrad over pixels; npix = pixels of the iterated grid (per class)
    def conv_fwd(x, w, s, p):
        ho = (x.shape[2] + 2 * p - w.shape[2]) // s + 1
        wo = (x.shape[3] + 2 * p - w.shape[3]) // s + 1
        on = tc2(w.shape[1], w.shape[0], x.shape[0] * ho * wo)
        return orig["conv2d_fwd"](r(x, on), w_(w, on), s, p)

    def conv_dgrad(dy, w, xs, s, p):
        on = tc2(w.shape[0], w.shape[1], xs[0] * -(-xs[2] // s) * -(-xs[3] // s))
        return orig["conv2d_dgrad"](r(dy, on), w_(w, on), xs, s, p)

    def full_fwd(x, w, s, p, adj=0):
        on = tc2(w.shape[0], w.shape[1], x.shape[0] * x.shape[2] * x.shape[3])
        return orig["fullconv2d_fwd"](r(x, on), w_(w, on), s, p, adj)

    def full_dgrad(dy, w, s, p):
        hi = (dy.shape[2] + 2 * p - w.shape[2]) // s + 1
        wi = (dy.shape[3] + 2 * p - w.shape[3]) // s + 1
        on = tc2(w.shape[1], w.shape[0], dy.shape[0] * hi * wi)
        return orig["fullconv2d_dgrad"](r(dy, on), w_(w, on), s, p)

    _ops.conv2d_fwd = conv_fwd
    _ops.conv2d_dgrad = conv_dgrad
    _ops.conv2d_wgrad = lambda x, dy, ws, s, p: orig["conv2d_wgrad"](r(x, _wg(ws)), r(dy, _wg(ws)), ws, s, p)
    _ops.fullconv2d_fwd = full_fwd
    _ops.fullconv2d_dgrad = full_dgrad
    _ops.fullconv2d_wgrad = lambda x, dy, ws, s, p: orig["fullconv2d_wgrad"](r(x, _wg(ws)), r(dy, _wg(ws)), ws, s, p)

    try:
        yield
    finally:
        for k, v in orig.items():
            setattr(_ops, k, v)


def l2_err(a, b):
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    n = np.linalg.norm(b)
    return np.linalg.norm(a - b) / n if n > 0 else np.linalg.norm(a - b)
