"""Functional restatement of the Torch7 ops on the DCGAN-SR hot path (oracle; test infra only).

All tensors are NCHW ``torch`` CPU tensors; the dtype of the inputs (float64 = ground
truth, float32 = "what Torch7 fp32 gives up to summation order") is preserved.
Each function cites the reference call site it stands in for and the Torch7 semantics
it restates.  PARITY UNPINNED: see ``oracle/__init__.py``.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F
from torch.nn import grad as _g

# ----------------------------------------------------------------------------------------
# nn.SpatialConvolution(nIn,nOut,kW,kH,dW,dH,padW,padH)   train.lua:108,111,121-133
# weight [nOut][nIn][kH][kW]; cross-correlation; bias-free (weights_init -> noBias, :42-51)
# ----------------------------------------------------------------------------------------


def conv2d_fwd(x, w, stride, pad):
    return F.conv2d(x, w, None, stride=stride, padding=pad)


def conv2d_dgrad(dy, w, x_shape, stride, pad):
    """updateGradInput of SpatialConvolution: gradient w.r.t. the input."""
    return _g.conv2d_input(list(x_shape), w, dy, stride=stride, padding=pad)


def conv2d_wgrad(x, dy, w_shape, stride, pad):
    """accGradParameters of SpatialConvolution (scale 1); caller accumulates (+=)."""
    return _g.conv2d_weight(x, list(w_shape), dy, stride=stride, padding=pad)


# ----------------------------------------------------------------------------------------
# nn.SpatialFullConvolution(nIn,nOut,kW,kH,dW,dH,padW,padH,adjW,adjH)   train.lua:99-105
# weight [nIn][nOut][kH][kW]; Hout=(H-1)s-2p+k+adj; == gradient of conv w.r.t. its input.
# ----------------------------------------------------------------------------------------


def fullconv2d_fwd(x, w, stride, pad, adj=0):
    return F.conv_transpose2d(x, w, None, stride=stride, padding=pad, output_padding=adj)


def fullconv2d_dgrad(dy, w, stride, pad):
    """Gradient of a transposed conv w.r.t. its input is a plain strided conv of dy."""
    return F.conv2d(dy, w, None, stride=stride, padding=pad)


def fullconv2d_wgrad(x, dy, w_shape, stride, pad):
    # y = conv_transpose(x, w)  <=>  x plays the role of "dy" and dy of "x" of a conv with
    # weight [nIn][nOut] read as [out=nIn][in=nOut].
    return _g.conv2d_weight(dy, list(w_shape), x, stride=stride, padding=pad)


# ----------------------------------------------------------------------------------------
# nn.SpatialBatchNormalization(C) eps 1e-5, momentum 0.1, affine     train.lua:100,103,...
# ----------------------------------------------------------------------------------------


def bn_fwd_train(x, gamma, beta, running_mean, running_var, eps=1e-5, momentum=0.1):
    """Training-mode forward.  Returns (y, save_mean, save_invstd, new_rm, new_rv).

    Biased variance normalises; the *unbiased* one (n/(n-1)) goes into running_var;
    ``save_std`` of Torch7 holds invstd.
    """
    n = x.shape[0] * x.shape[2] * x.shape[3]
    mean = x.mean(dim=(0, 2, 3))
    var_b = ((x - mean[None, :, None, None]) ** 2).mean(dim=(0, 2, 3))
    invstd = 1.0 / torch.sqrt(var_b + eps)
    y = (x - mean[None, :, None, None]) * invstd[None, :, None, None]
    y = y * gamma[None, :, None, None] + beta[None, :, None, None]
    unbiased = var_b * (n / max(n - 1, 1))
    new_rm = (1 - momentum) * running_mean + momentum * mean
    new_rv = (1 - momentum) * running_var + momentum * unbiased
    return y, mean, invstd, new_rm, new_rv


def bn_bwd(x, dy, gamma, save_mean, save_invstd):
    """Training-mode backward.  Returns (dx, dgamma, dbeta) (param grads un-accumulated).

    dx = (dy - mean(dy) - xhat * mean(dy*xhat)) * gamma * invstd ; dgamma = sum(dy*xhat);
    dbeta = sum(dy).  ``updateGradInput`` alone (train.lua:268) uses the *current* gamma
    with the *saved* statistics.
    """
    n = x.shape[0] * x.shape[2] * x.shape[3]
    xhat = (x - save_mean[None, :, None, None]) * save_invstd[None, :, None, None]
    sum_dy = dy.sum(dim=(0, 2, 3))
    sum_dy_xhat = (dy * xhat).sum(dim=(0, 2, 3))
    dx = (dy - (sum_dy / n)[None, :, None, None] - xhat * (sum_dy_xhat / n)[None, :, None, None])
    dx = dx * (gamma * save_invstd)[None, :, None, None]
    return dx, sum_dy_xhat, sum_dy


# ----------------------------------------------------------------------------------------
# activations: nn.ReLU(true), nn.LeakyReLU(0.2,true), nn.Tanh, nn.Sigmoid   train.lua:100..134
# In-place modules: backward tests the (already activated) output.
# ----------------------------------------------------------------------------------------

ACT_NONE, ACT_RELU, ACT_LRELU, ACT_TANH, ACT_SIGMOID = 0, 1, 2, 3, 4


def act_fwd(x, kind, negval=0.2):
    if kind == ACT_NONE:
        return x
    if kind == ACT_RELU:
        return torch.clamp(x, min=0)
    if kind == ACT_LRELU:
        return torch.where(x > 0, x, x * negval)
    if kind == ACT_TANH:
        return torch.tanh(x)
    if kind == ACT_SIGMOID:
        return torch.sigmoid(x)
    raise ValueError(kind)


def act_bwd(y, dy, kind, negval=0.2):
    """Gradient through the activation given its *output* ``y``."""
    if kind == ACT_NONE:
        return dy
    if kind == ACT_RELU:
        return torch.where(y > 0, dy, torch.zeros_like(dy))
    if kind == ACT_LRELU:
        return torch.where(y > 0, dy, dy * negval)
    if kind == ACT_TANH:
        return dy * (1 - y * y)
    if kind == ACT_SIGMOID:
        return dy * y * (1 - y)
    raise ValueError(kind)


# ----------------------------------------------------------------------------------------
# nn.SpatialUpSamplingNearest(2)   train-gray.lua:104, train-gray-patch.lua:56
# ----------------------------------------------------------------------------------------


def upnearest_fwd(x, scale=2):
    return x.repeat_interleave(scale, dim=2).repeat_interleave(scale, dim=3)


def upnearest_bwd(dy, scale=2):
    n, c, h, w = dy.shape
    return dy.reshape(n, c, h // scale, scale, w // scale, scale).sum(dim=(3, 5))


# ----------------------------------------------------------------------------------------
# 2x2 box down-sample loop   train.lua:225-230, train-gray-patch.lua:287-293
# (a[2i-1,2j-1] + a[2i,2j-1] + a[2i-1,2j] + a[2i,2j]) / 4   -- same association order
# ----------------------------------------------------------------------------------------


def avgpool2_fwd(x):
    a = x[:, :, 0::2, 0::2]
    b = x[:, :, 1::2, 0::2]
    c = x[:, :, 0::2, 1::2]
    d = x[:, :, 1::2, 1::2]
    return (((a + b) + c) + d) / 4


# ----------------------------------------------------------------------------------------
# criteria   nn.BCECriterion (train-gray-patch.lua:113), nn.MSECriterion (train.lua:142)
# sizeAverage=true.  BCE eps=1e-12 and is evaluated per element in double by THNN.
# ----------------------------------------------------------------------------------------

BCE_EPS = 1e-12


def bce_fwd(x, t):
    xd, td = x.double(), t.double()
    loss = -(td * torch.log(xd + BCE_EPS) + (1 - td) * torch.log(1 - xd + BCE_EPS)).sum() / x.numel()
    return float(loss)


def bce_bwd(x, t):
    xd, td = x.double(), t.double()
    g = -(td - xd) / ((1 - xd + BCE_EPS) * (xd + BCE_EPS)) / x.numel()
    return g.to(x.dtype)


def mse_fwd(x, t):
    d = (x.double() - t.double())
    return float((d * d).sum() / x.numel())


def mse_bwd(x, t):
    return ((2.0 / x.numel()) * (x - t)).to(x.dtype)


def pixel_mse_per_sample(real, fake, divisor):
    """calMSE loop (train.lua:193-195,237-239): per-sample sum((r-f)^2)/divisor.

    divisor = 4*C*H*W in train.lua, H*W in train-gray.lua:199-201 (img:size(3)*size(4)).
    """
    d = real - fake
    return (d * d).reshape(real.shape[0], -1).sum(dim=1) / divisor


# ----------------------------------------------------------------------------------------
# optim.adam(opfunc, x, config)   train.lua:280,283 ; state inside the config table
# x -= lr * sqrt(1-b2^t)/(1-b1^t) * m / (sqrt(v) + eps)      (NOT torch.optim.Adam)
# ----------------------------------------------------------------------------------------


def adam_step(p, g, m, v, t, lr=2e-4, beta1=0.5, beta2=0.999, eps=1e-8):
    """One Torch7 ``optim.adam`` update, in place on ``p, m, v``; returns the new ``t``."""
    t = t + 1
    m.mul_(beta1).add_(g, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    denom = v.sqrt().add_(eps)
    bc1 = 1 - beta1 ** t
    bc2 = 1 - beta2 ** t
    step = lr * math.sqrt(bc2) / bc1
    p.addcdiv_(m, denom, value=-step)
    return t
