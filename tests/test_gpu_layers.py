"""Per-layer parity: CUDA path (through the C ABI) vs the float64 oracle, strict fp32 <= 1e-5."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import naive, ops
from util import STRICT_TOL, ptr, rel_err, rng, t64

pytestmark = pytest.mark.gpu

# (n, cin, h, w, cout, k, s, p)
CONV_SHAPES = [
    (2, 3, 16, 16, 8, 4, 2, 1),        # first D layer style (Ci = 3)
    (3, 1, 8, 8, 64, 3, 1, 0),         # patch-D layer 1 (Ci = 1)
    (4, 64, 6, 6, 128, 3, 1, 0),       # patch-D layer 2 of C1a
    (2, 64, 32, 32, 128, 4, 2, 1),     # D layer 2 of C3a, batch reduced
    (2, 12, 16, 16, 3, 4, 2, 1),       # G last layer (Co = 3, Ci = 12)
    (5, 16, 8, 8, 1, 4, 2, 1),         # Co = 1
    (3, 512, 4, 4, 1, 4, 1, 0),        # D final layer
    (2, 24, 10, 14, 12, 4, 2, 1),      # ragged: non-square, Ci = 24, Co = 12
    (1, 8, 7, 9, 20, 3, 2, 1),         # odd sizes, stride 2 with k 3
    (2, 4, 5, 5, 6, 5, 1, 2),          # 5x5 kernel
    (2, 256, 2, 2, 1, 2, 1, 0),        # patch-D final 2x2
    (3, 12, 96, 80, 3, 4, 2, 1),       # G last layer at size: > 148*32 output pixels -> the pixel-per-thread thin-output kernel
    (2, 16, 128, 96, 1, 4, 2, 1),      # same with Co = 1 (train-gray.lua:116)
]


def _lib(ctx):
    return ctx.lib, ctx.h


def _chk(ctx, rc):
    from dcgan_super_resolution_b200 import _lib as L
    L.check(rc, ctx.h)


@pytest.mark.parametrize("shape", CONV_SHAPES)
def test_conv2d_fwd_dgrad_wgrad(ctx, shape):
    n, cin, h, w, cout, k, s, p = shape
    r = rng(hash(shape) % 2**31)
    x = r.standard_normal((n, cin, h, w)).astype(np.float32)
    wt = (r.standard_normal((cout, cin, k, k)) * 0.1).astype(np.float32)
    ho, wo = (h + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1
    dy = r.standard_normal((n, cout, ho, wo)).astype(np.float32)
    lib, hctx = _lib(ctx)
    args = (n, cin, h, w, cout, k, s, p)

    y = np.empty((n, cout, ho, wo), np.float32)
    _chk(ctx, lib.dcgansr_conv2d_fwd(hctx, ptr(x), ptr(wt), ptr(y), *args))
    ref = ops.conv2d_fwd(t64(x), t64(wt), s, p).numpy()
    assert rel_err(y, ref) <= STRICT_TOL
    assert rel_err(naive.conv2d_fwd(x.astype(np.float64), wt.astype(np.float64), s, p), ref) < 1e-12

    dx = np.empty_like(x)
    _chk(ctx, lib.dcgansr_conv2d_dgrad(hctx, ptr(dy), ptr(wt), ptr(dx), *args))
    ref = ops.conv2d_dgrad(t64(dy), t64(wt), x.shape, s, p).numpy()
    assert rel_err(dx, ref) <= STRICT_TOL

    dw = np.empty_like(wt)
    _chk(ctx, lib.dcgansr_conv2d_wgrad(hctx, ptr(x), ptr(dy), ptr(dw), *args))
    ref = ops.conv2d_wgrad(t64(x), t64(dy), wt.shape, s, p).numpy()
    assert rel_err(dw, ref) <= STRICT_TOL


FULL_SHAPES = [
    (2, 3, 8, 8, 96, 4, 2, 1),         # train.lua G layer 1 (Ci = 3)
    (2, 1, 8, 8, 64, 4, 2, 1),         # gray G layer 1 (Ci = 1)
    (2, 96, 8, 8, 48, 4, 2, 1),        # train.lua G layer 2
    (3, 48, 6, 10, 24, 4, 2, 1),       # non-square
    (2, 64, 16, 16, 32, 4, 2, 1),      # C2 G layer 3, spatial reduced
    (2, 8, 5, 5, 4, 3, 1, 1),          # stride 1 full conv
    (1, 4, 4, 4, 5, 5, 2, 2),          # k5 s2 (uneven taps per class), Co = 5
    (2, 3, 6, 6, 1024, 4, 2, 1),       # C5 G layer 1 (FC 3->1024): the thin wgrad's block reduction at its 48 KB limit
]


@pytest.mark.parametrize("shape", FULL_SHAPES)
def test_fullconv2d_fwd_dgrad_wgrad(ctx, shape):
    n, cin, h, w, cout, k, s, p = shape
    r = rng(hash(shape) % 2**31)
    x = r.standard_normal((n, cin, h, w)).astype(np.float32)
    wt = (r.standard_normal((cin, cout, k, k)) * 0.1).astype(np.float32)
    ho, wo = (h - 1) * s - 2 * p + k, (w - 1) * s - 2 * p + k
    dy = r.standard_normal((n, cout, ho, wo)).astype(np.float32)
    lib, hctx = _lib(ctx)
    args = (n, cin, h, w, cout, k, s, p)

    y = np.empty((n, cout, ho, wo), np.float32)
    _chk(ctx, lib.dcgansr_fullconv2d_fwd(hctx, ptr(x), ptr(wt), ptr(y), *args))
    ref = ops.fullconv2d_fwd(t64(x), t64(wt), s, p).numpy()
    assert rel_err(y, ref) <= STRICT_TOL
    assert rel_err(naive.fullconv2d_fwd(x.astype(np.float64), wt.astype(np.float64), s, p), ref) < 1e-12

    dx = np.empty_like(x)
    _chk(ctx, lib.dcgansr_fullconv2d_dgrad(hctx, ptr(dy), ptr(wt), ptr(dx), *args))
    ref = ops.fullconv2d_dgrad(t64(dy), t64(wt), s, p).numpy()
    assert rel_err(dx, ref) <= STRICT_TOL

    dw = np.empty_like(wt)
    _chk(ctx, lib.dcgansr_fullconv2d_wgrad(hctx, ptr(x), ptr(dy), ptr(dw), *args))
    ref = ops.fullconv2d_wgrad(t64(x), t64(dy), wt.shape, s, p).numpy()
    assert rel_err(dw, ref) <= STRICT_TOL


@pytest.mark.parametrize("shape", [(4, 8, 6, 6), (3, 12, 5, 7), (2, 3, 9, 9), (64, 128, 4, 4), (2, 1, 8, 8), (1, 24, 33, 17)])
def test_batchnorm_fwd_bwd(ctx, shape):
    n, c, h, w = shape
    r = rng(7 + c)
    x = (r.standard_normal(shape) * 2 + 0.5).astype(np.float32)
    gamma = (1 + 0.02 * r.standard_normal(c)).astype(np.float32)
    beta = (0.1 * r.standard_normal(c)).astype(np.float32)
    rm = (0.1 * r.standard_normal(c)).astype(np.float32)
    rv = (1 + 0.1 * r.random(c)).astype(np.float32)
    dy = r.standard_normal(shape).astype(np.float32)
    lib, hctx = _lib(ctx)
    y = np.empty(shape, np.float32)
    sm, si = np.empty(c, np.float32), np.empty(c, np.float32)
    rm2, rv2 = rm.copy(), rv.copy()
    _chk(ctx, lib.dcgansr_bn_fwd_train(hctx, ptr(x), ptr(gamma), ptr(beta), ptr(rm2), ptr(rv2), ptr(y), ptr(sm), ptr(si),
                                       n, c, h, w, 1e-5, 0.1))
    ry, rmean, rinv, nrm, nrv = ops.bn_fwd_train(t64(x), t64(gamma), t64(beta), t64(rm), t64(rv), 1e-5, 0.1)
    assert rel_err(y, ry.numpy()) <= STRICT_TOL
    assert rel_err(sm, rmean.numpy()) <= STRICT_TOL
    assert rel_err(si, rinv.numpy()) <= STRICT_TOL
    assert rel_err(rm2, nrm.numpy()) <= STRICT_TOL
    assert rel_err(rv2, nrv.numpy()) <= STRICT_TOL

    dx = np.empty(shape, np.float32)
    dg, db = np.empty(c, np.float32), np.empty(c, np.float32)
    _chk(ctx, lib.dcgansr_bn_bwd(hctx, ptr(x), ptr(dy), ptr(gamma), ptr(sm), ptr(si), ptr(dx), ptr(dg), ptr(db), n, c, h, w))
    rdx, rdg, rdb = ops.bn_bwd(t64(x), t64(dy), t64(gamma), rmean, rinv)
    assert rel_err(dx, rdx.numpy()) <= STRICT_TOL
    assert rel_err(dg, rdg.numpy()) <= STRICT_TOL
    assert rel_err(db, rdb.numpy()) <= STRICT_TOL


@pytest.mark.parametrize("kind", [1, 2, 3, 4])
@pytest.mark.parametrize("count", [1, 7, 1024, 4099])
def test_activations(ctx, kind, count):
    r = rng(kind * 100 + count)
    x = (r.standard_normal(count) * 2).astype(np.float32)
    dy = r.standard_normal(count).astype(np.float32)
    lib, hctx = _lib(ctx)
    y = np.empty_like(x)
    _chk(ctx, lib.dcgansr_act_fwd(hctx, ptr(x), ptr(y), count, kind, 0.2))
    ref = ops.act_fwd(t64(x), kind, 0.2).numpy()
    assert rel_err(y, ref) <= STRICT_TOL
    dx = np.empty_like(x)
    _chk(ctx, lib.dcgansr_act_bwd(hctx, ptr(y), ptr(dy), ptr(dx), count, kind, 0.2))
    ref = ops.act_bwd(t64(y), t64(dy), kind, 0.2).numpy()
    assert rel_err(dx, ref) <= STRICT_TOL


def test_act_empty(ctx):
    lib, hctx = _lib(ctx)
    x = np.zeros(1, np.float32)
    _chk(ctx, lib.dcgansr_act_fwd(hctx, ptr(x), ptr(x), 0, 1, 0.2))


@pytest.mark.parametrize("shape", [(2, 1, 4, 4), (3, 3, 5, 6), (2, 8, 3, 3)])
def test_upnearest_avgpool(ctx, shape):
    n, c, h, w = shape
    r = rng(11)
    x = r.standard_normal(shape).astype(np.float32)
    lib, hctx = _lib(ctx)
    y = np.empty((n, c, 2 * h, 2 * w), np.float32)
    _chk(ctx, lib.dcgansr_upnearest2_fwd(hctx, ptr(x), ptr(y), n, c, h, w))
    assert np.array_equal(y, ops.upnearest_fwd(torch.from_numpy(x), 2).numpy())       # bit exact: pure copy
    dy = r.standard_normal(y.shape).astype(np.float32)
    dx = np.empty_like(x)
    _chk(ctx, lib.dcgansr_upnearest2_bwd(hctx, ptr(dy), ptr(dx), n, c, h, w))
    assert rel_err(dx, ops.upnearest_bwd(t64(dy), 2).numpy()) <= STRICT_TOL
    big = r.standard_normal((n, c, 2 * h, 2 * w)).astype(np.float32)
    small = np.empty((n, c, h, w), np.float32)
    _chk(ctx, lib.dcgansr_avgpool2_fwd(hctx, ptr(big), ptr(small), n, c, 2 * h, 2 * w))
    # same association order as the Lua loop -> bit exact against the float32 oracle
    assert np.array_equal(small, ops.avgpool2_fwd(torch.from_numpy(big)).numpy())


@pytest.mark.parametrize("count", [1, 64, 1280, 40000])
def test_criteria(ctx, count):
    r = rng(count)
    x = r.uniform(0.02, 0.98, count).astype(np.float32)
    t = (r.random(count) > 0.5).astype(np.float32)
    lib, hctx = _lib(ctx)
    loss = np.empty(1, np.float32)
    dx = np.empty(count, np.float32)
    _chk(ctx, lib.dcgansr_bce(hctx, ptr(x), ptr(t), count, ptr(loss), ptr(dx)))
    assert abs(loss[0] - ops.bce_fwd(t64(x), t64(t))) <= STRICT_TOL * abs(loss[0])
    assert rel_err(dx, ops.bce_bwd(t64(x), t64(t)).numpy()) <= STRICT_TOL
    tt = r.standard_normal(count).astype(np.float32)
    _chk(ctx, lib.dcgansr_mse(hctx, ptr(x), ptr(tt), count, ptr(loss), ptr(dx)))
    assert abs(loss[0] - ops.mse_fwd(t64(x), t64(tt))) <= STRICT_TOL * abs(loss[0])
    assert rel_err(dx, ops.mse_bwd(t64(x), t64(tt)).numpy()) <= STRICT_TOL


def test_bce_saturated(ctx):
    """x = 0 / 1 exactly: eps 1e-12 keeps the loss finite (Torch7 BCECriterion, not torch.nn.BCELoss)."""
    x = np.array([0.0, 1.0, 0.0, 1.0], np.float32)
    t = np.array([0.0, 1.0, 1.0, 0.0], np.float32)
    lib, hctx = _lib(ctx)
    loss = np.empty(1, np.float32)
    dx = np.empty(4, np.float32)
    _chk(ctx, lib.dcgansr_bce(hctx, ptr(x), ptr(t), 4, ptr(loss), ptr(dx)))
    ref = ops.bce_fwd(t64(x), t64(t))
    assert np.isfinite(loss[0]) and abs(loss[0] - ref) <= 1e-5 * abs(ref)


def test_pixel_mse(ctx):
    r = rng(5)
    a = r.standard_normal((6, 3, 8, 8)).astype(np.float32)
    b = r.standard_normal((6, 3, 8, 8)).astype(np.float32)
    out = np.empty(6, np.float32)
    lib, hctx = _lib(ctx)
    _chk(ctx, lib.dcgansr_pixel_mse_per_sample(hctx, ptr(a), ptr(b), ptr(out), 6, 3 * 8 * 8, 4.0 * 3 * 8 * 8))
    ref = ops.pixel_mse_per_sample(t64(a), t64(b), 4.0 * 3 * 8 * 8).numpy()
    assert rel_err(out, ref) <= STRICT_TOL


@pytest.mark.parametrize("count", [5, 4096, 100003])
def test_adam(ctx, count):
    r = rng(count)
    p = r.standard_normal(count).astype(np.float32)
    g = r.standard_normal(count).astype(np.float32)
    m = (0.1 * r.standard_normal(count)).astype(np.float32)
    v = (0.01 * r.random(count)).astype(np.float32)
    pr, mr, vr = t64(p).clone(), t64(m).clone(), t64(v).clone()
    tnew = ops.adam_step(pr, t64(g), mr, vr, 3, 2e-4, 0.5, 0.999, 1e-8)
    assert tnew == 4
    lib, hctx = _lib(ctx)
    p2, m2, v2 = p.copy(), m.copy(), v.copy()
    _chk(ctx, lib.dcgansr_adam_step(hctx, ptr(p2), ptr(g), ptr(m2), ptr(v2), count, 3, 2e-4, 0.5, 0.999, 1e-8))
    assert rel_err(p2 - p, (pr - t64(p)).numpy()) <= 1e-3      # the update itself (fp32 cancellation bound)
    assert rel_err(p2, pr.numpy()) <= STRICT_TOL
    assert rel_err(m2, mr.numpy()) <= STRICT_TOL
    assert rel_err(v2, vr.numpy()) <= STRICT_TOL


def test_errors_are_loud(ctx):
    from dcgan_super_resolution_b200 import DcgansrError
    lib, hctx = _lib(ctx)
    x = np.zeros(4, np.float32)
    rc = lib.dcgansr_conv2d_fwd(hctx, ptr(x), ptr(x), ptr(x), 1, 1, 2, 2, 1, 3, 1, 0)   # kernel larger than input
    assert rc != 0
    with pytest.raises(DcgansrError):
        _chk(ctx, rc)
