"""CPU tests of the oracle itself: torch-functional restatement vs the independent naive numpy one, the
committed golden vectors (regression pin), Torch7-specific semantics (App. C of SURVEY.md)."""
import math
import os
import sys

import numpy as np
import torch

from oracle import naive, nets, ops
from oracle import step as ostep

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden as mg  # noqa: E402
from util import oracle_net, rel_err, rng, t64  # noqa: E402


def test_conv_matches_naive():
    r = rng(1)
    for (n, ci, h, w, co, k, s, p) in [(2, 3, 9, 7, 4, 4, 2, 1), (1, 2, 6, 6, 3, 3, 1, 0), (2, 1, 5, 5, 2, 5, 1, 2)]:
        x = r.standard_normal((n, ci, h, w))
        wt = r.standard_normal((co, ci, k, k))
        y = ops.conv2d_fwd(t64(x), t64(wt), s, p).numpy()
        assert rel_err(naive.conv2d_fwd(x, wt, s, p), y) < 1e-12
        dy = r.standard_normal(y.shape)
        assert rel_err(naive.conv2d_dgrad(dy, wt, x.shape, s, p), ops.conv2d_dgrad(t64(dy), t64(wt), x.shape, s, p).numpy()) < 1e-12
        assert rel_err(naive.conv2d_wgrad(x, dy, k, s, p), ops.conv2d_wgrad(t64(x), t64(dy), wt.shape, s, p).numpy()) < 1e-12


def test_fullconv_matches_naive_and_is_conv_adjoint():
    r = rng(2)
    for (n, ci, h, w, co, k, s, p) in [(2, 3, 4, 5, 4, 4, 2, 1), (1, 2, 3, 3, 3, 3, 1, 1), (1, 2, 3, 3, 2, 5, 2, 2)]:
        x = r.standard_normal((n, ci, h, w))
        wt = r.standard_normal((ci, co, k, k))
        y = ops.fullconv2d_fwd(t64(x), t64(wt), s, p).numpy()
        assert rel_err(naive.fullconv2d_fwd(x, wt, s, p), y) < 1e-12
        # <fullconv(x), dy> == <x, fullconv_dgrad(dy)> ; and the weight gradient is the derivative of that form
        dy = r.standard_normal(y.shape)
        lhs = float((y * dy).sum())
        rhs = float((x * ops.fullconv2d_dgrad(t64(dy), t64(wt), s, p).numpy()).sum())
        assert abs(lhs - rhs) < 1e-9 * abs(lhs)
        dw = ops.fullconv2d_wgrad(t64(x), t64(dy), wt.shape, s, p).numpy()
        assert abs(float((dw * wt).sum()) - lhs) < 1e-9 * abs(lhs)


def test_batchnorm_semantics():
    r = rng(3)
    x = t64(r.standard_normal((5, 4, 3, 3)) * 3 + 1)
    g, b = t64(r.standard_normal(4)), t64(r.standard_normal(4))
    y, mean, invstd, rm, rv = ops.bn_fwd_train(x, g, b, torch.zeros(4, dtype=torch.float64), torch.ones(4, dtype=torch.float64))
    n = 45
    xs = x.permute(1, 0, 2, 3).reshape(4, -1)
    assert torch.allclose(mean, xs.mean(1))
    assert torch.allclose(invstd, 1 / torch.sqrt(xs.var(1, unbiased=False) + 1e-5))
    assert torch.allclose(rv, 0.9 + 0.1 * xs.var(1, unbiased=True))       # unbiased into running_var
    assert torch.allclose(rm, 0.1 * xs.mean(1))
    # backward vs autograd
    xa = x.clone().requires_grad_(True)
    ga, ba = g.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ya = torch.nn.functional.batch_norm(xa, None, None, ga, ba, True, 0.1, 1e-5)
    dy = t64(r.standard_normal(tuple(y.shape)))
    ya.backward(dy)
    dx, dg, db = ops.bn_bwd(x, dy, g, mean, invstd)
    assert torch.allclose(dx, xa.grad, atol=1e-10) and torch.allclose(dg, ga.grad) and torch.allclose(db, ba.grad)
    assert n == xs.shape[1]


def test_bce_is_torch7_not_pytorch():
    x = torch.tensor([0.0, 1.0, 0.3], dtype=torch.float64)
    t = torch.tensor([1.0, 0.0, 1.0], dtype=torch.float64)
    # eps 1e-12 inside the log: log(1e-12) = -27.63, where torch.nn.BCELoss clamps at -100
    want = -(math.log(1e-12) + math.log(1e-12) + math.log(0.3 + 1e-12)) / 3
    assert abs(ops.bce_fwd(x, t) - want) < 1e-12
    g = ops.bce_bwd(x, t)
    assert abs(float(g[2]) - (-(1 - 0.3) / ((1 - 0.3 + 1e-12) * (0.3 + 1e-12)) / 3)) < 1e-12


def test_adam_is_torch7_not_pytorch():
    p = torch.tensor([1.0], dtype=torch.float64)
    g = torch.tensor([0.5], dtype=torch.float64)
    m, v = torch.zeros(1, dtype=torch.float64), torch.zeros(1, dtype=torch.float64)
    t = ops.adam_step(p, g, m, v, 0, lr=2e-4, beta1=0.5, beta2=0.999, eps=1e-8)
    assert t == 1
    mm, vv = 0.5 * 0.5, 0.001 * 0.25
    want = 1.0 - 2e-4 * math.sqrt(1 - 0.999) / (1 - 0.5) * mm / (math.sqrt(vv) + 1e-8)     # eps OUTSIDE the bias-corrected sqrt
    assert abs(float(p) - want) < 1e-15


def test_sequential_param_order_and_last_forward_wins():
    from dcgan_super_resolution_b200 import models
    net = oracle_net(models.dcgan64_D(3, 4), 1)
    names = [n for n, _, _ in net.param_list()]
    assert names == ["weight", "weight", "weight", "bias", "weight", "weight", "bias", "weight", "weight", "bias", "weight"]
    r = rng(5)
    a, b = t64(r.standard_normal((2, 3, 64, 64))), t64(r.standard_normal((2, 3, 64, 64)))
    net.forward(a)
    out_b = net.forward(b).clone()
    assert torch.equal(net.output, out_b)                       # netD.output is the LAST forward (train.lua:265)
    assert net.output.shape == (2, 1)                           # View(1):setNumInputDims(3)
    assert net.num_params() == sum(p.numel() for _, p, _ in net.param_list())


def test_oracle_reproduces_golden_layers():
    gold = np.load(os.path.join(HERE, "golden", "layers.npz"))
    for name in mg.LAYER_CASES:
        y, dx, dw = mg.layer_outputs(name)
        for k, v in ((".y", y), (".dx", dx), (".dw", dw)):
            assert rel_err(v, gold[name + k]) < 1e-12, name + k
    gm = np.load(os.path.join(HERE, "golden", "misc.npz"))
    for k, v in mg.misc_outputs().items():
        assert rel_err(v, gm[k]) < 1e-12, k


def test_oracle_reproduces_golden_steps():
    for name in mg.STEP_CASES:
        gold = np.load(os.path.join(HERE, "golden", f"step_{name}.npz"))
        out = mg.step_outputs(name)
        assert rel_err(out["losses"], gold["losses"]) < 1e-10
        assert rel_err(out["pG"], gold["pG"]) < 1e-10 and rel_err(out["pD"], gold["pD"]) < 1e-10


def test_float32_oracle_close_to_float64():
    out64 = mg.step_outputs("bce_patch", torch.float64)
    out32 = mg.step_outputs("bce_patch", torch.float32)
    assert rel_err(out32["losses"], out64["losses"]) < 1e-4


def test_step_is_data_parallel_decomposable():
    """Shard -> local backward with the criterion divided by the GLOBAL count -> summed gradients == big batch,
    when BN statistics are global (sync_bn) -- emulated here with a BN-free pair of nets."""
    specsD = [dict(kind="conv", cin=1, cout=4, k=4, s=2, p=1), dict(kind="lrelu", negval=0.2),
              dict(kind="conv", cin=4, cout=1, k=4, s=1, p=0), dict(kind="sigmoid"), dict(kind="view")]
    D = oracle_net(specsD, 3)
    r = rng(6)
    x = t64(r.uniform(0, 1, (8, 1, 8, 8)))
    out = D.forward(x)
    lab = torch.ones_like(out)
    D.zero_grad_parameters()
    D.backward(x, ops.bce_bwd(out, lab))
    full = D.get_flat_grads().clone()
    acc = torch.zeros_like(full)
    for rank in range(2):
        xs = x[rank * 4:(rank + 1) * 4]
        o = D.forward(xs)
        g = -(torch.ones_like(o) - o) / ((1 - o + ops.BCE_EPS) * (o + ops.BCE_EPS)) / 8      # global n = 8
        D.zero_grad_parameters()
        D.backward(xs, g)
        acc += D.get_flat_grads()
    assert torch.allclose(acc, full, atol=1e-12)


def test_sync_bn_sums_over_shards_equal_the_big_batch():
    """What sync_bn=1 exchanges between ranks (kernels_peer.cu / ncclAllReduce): per-shard (sum x, sum x^2) forward and
    (sum g, sum g*xhat) backward, added in rank order and finalised as var = sum x^2 / n - mean^2 with the GLOBAL n -- the
    arithmetic of bn_finalize_peer_kernel and bn_bwd_apply -- reproduces SpatialBatchNormalization on the concatenated
    batch (train.lua:100-109 semantics: biased variance normalises, unbiased n/(n-1) goes to running_var)."""
    r = rng(11)
    W, b = 4, 3
    x = t64(r.standard_normal((W * b, 5, 4, 6)) * 2 + 0.5)
    dy = t64(r.standard_normal(tuple(x.shape)))
    g, be = t64(r.standard_normal(5)), t64(r.standard_normal(5))
    rm0, rv0 = t64(r.standard_normal(5)), t64(r.uniform(0.5, 2.0, 5))
    y, mean, invstd, rm, rv = ops.bn_fwd_train(x, g, be, rm0, rv0)
    dx, dg, db = ops.bn_bwd(x, dy, g, mean, invstd)
    # forward: the ranks' sums, rank order
    s1, s2 = torch.zeros(5, dtype=torch.float64), torch.zeros(5, dtype=torch.float64)
    for k in range(W):
        xs = x[k * b:(k + 1) * b]
        s1 = s1 + xs.sum(dim=(0, 2, 3))
        s2 = s2 + (xs * xs).sum(dim=(0, 2, 3))
    n = float(W * b * 4 * 6)
    m = s1 / n
    var = torch.clamp(s2 / n - m * m, min=0.0)
    inv = 1.0 / torch.sqrt(var + 1e-5)
    assert torch.allclose(m, mean, atol=1e-12) and torch.allclose(inv, invstd, rtol=1e-10)
    assert torch.allclose(0.9 * rm0 + 0.1 * m, rm, atol=1e-12)
    assert torch.allclose(0.9 * rv0 + 0.1 * var * (n / (n - 1.0)), rv, rtol=1e-10)
    # every rank normalises its shard with the global statistics: the concatenation is the big-batch output
    ys = torch.cat([(x[k * b:(k + 1) * b] - m[None, :, None, None]) * inv[None, :, None, None] * g[None, :, None, None] + be[None, :, None, None]
                    for k in range(W)])
    assert torch.allclose(ys, y, atol=1e-10)
    # backward: the ranks' (sum g, sum g*xhat); dgamma / dbeta are the LOCAL sums (the gradient all-reduce adds them later)
    t1, t2 = torch.zeros(5, dtype=torch.float64), torch.zeros(5, dtype=torch.float64)
    loc = []
    for k in range(W):
        xs, ds = x[k * b:(k + 1) * b], dy[k * b:(k + 1) * b]
        xh = (xs - m[None, :, None, None]) * inv[None, :, None, None]
        a1, a2 = ds.sum(dim=(0, 2, 3)), (ds * xh).sum(dim=(0, 2, 3))
        loc.append((a1, a2))
        t1, t2 = t1 + a1, t2 + a2
    dxs = torch.cat([(dy[k * b:(k + 1) * b] - (t1 / n)[None, :, None, None]
                      - (x[k * b:(k + 1) * b] - m[None, :, None, None]) * inv[None, :, None, None] * (t2 / n)[None, :, None, None])
                     * (g * inv)[None, :, None, None] for k in range(W)])
    assert torch.allclose(dxs, dx, atol=1e-10)
    assert torch.allclose(sum(a for a, _ in loc), db, atol=1e-10) and torch.allclose(sum(a for _, a in loc), dg, atol=1e-10)
