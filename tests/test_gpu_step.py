"""Net-level and step-level parity: nn.Sequential forward/backward/updateGradInput and the fused
training step (fDx -> adam(D) -> fGx -> adam(G), train.lua:208-283) vs the float64 oracle."""
import numpy as np
import pytest
import torch

import dcgan_super_resolution_b200 as dsr
from dcgan_super_resolution_b200 import models
from oracle import step as ostep
from util import STRICT_TOL, oracle_net, ostep_cfg, rel_err, rng, smooth_images, t64

pytestmark = pytest.mark.gpu

NETS = {
    # name: (specs, input shape)
    "train_lua_G": (models.train_lua_G(3, 4), (3, 4, 4)),
    "train_gray_G": (models.train_gray_G(4), (1, 4, 4)),
    "train_gray_2_G": (models.train_gray_2_G(8), (1, 2, 2)),
    "train_gray_3_G": (models.train_gray_3_G(4), (1, 4, 4)),
    "patch_batch_G": (models.patch_batch_G(4), (1, 2, 2)),
    "dcgan64_D": (models.dcgan64_D(3, 8), (3, 64, 64)),
    "dcgan64_D_big": (models.dcgan64_D(1, 8), (1, 96, 96)),     # D does not end at 1x1
    "patch_D": (models.patch_D(16), (1, 8, 8)),
}


@pytest.mark.parametrize("name", sorted(NETS))
def test_net_forward_backward(ctx, name):
    specs, ishape = NETS[name]
    B = 5
    onet = oracle_net(specs, seed=4321)
    net = dsr.Sequential.from_specs(specs).cuda(ctx, ishape, B)
    assert net.num_params() == onet.num_params()
    flat = onet.get_flat_params().numpy().astype(np.float32)
    net.set_params(flat)
    assert np.array_equal(net.get_params(), flat)
    r = rng(1234)
    x = r.uniform(-1, 1, (B,) + ishape).astype(np.float32)
    y = net.forward(x)
    ry = onet.forward(t64(x))
    c, h, w = net.out_shape()
    assert rel_err(y.reshape(-1), ry.numpy().reshape(-1)) <= STRICT_TOL
    dy = r.standard_normal(y.shape).astype(np.float32)
    net.zeroGradParameters()
    onet.zero_grad_parameters()
    dx = net.backward(x, dy)
    rdx = onet.backward(t64(x), t64(dy).reshape(ry.shape))
    assert rel_err(dx, rdx.numpy()) <= STRICT_TOL
    g = net.get_grads()
    rg = onet.get_flat_grads().numpy()
    # per parameter tensor (each has its own scale)
    off = 0
    for _, p, _g in onet.param_list():
        n = p.numel()
        assert rel_err(g[off:off + n], rg[off:off + n]) <= STRICT_TOL, (name, off)
        off += n
    # grads accumulate (+=) over a second backward, running stats updated per forward (F6)
    net.backward(x, dy, need_dx=False)
    assert rel_err(net.get_grads(), 2 * rg) <= STRICT_TOL
    # updateGradInput: same dx, no parameter-gradient change
    dx2 = net.updateGradInput(x, dy)
    assert rel_err(dx2, rdx.numpy()) <= STRICT_TOL
    assert rel_err(net.get_grads(), 2 * rg) <= STRICT_TOL
    if net.num_bn_channels():
        rm, rv = net.get_bn_running()
        orm = torch.cat([m.running_mean for m in onet.bn_modules()]).numpy()
        orv = torch.cat([m.running_var for m in onet.bn_modules()]).numpy()
        assert rel_err(rm, orm) <= STRICT_TOL and rel_err(rv, orv) <= STRICT_TOL
    net.close()


STEP_CASES = {
    # name: (config name, ngf/ndf override builder, hr, batch)
    "bce_patch": dict(G=models.train_gray_3_G(4), D=models.patch_D(8), nc=1, hr=8, batch=16,
                      step=dict(family="bce", real_label=1.0, fake_label=0.0, gen_label=1.0), rng=(0.0, 1.0)),
    "mse_rgb": dict(G=models.train_lua_G(3, 4), D=models.dcgan64_D(3, 8), nc=3, hr=64, batch=6,
                    step=dict(family="mse", real_label=0.0, fake_label=0.0, gen_label=0.0, pixel_label=True,
                              pixel_div=4.0 * 3 * 64 * 64), rng=(-1.0, 1.0)),
    "mse_gray": dict(G=models.train_gray_G(4), D=models.dcgan64_D(1, 8), nc=1, hr=64, batch=4,
                     step=dict(family="mse", real_label=0.001, fake_label=0.0, gen_label=0.0, pixel_label=True,
                               pixel_div=64.0 * 64.0), rng=(-1.0, 1.0)),
    "bce_patch_batch": dict(G=models.patch_batch_G(4), D=models.patch_D(8), nc=1, hr=16, batch=8,
                            step=dict(family="bce", real_label=1.0, fake_label=0.0, gen_label=1.0), rng=(0.0, 1.0)),
}


def _build(ctx, case, use_oracle_dtype=torch.float64, paired=False):
    """paired: create D for 2B samples, which makes the fused step run D(real) and D(fake) as ONE grouped pass."""
    nc, hr, B = case["nc"], case["hr"], case["batch"]
    oG, oD = oracle_net(case["G"], 4321, use_oracle_dtype), oracle_net(case["D"], 8765, use_oracle_dtype)
    G = dsr.Sequential.from_specs(case["G"]).cuda(ctx, (nc, hr // 2, hr // 2), B)
    D = dsr.Sequential.from_specs(case["D"]).cuda(ctx, (nc, hr, hr), 2 * B if paired else B)
    G.set_params(oG.get_flat_params().numpy().astype(np.float32))
    D.set_params(oD.get_flat_params().numpy().astype(np.float32))
    return oG, oD, G, D


def _sync_from_oracle(net, onet, st):
    net.set_params(onet.get_flat_params().numpy().astype(np.float32))
    net.set_adam_state(st.m.numpy().astype(np.float32), st.v.numpy().astype(np.float32), st.t)


# Data seeds of the step tests: with ~1e5..1e6 ReLU / LeakyReLU inputs per step some pre-activation can sit within float32
# rounding of the kink, and then ANY two float32 evaluations (the float32 and the float64 oracle included) differ at the
# 1e-4..1e-2 level in that step (scripts/find_step_seeds_gpu.py on mse_rgb: 13 of 16 consecutive seeds show such an event for
# the library's summation order, 1004 / 1005 / 1014 do not).  The seeds below are ones where neither the float32 ORACLE
# (scripts/find_step_seeds.py, CPU) nor the library (scripts/find_step_seeds_gpu.py, B200) has a kink event on any of the four
# iterations, so the strict bound can be demanded of every iteration.
STEP_SEEDS = {"bce_patch": 1033, "bce_patch_batch": 1006, "mse_gray": 1030, "mse_rgb": 1005}


@pytest.mark.parametrize("paired", [False, True])
@pytest.mark.parametrize("name", sorted(STEP_CASES))
def test_train_step_parity(ctx, name, paired):
    """Every step starts from the oracle's exact state (parameters + Adam moments), so each iteration is an independent
    single-step parity check at t = 1..4 (bias correction, accumulated moments).  The strict bound holds on EVERY iteration:
    losses within 1e-5, gradients / Adam moments / parameters within 5e-5 (max-norm over the flat vector).  Parameters: the
    early Adam update is lr * g / (|g| + ~3e-7), so for the elements whose gradient is of that size an absolute gradient error
    dg moves the parameter by lr * dg / 1e-6 -- a 1e-6-class gradient error becomes a 1e-5-class parameter error there, in any
    float32 implementation."""
    case = STEP_CASES[name]
    oG, oD, G, D = _build(ctx, case, paired=paired)
    B, nc, hr = case["batch"], case["nc"], case["hr"]
    ocfg = ostep_cfg(case["step"])
    cfg = dsr.make_step_cfg(**case["step"])
    stG, stD = ostep.new_adam_state(oG), ostep.new_adam_state(oD)
    r = rng(STEP_SEEDS[name])
    report = []
    for it in range(4):
        real = smooth_images(r, (B, nc, hr, hr), *case["rng"])
        trace = {}
        oerr = ostep.train_step(oG, oD, stG, stD, torch.from_numpy(real), ocfg, trace)
        err = dsr.train_step(ctx, G, D, cfg, real)
        for a, b in zip(err, oerr):
            assert abs(a - b) <= 1e-5 * max(abs(b), 1e-3), (name, it, err, oerr)
        eD = rel_err(D.get_grads(), trace["gradD"].numpy())
        eG = rel_err(G.get_grads(), trace["gradG"].numpy())
        pD = rel_err(D.get_params(), oD.get_flat_params().numpy())
        pG = rel_err(G.get_params(), oG.get_flat_params().numpy())
        m, v, t = D.get_adam_state()
        eM = rel_err(m, stD.m.numpy())
        assert t == it + 1
        if D.num_bn_channels():
            rm, rv = D.get_bn_running()
            orm = torch.cat([m.running_mean for m in oD.bn_modules()]).numpy()
            orv = torch.cat([m.running_var for m in oD.bn_modules()]).numpy()
            assert rel_err(rm, orm) <= STRICT_TOL and rel_err(rv, orv) <= STRICT_TOL, (name, it, "BN running statistics")
        report.append((it, eD, eG, pD, pG, eM))
        assert max(eD, eG, eM, pD, pG) <= 5 * STRICT_TOL, (name, report)
        _sync_from_oracle(G, oG, stG)
        _sync_from_oracle(D, oD, stD)
    G.close(); D.close()


@pytest.mark.parametrize("name", sorted(STEP_CASES))
def test_train_steps_free_running(ctx, name):
    """Four consecutive steps WITHOUT re-synchronising from the oracle (drift accumulates through the parameters, the Adam
    moments and D's BN running statistics): losses within 1e-5 at every step, parameters within 5e-5 after the fourth (the
    float32 oracle itself is 1e-6..7e-6 away from the float64 one there, scripts/find_step_seeds.py)."""
    case = STEP_CASES[name]
    oG, oD, G, D = _build(ctx, case, paired=True)
    B, nc, hr = case["batch"], case["nc"], case["hr"]
    ocfg = ostep_cfg(case["step"])
    cfg = dsr.make_step_cfg(**case["step"])
    stG, stD = ostep.new_adam_state(oG), ostep.new_adam_state(oD)
    r = rng(STEP_SEEDS[name])
    for it in range(4):
        real = smooth_images(r, (B, nc, hr, hr), *case["rng"])
        oerr = ostep.train_step(oG, oD, stG, stD, torch.from_numpy(real), ocfg)
        err = dsr.train_step(ctx, G, D, cfg, real)
        for a, b in zip(err, oerr):
            assert abs(a - b) <= 1e-5 * max(abs(b), 1e-3), (name, it, err, oerr)
    pD = rel_err(D.get_params(), oD.get_flat_params().numpy())
    pG = rel_err(G.get_params(), oG.get_flat_params().numpy())
    assert max(pD, pG) <= 5 * STRICT_TOL, (name, pD, pG)
    G.close(); D.close()


def test_stale_activation_step_differs_from_fresh(ctx):
    """F5: fGx reuses the pre-Adam D activations (train.lua:264-270).  A 'fresh' third D forward gives G
    gradients that differ by ~28 %; the library must follow the stale reference semantics."""
    from oracle import ops
    case = STEP_CASES["bce_patch"]
    B, nc, hr = case["batch"], case["nc"], case["hr"]
    ocfg = ostep_cfg(case["step"])
    cfg = dsr.make_step_cfg(**case["step"])
    oG, oD, G, D = _build(ctx, case)
    real = smooth_images(rng(99), (B, nc, hr, hr), 0.0, 1.0)
    trace = {}
    ostep.train_step(oG, oD, ostep.new_adam_state(oG), ostep.new_adam_state(oD), torch.from_numpy(real), ocfg, trace)
    dsr.train_step(ctx, G, D, cfg, real)
    stale = trace["gradG"].numpy()
    fake = trace["fake"]
    out = oD.forward(fake)                                   # fresh variant: post-Adam D forward on fake
    dfdg = oD.update_grad_input(fake, ops.bce_bwd(out, torch.full_like(out, 1.0)))
    oG2 = oracle_net(case["G"], 4321)
    oG2.forward(trace["lr"])
    oG2.zero_grad_parameters()
    oG2.backward(trace["lr"], dfdg)
    fresh = oG2.get_flat_grads().numpy()
    ours = G.get_grads()
    assert rel_err(fresh, stale) > 0.1                       # the two semantics really differ
    # decisive either way; 1e-2 (not 5e-5) because this data seed has a ReLU-kink element (the float32 ORACLE
    # itself is 5e-3 away from the float64 one on it)
    assert rel_err(ours, stale) <= 1e-2 and rel_err(ours, fresh) > 0.1
    G.close(); D.close()


def test_staged_and_graph_step_match_host_step(ctx):
    case = STEP_CASES["bce_patch"]
    B, nc, hr = case["batch"], case["nc"], case["hr"]
    real = smooth_images(rng(3), (B, nc, hr, hr), 0.0, 1.0)
    cfg = dsr.make_step_cfg(**case["step"])
    res = []
    gctx = dsr.Context(device=0, precision="strict", use_graph=True)
    for c, staged in ((ctx, False), (ctx, True), (gctx, True)):
        _, _, G, D = _build(c, case)
        losses = []
        for it in range(3):
            if staged:
                dsr.stage_batch(c, D, real, 0)
                losses.append(dsr.train_step_staged(c, G, D, cfg, 0, B, want_losses=True))
            else:
                losses.append(dsr.train_step(c, G, D, cfg, real))
        res.append((losses, G.get_params(), D.get_params()))
        G.close(); D.close()
    gctx.close()
    for other in res[1:]:
        assert np.array_equal(np.array(res[0][0]), np.array(other[0]))    # deterministic: bit-identical
        assert np.array_equal(res[0][1], other[1]) and np.array_equal(res[0][2], other[2])


def _trajectory(ctx, case, steps):
    oG, oD, G, D = _build(ctx, case, torch.float32, paired=True)
    B, nc, hr = case["batch"], case["nc"], case["hr"]
    ocfg = ostep_cfg(case["step"])
    cfg = dsr.make_step_cfg(**case["step"])
    stG, stD = ostep.new_adam_state(oG), ostep.new_adam_state(oD)
    r = rng(2024)
    pool = [smooth_images(r, (B, nc, hr, hr), *case["rng"]) for _ in range(8)]
    ours, ref = [], []
    for it in range(steps):
        real = pool[it % 8]
        oerr = ostep.train_step(oG, oD, stG, stD, torch.from_numpy(real), ocfg)
        err = dsr.train_step(ctx, G, D, cfg, real)
        ours.append((err[0] + err[1], err[2]))
        ref.append((oerr[0] + oerr[1], oerr[2]))
    G.close(); D.close()
    return np.array(ours), np.array(ref)


def test_loss_trajectory_200_steps(ctx):
    """north_star: Err_D / Err_G trajectories over 200 free-running steps within 1 % of the reference path
    (float32 oracle), on the train-gray.lua graph (the bench workload's graph; its float32-vs-float64 oracle
    spread over 200 steps is 0.25 %, i.e. the comparison is meaningful)."""
    ours, ref = _trajectory(ctx, STEP_CASES["mse_gray"], 200)
    rel = np.abs(ours - ref) / np.maximum(np.abs(ref), 1e-6)
    assert rel.max() <= 0.01, (rel.max(), rel.argmax())


def test_loss_trajectory_chaotic_config(ctx):
    """train-gray-patch.lua graph at toy size: the float32 and float64 ORACLES already drift apart by 10-15 %
    pointwise within 200 steps (BN + ReLU kinks + Adam's sign-like early updates), so pointwise 1 % is not a
    property any float32 implementation can have here.  Checked instead: the first steps pointwise, and the
    200-step mean losses."""
    ours, ref = _trajectory(ctx, STEP_CASES["bce_patch"], 200)
    rel = np.abs(ours - ref) / np.maximum(np.abs(ref), 1e-6)
    assert rel[:5].max() <= 0.01, rel[:5]
    mean_rel = np.abs(ours.mean(axis=0) - ref.mean(axis=0)) / np.abs(ref.mean(axis=0))
    assert mean_rel.max() <= 0.03, mean_rel
