"""Generates tests/golden/*.npz from the CPU oracle (float64), seeded.  PARITY UNPINNED: the reference
(Lua/Torch7) cannot run here and ships no vectors, so these pin the ORACLE (and, through the GPU tests,
the library) against regressions -- they are not Torch7 outputs.   Run: python tests/golden/make_golden.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from dcgan_super_resolution_b200 import models  # noqa: E402  (layer specs only: pure data)
from oracle import ops, step as ostep  # noqa: E402
from util import oracle_net, ostep_cfg, rng, smooth_images, t64  # noqa: E402

LAYER_CASES = {
    # name: (kind, n, cin, h, w, cout, k, s, p)
    "conv_d2": ("conv", 2, 16, 8, 8, 32, 4, 2, 1),
    "conv_patch": ("conv", 3, 8, 6, 6, 16, 3, 1, 0),
    "conv_rgb_in": ("conv", 2, 3, 8, 8, 8, 4, 2, 1),
    "full_g2": ("fullconv", 2, 16, 4, 4, 8, 4, 2, 1),
    "full_gray_in": ("fullconv", 2, 1, 4, 4, 8, 4, 2, 1),
}

STEP_CASES = {
    "bce_patch": dict(G=models.train_gray_3_G(4), D=models.patch_D(8), nc=1, hr=8, batch=16,
                      step=dict(family="bce", real_label=1.0, fake_label=0.0, gen_label=1.0), rng=(0.0, 1.0)),
    "mse_gray": dict(G=models.train_gray_G(4), D=models.dcgan64_D(1, 8), nc=1, hr=64, batch=4,
                     step=dict(family="mse", real_label=0.001, fake_label=0.0, gen_label=0.0, pixel_label=True,
                               pixel_div=64.0 * 64.0), rng=(-1.0, 1.0)),
}
N_STEPS = 3


def layer_inputs(name):
    kind, n, cin, h, w, cout, k, s, p = LAYER_CASES[name]
    r = rng(sum(map(ord, name)))
    x = r.standard_normal((n, cin, h, w)).astype(np.float32)
    wshape = (cout, cin, k, k) if kind == "conv" else (cin, cout, k, k)
    wt = (0.1 * r.standard_normal(wshape)).astype(np.float32)
    if kind == "conv":
        ho, wo = (h + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1
    else:
        ho, wo = (h - 1) * s - 2 * p + k, (w - 1) * s - 2 * p + k
    dy = r.standard_normal((n, cout, ho, wo)).astype(np.float32)
    return x, wt, dy


def layer_outputs(name):
    kind, n, cin, h, w, cout, k, s, p = LAYER_CASES[name]
    x, wt, dy = layer_inputs(name)
    X, W, DY = t64(x), t64(wt), t64(dy)
    if kind == "conv":
        return ops.conv2d_fwd(X, W, s, p).numpy(), ops.conv2d_dgrad(DY, W, X.shape, s, p).numpy(), \
            ops.conv2d_wgrad(X, DY, W.shape, s, p).numpy()
    return ops.fullconv2d_fwd(X, W, s, p).numpy(), ops.fullconv2d_dgrad(DY, W, s, p).numpy(), \
        ops.fullconv2d_wgrad(X, DY, W.shape, s, p).numpy()


def misc_inputs():
    r = rng(77)
    return dict(bn_x=(r.standard_normal((4, 8, 5, 5)) * 2 + 0.5).astype(np.float32),
                bn_gamma=(1 + 0.02 * r.standard_normal(8)).astype(np.float32),
                bn_beta=(0.1 * r.standard_normal(8)).astype(np.float32),
                bn_dy=r.standard_normal((4, 8, 5, 5)).astype(np.float32),
                crit_x=r.uniform(0.02, 0.98, 96).astype(np.float32),
                crit_t=(r.random(96) > 0.5).astype(np.float32),
                adam_p=r.standard_normal(257).astype(np.float32), adam_g=r.standard_normal(257).astype(np.float32),
                adam_m=(0.1 * r.standard_normal(257)).astype(np.float32), adam_v=(0.01 * r.random(257)).astype(np.float32))


def misc_outputs():
    i = misc_inputs()
    y, mean, invstd, rm, rv = ops.bn_fwd_train(t64(i["bn_x"]), t64(i["bn_gamma"]), t64(i["bn_beta"]), torch.zeros(8, dtype=torch.float64),
                                               torch.ones(8, dtype=torch.float64))
    dx, dg, db = ops.bn_bwd(t64(i["bn_x"]), t64(i["bn_dy"]), t64(i["bn_gamma"]), mean, invstd)
    p, m, v = t64(i["adam_p"]).clone(), t64(i["adam_m"]).clone(), t64(i["adam_v"]).clone()
    ops.adam_step(p, t64(i["adam_g"]), m, v, 3)
    return dict(bn_y=y.numpy(), bn_mean=mean.numpy(), bn_invstd=invstd.numpy(), bn_rm=rm.numpy(), bn_rv=rv.numpy(),
                bn_dx=dx.numpy(), bn_dgamma=dg.numpy(), bn_dbeta=db.numpy(),
                bce=np.array([ops.bce_fwd(t64(i["crit_x"]), t64(i["crit_t"]))]), bce_dx=ops.bce_bwd(t64(i["crit_x"]), t64(i["crit_t"])).numpy(),
                mse=np.array([ops.mse_fwd(t64(i["crit_x"]), t64(i["crit_t"]))]), mse_dx=ops.mse_bwd(t64(i["crit_x"]), t64(i["crit_t"])).numpy(),
                adam_p=p.numpy(), adam_m=m.numpy(), adam_v=v.numpy())


def step_batches(name):
    case = STEP_CASES[name]
    r = rng(1234)
    return [smooth_images(r, (case["batch"], case["nc"], case["hr"], case["hr"]), *case["rng"]) for _ in range(N_STEPS)]


def step_outputs(name, dtype=torch.float64):
    case = STEP_CASES[name]
    oG, oD = oracle_net(case["G"], 4321, dtype), oracle_net(case["D"], 8765, dtype)
    stG, stD = ostep.new_adam_state(oG), ostep.new_adam_state(oD)
    cfg = ostep_cfg(case["step"])
    losses = []
    for real in step_batches(name):
        losses.append(ostep.train_step(oG, oD, stG, stD, torch.from_numpy(real), cfg))
    return dict(losses=np.array(losses, np.float64), pG=oG.get_flat_params().double().numpy(), pD=oD.get_flat_params().double().numpy())


def main():
    out = {}
    for name in LAYER_CASES:
        y, dx, dw = layer_outputs(name)
        out[name + ".y"], out[name + ".dx"], out[name + ".dw"] = y, dx, dw
    np.savez_compressed(os.path.join(HERE, "layers.npz"), **out)
    np.savez_compressed(os.path.join(HERE, "misc.npz"), **misc_outputs())
    for name in STEP_CASES:
        np.savez_compressed(os.path.join(HERE, f"step_{name}.npz"), **step_outputs(name))
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main()
