"""The DCGAN-SR training step restated on CPU (oracle; test infrastructure only).

Follows ``fDx`` / ``fGx`` / the train loop of the reference:

* ``train.lua:208-253`` (fDx), ``:256-272`` (fGx), ``:275-283`` (order: Adam(D) *then* fGx);
* BCE family: ``train-gray-patch.lua:236-325`` (labels 1 / 0 / 1);
* MSE family: ``train.lua`` (labels 0 / per-sample pixel MSE / 0, divisor 4*C*H*W),
  ``train-gray.lua:232,265,282`` (0.001 / pixel MSE with divisor H*W / 0).

Quirks reproduced: F5 (stale-activation G step: ``netD.output`` and all of D's cached
activations / BN statistics come from the pre-Adam fake forward, but the dgrad through D
uses the post-Adam weights), F6 (two separate D passes with separate BN statistics, grads
accumulated into one zeroed buffer, running stats updated twice).

PARITY UNPINNED: see ``oracle/__init__.py``.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import torch

from . import ops
from .nets import Sequential


@dataclass
class StepCfg:
    family: str = "bce"           # "bce" | "mse"
    real_label: float = 1.0       # target of D(real)
    fake_label: float = 0.0       # target of D(fake) (BCE family; MSE family uses pixel MSE)
    gen_label: float = 1.0        # target of D(fake) in fGx
    pixel_label: bool = False     # MSE family: D(fake) target = per-sample pixel MSE
    pixel_div: float = 1.0        # divisor of the per-sample squared error sum
    lr: float = 2e-4
    beta1: float = 0.5
    beta2: float = 0.999
    eps: float = 1e-8
    # data parallel emulation: the *global* element count the criteria average over.
    # None = local (single rank).
    global_scale: float | None = None


@dataclass
class AdamState:
    m: torch.Tensor
    v: torch.Tensor
    t: int = 0


def new_adam_state(net: Sequential) -> AdamState:
    n = net.num_params()
    return AdamState(torch.zeros(n, dtype=net.dtype), torch.zeros(n, dtype=net.dtype), 0)


def _crit(family, out, label):
    if family == "bce":
        return ops.bce_fwd(out, label), ops.bce_bwd(out, label)
    return ops.mse_fwd(out, label), ops.mse_bwd(out, label)


def _labels_like(out, value_or_vec, batch):
    """``label`` has one entry per element of D's output (SURVEY 8(d) shape decisions):
    a constant fill, or a per-sample value replicated over D's h x w outputs."""
    n = out.numel()
    if torch.is_tensor(value_or_vec):
        per = n // batch
        return value_or_vec.to(out.dtype).reshape(batch, 1).expand(batch, per).reshape(out.shape)
    return torch.full_like(out, float(value_or_vec))


def adam_apply(net: Sequential, st: AdamState, cfg: StepCfg):
    p = net.get_flat_params()
    g = net.get_flat_grads()
    st.t = ops.adam_step(p, g, st.m, st.v, st.t, cfg.lr, cfg.beta1, cfg.beta2, cfg.eps)
    net.set_flat_params(p)


def train_step(netG: Sequential, netD: Sequential, stG: AdamState, stD: AdamState,
               real: torch.Tensor, cfg: StepCfg, trace: dict | None = None):
    """One iteration of the reference loop.  Returns (errD_real, errD_fake, errG)."""
    B = real.shape[0]
    real = real.to(netD.dtype)
    margin = [float("inf")]

    def probe(net):
        if trace is not None:
            margin[0] = min(margin[0], net.kink_margin())

    # ---------------- fDx (train.lua:208-253) ----------------
    netD.zero_grad_parameters()
    out = netD.forward(real)
    probe(netD)
    lab = _labels_like(out, cfg.real_label, B)
    errD_real, df_do = _crit(cfg.family, out, lab)
    netD.backward(real, df_do)

    lr_img = ops.avgpool2_fwd(real)                      # train.lua:225-230
    fake = netG.forward(lr_img)                          # :233-234
    probe(netG)

    if cfg.pixel_label:
        pm = ops.pixel_mse_per_sample(real, fake, cfg.pixel_div)   # :237-239
        fake_target = pm
    else:
        fake_target = cfg.fake_label

    out = netD.forward(fake)                             # :242-243 (cached acts = FAKE pass)
    probe(netD)
    lab = _labels_like(out, fake_target, B)
    errD_fake, df_do = _crit(cfg.family, out, lab)
    netD.backward(fake, df_do)                           # grads accumulate (F6)
    if trace is not None:
        trace["gradD"] = netD.get_flat_grads().clone()
        trace["fake"] = fake.clone()
        trace["lr"] = lr_img.clone()
        trace["outD_fake"] = out.clone()

    adam_apply(netD, stD, cfg)                           # optim.adam(fDx, ...) :280

    # ---------------- fGx (train.lua:256-272) ----------------
    netG.zero_grad_parameters()
    out = netD.output                                    # STALE pre-Adam output (F5)
    lab = _labels_like(out, cfg.gen_label, B)
    errG, df_do = _crit(cfg.family, out, lab)
    df_dg = netD.update_grad_input(fake, df_do)          # post-Adam weights, stale acts
    netG.backward(lr_img, df_dg)
    if trace is not None:
        trace["gradG"] = netG.get_flat_grads().clone()
        trace["df_dg"] = df_dg.clone()
        trace["kink_margin"] = margin[0]

    adam_apply(netG, stG, cfg)                           # optim.adam(fGx, ...) :283
    return errD_real, errD_fake, errG
