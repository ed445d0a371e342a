"""GPU diagnostic: FAST_TF32 nets vs the float64 oracle, exact and with TF32-emulated operands (trunc / rna)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import contextlib
import dcgan_super_resolution_b200 as dsr
from dcgan_super_resolution_b200 import models
from util import oracle_net, rng, t64, rel_err, l2_err, tf32_oracle, trunc_tf32, rna_tf32

NETS = {
    "train_lua_G": (models.train_lua_G(3, 8), (3, 8, 8)),
    "train_gray_G": (models.train_gray_G(16), (1, 8, 8)),
    "train_gray_3_G": (models.train_gray_3_G(8), (1, 4, 4)),
    "dcgan64_D": (models.dcgan64_D(3, 16), (3, 64, 64)),
    "patch_D": (models.patch_D(16), (1, 8, 8)),
}
ctx = dsr.Context(device=0, precision="tf32")
for name, (specs, ishape) in NETS.items():
    B = 4
    r = rng(1234)
    x = r.uniform(-1, 1, (B,) + ishape).astype(np.float32)
    net = dsr.Sequential.from_specs(specs).cuda(ctx, ishape, B)
    onet = oracle_net(specs, seed=4321)
    net.set_params(onet.get_flat_params().numpy().astype(np.float32))
    y = net.forward(x)
    dy = r.standard_normal(y.shape).astype(np.float32)
    net.zeroGradParameters()
    dx = net.backward(x, dy)
    g = net.get_grads()
    for mode, cm in (("exact", contextlib.nullcontext()), ("trunc", tf32_oracle(trunc_tf32)), ("rna", tf32_oracle(rna_tf32))):
        with cm:
            onet = oracle_net(specs, seed=4321)
            ry = onet.forward(t64(x))
            onet.zero_grad_parameters()
            rdx = onet.backward(t64(x), t64(dy).reshape(ry.shape))
            rg = onet.get_flat_grads().numpy()
        print(f"{name:16s} vs {mode:5s}: y {rel_err(y.reshape(-1), ry.numpy().reshape(-1)):.2e} dx {rel_err(dx, rdx.numpy()):.2e} (l2 {l2_err(dx, rdx.numpy()):.2e}) "
              f"grads {rel_err(g, rg):.2e} (l2 {l2_err(g, rg):.2e})")
    net.close()
ctx.close()
