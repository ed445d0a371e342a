"""GPU diagnostic: per-iteration, per-tensor error of the fused step vs the float64 oracle."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import dcgan_super_resolution_b200 as dsr
from oracle import step as ostep
from util import *
import test_gpu_step as T

ctx = dsr.Context(0, "strict")
names = sys.argv[1:] or ["mse_rgb"]
for name in names:
    case = T.STEP_CASES[name]
    oG, oD, G, D = T._build(ctx, case)
    B, nc, hr = case["batch"], case["nc"], case["hr"]
    ocfg = ostep_cfg(case["step"]); cfg = dsr.make_step_cfg(**case["step"])
    stG, stD = ostep.new_adam_state(oG), ostep.new_adam_state(oD)
    r = rng(1234)
    def per_tensor(net, a, b):
        off = 0; out = []
        for (nm, p, g), spec in zip(net.param_list(), range(10**6)):
            n = p.numel(); out.append((nm, tuple(p.shape), float(rel_err(a[off:off+n], b[off:off+n])), float(np.max(np.abs(b[off:off+n]))))); off += n
        return out
    for it in range(4):
        real = smooth_images(r, (B, nc, hr, hr), *case["rng"])
        tr = {}
        oerr = ostep.train_step(oG, oD, stG, stD, torch.from_numpy(real), ocfg, tr)
        err = dsr.train_step(ctx, G, D, cfg, real)
        print(name, it, "loss", err, oerr)
        print("  gradD", rel_err(D.get_grads(), tr["gradD"].numpy()), "gradG", rel_err(G.get_grads(), tr["gradG"].numpy()),
              "pD", rel_err(D.get_params(), oD.get_flat_params().numpy()), "pG", rel_err(G.get_params(), oG.get_flat_params().numpy()))
        for row in per_tensor(oG, G.get_grads(), tr["gradG"].numpy()): print("   gG", row)
        for row in per_tensor(oD, D.get_grads(), tr["gradD"].numpy()): print("   gD", row)
        for row in per_tensor(oD, D.get_params(), oD.get_flat_params().numpy()): print("   pD", row)
