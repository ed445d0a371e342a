"""Data-seed search for tests/test_gpu_step.py::test_train_step_parity (CPU, oracle only).

A training step of these nets has ~1e5..1e6 ReLU / LeakyReLU inputs; when one of them lies within float32 rounding of the
kink, ANY two float32 evaluations differ at the 1e-3 level in that step.  The strict-mode step test must hold the 1e-5
class bound on EVERY iteration, so it uses data seeds for which the float32 ORACLE itself stays within the bound of the
float64 oracle on all four iterations (state re-synchronised after each step, as in the test), with the largest kink margin.
Prints the chosen seed per case."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
from oracle import step as ostep
from util import oracle_net, ostep_cfg, rel_err, rng, smooth_images
from test_gpu_step import STEP_CASES

torch.set_num_threads(8)
for name, case in sorted(STEP_CASES.items()):
    B, nc, hr = case["batch"], case["nc"], case["hr"]
    best = None
    for seed in range(1000, 1040):
        o64 = (oracle_net(case["G"], 4321), oracle_net(case["D"], 8765))
        o32 = (oracle_net(case["G"], 4321, torch.float32), oracle_net(case["D"], 8765, torch.float32))
        st64 = (ostep.new_adam_state(o64[0]), ostep.new_adam_state(o64[1]))
        st32 = (ostep.new_adam_state(o32[0]), ostep.new_adam_state(o32[1]))
        r = rng(seed)
        worst, margin = 0.0, float("inf")
        for it in range(4):
            real = smooth_images(r, (B, nc, hr, hr), *case["rng"])
            t64, t32 = {}, {}
            ostep.train_step(o64[0], o64[1], st64[0], st64[1], torch.from_numpy(real), ostep_cfg(case["step"]), t64)
            ostep.train_step(o32[0], o32[1], st32[0], st32[1], torch.from_numpy(real), ostep_cfg(case["step"]), t32)
            eD = rel_err(t32["gradD"].numpy(), t64["gradD"].numpy())
            eG = rel_err(t32["gradG"].numpy(), t64["gradG"].numpy())
            pD = rel_err(o32[1].get_flat_params().numpy(), o64[1].get_flat_params().numpy())
            pG = rel_err(o32[0].get_flat_params().numpy(), o64[0].get_flat_params().numpy())
            worst = max(worst, eD, eG, pD, pG)
            margin = min(margin, t64["kink_margin"])
            for k in range(2):          # re-sync the float32 side from the float64 state
                o32[k].set_flat_params(o64[k].get_flat_params().to(torch.float32))
                st32[k].m, st32[k].v, st32[k].t = st64[k].m.to(torch.float32), st64[k].v.to(torch.float32), st64[k].t
        if worst <= 1e-5 and (best is None or margin > best[1]):
            best = (seed, margin, worst)
    print(name, "seed", best)
