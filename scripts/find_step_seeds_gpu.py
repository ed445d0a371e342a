"""GPU side of the step-test seed selection (see scripts/find_step_seeds.py): for candidate data seeds, run the library against
the float64 oracle exactly as tests/test_gpu_step.py does and print the worst error per seed, so the committed seeds are ones
on which no ReLU / LeakyReLU input sits within float32 rounding of its kink for the library's summation order either."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import dcgan_super_resolution_b200 as dsr
from oracle import step as ostep
from util import ostep_cfg, rel_err, rng, smooth_images
from test_gpu_step import STEP_CASES, _build, _sync_from_oracle

cases = sys.argv[1].split(",") if len(sys.argv) > 1 else sorted(STEP_CASES)
lo, hi = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1000, 1030)
ctx = dsr.Context(device=0, precision="strict")
for name in cases:
    case = STEP_CASES[name]
    B, nc, hr = case["batch"], case["nc"], case["hr"]
    ocfg, cfg = ostep_cfg(case["step"]), dsr.make_step_cfg(**case["step"])
    for seed in range(lo, hi):
        worst = 0.0
        for paired in (False, True):
            oG, oD, G, D = _build(ctx, case, paired=paired)
            stG, stD = ostep.new_adam_state(oG), ostep.new_adam_state(oD)
            r = rng(seed)
            for it in range(4):
                real = smooth_images(r, (B, nc, hr, hr), *case["rng"])
                trace = {}
                oerr = ostep.train_step(oG, oD, stG, stD, torch.from_numpy(real), ocfg, trace)
                err = dsr.train_step(ctx, G, D, cfg, real)
                el = max(abs(a - b) / max(abs(b), 1e-3) for a, b in zip(err, oerr))
                e = max(rel_err(D.get_grads(), trace["gradD"].numpy()), rel_err(G.get_grads(), trace["gradG"].numpy()),
                        rel_err(D.get_params(), oD.get_flat_params().numpy()), rel_err(G.get_params(), oG.get_flat_params().numpy()), 5 * el)
                worst = max(worst, e)
                _sync_from_oracle(G, oG, stG)
                _sync_from_oracle(D, oD, stD)
            G.close(); D.close()
        # free-running, paired
        oG, oD, G, D = _build(ctx, case, paired=True)
        stG, stD = ostep.new_adam_state(oG), ostep.new_adam_state(oD)
        r = rng(seed)
        fr = 0.0
        for it in range(4):
            real = smooth_images(r, (B, nc, hr, hr), *case["rng"])
            oerr = ostep.train_step(oG, oD, stG, stD, torch.from_numpy(real), ocfg)
            err = dsr.train_step(ctx, G, D, cfg, real)
            fr = max(fr, 5 * max(abs(a - b) / max(abs(b), 1e-3) for a, b in zip(err, oerr)))
        fr = max(fr, rel_err(D.get_params(), oD.get_flat_params().numpy()), rel_err(G.get_params(), oG.get_flat_params().numpy()))
        G.close(); D.close()
        print(f"{name} seed {seed} resync_worst {worst:.2e} free_worst {fr:.2e}", flush=True)
ctx.close()
