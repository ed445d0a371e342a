"""CPU experiment: 200 free-running steps of the float32 oracle with TF32-rounded conv operands vs the plain float32 oracle."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.dirname(__file__))
from oracle import step as ostep
from util import oracle_net, rng, smooth_images, ostep_cfg
from dcgan_super_resolution_b200 import models
import tf32_conditioning as tc  # noqa (runs the net table once; cheap)

case = dict(G=models.train_gray_G(4), D=models.dcgan64_D(1, 8), nc=1, hr=64, batch=4,
            step=dict(family="mse", real_label=0.001, fake_label=0.0, gen_label=0.0, pixel_label=True, pixel_div=64.0 * 64.0), rng=(-1.0, 1.0))


def run(mode, steps=200):
    if mode != "exact":
        tc.patch(tc.trunc_tf32 if mode == "trunc" else tc.rna_tf32)
    oG, oD = oracle_net(case["G"], 4321, torch.float32), oracle_net(case["D"], 8765, torch.float32)
    stG, stD = ostep.new_adam_state(oG), ostep.new_adam_state(oD)
    r = rng(2024)
    B, nc, hr = case["batch"], case["nc"], case["hr"]
    pool = [smooth_images(r, (B, nc, hr, hr), *case["rng"]) for _ in range(8)]
    out = []
    for it in range(steps):
        e = ostep.train_step(oG, oD, stG, stD, torch.from_numpy(pool[it % 8]), ostep_cfg(case["step"]))
        out.append((e[0] + e[1], e[2]))
    tc.unpatch()
    return np.array(out)

t0 = time.time()
ref = run("exact")
print("time", time.time() - t0)
for mode in ("trunc", "rna"):
    o = run(mode)
    rel = np.abs(o - ref) / np.maximum(np.abs(ref), 1e-6)
    print(mode, "max rel", rel.max(), "at", rel.argmax(), "first>1%:", np.argmax(rel.max(axis=1) > 0.01), "mean-rel", np.abs(o.mean(0) - ref.mean(0)) / np.abs(ref.mean(0)),
          "median rel", np.median(rel), "p90", np.quantile(rel, 0.9))
