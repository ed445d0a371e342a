"""CPU experiment (test-infrastructure side): how far does a *correct* TF32 implementation drift from the float64 oracle
on whole nets?  The float64 oracle is re-run with every convolution operand rounded to TF32 (activations truncated like
the tensor core does, weights round-to-nearest like pack_taps_tc does); the difference to the unrounded oracle is the
conditioning floor the FAST_TF32 net-level tests have to allow for."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
from oracle import ops
from dcgan_super_resolution_b200 import models
from util import oracle_net, rng, t64, rel_err


def trunc_tf32(x):
    a = x.to(torch.float32).contiguous().view(torch.int32)
    return (a & ~0x1FFF).view(torch.float32).to(x.dtype)


def rna_tf32(x):
    a = x.to(torch.float32).contiguous().view(torch.int32)
    return ((a + 0x1000) & ~0x1FFF).view(torch.float32).to(x.dtype)


_orig = {k: getattr(ops, k) for k in ("conv2d_fwd", "conv2d_dgrad", "conv2d_wgrad", "fullconv2d_fwd", "fullconv2d_dgrad", "fullconv2d_wgrad")}


def patch(act_round):
    ops.conv2d_fwd = lambda x, w, s, p: _orig["conv2d_fwd"](act_round(x), rna_tf32(w), s, p)
    ops.conv2d_dgrad = lambda dy, w, xs, s, p: _orig["conv2d_dgrad"](act_round(dy), rna_tf32(w), xs, s, p)
    ops.conv2d_wgrad = lambda x, dy, ws, s, p: _orig["conv2d_wgrad"](act_round(x), act_round(dy), ws, s, p)
    ops.fullconv2d_fwd = lambda x, w, s, p, adj=0: _orig["fullconv2d_fwd"](act_round(x), rna_tf32(w), s, p, adj)
    ops.fullconv2d_dgrad = lambda dy, w, s, p: _orig["fullconv2d_dgrad"](act_round(dy), rna_tf32(w), s, p)
    ops.fullconv2d_wgrad = lambda x, dy, ws, s, p: _orig["fullconv2d_wgrad"](act_round(x), act_round(dy), ws, s, p)


def unpatch():
    for k, v in _orig.items():
        setattr(ops, k, v)


def l2(a, b):
    a = np.asarray(a, np.float64).ravel(); b = np.asarray(b, np.float64).ravel()
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


NETS = {
    "train_lua_G": (models.train_lua_G(3, 8), (3, 8, 8)),
    "train_gray_G": (models.train_gray_G(16), (1, 8, 8)),
    "train_gray_3_G": (models.train_gray_3_G(8), (1, 4, 4)),
    "dcgan64_D": (models.dcgan64_D(3, 16), (3, 64, 64)),
    "patch_D": (models.patch_D(16), (1, 8, 8)),
}

for name, (specs, ishape) in NETS.items():
    B = 4
    res = {}
    for mode in ("exact", "trunc", "rna"):
        if mode != "exact":
            patch(trunc_tf32 if mode == "trunc" else rna_tf32)
        onet = oracle_net(specs, seed=4321)
        r = rng(1234)
        x = r.uniform(-1, 1, (B,) + ishape).astype(np.float32)
        y = onet.forward(t64(x))
        dy = r.standard_normal(tuple(y.shape)).astype(np.float32)
        onet.zero_grad_parameters()
        dx = onet.backward(t64(x), t64(dy).reshape(y.shape))
        res[mode] = (y.numpy().copy(), dx.numpy().copy(), onet.get_flat_grads().numpy().copy())
        unpatch()
    for mode in ("trunc", "rna"):
        print(f"{name:16s} {mode:5s} max-norm: y {rel_err(res[mode][0], res['exact'][0]):.2e} dx {rel_err(res[mode][1], res['exact'][1]):.2e} "
              f"grads {rel_err(res[mode][2], res['exact'][2]):.2e} | L2: y {l2(res[mode][0], res['exact'][0]):.2e} dx {l2(res[mode][1], res['exact'][1]):.2e} "
              f"grads {l2(res[mode][2], res['exact'][2]):.2e}")
