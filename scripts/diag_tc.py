"""GPU diagnostic: tcgen05 tapconv (FAST_TF32) vs the float64 oracle on a list of shapes."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import dcgan_super_resolution_b200 as dsr
from dcgan_super_resolution_b200 import _lib as L
from oracle import ops
from util import *

ctx = dsr.Context(0, "tf32")
lib, h = ctx.lib, ctx.h
def run(kind, what, n, cin, hh, w, cout, k, s, p):
    r = rng(1)
    x = r.standard_normal((n, cin, hh, w)).astype(np.float32)
    full = kind == "full"
    wt = (0.1 * r.standard_normal((cin, cout, k, k) if full else (cout, cin, k, k))).astype(np.float32)
    ho, wo = ((hh - 1) * s - 2 * p + k, (w - 1) * s - 2 * p + k) if full else ((hh + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1)
    dy = r.standard_normal((n, cout, ho, wo)).astype(np.float32)
    X, W, DY = t64(x), t64(wt), t64(dy)
    pre = "dcgansr_fullconv2d_" if full else "dcgansr_conv2d_"
    if what == "fwd":
        out = np.empty((n, cout, ho, wo), np.float32)
        L.check(getattr(lib, pre + "fwd")(h, ptr(x), ptr(wt), ptr(out), n, cin, hh, w, cout, k, s, p), h)
        ref = (ops.fullconv2d_fwd(X, W, s, p) if full else ops.conv2d_fwd(X, W, s, p)).numpy()
    elif what == "wgrad":
        out = np.empty_like(wt)
        L.check(getattr(lib, pre + "wgrad")(h, ptr(x), ptr(dy), ptr(out), n, cin, hh, w, cout, k, s, p), h)
        ref = (ops.fullconv2d_wgrad(X, DY, wt.shape, s, p) if full else ops.conv2d_wgrad(X, DY, wt.shape, s, p)).numpy()
    else:
        out = np.empty_like(x)
        L.check(getattr(lib, pre + "dgrad")(h, ptr(dy), ptr(wt), ptr(out), n, cin, hh, w, cout, k, s, p), h)
        ref = (ops.fullconv2d_dgrad(DY, W, s, p) if full else ops.conv2d_dgrad(DY, W, x.shape, s, p)).numpy()
    e = rel_err(out, ref)
    print(f"{kind:5s} {what:5s} n{n} ci{cin} {hh}x{w} co{cout} k{k}s{s}p{p}: rel_err {e:.3e}", "OK" if e < 2e-3 else "FAIL", flush=True)

cases = [
    ("conv", 2, 32, 8, 8, 32, 3, 1, 1),       # si=1, KB=32
    ("conv", 2, 64, 32, 32, 128, 4, 2, 1),    # si=2 (5-D view)
    ("conv", 4, 64, 6, 6, 128, 3, 1, 0),
    ("conv", 2, 16, 16, 16, 32, 4, 2, 1),     # KB=16
    ("conv", 2, 8, 16, 16, 16, 4, 2, 1),      # KB=8
    ("conv", 3, 512, 4, 4, 1, 4, 1, 0),       # Co=1
    ("conv", 2, 24, 10, 14, 12, 4, 2, 1),     # KB=8, Co=12, ragged
    ("full", 2, 64, 16, 16, 32, 4, 2, 1),
    ("full", 2, 96, 8, 8, 48, 4, 2, 1),
    ("full", 3, 48, 6, 10, 24, 4, 2, 1),
    ("conv", 2, 128, 16, 16, 256, 4, 2, 1),   # 2 N tiles
]
for c in cases:
    for what in (sys.argv[1:] or ("fwd", "dgrad", "wgrad")):
        run(c[0], what, *c[1:])
