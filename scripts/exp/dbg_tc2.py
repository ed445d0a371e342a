import os, sys
os.environ["DCGANSR_TC2"] = "2"; os.environ["DCGANSR_NO_HALO"] = "1"
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
import dcgan_super_resolution_b200 as dsr
from dcgan_super_resolution_b200 import _lib as L
from util import ptr
ctx = dsr.Context(device=0, precision="tf32")
for (kind, n, cin, h, w, cout, k, s, p) in [("conv", 75, 64, 64, 64, 128, 4, 2, 1), ("conv", 37, 128, 32, 32, 256, 4, 2, 1)]:
    x = np.zeros((n, cin, h, w), np.float32); wt = np.zeros((cout, cin, k, k), np.float32)
    ho = (h + 2 * p - k) // s + 1
    y = np.empty((n, cout, ho, ho), np.float32)
    ctx.profile_begin()
    L.check(ctx.lib.dcgansr_conv2d_fwd(ctx.h, ptr(x), ptr(wt), ptr(y), n, cin, h, w, cout, k, s, p), ctx.h)
    print("fwd", ctx.profile_end())
    ctx.profile_begin()
    L.check(ctx.lib.dcgansr_conv2d_dgrad(ctx.h, ptr(y), ptr(wt), ptr(x), n, cin, h, w, cout, k, s, p), ctx.h)
    print("dgrad", ctx.profile_end())
