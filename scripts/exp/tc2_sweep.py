"""A/B of the one-tile per-tap kernel (DCGANSR_TC2=0) against the persistent wide-tile kernel (forced, DCGANSR_TC2=2) per layer and batch."""
import ctypes, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import dcgan_super_resolution_b200 as dsr
from dcgan_super_resolution_b200 import _lib as L
ctx = dsr.Context(device=0, precision="tf32")
LAYERS = [  # full, cin, h, cout, k, s, p
    (0, 64, 32, 128, 4, 2, 1), (0, 128, 16, 256, 4, 2, 1), (0, 256, 8, 512, 4, 2, 1),        # D at 64^2 input (C2 / C3a)
    (0, 64, 64, 128, 4, 2, 1), (0, 128, 32, 256, 4, 2, 1), (0, 256, 16, 512, 4, 2, 1),       # D at 128^2 input (C3b)
    (0, 64, 30, 128, 3, 1, 0), (0, 128, 28, 256, 3, 1, 0),                                    # patch-D (C1b / C4)
    (1, 64, 32, 32, 4, 2, 1), (0, 32, 128, 64, 4, 2, 1),                                      # C4 G inner layers (ngf 16)
]
for (full, cin, h, cout, k, s, p) in LAYERS:
    for n in (64, 128, 256, 512):
        row = []
        for what in (0, 1):
            for mode in ("0", "2"):
                os.environ["DCGANSR_TC2"] = mode
                ms = ctypes.c_float()
                L.check(ctx.lib.dcgansr_bench_conv(ctx.h, full, what, n, cin, h, h, cout, k, s, p, 10, ctypes.byref(ms)), ctx.h)
                row.append(ms.value * 1e3)
        print(f"{'FC' if full else 'C '} {cin:4d}->{cout:4d} {h:3d} k{k}s{s} n={n:4d}  fwd old {row[0]:7.1f} wide {row[1]:7.1f} ({row[0] / row[1]:4.2f}x)   "
              f"dgrad old {row[2]:7.1f} wide {row[3]:7.1f} ({row[2] / row[3]:4.2f}x)", flush=True)
ctx.close()
