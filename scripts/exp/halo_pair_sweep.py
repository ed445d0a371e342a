"""A/B/C of the weights-resident halo kernel within one process: DCGANSR_HALO_PAIR=0 (single CTA), 1 (policy), 2 (pairs wherever possible)."""
import ctypes, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import dcgan_super_resolution_b200 as dsr
from dcgan_super_resolution_b200 import _lib as L
ctx = dsr.Context(device=0, precision="tf32")
LAYERS = [  # full, cin, h, cout, k, s, p, batch
    (1, 64, 128, 32, 4, 2, 1, 64), (0, 32, 256, 16, 4, 2, 1, 64),                                      # C2 G
    (1, 96, 128, 48, 4, 2, 1, 128), (1, 48, 256, 24, 4, 2, 1, 128), (0, 24, 512, 12, 4, 2, 1, 128),   # C3b G
    (1, 96, 64, 48, 4, 2, 1, 128), (1, 48, 128, 24, 4, 2, 1, 128),                                     # C3a G
    (1, 64, 64, 32, 4, 2, 1, 64), (1, 32, 128, 16, 4, 2, 1, 64), (0, 16, 256, 32, 4, 2, 1, 64), (0, 32, 128, 64, 4, 2, 1, 64),  # C4 G
]
for rep in range(2):
    for (full, cin, h, cout, k, s, p, n) in LAYERS:
        row = []
        for what in (0, 1):
            for mode in ("0", "1", "2"):
                os.environ["DCGANSR_HALO_PAIR"] = mode
                ms = ctypes.c_float()
                L.check(ctx.lib.dcgansr_bench_conv(ctx.h, full, what, n, cin, h, h, cout, k, s, p, 5, ctypes.byref(ms)), ctx.h)
                row.append(ms.value * 1e3)
        print(f"{'FC' if full else 'C '} {cin:3d}->{cout:3d} {h:3d} n={n:3d}  fwd single {row[0]:7.1f} policy {row[1]:7.1f} pairs {row[2]:7.1f}   dgrad single {row[3]:7.1f} policy {row[4]:7.1f} pairs {row[5]:7.1f}", flush=True)
ctx.close()
