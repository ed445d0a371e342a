"""A/B of the wide-tile kernel's single-CTA 128 x 256 items (DCGANSR_TC2_PAIR=0) against CTA pairs (cta_group::2, forced) per layer / batch."""
import ctypes, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import dcgan_super_resolution_b200 as dsr
from dcgan_super_resolution_b200 import _lib as L
ctx = dsr.Context(device=0, precision="tf32")
LAYERS = [  # full, cin, h, cout, k, s, p, batches
    (0, 128, 32, 256, 4, 2, 1, (64, 128, 256)), (0, 256, 16, 512, 4, 2, 1, (64, 128, 256)),     # D at 128^2 input (C3b)
    (0, 128, 128, 256, 4, 2, 1, (16, 64)), (1, 256, 64, 128, 4, 2, 1, (16, 64)),                 # C1b G: C 128->256 fwd, FC 256->128 dgrad
    (1, 1024, 256, 512, 4, 2, 1, (2,)), (1, 512, 512, 256, 4, 2, 1, (1,)),                       # C5 G (micro-batch slices)
    (0, 128, 28, 256, 3, 1, 0, (64, 128)),
]
os.environ["DCGANSR_TC2"] = "2"
for (full, cin, h, cout, k, s, p, batches) in LAYERS:
    for n in batches:
        row = []
        for what in (0, 1):
            for mode in ("0", "2"):
                os.environ["DCGANSR_TC2_PAIR"] = mode
                ms = ctypes.c_float()
                L.check(ctx.lib.dcgansr_bench_conv(ctx.h, full, what, n, cin, h, h, cout, k, s, p, 5, ctypes.byref(ms)), ctx.h)
                row.append(ms.value * 1e3)
        ho = (h - 1) * s - 2 * p + k if full else (h + 2 * p - k) // s + 1
        gf = 2.0 * n * (h * h if full else ho * ho) * cin * cout * k * k / 1e9
        print(f"{'FC' if full else 'C '} {cin:4d}->{cout:4d} {h:3d} k{k}s{s} n={n:4d} {gf:8.1f} GF  fwd single {row[0]:8.1f} pair {row[1]:8.1f} ({row[0] / row[1]:4.2f}x, {gf / row[1] / 1e3:5.0f} TF/s)   "
              f"dgrad single {row[2]:8.1f} pair {row[3]:8.1f} ({row[2] / row[3]:4.2f}x, {gf / row[3] / 1e3:5.0f} TF/s)", flush=True)
ctx.close()
