"""A/B of wgrad_tc (DCGANSR_WGRAD_PAIR=0) against the CTA-pair wgrad kernel (forced) per layer / batch, in one process."""
import ctypes, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import dcgan_super_resolution_b200 as dsr
from dcgan_super_resolution_b200 import _lib as L
ctx = dsr.Context(device=0, precision="tf32")
LAYERS = [  # full, cin, h, cout, k, s, p, batches
    (1, 256, 64, 128, 4, 2, 1, (16, 64)), (0, 128, 128, 256, 4, 2, 1, (16, 64)),               # C1b G: FC 256->128, C 128->256
    (0, 128, 32, 256, 4, 2, 1, (128, 256)), (0, 256, 16, 512, 4, 2, 1, (128, 256)),            # D at C3b
    (0, 128, 16, 256, 4, 2, 1, (128,)), (0, 256, 8, 512, 4, 2, 1, (128,)),                     # D at C2 / C3a
    (0, 128, 28, 256, 3, 1, 0, (64, 128)),                                                     # patch-D (C1b / C4)
    (1, 1024, 256, 512, 4, 2, 1, (2,)), (1, 512, 512, 256, 4, 2, 1, (1,)),                     # C5 G (micro-batch slices)
]
for (full, cin, h, cout, k, s, p, batches) in LAYERS:
    for n in batches:
        row = []
        for mode in ("0", "2"):
            os.environ["DCGANSR_WGRAD_PAIR"] = mode
            ms = ctypes.c_float()
            L.check(ctx.lib.dcgansr_bench_conv(ctx.h, full, 2, n, cin, h, h, cout, k, s, p, 5, ctypes.byref(ms)), ctx.h)
            row.append(ms.value * 1e3)
        ho = (h - 1) * s - 2 * p + k if full else (h + 2 * p - k) // s + 1
        gf = 2.0 * n * (h * h if full else ho * ho) * cin * cout * k * k / 1e9
        print(f"{'FC' if full else 'C '} {cin:4d}->{cout:4d} {h:3d} k{k}s{s} n={n:4d} {gf:7.1f} GF  wgrad single {row[0]:8.1f} ({gf / row[0]:5.1f} TF/s)  pair {row[1]:8.1f} ({gf / row[1]:5.1f} TF/s)  {row[0] / row[1]:4.2f}x", flush=True)
ctx.close()
