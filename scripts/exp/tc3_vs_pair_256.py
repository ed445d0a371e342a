"""256-cout layers on full tiles: per-tap CTA-pair kernel (DCGANSR_TC3=0) against the halo-tile pair kernel (policy), one process."""
import ctypes, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import dcgan_super_resolution_b200 as dsr
from dcgan_super_resolution_b200 import _lib as L
ctx = dsr.Context(device=0, precision="tf32")
LAYERS = [("FC 1024->512 (C5)", (1, 1024, 256, 512, 4, 2, 1, 2)), ("FC 512->256 (C5)", (1, 512, 512, 256, 4, 2, 1, 1)), ("C 256->128 (C5)", (0, 256, 1024, 128, 4, 2, 1, 1)),
          ("C 128->256 (C1b)", (0, 128, 128, 256, 4, 2, 1, 64)), ("FC 256->128 (C1b)", (1, 256, 64, 128, 4, 2, 1, 64)), ("D C 128->256 (C3b)", (0, 128, 32, 256, 4, 2, 1, 256))]
for name, (full, cin, h, cout, k, s, p, n) in LAYERS:
    row = []
    for what in (0, 1):
        for mode in ("0", "1"):
            os.environ["DCGANSR_TC3"] = mode
            ms = ctypes.c_float()
            L.check(ctx.lib.dcgansr_bench_conv(ctx.h, full, what, n, cin, h, h, cout, k, s, p, 3, ctypes.byref(ms)), ctx.h)
            row.append(ms.value * 1e3)
    ho = (h - 1) * s - 2 * p + k if full else (h + 2 * p - k) // s + 1
    gf = 2.0 * n * (h * h if full else ho * ho) * cin * cout * k * k / 1e9
    print(f"{name:20s} {gf:7.1f} GF  fwd per-tap pair {row[0]:8.1f} ({gf / row[0]:5.2f} TF/s x1000)  halo-tile {row[1]:8.1f} ({gf / row[1]:5.2f})   dgrad per-tap pair {row[2]:8.1f} ({gf / row[2]:5.2f})  halo-tile {row[3]:8.1f} ({gf / row[3]:5.2f})", flush=True)
ctx.close()
