"""Ring-depth experiment for the halo-tile pair kernel: plane buffers (DCGANSR_TC3_NA) x weight stages (DCGANSR_TC3_NB), one process."""
import ctypes, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import dcgan_super_resolution_b200 as dsr
from dcgan_super_resolution_b200 import _lib as L
ctx = dsr.Context(device=0, precision="tf32")
LAYERS = [("FC 256->128 fwd (2x2 classes, 128 couts)", (1, 256, 64, 128, 4, 2, 1, 64, 0)), ("FC 256->128 dgrad (16 taps, 256 couts)", (1, 256, 64, 128, 4, 2, 1, 64, 1)),
          ("C 64->128 fwd (16 taps, 128 couts)", (0, 64, 256, 128, 4, 2, 1, 64, 0))]
for name, (full, cin, h, cout, k, s, p, n, what) in LAYERS:
    for na in (2, 3, 4, 5, 8):
        row = []
        for nb in (4, 8, 12, 16, 32):
            os.environ["DCGANSR_TC3_NA"] = str(na); os.environ["DCGANSR_TC3_NB"] = str(nb)
            ms = ctypes.c_float()
            L.check(ctx.lib.dcgansr_bench_conv(ctx.h, full, what, n, cin, h, h, cout, k, s, p, 5, ctypes.byref(ms)), ctx.h)
            row.append(f"nb<={nb}: {ms.value * 1e3:6.1f}")
        print(f"{name:42s} na<={na}  " + "  ".join(row), flush=True)
ctx.close()
