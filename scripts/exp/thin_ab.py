"""In-process A/B of the thin-kernel switches (DCGANSR_THIN_PAD34, DCGANSR_THIN_PX2) on the RGB layers of train.lua at C3b size."""
import ctypes, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import dcgan_super_resolution_b200 as dsr
from dcgan_super_resolution_b200 import _lib as L
ctx = dsr.Context(device=0, precision="tf32")
LAYERS = [("C 12->3 wgrad", (0, 12, 256, 3, 4, 2, 1, 128, 2)), ("D C 3->64 wgrad", (0, 3, 128, 64, 4, 2, 1, 256, 2)), ("FC 3->96 wgrad", (1, 3, 64, 96, 4, 2, 1, 128, 2)),
          ("D C 3->64 fwd", (0, 3, 128, 64, 4, 2, 1, 256, 0)), ("C 12->3 dgrad", (0, 12, 256, 3, 4, 2, 1, 128, 1))]
env = sys.argv[1] if len(sys.argv) > 1 else "DCGANSR_THIN_PAD34"
for rep in range(2):
    for name, (full, cin, h, cout, k, s, p, n, what) in LAYERS:
        row = []
        for mode in ("0", "1"):
            os.environ[env] = mode
            ms = ctypes.c_float()
            L.check(ctx.lib.dcgansr_bench_conv(ctx.h, full, what, n, cin, h, h, cout, k, s, p, 5, ctypes.byref(ms)), ctx.h)
            row.append(ms.value * 1e3)
        print(f"{name:18s} {env}=0 {row[0]:7.1f} us   =1 {row[1]:7.1f} us   {row[0] / row[1]:4.2f}x", flush=True)
ctx.close()
