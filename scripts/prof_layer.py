"""One convolution op on device-resident tensors (for ncu):  python scripts/prof_layer.py full|conv what n cin h w cout k s p [iters]"""
import ctypes, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import dcgan_super_resolution_b200 as dsr
from dcgan_super_resolution_b200 import _lib as L
full = sys.argv[1] == "full"
what, n, cin, h, w, cout, k, s, p = [int(x) for x in sys.argv[2:11]]
iters = int(sys.argv[11]) if len(sys.argv) > 11 else 1
ctx = dsr.Context(device=0, precision=os.environ.get("DCGANSR_PRECISION", "tf32"))
ms = ctypes.c_float()
L.check(ctx.lib.dcgansr_bench_conv(ctx.h, int(full), what, n, cin, h, w, cout, k, s, p, iters, ctypes.byref(ms)), ctx.h)
print("ms", ms.value)
ctx.close()
