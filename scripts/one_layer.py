"""One convolution op on the device (2 warm-ups + `iters` launches) for single-kernel ncu captures.

  python scripts/one_layer.py <full 0|1> <cin> <h> <cout> <k> <s> <p> <batch> <what: 0 fwd | 1 dgrad | 2 wgrad> [iters]"""
import ctypes, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import dcgan_super_resolution_b200 as dsr
from dcgan_super_resolution_b200 import _lib as L
full, cin, h, cout, k, s, p, n, what = (int(a) for a in sys.argv[1:10])
iters = int(sys.argv[10]) if len(sys.argv) > 10 else 1
ctx = dsr.Context(device=0, precision=os.environ.get("DCGANSR_PRECISION", "tf32"))
ms = ctypes.c_float()
L.check(ctx.lib.dcgansr_bench_conv(ctx.h, full, what, n, cin, h, h, cout, k, s, p, iters, ctypes.byref(ms)), ctx.h)
print(f"{'FC' if full else 'C'} {cin}->{cout} {h} n={n} what={what}: {ms.value * 1e3:.1f} us")
ctx.close()
