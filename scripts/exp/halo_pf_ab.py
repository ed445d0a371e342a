"""In-process A/B of the halo kernel's L2 prefetch distance (DCGANSR_HALO_PFDIST: unset = ring depth in tiles + 1, 0 = no prefetch)
on the six halo launches of a C3b step (train.lua generator at 64x64 -> 128x128, batch 128)."""
import ctypes, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import dcgan_super_resolution_b200 as dsr
from dcgan_super_resolution_b200 import _lib as L
ctx = dsr.Context(device=0, precision="tf32")
LAYERS = [("FC 96->48 fwd", (1, 96, 128, 48, 4, 2, 1, 128, 0)), ("FC 48->24 fwd", (1, 48, 256, 24, 4, 2, 1, 128, 0)),
          ("C 24->12 fwd", (0, 24, 512, 12, 4, 2, 1, 128, 0)), ("C 24->12 dgrad", (0, 24, 512, 12, 4, 2, 1, 128, 1)),
          ("FC 48->24 dgrad", (1, 48, 256, 24, 4, 2, 1, 128, 1)), ("FC 96->48 dgrad", (1, 96, 128, 48, 4, 2, 1, 128, 1))]
if len(sys.argv) > 1 and sys.argv[1] == "small":      # the halo launches of C2 (batch 64) and C4 (64 per GPU)
    LAYERS = [("C2 FC 64->32 fwd", (1, 64, 128, 32, 4, 2, 1, 64, 0)), ("C2 FC 64->32 dgrad", (1, 64, 128, 32, 4, 2, 1, 64, 1)),
              ("C2 C 32->16 fwd", (0, 32, 256, 16, 4, 2, 1, 64, 0)), ("C2 C 32->16 dgrad", (0, 32, 256, 16, 4, 2, 1, 64, 1)),
              ("C4 FC 32->16 fwd", (1, 32, 128, 16, 4, 2, 1, 64, 0)), ("C4 C 16->32 fwd", (0, 16, 256, 32, 4, 2, 1, 64, 0))]
if len(sys.argv) > 1 and sys.argv[1] == "onetile":    # layers whose plane ring holds at most one tile (C2 / C4 / C1b)
    LAYERS = [("C2 FC 64->32 dgrad", (1, 64, 128, 32, 4, 2, 1, 64, 1)), ("C4 C 32->64 fwd", (0, 32, 128, 64, 4, 2, 1, 64, 0)),
              ("C1b FC 128->64 fwd", (1, 128, 128, 64, 4, 2, 1, 64, 0)), ("C1b FC 128->64 dgr", (1, 128, 128, 64, 4, 2, 1, 64, 1)),
              ("C1b C 64->128 fwd", (0, 64, 256, 128, 4, 2, 1, 64, 0)), ("C1b C 64->128 dgr", (0, 64, 256, 128, 4, 2, 1, 64, 1))]
MODES = [None, "0", "1", "2", "3", "4"]      # unset = the policy of k_tapconv_halo
for rep in range(2):
    for name, (full, cin, h, cout, k, s, p, n, what) in LAYERS:
        row = []
        for mode in MODES:
            if mode is None:
                os.environ.pop("DCGANSR_HALO_PFDIST", None)
            else:
                os.environ["DCGANSR_HALO_PFDIST"] = mode
            ms = ctypes.c_float()
            L.check(ctx.lib.dcgansr_bench_conv(ctx.h, full, what, n, cin, h, h, cout, k, s, p, 5, ctypes.byref(ms)), ctx.h)
            row.append(ms.value * 1e3)
        print(f"{name:18s} " + "  ".join(f"pf={m or 'dflt'}: {v:7.1f} us" for m, v in zip(MODES, row)), flush=True)
os.environ.pop("DCGANSR_HALO_PFDIST", None)
ctx.close()
