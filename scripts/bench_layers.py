"""Per-layer device timing of every convolution op of a config against its roofline (dcgansr_bench_conv).

  python scripts/bench_layers.py [C2] [--precision tf32|strict] [--iters 10] [--json out.json]

min time = max(flops / TF32 peak, min bytes / HBM peak) with min bytes = in + out + weights (fp32)."""
import argparse, ctypes, json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import dcgan_super_resolution_b200 as dsr
from dcgan_super_resolution_b200 import _lib as L
from dcgan_super_resolution_b200 import models

ap = argparse.ArgumentParser()
ap.add_argument("workload", nargs="?", default="C2")
ap.add_argument("--precision", default="tf32")
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--batch", type=int, default=0)
ap.add_argument("--json", default="")
args = ap.parse_args()
cfg = models.config(args.workload)
B = args.batch or cfg["batch"]
peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))) if os.path.exists(
    os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
HBM = peaks["hbm_gbs"] * 1e9
TF32 = peaks["bf16_tflops"] * 1e12 / 2          # dense TF32 = half the bf16 rate
ctx = dsr.Context(device=0, precision=args.precision)
rows = []
tot = tot_min = 0.0
for net, specs, (c, h, w) in (("G", cfg["G"], (cfg["nc"], cfg["hr"] // 2, cfg["hr"] // 2)), ("D", cfg["D"], (cfg["nc"], cfg["hr"], cfg["hr"]))):
    first = True
    for s in specs:
        if s["kind"] == "upnearest":
            h *= 2; w *= 2
            continue
        if s["kind"] not in ("conv", "fullconv"):
            continue
        full = s["kind"] == "fullconv"
        k, st, p, cout = s["k"], s["s"], s["p"], s["cout"]
        ho, wo = ((h - 1) * st - 2 * p + k, (w - 1) * st - 2 * p + k) if full else ((h + 2 * p - k) // st + 1, (w + 2 * p - k) // st + 1)
        flops = 2.0 * B * (h * w if full else ho * wo) * c * cout * k * k
        bytes_ = 4.0 * (B * c * h * w + B * cout * ho * wo + c * cout * k * k)
        tmin = max(flops / TF32, bytes_ / HBM) * 1e3
        for what, name in ((0, "fwd"), (1, "dgrad"), (2, "wgrad")):
            if what == 1 and first and net == "G":
                continue   # dead dgrad of G's first layer (never run in the step)
            ms = ctypes.c_float()
            L.check(ctx.lib.dcgansr_bench_conv(ctx.h, int(full), what, B, c, h, w, cout, k, st, p, args.iters, ctypes.byref(ms)), ctx.h)
            mult = {"G": {0: 1, 1: 1, 2: 1}, "D": {0: 2, 1: 3 if not first else 1, 2: 2}}[net][what]
            if net == "D" and first and what == 1:
                mult = 1      # only the fGx walk needs D's input gradient
            rows.append(dict(net=net, layer=f"{'FC' if full else 'C'} {c}->{cout} {h}->{ho}", op=name, ms=ms.value, min_ms=tmin,
                             frac=tmin / ms.value, gflop=flops / 1e9, mbytes=bytes_ / 1e6, per_step=mult))
            tot += ms.value * mult
            tot_min += tmin * mult
            print(f"{net} {'FC' if full else 'C '} {c:4d}->{cout:4d} {h:4d}->{ho:4d} {name:5s} {ms.value * 1e3:9.1f} us  min {tmin * 1e3:8.1f} us  "
                  f"frac {tmin / ms.value:5.2f}  x{mult}  ({flops / 1e9:7.2f} GF, {bytes_ / 1e6:7.1f} MB)", flush=True)
        c, h, w = cout, ho, wo
        first = False
print(f"conv ops per step: {tot:.3f} ms measured, {tot_min:.3f} ms roofline minimum")
if args.json:
    json.dump(dict(workload=args.workload, batch=B, precision=args.precision, rows=rows, total_ms=tot, total_min_ms=tot_min), open(args.json, "w"), indent=1)
ctx.close()
