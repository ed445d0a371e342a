#!/usr/bin/env python
"""bench.py -- SR-GAN train patches/s of the DCGAN-SR training step (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C2] [--precision strict|tf32]

A "step" is one full iteration of the reference loop (fDx -> adam(D) -> fGx -> adam(G), train.lua:208-283)
over one synthetic batch.  N=1 workload: BASELINE.json configs[1] = train-gray.lua, 64x64 gray images,
batchSize 64 ("C2").  N>1 (torchrun, one rank per GPU): the same per-GPU batch on every rank (weak scaling),
gradients / losses all-reduced by NCCL inside libdcgansr.so.

One JSON line on rank 0.  `value`: inputs resident in HBM (staged batches), CUDA-event timed on the library's
stream, max over ranks.  `e2e`: the same step through dcgansr_train_step with pinned HOST batches (H2D copy
inside the timed region) and the three loss scalars read back every step.  `roofline`: dominant kernel, timed
live with per-launch CUDA events (dcgansr_profile_begin/end) in extra steps right after the timed region.
`cpu_baseline` / `--impl reference`: the oracle's float32 restatement of the Torch7 step on the host cores
(Torch7 itself cannot run: no Lua in the image) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "SR-GAN train patches/sec (G+D fwd+bwd+adam)"
UNIT = "patches/s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=float(d["hbm_gbs"]), tf=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                    tf_burst=float(d["bf16_tflops"]), source="measured")
    return dict(hbm=6650.0, tf=1400.0, tf_burst=1590.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark(self):
        return len(self.lines)

    def stop(self, lo=0, hi=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = self.lines[lo:hi] or self.lines
        for ln in rows:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def synth_batches(cfg, batch, n, seed):
    import numpy as np
    rng = np.random.Generator(np.random.Philox(seed))
    lo, hi = cfg["data_range"]
    return [rng.uniform(lo, hi, size=(batch, cfg["nc"], cfg["hr"], cfg["hr"])).astype(np.float32) for _ in range(n)]


def oracle_step_rate(workload, sample_batch, min_seconds, max_steps, threads=None):
    """Times the oracle's float32 step (test infrastructure, used here ONLY as the CPU baseline)."""
    import numpy as np
    import torch
    from dcgan_super_resolution_b200 import models
    from oracle import nets as onets
    from oracle import step as ostep
    torch.set_num_threads(threads or os.cpu_count() or 1)      # torchrun pins OMP_NUM_THREADS=1: undo for the CPU arm
    cfg = models.config(workload)
    oG = onets.weights_init(onets.Sequential(cfg["G"], torch.float32), 4321)
    oD = onets.weights_init(onets.Sequential(cfg["D"], torch.float32), 8765)
    stG, stD = ostep.new_adam_state(oG), ostep.new_adam_state(oD)
    scfg = ostep.StepCfg(**cfg["step"])
    data = [torch.from_numpy(b) for b in synth_batches(cfg, sample_batch, 2, 1234)]
    ostep.train_step(oG, oD, stG, stD, data[0], scfg)          # warm-up
    t0 = time.perf_counter()
    n = 0
    while n < max_steps:
        ostep.train_step(oG, oD, stG, stD, data[n % 2], scfg)
        n += 1
        if time.perf_counter() - t0 >= min_seconds:
            break
    dt = time.perf_counter() - t0
    return sample_batch * n / dt, n, dt, torch.get_num_threads()


def reference_arm(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return 0
    from dcgan_super_resolution_b200 import models
    cfg = models.config(args.workload)
    sb = min(args.cpu_sample_batch, cfg["batch"])
    # each "step" = one oracle step on the bounded sample; K steps, W warm-ups
    import torch
    from oracle import nets as onets
    from oracle import step as ostep
    torch.set_num_threads(os.cpu_count() or 1)                  # all host threads (torchrun pins OMP_NUM_THREADS=1)
    oG = onets.weights_init(onets.Sequential(cfg["G"], torch.float32), 4321)
    oD = onets.weights_init(onets.Sequential(cfg["D"], torch.float32), 8765)
    stG, stD = ostep.new_adam_state(oG), ostep.new_adam_state(oD)
    scfg = ostep.StepCfg(**cfg["step"])
    data = [torch.from_numpy(b) for b in synth_batches(cfg, sb, 2, 1234)]
    steps = max(1, min(args.steps, 6))
    warm = max(1, min(args.warmup, 2))
    for i in range(warm):
        ostep.train_step(oG, oD, stG, stD, data[i % 2], scfg)
    t0 = time.perf_counter()
    for i in range(steps):
        ostep.train_step(oG, oD, stG, stD, data[i % 2], scfg)
    dt = time.perf_counter() - t0
    val = sb * steps / dt
    cores = torch.get_num_threads()
    sample = (f"oracle float32 restatement (PyTorch-CPU, oneDNN, {cores} threads) of the {args.workload} step on "
              f"{sb}-sample batches, {steps} steps; Torch7 itself cannot run here (no Lua)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": workload_name(args.workload, cfg, sb), "cpu_sample_batch": sb},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))
    return 0


def workload_name(name, cfg, batch):
    src = {"C1a": "train-gray-patch.lua", "C1b": "train-gray-patch.lua (32x32 patches, ngf=ndf=64)", "C2": "train-gray.lua",
           "C3a": "train.lua (32->64)", "C3b": "train.lua (64->128)", "C4": "train-gray-patch-batch-overlap.lua (32x32)",
           "C4a": "train-gray-patch-batch-overlap.lua (8x8)", "C5": "train.lua scaled (128->256, ngf=ndf=128)"}[name]
    return f"{name}: {src}, nc={cfg['nc']}, {cfg['hr'] // 2}x{cfg['hr'] // 2}->{cfg['hr']}x{cfg['hr']}, per-GPU batch {batch}, {cfg['step']['family']} family"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2")
    ap.add_argument("--precision", default=os.environ.get("DCGANSR_PRECISION", "tf32"), choices=["strict", "tf32"])
    ap.add_argument("--graph", type=int, default=1, help="replay the step as a CUDA graph (one cached graph per staged batch)")
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch override")
    ap.add_argument("--sync-bn", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-batch", type=int, default=64, help="batch of the CPU arms (bounded sample of the workload)")
    ap.add_argument("--profile-steps", type=int, default=2)
    ap.add_argument("--profile-out", default="", help="write the full per-kernel profile table (JSON) here")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    if args.warmup < 3:
        args.warmup = 3

    import numpy as np
    import torch

    import dcgan_super_resolution_b200 as dsr
    from dcgan_super_resolution_b200 import init, models, parallel

    rank, local_rank, world = parallel.env_rank()
    if world != args.gpus and world > 1:
        args.gpus = world
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    cfg = models.config(args.workload)
    B = args.batch or cfg["batch"]
    ctx = dsr.Context(device=local_rank, precision=args.precision, world_size=world, rank=rank, sync_bn=bool(args.sync_bn),
                      use_graph=bool(args.graph))
    if world > 1:
        parallel.exchange_unique_id(ctx, dist, device=torch.device("cuda", local_rank))
    G = dsr.Sequential.from_specs(cfg["G"]).cuda(ctx, (cfg["nc"], cfg["hr"] // 2, cfg["hr"] // 2), B)
    # D holds 2B samples: the step then runs D(real) and D(fake) as one grouped pass (dcgansr.cu:step_body)
    D = dsr.Sequential.from_specs(cfg["D"]).cuda(ctx, (cfg["nc"], cfg["hr"], cfg["hr"]), 2 * B)
    G.set_params(init.weights_init(cfg["G"], 4321))
    D.set_params(init.weights_init(cfg["D"], 8765))
    scfg = dsr.make_step_cfg(**cfg["step"])

    NPOOL = 8
    pool = synth_batches(cfg, B, NPOOL, 1234 + rank)
    for i, b in enumerate(pool):
        dsr.stage_batch(ctx, D, b, i)
    pinned = [torch.from_numpy(b).pin_memory() for b in pool]
    ctx.synchronize()

    def barrier():
        ctx.synchronize()
        if dist is not None:
            dist.barrier()
        ctx.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=torch.device("cuda", local_rank))
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- device-resident leg ----------------
    # The library caches one CUDA graph per staged batch (the batch pointer is baked into the graph): visit every slot once so
    # no capture / instantiation (tens of ms each with NCCL nodes) falls into the timed region, then the W warm-up steps.
    prime = NPOOL if args.graph else 0
    for i in range(prime):
        dsr.train_step_staged(ctx, G, D, scfg, i, B)
    for i in range(args.warmup):
        dsr.train_step_staged(ctx, G, D, scfg, i % NPOOL, B)
    sampler = ClockSampler(local_rank)
    if rank == 0 and not os.environ.get("DCGANSR_NO_SAMPLER"):
        sampler.start()
        time.sleep(0.25)
    barrier()                      # AFTER the sampler start-up: every rank enters the timed region together
    l0 = ctx.launch_count()
    lo_mark = sampler.mark()
    ctx.timer_begin()
    for i in range(args.steps):
        dsr.train_step_staged(ctx, G, D, scfg, i % NPOOL, B)
    ms = ctx.timer_end()
    barrier()
    hi_mark = sampler.mark()
    launches = ctx.launch_count() - l0
    ms = max_over_ranks(ms)
    clocks = sampler.stop(lo_mark, hi_mark) if rank == 0 else None
    value = world * B * args.steps / (ms * 1e-3)

    # ---------------- end-to-end leg: pinned host batch in, losses out, every step ----------------
    losses = (ctypes.c_float * 3)()
    for i in range(2):
        dsr.nn.train_step_ptr(ctx, G, D, scfg, pinned[i % NPOOL].data_ptr(), B, losses)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        dsr.nn.train_step_ptr(ctx, G, D, scfg, pinned[i % NPOOL].data_ptr(), B, losses)
    ctx.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_value = world * B * args.steps / e2e_s
    last_losses = [float(x) for x in losses]

    # ---------------- per-kernel profile (extra steps, same stream, CUDA events per launch) --------
    prof = []
    if args.profile_steps > 0:                       # eager launches with per-launch events (the graph is bypassed while profiling)
        ctx.profile_begin()
        for i in range(args.profile_steps):
            dsr.train_step_staged(ctx, G, D, scfg, i % NPOOL, B)
        prof = ctx.profile_end()
    barrier()
    if args.profile_out and rank == 0:
        with open(args.profile_out, "w") as f:
            json.dump({"profile_steps": args.profile_steps, "kernels": prof}, f, indent=1)

    if os.environ.get("DCGANSR_DEBUG_LEGS"):
        for label, sync_each in (("staged-async", False), ("staged-sync-each-step", True)):
            barrier()
            ctx.timer_begin()
            for i in range(args.steps):
                dsr.train_step_staged(ctx, G, D, scfg, i % NPOOL, B, want_losses=sync_each)
            t = ctx.timer_end()
            barrier()
            sys.stderr.write(f"[rank {rank}] {label}: {t / args.steps:.3f} ms/step\n")

    def teardown():
        # orderly: every rank destroys its communicator at the same point, then leaves without waiting for
        # interpreter-exit destructors (a lingering NCCL / process-group thread must never hang the driver)
        sys.stdout.flush()
        sys.stderr.flush()
        barrier()
        try:
            G.close()
            D.close()
            ctx.close()
        finally:
            os._exit(0)

    if rank != 0:
        teardown()

    peaks = load_peaks()
    total_ms = sum(p["ms"] for p in prof) or 1.0
    roof = None
    # the library reports one record per (kernel, algorithmic work of the launch); the dominant KERNEL is the one with the
    # largest summed time, its achieved rate = summed algorithmic work / summed launch time (i.e. per average launch)
    by_name = {}
    for p in prof:
        a = by_name.setdefault(p["name"], {"kernel": p["name"], "kind": p["kind"], "ms": 0.0, "launches": 0, "work": 0.0})
        a["ms"] += p["ms"]
        a["launches"] += p["launches"]
        a["work"] += p["work"] * p["launches"]
    top = []
    for a in sorted(by_name.values(), key=lambda x: -x["ms"])[:10]:
        avg = a["ms"] / a["launches"]
        wpl = a["work"] / a["launches"]
        if a["kind"] == "flops":
            # dense TF32 runs at half the bf16 rate the peaks file holds
            ach, peak, unit, bound = wpl / (avg * 1e-3) / 1e12, peaks["tf"] / (2.0 if args.precision == "tf32" else 1.0), "TFLOP/s", "tensor"
        else:
            ach, peak, unit, bound = wpl / (avg * 1e-3) / 1e9, peaks["hbm"], "GB/s", "hbm"
        top.append({"kernel": a["kernel"], "work_per_launch": wpl, "launches_per_step": a["launches"] / max(args.profile_steps, 1),
                    "avg_ms": avg, "share": a["ms"] / total_ms, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak, "bound": bound})
    if top:
        t = top[0]
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get(t["kernel"], {}).get("dram_bytes")
            except Exception:
                traffic = None
        roof = {"bound": t["bound"], "achieved": t["achieved"], "peak": t["peak"], "unit": t["unit"], "frac": t["frac"], "traffic": traffic,
                "kernel": t["kernel"], "share_of_step": t["share"], "avg_launch_ms": t["avg_ms"], "work_per_launch": t["work_per_launch"],
                "launches_per_step": t["launches_per_step"],
                "peak_source": peaks["source"] + (" (copy bandwidth)" if t["bound"] == "hbm" else
                                                  " (sustained bf16 GEMM / 2: kind::tf32 runs at half the bf16 rate)" if args.precision == "tf32"
                                                  else " (sustained bf16 GEMM; strict mode computes in fp32 FFMA)"),
                "traffic_note": "dram__bytes_read+write of the heaviest launch of this kernel in profiles/ (ncu --set full)"}

    cpu = None
    if not args.no_cpu_baseline:
        try:
            v, n, dt, cores = oracle_step_rate(args.workload, min(args.cpu_sample_batch, B), 10.0, 40)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"oracle float32 step (PyTorch-CPU restatement of the Torch7 path), {min(args.cpu_sample_batch, B)}-sample "
                             f"batches of {args.workload}, {n} steps in {dt:.1f} s after 1 warm-up"}
        except Exception as e:   # the baseline must never kill the GPU number
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {e!r}"}

    flops = models.step_flops(cfg, B)
    act_bytes = 0
    for specs, (c, h, w) in ((cfg["G"], (cfg["nc"], cfg["hr"] // 2, cfg["hr"] // 2)), (cfg["D"], (cfg["nc"], cfg["hr"], cfg["hr"]))):
        net = dsr.Sequential.from_specs(specs).cuda(None, (c, h, w), B)
        # rough: sum of conv outputs, fp32
        cc, hh, ww = c, h, w
        for s in specs:
            if s["kind"] == "conv":
                hh = (hh + 2 * s["p"] - s["k"]) // s["s"] + 1
                ww = (ww + 2 * s["p"] - s["k"]) // s["s"] + 1
                cc = s["cout"]
                act_bytes += 4 * B * cc * hh * ww
            elif s["kind"] == "fullconv":
                hh = (hh - 1) * s["s"] - 2 * s["p"] + s["k"]
                ww = (ww - 1) * s["s"] - 2 * s["p"] + s["k"]
                cc = s["cout"]
                act_bytes += 4 * B * cc * hh * ww
            elif s["kind"] == "upnearest":
                hh *= 2
                ww *= 2
        net.close()
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "tf32" if args.precision == "tf32" else "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.workload, cfg, B), "global_batch": world * B, "precision": args.precision,
                   "cuda_graph": bool(args.graph), "graph_prime_steps": prime, "sync_bn": bool(args.sync_bn), "parallelism": f"dp{world}",
                   "l2": f"no flush: per-step working set (conv outputs {act_bytes / 1e6:.0f} MB fp32 per rank, 8 rotating input batches) exceeds the 126 MB L2",
                   "algorithmic_gflop_per_step_per_gpu": flops / 1e9},
        "step_tflops": world * flops / (ms / args.steps * 1e-3) / 1e12,
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(pool[0].nbytes), "d2h_bytes_per_step": 12,
                "ms_per_step": 1e3 * e2e_s / args.steps, "last_losses": last_losses},
        "gpu_launches": int(launches),
        "roofline": roof,
        "kernels": top,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    teardown()
    return 0


if __name__ == "__main__":
    sys.exit(main())
