#!/usr/bin/env python
"""bench.py -- SR-GAN train patches/s of the DCGAN-SR training step (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C3b] [--precision tf32|strict]

A "step" is one full iteration of the reference loop (fDx -> adam(D) -> fGx -> adam(G), train.lua:208-283) over one
synthetic batch.  Headline workload at N=1: C3b = BASELINE.json configs[2], train.lua scaled to RGB 64x64 -> 128x128,
batchSize 128 -- the largest configuration BASELINE tags for a single B200.  N>1 (torchrun, one rank per GPU): the same
per-GPU batch on every rank (weak scaling); gradients / BN running statistics / losses are all-reduced by NCCL inside
libdcgansr.so (the control plane of this script -- barriers, max over ranks -- is gloo, so every NCCL communicator of the
run is the library's own).

ONE JSON line on rank 0:
  value / ms_per_step   staged (HBM-resident) batches, CUDA events on the library stream, max over ranks
  e2e                   the same step through dcgansr_train_step with pinned HOST batches (H2D copy + layout change inside
                        the timed region) and the three loss scalars read back every step
  strict                the headline workload in DCGANSR_STRICT_FP32 (fp32 FFMA kernels, the 1e-5 acceptance mode)
  roofline / kernels    per-launch CUDA events (dcgansr_profile_begin/end) in extra steps right after the timed region
  workloads[]           the other BASELINE configurations on the same GPU(s): N=1 -> C1b, C2, C3a, C4 (64 per GPU);
                        N>1 -> the BASELINE data-parallel wordings: C3b strong-scaled (128 / N per GPU), C4 (512 / N per GPU)
  dp_check              N>1: bucket-overlapped all-reduce bit-identical to one all-reduce per net, graph replay identical,
                        sync_bn over N shards == the single-GPU step on the concatenated batch
  cpu_baseline          the oracle's float32 restatement of the Torch7 step on the host cores (Torch7 itself cannot run: no Lua
                        in the image): all threads + oneDNN, all threads + slow_conv2d (the THNN descendant), 1 thread
`--impl reference` times that CPU restatement alone (rank 0; the other ranks exit 0).
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
HEADLINE = "C3b"
# per-sample cost of the CPU restatement differs by ~3 orders of magnitude between workloads: bounded samples (batch) per workload
CPU_SAMPLE = {"C1a": 64, "C1b": 8, "C2": 32, "C3a": 16, "C3b": 8, "C4": 32, "C4a": 64, "C5": 1}
WORKLOAD_SRC = {"C1a": "train-gray-patch.lua", "C1b": "train-gray-patch.lua (32x32 patches, ngf=ndf=64)", "C2": "train-gray.lua",
                "C3a": "train.lua (32->64)", "C3b": "train.lua (64->128)", "C4": "train-gray-patch-batch-overlap.lua (32x32)",
                "C4a": "train-gray-patch-batch-overlap.lua (8x8)", "C5": "train.lua scaled (128->256, ngf=ndf=128)"}


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


def workload_name(name, cfg, batch):
    return (f"{name}: {WORKLOAD_SRC[name]}, nc={cfg['nc']}, {cfg['hr'] // 2}x{cfg['hr'] // 2}->{cfg['hr']}x{cfg['hr']}, "
            f"per-GPU batch {batch}, {cfg['step']['family']} family")


def make_config(workload, cfg, B, world):
    """The `config` object is identical in the GPU arm and in the reference arm (what is measured, not how)."""
    return {"workload": workload_name(workload, cfg, B), "global_batch": world * B, "parallelism": f"dp{world}"}


# ------------------------------------------------------------------------------------------------------------------
# CPU arms: the oracle's float32 step (test infrastructure, used here ONLY as the reported CPU baseline)
# ------------------------------------------------------------------------------------------------------------------
class OracleStepper:
    def __init__(self, workload, sample_batch, threads=None, mkldnn=True):
        import torch
        from dcgan_super_resolution_b200 import models
        from oracle import nets as onets
        from oracle import step as ostep
        self.torch, self.ostep = torch, ostep
        torch.set_num_threads(threads or os.cpu_count() or 1)      # torchrun pins OMP_NUM_THREADS=1: undo for the CPU arm
        self.mkldnn = mkldnn
        cfg = models.config(workload)
        self.sb = sample_batch
        self.oG = onets.weights_init(onets.Sequential(cfg["G"], torch.float32), 4321)
        self.oD = onets.weights_init(onets.Sequential(cfg["D"], torch.float32), 8765)
        self.stG, self.stD = ostep.new_adam_state(self.oG), ostep.new_adam_state(self.oD)
        self.scfg = ostep.StepCfg(**cfg["step"])
        self.data = [torch.from_numpy(b) for b in synth_batches(cfg, sample_batch, 2, 1234)]
        self.i = 0

    def step(self):
        with self.torch.backends.mkldnn.flags(enabled=self.mkldnn):
            self.ostep.train_step(self.oG, self.oD, self.stG, self.stD, self.data[self.i % 2], self.scfg)
        self.i += 1

    def rate(self, min_seconds, max_steps, warmup=1):
        for _ in range(warmup):
            self.step()
        t0 = time.perf_counter()
        n = 0
        while n < max_steps:
            self.step()
            n += 1
            if time.perf_counter() - t0 >= min_seconds:
                break
        dt = time.perf_counter() - t0
        return self.sb * n / dt, n, dt, self.torch.get_num_threads()


def cpu_baseline(workload, B):
    """oneDNN / all threads (the line's cpu_baseline.value), plus the two variants BASELINE.md section 3 names."""
    import torch
    sb = min(CPU_SAMPLE.get(workload, 8), B)
    out = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": ""}
    try:
        v, n, dt, cores = OracleStepper(workload, sb).rate(10.0, 40)
        out.update(value=v, cores=cores,
                   sample=f"oracle float32 step (PyTorch-CPU restatement of the Torch7 path, oneDNN, {cores} threads), {sb}-sample "
                          f"batches of {workload}, {n} steps in {dt:.1f} s after 1 warm-up; torch {torch.__version__}, os.cpu_count()={os.cpu_count()}")
        variants = []
        sb2 = max(1, sb // 2)
        v2, n2, dt2, c2 = OracleStepper(workload, sb2, mkldnn=False).rate(5.0, 10)
        variants.append({"name": "slow_conv2d (mkldnn off: ATen's direct port of THNN SpatialConvolutionMM, closest to torch7 nn)", "value": v2,
                         "cores": c2, "sample": f"{sb2}-sample batches, {n2} steps in {dt2:.1f} s"})
        sb3 = max(1, sb // 4)
        v3, n3, dt3, c3 = OracleStepper(workload, sb3, threads=1).rate(5.0, 10)
        variants.append({"name": "1 thread, oneDNN (the reference pins torch.setnumthreads(1), train.lua:33)", "value": v3, "cores": c3,
                         "sample": f"{sb3}-sample batches, {n3} steps in {dt3:.1f} s"})
        out["variants"] = variants
    except Exception as e:   # the baseline must never kill the GPU number
        out["sample"] += f" failed: {e!r}"
    return out


def reference_arm(args):
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if rank != 0:
        return 0
    import torch
    from dcgan_super_resolution_b200 import models
    cfg = models.config(args.workload)
    B = args.batch or cfg["batch"]
    sb = min(args.cpu_sample_batch or CPU_SAMPLE.get(args.workload, 8), B)
    st = OracleStepper(args.workload, sb)
    for _ in range(args.warmup):
        st.step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        st.step()
    dt = time.perf_counter() - t0
    val = sb * args.steps / dt
    cores = torch.get_num_threads()
    sample = (f"oracle float32 restatement (PyTorch-CPU, oneDNN, {cores} threads) of the {args.workload} step on "
              f"{sb}-sample batches, {args.steps} steps after {args.warmup} warm-ups; Torch7 itself cannot run here (no Lua)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": max(args.gpus, world), "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": make_config(args.workload, cfg, B, max(args.gpus, world)),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))
    return 0


# ------------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------------
class Env:
    """rank / world, the gloo control plane and the helpers every leg uses."""

    def __init__(self):
        from dcgan_super_resolution_b200 import parallel
        self.rank, self.local_rank, self.world = parallel.env_rank()
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            os.environ.setdefault("GLOO_SOCKET_IFNAME", "lo")
            dist.init_process_group("gloo", rank=self.rank, world_size=self.world)
            self.dist = dist

    def barrier(self, ctx=None):
        if ctx is not None:
            ctx.synchronize()
        if self.dist is not None:
            self.dist.barrier()

    def max_over_ranks(self, x):
        if self.dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def new_ctx(self, precision, graph=True, sync_bn=False, world=None):
        import dcgan_super_resolution_b200 as dsr
        from dcgan_super_resolution_b200 import parallel
        w = self.world if world is None else world
        ctx = dsr.Context(device=self.local_rank, precision=precision, world_size=w, rank=self.rank if w > 1 else 0,
                          sync_bn=sync_bn, use_graph=graph)
        if w > 1:
            parallel.exchange_unique_id(ctx, self.dist)
        return ctx

    def close(self):
        if self.dist is not None:
            self.dist.barrier()
            self.dist.destroy_process_group()
            self.dist = None


def aggregate_profile(prof, profile_steps, peaks, precision):
    """One record per kernel name: summed time, average launch, achieved rate against the measured peak."""
    total_ms = sum(p["ms"] for p in prof) or 1.0
    by_name = {}
    for p in prof:
        a = by_name.setdefault(p["name"], {"kernel": p["name"], "kind": p["kind"], "ms": 0.0, "launches": 0, "work": 0.0})
        a["ms"] += p["ms"]
        a["launches"] += p["launches"]
        a["work"] += p["work"] * p["launches"]
    top = []
    for a in sorted(by_name.values(), key=lambda x: -x["ms"]):
        avg = a["ms"] / a["launches"]
        wpl = a["work"] / a["launches"]
        if a["kind"] == "flops":
            # the profile launches are timed one by one (a kernel alone between two events): burst peak; dense TF32 = bf16 / 2
            ach, peak, unit, bound = wpl / (avg * 1e-3) / 1e12, peaks["tf_burst"] / (2.0 if precision == "tf32" else 1.0), "TFLOP/s", "tensor"
        else:
            ach, peak, unit, bound = wpl / (avg * 1e-3) / 1e9, peaks["hbm"], "GB/s", "hbm"
        top.append({"kernel": a["kernel"], "work_per_launch": wpl, "launches_per_step": a["launches"] / max(profile_steps, 1),
                    "avg_ms": avg, "ms_per_step": a["ms"] / max(profile_steps, 1), "share": a["ms"] / total_ms, "achieved": ach,
                    "peak": peak, "unit": unit, "frac": ach / peak, "bound": bound})
    return top, total_ms / max(profile_steps, 1)


def roofline_of(top, peaks, precision, workload=None, batch=None):
    if not top:
        return None
    t = top[0]
    # traffic: dram__bytes_read + dram__bytes_write PER LAUNCH (like `achieved`): the mean over this kernel's launches in the
    # committed per-launch ncu capture of one step -- which exists for the headline workload at its own batch only
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp) and workload == HEADLINE and precision == "tf32":
        try:
            from dcgan_super_resolution_b200 import models
            if batch == models.config(HEADLINE)["batch"]:
                rows = [r for r in json.load(open(tp)).get("_all", {}).get(t["kernel"], []) if r.get("from") == f"{HEADLINE} step"]
                if rows:
                    traffic = sum(r["dram_bytes"] for r in rows) / len(rows)
        except Exception:
            traffic = None
    return {"bound": t["bound"], "achieved": t["achieved"], "peak": t["peak"], "unit": t["unit"], "frac": t["frac"], "traffic": traffic,
            "kernel": t["kernel"], "share_of_step": t["share"], "avg_launch_ms": t["avg_ms"], "work_per_launch": t["work_per_launch"],
            "launches_per_step": t["launches_per_step"],
            "peak_source": peaks["source"] + (" (copy bandwidth)" if t["bound"] == "hbm" else
                                              " (burst bf16 GEMM / 2: kind::tf32 runs at half the bf16 rate; the cta_group::2 kernels exceed what cuBLAS "
                                              "reached -- the pipe's own TF32 ceiling is 1.09-1.16 PFLOP/s at 1.8-1.96 GHz, see sm__pipe_tensor_cycles_active "
                                              "in profiles/r2_ncu_full.csv)" if precision == "tf32"
                                              else " (burst bf16 GEMM; strict mode computes in fp32 FFMA)"),
            "traffic_note": "mean dram__bytes_read+write per launch of this kernel over one step of this workload (ncu per-launch capture, "
                            "profiles/r2_step_metrics_c3b.csv -> profiles/traffic.json); null for workloads without a committed capture"}


def conv_out_bytes(cfg, B):
    act = 0
    for specs, (c, h, w) in ((cfg["G"], (cfg["nc"], cfg["hr"] // 2, cfg["hr"] // 2)), (cfg["D"], (cfg["nc"], cfg["hr"], cfg["hr"]))):
        hh, ww = h, w
        for s in specs:
            if s["kind"] == "conv":
                hh = (hh + 2 * s["p"] - s["k"]) // s["s"] + 1
                ww = (ww + 2 * s["p"] - s["k"]) // s["s"] + 1
                act += 4 * B * s["cout"] * hh * ww
            elif s["kind"] == "fullconv":
                hh = (hh - 1) * s["s"] - 2 * s["p"] + s["k"]
                ww = (ww - 1) * s["s"] - 2 * s["p"] + s["k"]
                act += 4 * B * s["cout"] * hh * ww
            elif s["kind"] == "upnearest":
                hh *= 2
                ww *= 2
    return act


def run_workload(env, ctx, workload, B, steps, warmup, precision, legs=("staged", "e2e", "profile"), sample_clocks=False,
                 profile_steps=2, profile_out=""):
    """Builds the workload's nets on ctx, runs the requested legs, destroys the nets.  Returns a dict."""
    import torch

    import dcgan_super_resolution_b200 as dsr
    from dcgan_super_resolution_b200 import init, models
    cfg = models.config(workload)
    world, rank = env.world, env.rank
    # a workload may ask for a micro-batched generator (C5): netG is created for the micro-batch, the step still sees B samples
    gmb = cfg.get("g_microbatch", 0)
    gB = gmb if gmb and B > gmb and B % gmb == 0 else B
    G = dsr.Sequential.from_specs(cfg["G"]).cuda(ctx, (cfg["nc"], cfg["hr"] // 2, cfg["hr"] // 2), gB)
    # D holds 2B samples: the step then runs D(real) and D(fake) as one grouped pass (dcgansr.cu:step_body)
    D = dsr.Sequential.from_specs(cfg["D"]).cuda(ctx, (cfg["nc"], cfg["hr"], cfg["hr"]), 2 * B)
    G.set_params(init.weights_init(cfg["G"], 4321))
    D.set_params(init.weights_init(cfg["D"], 8765))
    scfg = dsr.make_step_cfg(**cfg["step"])
    NPOOL = 8 if B * cfg["nc"] * cfg["hr"] ** 2 * 4 <= (64 << 20) else (4 if gB == B else 2)
    pool = synth_batches(cfg, B, NPOOL, 1234 + rank)
    for i, b in enumerate(pool):
        dsr.stage_batch(ctx, D, b, i)
    ctx.synchronize()
    res = {"workload": workload_name(workload, cfg, B), "global_batch": world * B, "precision": precision, "steps": steps, "warmup": warmup,
           "generator_microbatch": gB if gB != B else None}
    flops = models.step_flops(cfg, B)
    res["algorithmic_gflop_per_step_per_gpu"] = flops / 1e9

    if "staged" in legs:
        # the library caches one CUDA graph per staged batch (the batch pointer is baked into the graph): visit every slot
        # once so no capture / instantiation falls into the timed region, then the W warm-up steps
        for i in range(NPOOL):
            dsr.train_step_staged(ctx, G, D, scfg, i, B)
        for i in range(warmup):
            dsr.train_step_staged(ctx, G, D, scfg, i % NPOOL, B)
        sampler = ClockSampler(env.local_rank)
        if sample_clocks and rank == 0 and not os.environ.get("DCGANSR_NO_SAMPLER"):
            sampler.start()
            time.sleep(0.25)
        env.barrier(ctx)               # AFTER the sampler start-up: every rank enters the timed region together
        l0 = ctx.launch_count()
        lo_mark = sampler.mark()
        ctx.timer_begin()
        for i in range(steps):
            dsr.train_step_staged(ctx, G, D, scfg, i % NPOOL, B)
        ms = ctx.timer_end()
        env.barrier(ctx)
        hi_mark = sampler.mark()
        res["gpu_launches"] = int(ctx.launch_count() - l0)
        ms = env.max_over_ranks(ms)
        if sample_clocks and rank == 0:
            res["clocks"] = sampler.stop(lo_mark, hi_mark)
        res["ms_per_step"] = ms / steps
        res["value"] = world * B * steps / (ms * 1e-3)
        res["step_tflops"] = world * flops / (ms / steps * 1e-3) / 1e12
        res["graph_prime_steps"] = NPOOL

    if "e2e" in legs:
        # pinned host batch in (H2D + NCHW -> NHWC inside the call), the three loss scalars out, every step
        pinned = [torch.from_numpy(b).pin_memory() for b in pool]
        losses = (ctypes.c_float * 3)()
        for i in range(2):
            dsr.nn.train_step_ptr(ctx, G, D, scfg, pinned[i % NPOOL].data_ptr(), B, losses)
        env.barrier(ctx)
        t0 = time.perf_counter()
        for i in range(steps):
            dsr.nn.train_step_ptr(ctx, G, D, scfg, pinned[i % NPOOL].data_ptr(), B, losses)
        ctx.synchronize()
        e2e_s = env.max_over_ranks(time.perf_counter() - t0)
        env.barrier(ctx)
        res["e2e"] = {"value": world * B * steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(pool[0].nbytes), "d2h_bytes_per_step": 12,
                      "ms_per_step": 1e3 * e2e_s / steps, "last_losses": [float(x) for x in losses]}
        del pinned

    if "profile" in legs and profile_steps > 0:
        # eager launches with per-launch events on the library stream (the graph is bypassed while profiling)
        ctx.profile_begin()
        for i in range(profile_steps):
            dsr.train_step_staged(ctx, G, D, scfg, i % NPOOL, B)
        prof = ctx.profile_end()
        env.barrier(ctx)
        pdir = os.environ.get("DCGANSR_PROFILE_DIR")
        if not profile_out and pdir:
            profile_out = os.path.join(pdir, f"prof_{workload}_b{B}_{precision}_dp{world}.json")
        if profile_out and rank == 0:
            with open(profile_out, "w") as f:
                json.dump({"workload": workload, "batch": B, "profile_steps": profile_steps, "kernels": prof}, f, indent=1)
        peaks = load_peaks()
        top, eager_ms = aggregate_profile(prof, profile_steps, peaks, precision)
        res["kernels"] = top[:10]
        res["roofline"] = roofline_of(top, peaks, precision, workload, B)
        res["eager_profile_ms_per_step"] = eager_ms
        res["launches_per_step_eager"] = sum(t["launches_per_step"] for t in top)
    res["conv_out_mb_per_rank"] = conv_out_bytes(cfg, B) / 1e6
    G.close()
    D.close()
    return res


def sync_bn_timing(env, workload="C4", B=64, steps=20):
    """ms per sync_bn=1 step of a BatchNorm-heavy small-step workload, statistics through the peer-memory kernel vs ncclAllReduce."""
    out = {"workload": f"{workload} at {B} per GPU, sync_bn=1, CUDA graph, {steps} steps"}
    for name, peer in (("nccl", False), ("peer", True)) * 2:
        name = name + ("_again" if name in out else "")
        if not peer:
            os.environ["DCGANSR_PEER_AR"] = "0"
        try:
            ctx = env.new_ctx("tf32", graph=True, sync_bn=True)
            r = run_workload(env, ctx, workload, B, steps, 3, "tf32", legs=("staged",))
            ctx.close()
            out[name] = r["ms_per_step"]
        finally:
            os.environ.pop("DCGANSR_PEER_AR", None)
    return out


def dp_check(env, quick=False):
    """The 2-GPU checks of scripts/dp_check.py at whatever N the run has (the driver's pytest box has one GPU).
    quick: only the sync_bn / peer-memory all-reduce part."""
    import numpy as np

    import dcgan_super_resolution_b200 as dsr
    from dcgan_super_resolution_b200 import init, models
    cfg = models.config("C2")
    B = 8
    specsG, specsD = models.train_gray_G(8), models.dcgan64_D(1, 8)
    step = dsr.make_step_cfg(**cfg["step"])
    rng = np.random.Generator(np.random.Philox(99))
    full = [rng.uniform(-1, 1, (env.world * B, 1, 64, 64)).astype(np.float32) for _ in range(3)]

    def run(world, precision, sync_bn, no_overlap=False, batch=B, shard=True, graph=False, peer=True, info=None):
        if no_overlap:
            os.environ["DCGANSR_NO_OVERLAP"] = "1"
        if not peer:
            os.environ["DCGANSR_PEER_AR"] = "0"
        try:
            ctx = env.new_ctx(precision, graph=graph, sync_bn=sync_bn, world=world)
            if info is not None:
                info["peer"] = ctx.comm_peer_enabled()
            G = dsr.Sequential.from_specs(specsG).cuda(ctx, (1, 32, 32), batch)
            D = dsr.Sequential.from_specs(specsD).cuda(ctx, (1, 64, 64), 2 * batch)
            G.set_params(init.weights_init(specsG, 4321))
            D.set_params(init.weights_init(specsD, 8765))
            losses = []
            for x in full:
                xb = x[env.rank * batch:(env.rank + 1) * batch] if shard else x
                losses.append(dsr.train_step(ctx, G, D, step, xb))
            out = (G.get_params(), D.get_params(), losses)
            G.close()
            D.close()
            ctx.close()
            return out
        finally:
            os.environ.pop("DCGANSR_NO_OVERLAP", None)
            os.environ.pop("DCGANSR_PEER_AR", None)

    res = {"world": env.world, "batch_per_rank": B}
    for prec in (() if quick else ("strict", "tf32")):
        a = run(env.world, prec, False)
        b = run(env.world, prec, False, no_overlap=True)
        c = run(env.world, prec, False, graph=True)
        res[f"bucket_bit_identical_{prec}"] = bool(np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]))
        # (bit identity holds at N = 2, where a sum of two terms has one order; from N = 4 up NCCL may pick another algorithm --
        #  another summation order -- for the bucket-sized messages than for the whole gradient vector: report the size of it)
        res[f"bucket_max_rel_diff_{prec}"] = float(max(np.max(np.abs(a[0] - b[0])) / np.max(np.abs(b[0])), np.max(np.abs(a[1] - b[1])) / np.max(np.abs(b[1]))))
        res[f"graph_replay_identical_{prec}"] = bool(np.array_equal(a[0], c[0]) and np.array_equal(a[1], c[1]))
    info = {}
    a = run(env.world, "strict", True, info=info)
    # sync_bn statistics: one-shot NVLink peer-memory all-reduce fused into the BatchNorm statistics tail (kernels_peer.cu) against
    # ncclAllReduce of the same sums (DCGANSR_PEER_AR=0), eager and under graph replay (the call counter lives on the device)
    res["sync_bn_peer_allreduce"] = bool(info.get("peer"))
    if info.get("peer"):
        n = run(env.world, "strict", True, peer=False)
        g = run(env.world, "strict", True, graph=True)
        res["sync_bn_peer_vs_nccl_max_rel_diff"] = float(max(np.max(np.abs(a[0] - n[0])) / np.max(np.abs(n[0])),
                                                             np.max(np.abs(a[1] - n[1])) / np.max(np.abs(n[1]))))
        res["sync_bn_peer_graph_replay_identical"] = bool(np.array_equal(a[0], g[0]) and np.array_equal(a[1], g[1]))
        res["sync_bn_step_ms"] = sync_bn_timing(env)
    if env.rank == 0:
        ref = run(1, "strict", False, batch=env.world * B, shard=False)
        eG = float(np.max(np.abs(a[0] - ref[0])) / np.max(np.abs(ref[0])))
        eD = float(np.max(np.abs(a[1] - ref[1])) / np.max(np.abs(ref[1])))
        res["sync_bn_rel_err"] = max(eG, eD)
        res["sync_bn_losses"] = [float(x) for x in a[2][-1]]
        res["single_gpu_losses"] = [float(x) for x in ref[2][-1]]
    env.barrier()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=HEADLINE)
    ap.add_argument("--precision", default=os.environ.get("DCGANSR_PRECISION", "tf32"), choices=["strict", "tf32"])
    ap.add_argument("--graph", type=int, default=1, help="replay the step as a CUDA graph (one cached graph per staged batch)")
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch override")
    ap.add_argument("--sync-bn", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the workloads[] / strict / dp_check legs")
    ap.add_argument("--extra", default="", help="comma-separated workloads for workloads[] (default: the BASELINE configurations)")
    ap.add_argument("--cpu-sample-batch", type=int, default=0, help="batch of the CPU arms (bounded sample of the workload)")
    ap.add_argument("--legs", default="staged,e2e,profile", help="legs of the headline workload (C5: staged,profile keeps the run short)")
    ap.add_argument("--profile-steps", type=int, default=2)
    ap.add_argument("--profile-out", default="", help="write the full per-kernel profile table (JSON) here")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    if args.warmup < 3:
        args.warmup = 3

    from dcgan_super_resolution_b200 import models

    env = Env()
    rank, world = env.rank, env.world
    if world != args.gpus and world > 1:
        args.gpus = world
    cfg = models.config(args.workload)
    B = args.batch or cfg["batch"]

    ctx = env.new_ctx(args.precision, graph=bool(args.graph), sync_bn=bool(args.sync_bn))
    head = run_workload(env, ctx, args.workload, B, args.steps, args.warmup, args.precision, legs=tuple(args.legs.split(",")),
                        sample_clocks=True, profile_steps=args.profile_steps, profile_out=args.profile_out)

    extras, strict, dpc = [], None, None
    if not args.no_extra:
        esteps = max(3, min(args.steps, 10))
        if args.extra:
            todo = [(w, 0, "weak") for w in args.extra.split(",") if w]
        elif world == 1:
            todo = [("C1b", 0, "weak"), ("C2", 0, "weak"), ("C3a", 0, "weak"), ("C4", 64, "weak")]
        else:   # the BASELINE data-parallel wordings: batchSize 128 / 512 sharded over the GPUs
            todo = [("C3b", max(1, 128 // world), "strong"), ("C4", max(1, 512 // world), "strong")]
        for w, b, scal in todo:
            wc = models.config(w)
            r = run_workload(env, ctx, w, b or wc["batch"], esteps, 3, args.precision, legs=("staged", "profile"), profile_steps=1)
            r["scaling"] = scal
            extras.append(r)
    ctx.close()
    if not args.no_extra and args.precision != "strict":
        sctx = env.new_ctx("strict", graph=bool(args.graph), sync_bn=bool(args.sync_bn))
        strict = run_workload(env, sctx, args.workload, B, max(2, min(args.steps, 5)), 3, "strict", legs=("staged", "profile"), profile_steps=1)
        sctx.close()

    def make_line(dpc, cpu):
        return {
            "metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "tf32" if args.precision == "tf32" else "f32", "data": "synthetic",
            "config": make_config(args.workload, cfg, B, world),
            "run": {"precision": args.precision, "cuda_graph": bool(args.graph), "graph_prime_steps": head.get("graph_prime_steps"),
                    "sync_bn": bool(args.sync_bn), "generator_microbatch": head.get("generator_microbatch"), "control_plane": "gloo (the only NCCL communicators are libdcgansr's)" if world > 1 else "single process",
                    "l2": f"no flush: per-step working set (conv outputs {head['conv_out_mb_per_rank']:.0f} MB fp32 per rank, rotating staged input "
                          "batches) exceeds the 126 MB L2",
                    "algorithmic_gflop_per_step_per_gpu": head["algorithmic_gflop_per_step_per_gpu"]},
            "step_tflops": head["step_tflops"],
            "clocks": head.get("clocks"),
            "e2e": head.get("e2e"),
            "gpu_launches": head["gpu_launches"],
            "roofline": head.get("roofline"),
            "kernels": head.get("kernels"),
            "eager_profile_ms_per_step": head.get("eager_profile_ms_per_step"),
            "strict": None if strict is None else {k: strict.get(k) for k in ("value", "ms_per_step", "step_tflops", "steps", "gpu_launches", "roofline")},
            "workloads": [{k: r.get(k) for k in ("workload", "global_batch", "scaling", "steps", "ms_per_step", "value", "step_tflops", "gpu_launches",
                                                 "roofline", "eager_profile_ms_per_step")} for r in extras],
            "dp_check": dpc,
            "cpu_baseline": cpu,
        }

    if not args.no_extra and world > 1:
        # the checks run behind the measurements; should they ever raise or hang (a lost rank), the bench line still goes out
        def _give_up():
            if rank == 0:
                print(json.dumps(make_line({"error": "dp_check did not finish within 300 s"}, None)))
                sys.stdout.flush()
            os._exit(0)
        wd = threading.Timer(300.0, _give_up)
        wd.daemon = True
        wd.start()
        try:
            dpc = dp_check(env)
        except Exception as e:   # noqa: BLE001 -- reported in the line
            dpc = {"error": f"{type(e).__name__}: {e}"[:400]}
        wd.cancel()

    # every rank has destroyed its contexts (and with them the library's NCCL communicators); ranks != 0 leave before the
    # CPU leg starts, through the normal interpreter exit
    env.close()
    if rank != 0:
        return 0

    cpu = None if args.no_cpu_baseline else cpu_baseline(args.workload, B)
    print(json.dumps(make_line(dpc, cpu)))
    sys.stdout.flush()
    return 0


def _exit_watchdog(seconds=120.0):
    """The process leaves through the normal interpreter exit (atexit hooks run).  Should a finaliser ever hang, this daemon
    timer ends the process instead of the driver's time limit."""
    t = threading.Timer(seconds, lambda: os._exit(0))
    t.daemon = True
    t.start()


if __name__ == "__main__":
    rc = main()
    _exit_watchdog()
    sys.exit(rc)
