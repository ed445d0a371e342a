"""Turn gpurun_out/r2/ (written by scripts/capture_r2.sh on a B200) into the small, tracked summaries under profiles/.

  python scripts/summarize_r2.py --last-step <launches.csv>     -> "skip count" of the last training step in an ncu launch list
  python scripts/summarize_r2.py [gpurun_out/r2]                -> profiles/r2_*.{json,csv,md} + profiles/traffic.json
"""
import collections, csv, json, os, re, sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
OUT = os.path.join(ROOT, "profiles")


def short(name):
    name = re.sub(r"^void ", "", name)
    return re.sub(r"[<(].*", "", name)


def read_long_csv(path):
    """ncu --csv with --metrics: one row per (launch, metric).  -> ordered {launch id: {"name", "grid", "block", metric: (value, unit)}}"""
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    hdr = next(r for r in rows if r[0] == "ID")
    ci = {h: i for i, h in enumerate(hdr)}
    by = collections.OrderedDict()
    for r in rows:
        if not r[0].isdigit():
            continue
        d = by.setdefault(int(r[0]), {"name": short(r[ci["Kernel Name"]]), "grid": r[ci["Grid Size"]], "block": r[ci["Block Size"]]})
        d[r[ci["Metric Name"]]] = (r[ci["Metric Value"]], r[ci["Metric Unit"]])
    return by


def last_step(by):
    ids = list(by.keys())
    names = [by[i]["name"] for i in ids]
    adam = [k for k, n in enumerate(names) if n == "adam_kernel"]
    # a step ends with adam(G) + its weight repack; the step before it ended at adam[-3] (+ repack)
    end = adam[-1] + (2 if adam[-1] + 1 < len(names) and names[adam[-1] + 1].startswith("pack_all") else 1)
    beg = adam[-3] + (2 if names[adam[-3] + 1].startswith("pack_all") else 1)
    return beg, end - beg


def num(v):
    try:
        return float(v[0].replace(",", ""))
    except Exception:
        return None


def main():
    if len(sys.argv) >= 3 and sys.argv[1] == "--last-step":
        beg, cnt = last_step(read_long_csv(sys.argv[2]))
        print(beg, cnt)
        return
    src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "r2")
    os.makedirs(OUT, exist_ok=True)
    # ---- bench lines
    for f, dst in (("bench.json", "r2_bench.json"), ("bench_reference.json", "r2_bench_reference.json"), ("bench_c5.json", "r2_bench_c5.json")):
        p = os.path.join(src, f)
        if os.path.exists(p):
            lines = [l for l in open(p) if l.startswith("{")]
            if lines:
                open(os.path.join(OUT, dst), "w").write(lines[-1])
    # ---- launch list of the last step
    lp = os.path.join(src, "launches_c3b.csv")
    if os.path.exists(lp):
        by = read_long_csv(lp)
        beg, cnt = last_step(by)
        ids = list(by.keys())[beg:beg + cnt]
        tot = sum(num(by[i]["gpu__time_duration.sum"]) for i in ids)
        with open(os.path.join(OUT, "r2_launches_step.csv"), "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(["kernel", "grid", "block", "gpu__time_duration_ns"])
            for i in ids:
                w.writerow([by[i]["name"], by[i]["grid"], by[i]["block"], by[i]["gpu__time_duration.sum"][0]])
        agg = collections.OrderedDict()
        for i in ids:
            a = agg.setdefault(by[i]["name"], [0, 0.0])
            a[0] += 1
            a[1] += num(by[i]["gpu__time_duration.sum"])
        with open(os.path.join(OUT, "r2_launch_summary.md"), "w") as f:
            f.write("# ncu launch list of one training step (r2)\n\n"
                    "`ncu --metrics gpu__time_duration.sum --clock-control none` over `python bench.py --steps 1 --warmup 3 --no-cpu-baseline "
                    "--no-extra --profile-steps 0 --legs staged` (headline workload C3b: train.lua 64x64 -> 128x128, batch 128, FAST_TF32; the last "
                    "step of the run).  Per-launch times under ncu are cold-cache and serialised: compare SHARES with the event profile in "
                    "r2_bench.json (`kernels[]`).\n\n")
            f.write(f"{len(ids)} launches, {tot / 1e6:.3f} ms summed\n\n| kernel | launches | us | share |\n|---|---:|---:|---:|\n")
            for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
                f.write(f"| `{k}` | {n} | {t / 1e3:.1f} | {100 * t / tot:.1f} % |\n")
    # ---- per-launch metrics of that step
    keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum",
            "sm__cycles_elapsed.avg.per_second"]
    traffic = {}
    mul = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}

    def dump_metrics(path, dst, tag):
        by = read_long_csv(path)
        with open(os.path.join(OUT, dst), "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(["kernel", "grid", "block"] + keys + ["dram_GBps", "source"])
            for i, d in by.items():
                row = [d["name"], d["grid"], d["block"]]
                for k in keys:
                    v = d.get(k)
                    row.append(f"{v[0]} {v[1]}".strip() if v else "")
                rb, wb, t = d.get("dram__bytes_read.sum"), d.get("dram__bytes_write.sum"), d.get("gpu__time_duration.sum")
                gbps = ""
                if rb and wb and t:
                    tb = num(rb) * mul.get(rb[1], 1.0) + num(wb) * mul.get(wb[1], 1.0)
                    ns = num(t) * {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(t[1], 1.0)
                    gbps = f"{tb / ns:.0f}"
                    traffic.setdefault(d["name"].replace("_kernel", ""), []).append({"grid": d["grid"], "dram_bytes": tb, "us": ns / 1e3, "from": tag})
                row += [gbps, tag]
                w.writerow(row)

    sp = os.path.join(src, "step_metrics_c3b.csv")
    if os.path.exists(sp):
        dump_metrics(sp, "r2_step_metrics_c3b.csv", "C3b step")
    for extra in sorted(os.listdir(src)):
        if extra.startswith("layers_") and extra.endswith("_metrics.csv"):
            dump_metrics(os.path.join(src, extra), "r2_" + extra, extra[len("layers_"):-len("_metrics.csv")])
    # ---- --set full captures (raw pages, wide format: one row per launch)
    cols = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
            "smsp__inst_executed.sum", "sm__inst_executed_pipe_tensor.sum", "launch__cluster_dim_x"]
    fulls = sorted(f for f in os.listdir(src) if f.startswith("full_") and f.endswith(".csv"))
    if fulls:
        with open(os.path.join(OUT, "r2_ncu_full.csv"), "w", newline="") as f:
            w = csv.writer(f)
            first = True
            for fn in fulls:
                rr = [r for r in csv.reader(open(os.path.join(src, fn), errors="replace")) if len(r) > 5]
                if len(rr) < 3:
                    continue
                ci = {h: i for i, h in enumerate(rr[0])}
                use = [c for c in cols if c in ci]
                if first:
                    w.writerow(["capture"] + use)
                    w.writerow(["(unit)"] + [rr[1][ci[c]] for c in use])
                    first = False
                for r in rr[2:]:
                    w.writerow([fn[len("full_"):-4]] + [short(r[ci[c]]) if c == "Kernel Name" else r[ci[c]] for c in use])
                    try:
                        ur, uw = rr[1][ci["dram__bytes_read.sum"]], rr[1][ci["dram__bytes_write.sum"]]
                        tb = float(r[ci["dram__bytes_read.sum"]].replace(",", "")) * mul.get(ur, 1.0) + float(r[ci["dram__bytes_write.sum"]].replace(",", "")) * mul.get(uw, 1.0)
                        ut = rr[1][ci["gpu__time_duration.sum"]]
                        us = float(r[ci["gpu__time_duration.sum"]].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(ut, 1e-3)
                        traffic.setdefault(short(r[ci["Kernel Name"]]).replace("_kernel", ""), []).append(
                            {"grid": r[ci["Grid Size"]], "dram_bytes": tb, "us": us, "from": "--set full " + fn[len("full_"):-4]})
                    except Exception:
                        pass
    if traffic:
        # bench.py's roofline.traffic: dram bytes of the heaviest captured launch of the dominant kernel
        json.dump({k: max(v, key=lambda x: x["us"]) for k, v in traffic.items()} | {"_all": traffic}, open(os.path.join(OUT, "traffic.json"), "w"), indent=1)
    for fn in os.listdir(src):
        if fn.startswith("full_") and fn.endswith(".ncu-rep"):
            os.replace(os.path.join(src, fn), os.path.join(OUT, "r2_" + fn))
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
