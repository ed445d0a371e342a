"""Turn the ncu exports brought back in gpurun_out/ into the small, tracked summaries under profiles/.

  python scripts/summarize_profiles.py <tag> <launches.csv> <full_raw.csv> [<full_raw2.csv> ...]

Writes profiles/<tag>_launches_step.csv (every launch of ONE training step with its ncu gpu__time_duration),
profiles/<tag>_launch_summary.md (per-kernel share of the step), profiles/<tag>_ncu_full.csv (key `--set full`
metrics per captured launch) and profiles/traffic.json (dram bytes per launch of each kernel, read by bench.py)."""
import collections, csv, json, os, re, sys

tag, launches = sys.argv[1], sys.argv[2]
fulls = sys.argv[3:]
out = os.path.join(os.path.dirname(__file__), "..", "profiles")
os.makedirs(out, exist_ok=True)


def short(name):
    name = re.sub(r"^void ", "", name)
    return re.sub(r"\(.*", "", name)


rows = [r for r in csv.reader(open(launches)) if len(r) > 10 and r[0].isdigit()]
names = [short(r[4]) for r in rows]
adam = [i for i, n in enumerate(names) if n == "adam_kernel"]
# a step = ... adam(D) ... adam(G); take the last complete device-resident step before the e2e leg: between G-adams
a, b = adam[-7] + 1, adam[-5] + 1
# skip the weight repack launches that belong to the previous step's adam(G)
while a < b and names[a].startswith("pack_"):
    a += 1
step = rows[a:b]
tot = sum(float(r[-1]) for r in step)
with open(os.path.join(out, f"{tag}_launches_step.csv"), "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["kernel", "grid", "block", "gpu__time_duration_ns"])
    for r in step:
        w.writerow([short(r[4]), r[8], r[7], r[-1]])
agg = collections.OrderedDict()
for r in step:
    k = short(r[4])
    agg.setdefault(k, [0, 0.0])
    agg[k][0] += 1
    agg[k][1] += float(r[-1])
with open(os.path.join(out, f"{tag}_launch_summary.md"), "w") as f:
    f.write(f"# ncu launch list of one training step ({tag})\n\n")
    f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` over `python bench.py --steps 2 --warmup 3 --no-cpu-baseline "
            "--profile-steps 0` (C2, batch 64, FAST_TF32).  Per-launch times under ncu are cold-cache and serialised: compare SHARES.\n\n")
    f.write(f"{len(step)} launches, {tot / 1e3:.1f} us summed\n\n| kernel | launches | us | share |\n|---|---:|---:|---:|\n")
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        f.write(f"| `{k}` | {n} | {t / 1e3:.1f} | {100 * t / tot:.1f} % |\n")

cols = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum"]
traffic = {}
with open(os.path.join(out, f"{tag}_ncu_full.csv"), "w", newline="") as f:
    w = csv.writer(f)
    first = True
    for path in fulls:
        rr = list(csv.reader(open(path)))
        ci = {h: i for i, h in enumerate(rr[0])}
        use = [c for c in cols if c in ci]
        if first:
            w.writerow(use)
            w.writerow([rr[1][ci[c]] for c in use])
            first = False
        for r in rr[2:]:
            w.writerow([short(r[ci[c]]) if c == "Kernel Name" else r[ci[c]] for c in use])
            k = short(r[ci["Kernel Name"]]).replace("_kernel", "")
            unit_r, unit_w = rr[1][ci["dram__bytes_read.sum"]], rr[1][ci["dram__bytes_write.sum"]]
            mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            tb = float(r[ci["dram__bytes_read.sum"]]) * mul[unit_r] + float(r[ci["dram__bytes_write.sum"]]) * mul[unit_w]
            traffic.setdefault(k, []).append({"grid": r[ci["Grid Size"]], "dram_bytes": tb, "us": float(r[ci["gpu__time_duration.sum"]])})
# bench.py keys its kernels by name + work; keep the heaviest launch per kernel as the headline traffic figure
json.dump({k: max(v, key=lambda x: x["us"]) for k, v in traffic.items()} | {"_all": traffic}, open(os.path.join(out, "traffic.json"), "w"), indent=1)
print("wrote", sorted(os.listdir(out)))
