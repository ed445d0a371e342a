"""Per-kernel profile of one training step of a workload (CUDA events per launch, eager):  python scripts/step_profile.py C3a"""
import json, os, subprocess, sys, tempfile
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
wl = sys.argv[1] if len(sys.argv) > 1 else "C2"
out = os.path.join(tempfile.gettempdir(), f"prof_{wl}.json")
r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--workload", wl, "--no-cpu-baseline", "--profile-steps", "2", "--profile-out", out,
                    "--steps", "10"], capture_output=True, text=True)
line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
d = json.loads(line)
print(wl, "ms/step", round(d["ms_per_step"], 3), "patches/s", round(d["value"]), "e2e", round(d["e2e"]["value"]))
prof = json.load(open(out))
ks = prof["kernels"]
tot = sum(k["us_total"] for k in ks) if ks and "us_total" in ks[0] else None
print(list(ks[0].keys()) if ks else None)
for k in sorted(ks, key=lambda k: -k.get("us_total", 0))[:22]:
    print({a: (round(b, 2) if isinstance(b, float) else b) for a, b in k.items()})
