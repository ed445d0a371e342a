"""Print the hottest SASS instructions of an `ncu --page source --csv` export:  python scripts/ncu_hot.py file.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
sc = ci["# Samples"]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for idx, r in enumerate(rows[2:]):
    if len(r) <= sc:
        continue
    try:
        data.append((int(r[sc] or 0), idx, r))
    except ValueError:
        pass
tot = sum(d[0] for d in data)
print("total samples", tot, "instructions", len(data))
for n, idx, r in sorted(data, key=lambda x: -x[0])[:N]:
    top = sorted(((int(r[ci[s]] or 0), s) for s in stalls), reverse=True)[:2]
    print(f"{n:6d} {100.0 * n / max(tot, 1):5.1f}%  #{idx:4d} exec={r[ci['Instructions Executed']]:>8s} {r[ci['Source']].strip()[:70]:70s} {top}")
