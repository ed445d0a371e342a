"""Per-kernel counts of the SASS mnemonics that prove the Blackwell-native paths (B200_PROFILING.md): UTCHMMA (tcgen05.mma),
UTMALDG / UTMASTG / UTMAPF / UBLKCP (TMA), LDTM / STTM (tcgen05.ld / st), STG.*.256, SYS (system-scope STG / LDG / MEMBAR: the NVLink peer-memory exchange), HMMA (legacy, must be 0).
  python scripts/sass_opcodes.py > profiles/sass_opcodes.txt        (CPU only: cuobjdump on the built library)"""
import collections, os, re, subprocess, sys
lib = os.path.join(os.path.dirname(__file__), "..", "dcgan_super_resolution_b200", "libdcgansr.so")
out = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True, check=True).stdout
pats = collections.OrderedDict([("UTCHMMA", r"\bUTC[A-Z]*MMA"), ("UTCBAR", r"\bUTCBAR"), ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"),
                                ("UTMAPF", r"\bUTMAPF"), ("UBLKCP", r"\bUBLKCP"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"),
                                ("STG.256", r"\bSTG\.[A-Z0-9.]*256"), ("SYNCS", r"\bSYNCS"), ("SYS", r"\b(STG|LDG|MEMBAR)\.[A-Z0-9.]*SYS"), ("HMMA", r"\bHMMA"), ("FFMA", r"\bFFMA")])
rows, cur, cnt = [], None, None
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        if cur:
            rows.append((cur, cnt))
        cur, cnt = m.group(1), collections.Counter()
        continue
    if cur:
        for k, p in pats.items():
            if re.search(p, ln):
                cnt[k] += 1
if cur:
    rows.append((cur, cnt))
dem = subprocess.run(["c++filt"], input="\n".join(r[0] for r in rows), stdout=subprocess.PIPE, text=True).stdout.splitlines()
print("# SASS opcode counts per kernel of libdcgansr.so (cuobjdump -sass, sm_100a); legacy HMMA must be 0 everywhere")
print("# " + " ".join(f"{k:>8s}" for k in pats) + "  kernel")
tot = collections.Counter()
for (name, c), d in zip(rows, dem):
    tot.update(c)
    if any(c[k] for k in pats if k != "FFMA"):
        d = re.sub(r"\(.*", "", d.replace("(anonymous namespace)::", ""))
        print("  " + " ".join(f"{c[k]:8d}" for k in pats) + "  " + d)
print("  " + " ".join(f"{tot[k]:8d}" for k in pats) + "  TOTAL (all %d kernels)" % len(rows))
