"""Stall-reason breakdown of an ncu report for named source-line regions of gomp_kernel.cuh.
usage: python tools/ncu_regions.py report.ncu-rep name=a-b [name=a-b ...]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
agg = collections.defaultdict(collections.Counter)
inst = collections.Counter()
for r in rows:
    if r and r[0] == "Line No":
        hdr = r
        stall_idx = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None or not r:
        continue
    sh = len(r) - len(hdr)
    if r[0].strip().isdigit():
        ln = int(r[0])
        for i in stall_idx:
            try:
                agg[ln][hdr[i]] += int(r[i + sh])
            except Exception:
                pass
        try:
            inst[ln] += int(r[7 + sh])
        except Exception:
            pass
total = sum(sum(c.values()) for c in agg.values())
print("total samples", total)
for spec in sys.argv[2:]:
    name, rng = spec.split("=")
    a, b = (int(v) for v in rng.split("-"))
    c = collections.Counter()
    n = 0
    for l in range(a, b + 1):
        c.update(agg[l])
        n += inst[l]
    tot = sum(c.values()) or 1
    print(f"{name:22s} lines {a}-{b}: {100 * tot / total:5.1f}% of samples, {n} warp-instr, " +
          ", ".join(f"{k[6:]} {100 * v / tot:.0f}%" for k, v in c.most_common(6)))
