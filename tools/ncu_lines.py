"""Aggregate an ncu report's source page by CUDA source line: stall samples and executed instructions.
usage: python tools/ncu_lines.py report.ncu-rep [source-file] [top]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
srcf = sys.argv[2] if len(sys.argv) > 2 else "osqp-solver_b200/csrc/gomp_kernel.cuh"
top = int(sys.argv[3]) if len(sys.argv) > 3 else 45
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
src = open(srcf).read().split("\n")
hdr = None
d = collections.defaultdict(lambda: [0, 0])
sass = collections.defaultdict(list)
cur = None
for r in rows:
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or not r:
        continue
    sh = len(r) - len(hdr)
    if r[0].strip().isdigit():
        cur = int(r[0])
        try:
            d[cur][0] += int(r[6 + sh]); d[cur][1] += int(r[7 + sh])
        except Exception:
            pass
    elif cur is not None and len(r) > 7:
        try:
            smp = int(r[6 + sh]); ins = int(r[7 + sh])
            sass[cur].append((smp, ins, r[3 + sh].strip()))
        except Exception:
            pass
ts = sum(v[0] for v in d.values()) or 1
ti = sum(v[1] for v in d.values()) or 1
print(f"total samples {ts}, warp instructions {ti}")
for ln, (s, i) in sorted(d.items(), key=lambda kv: -kv[1][0])[:top]:
    text = src[ln - 1].strip()[:100] if 0 < ln <= len(src) else "?"
    print(f"{ln:5d} {100 * s / ts:5.1f}% samp {100 * i / ti:5.1f}% inst | {text}")
if len(sys.argv) > 4:
    ln = int(sys.argv[4])
    print(f"--- SASS of line {ln}")
    for smp, ins, t in sass[ln][:80]:
        print(f"   {smp:7d} {ins:9d}  {t}")
