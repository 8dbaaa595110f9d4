"""Per-source-line totals of an ncu report's source page (warp instructions, shared-memory wavefronts, stall samples), grouped
by line ranges.  usage: python tools/ncu_phase.py report.ncu-rep source.cu [begin-end:name ...]"""
import collections
import csv
import subprocess
import sys

rep, srcf = sys.argv[1], sys.argv[2]
ranges = []
for a in sys.argv[3:]:
    r, name = a.split(":")
    b, e = r.split("-")
    ranges.append((int(b), int(e), name))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
agg = collections.defaultdict(lambda: collections.Counter())
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
        continue
    if cur is None:
        continue
    def col(name):
        try:
            return float(r[hdr.index(name) + sh] or 0)
        except Exception:
            return 0.0
    a = agg[cur]
    a["inst"] += col("Instructions Executed")
    a["wf"] += col("L1 Wavefronts Shared")
    a["wf_ideal"] += col("L1 Wavefronts Shared Ideal")
    a["samples"] += col("# Samples")
    for k in hdr:
        if k.startswith("stall_") and "Not Issued" not in k:
            a[k] += col(k)
tot = collections.Counter()
for a in agg.values():
    tot.update(a)
print("total: warp inst %.0f, smem wavefronts %.0f (ideal %.0f), samples %.0f" % (tot["inst"], tot["wf"], tot["wf_ideal"], tot["samples"]))
src = open(srcf).read().split("\n")
if ranges:
    for b, e, name in ranges:
        c = collections.Counter()
        for ln, a in agg.items():
            if b <= ln <= e:
                c.update(a)
        st = sorted(((v, k) for k, v in c.items() if k.startswith("stall_")), reverse=True)[:4]
        print("%-18s lines %4d-%4d: inst %5.1f%%  wavefronts %5.1f%% (x%.2f ideal)  samples %5.1f%%  %s" % (
            name, b, e, 100 * c["inst"] / tot["inst"], 100 * c["wf"] / max(tot["wf"], 1), c["wf"] / max(c["wf_ideal"], 1), 100 * c["samples"] / tot["samples"],
            ", ".join("%s %.0f%%" % (k[6:], 100 * v / max(c["samples"], 1)) for v, k in st)))
else:
    for ln, a in sorted(agg.items(), key=lambda kv: -kv[1]["wf"])[:40]:
        print("%5d inst %5.1f%% wf %5.1f%% (x%.2f) samp %5.1f%% | %s" % (ln, 100 * a["inst"] / tot["inst"], 100 * a["wf"] / max(tot["wf"], 1), a["wf"] / max(a["wf_ideal"], 1),
                                                                     100 * a["samples"] / tot["samples"], src[ln - 1].strip()[:90] if 0 < ln <= len(src) else "?"))
