"""Per-kernel SASS mnemonic counts of libgomp_cuda.so (cuobjdump -sass): the instructions that prove what the kernels are built
from -- UBLKCP (1-D bulk TMA copies), SYNCS (mbarrier), UTCBAR / TCGEN05-class tensor-memory instructions (the two-slot kernel
keeps its factor in TMEM), DFMA / DMUL / DADD (fp64 pipe), LDS / STS (shared memory), SHFL, BAR.
usage: python tools/sass_counts.py [lib] > profiles/r02_sass_counts.txt"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "osqp-solver_b200", "libgomp_cuda.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
WATCH = ["UBLKCP", "UTMALDG", "SYNCS", "LDTM", "STTM", "UTCBAR", "UTCHMMA", "UTCQMMA", "DFMA", "DMUL", "DADD", "MUFU.RCP64H", "DMMA", "LDS", "STS", "LDG", "STG", "SHFL",
         "BAR", "ATOMG", "RED"]
arch, kernels, cur = None, collections.OrderedDict(), None
for line in out.splitlines():
    m = re.match(r"\s*arch = (\S+)", line)
    if m:
        arch = m.group(1)
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = kernels.setdefault(m.group(1), collections.Counter())
        cur["__arch__"] = arch
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur is not None:
        op = m.group(1)
        cur["__total__"] += 1
        for w in WATCH:
            if op == w or op.startswith(w + "."):
                cur[w] += 1
print(f"# {os.path.basename(lib)}: SASS mnemonic counts per kernel (cuobjdump -sass, CUDA 12.9); every cubin in the library is {sorted({c['__arch__'] for c in kernels.values()})}")
for name, c in kernels.items():
    if c["__total__"] < 200:
        continue
    print(f"\n{demangle(name)}\n  arch {c['__arch__']}, {c['__total__']} instructions")
    print("  " + "  ".join(f"{w} {c[w]}" for w in WATCH if c[w]))
