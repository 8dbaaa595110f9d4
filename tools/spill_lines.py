"""List local-memory (spill) instructions of a cubin by CUDA source line.
usage: python tools/spill_lines.py file.cubin   (cubin from `cuobjdump -xelf all object.o`)"""
import collections
import re
import subprocess
import sys

out = subprocess.run(["nvdisasm", "-g", "-c", sys.argv[1]], capture_output=True, text=True).stdout
cur, cnt = None, collections.Counter()
for line in out.splitlines():
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.search(r"\b(STL|LDL)\b", line)
    if m:
        cnt[(cur, m.group(1))] += 1
for (loc, op), c in sorted(cnt.items(), key=lambda kv: (kv[0][0] or ("", 0))):
    print(loc, op, c)
