"""Opcode mix (warp instructions executed) of the profiled kernel from a .ncu-rep.  Development aid."""
import csv
import subprocess
import sys
from collections import Counter

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]
ie = hdr.index("Instructions Executed")
c = Counter()
for r in rows[hi + 1:]:
    try:
        n = int(r[ie])
    except (ValueError, IndexError):
        continue
    toks = r[1].split()
    op = toks[1] if toks and toks[0].startswith("@") else toks[0]
    c[op.split(".")[0]] += n
tot = sum(c.values())
print("total", tot)
for op, n in c.most_common(30):
    print(f"{op:10s} {n:12d} {100.0 * n / tot:5.1f}%")
