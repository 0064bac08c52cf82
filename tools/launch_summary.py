"""Markdown table of per-kernel device time from an ncu launch list
(ncu --metrics gpu__time_duration.sum --csv --log-file list.csv ...).
usage: python tools/launch_summary.py list.csv > summary.md"""
import csv
import sys
from collections import OrderedDict

rows = []
with open(sys.argv[1]) as f:
    for r in csv.reader(f):
        if len(r) >= 15 and r[12] == "gpu__time_duration.sum":
            rows.append(r)
agg = OrderedDict()
for r in rows:
    name = r[4].split("(")[0] if r[4].startswith("mpb::") or " mpb::" in r[4] else r[4][:110]
    name = r[4] if len(r[4]) < 70 else name
    k = (name, r[8], r[7])
    ns = float(r[14].replace(",", ""))
    if r[13] == "us":
        ns *= 1e3
    elif r[13] == "ms":
        ns *= 1e6
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += ns
total = sum(v[1] for v in agg.values())
print("| kernel | launches | total ms | share | grid | block |")
print("|---|---:|---:|---:|---|---|")
for (name, grid, block), (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{name}` | {n} | {ns / 1e6:.3f} | {100 * ns / total:.1f}% | {grid} | {block} |")
print(f"\nTotal {total / 1e6:.1f} ms over {sum(v[0] for v in agg.values())} launches.")
