"""Per-source-line instruction counts / shared wavefronts / stall samples from a .ncu-rep
(captured with --import-source on, binary built with -lineinfo).  Development aid.
usage: python tools/ncu_lines.py report.ncu-rep [top-n]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
fname = ""
hdr = None
lines = []
total = 0
for r in csv.reader(out.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        ie, ws, wsi, smp = (hdr.index("Instructions Executed"), hdr.index("L1 Wavefronts Shared"),
                            hdr.index("L1 Wavefronts Shared Ideal"), hdr.index("# Samples"))
        continue
    if hdr is None or r[0] in ("", "Function Name"):
        continue
    try:
        n = int(r[ie])
    except ValueError:
        continue
    total += n
    num = lambda x: int(x) if x.lstrip("-").isdigit() else 0
    lines.append((n, num(r[smp]), num(r[ws]), num(r[wsi]), fname, r[0], r[1].strip()[:100]))
print("total warp instructions", total)
print(f"{'instr':>12} {'%':>5} {'samples':>8} {'sh.wavefr':>11} {'ideal':>11}  where")
for n, s, w, wi, f, ln, src in sorted(lines, reverse=True)[:top]:
    print(f"{n:12d} {100.0 * n / total:5.1f} {s:8d} {w:11d} {wi:11d}  {f}:{ln}  {src}")
