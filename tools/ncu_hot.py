#!/usr/bin/env python3
"""Top stalled instructions of one kernel from an ncu report (source page).
    python tools/ncu_hot.py gpurun_out/prof.ncu-rep sign_w [N]"""
import csv
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}", "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
isrc, ins = hdr.index("Source"), hdr.index("# Samples")
stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for k, r in enumerate(rows[hi + 1:]):
    if len(r) < len(hdr):
        continue
    try:
        n = int(r[ins])
    except ValueError:
        continue
    best = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stalls), reverse=True)[:2]
    data.append((n, k, r[isrc].strip()[:90], best))
tot = sum(d[0] for d in data)
print(f"{pat}: {len(data)} instructions, {tot} samples")
for n, k, src, best in sorted(data, reverse=True)[:top]:
    print(f"{100.0 * n / tot:5.2f}%  #{k:5d}  {src:90s} {best}")
