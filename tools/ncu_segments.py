#!/usr/bin/env python
"""Stall samples of a kernel's SASS, summed over program-order segments that end at a barrier / branch / mbarrier wait /
bulk copy: where in the item loop the time goes.  python tools/ncu_segments.py report.ncu-rep [min_pct]"""
import csv
import io
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, data = rows[1], rows[2:]
ci = {h: i for i, h in enumerate(hdr)}
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ci["# Samples"]]) for r in data)
minp = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
cur, cr, n0 = 0, {}, 0
fp = 0
print(f"# {sys.argv[1]}: {tot} samples, {len(data)} instructions")
for k, r in enumerate(data):
    src = r[ci["Source"]].strip()
    cur += int(r[ci["# Samples"]])
    for h in reasons:
        cr[h] = cr.get(h, 0) + int(r[ci[h]] or 0)
    toks = src.split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    if op[:4] in ("DFMA", "DADD", "DMUL", "FFMA", "FADD", "FMUL"):
        fp += 1
    if op.startswith(("BAR", "SYNCS", "BRA", "UBLKCP", "EXIT")) or k == len(data) - 1:
        if cur >= tot * minp / 100:
            top = sorted(cr.items(), key=lambda kv: -kv[1])[:4]
            print(f"{n0:5d}..{k:5d} {100 * cur / tot:6.2f}%  fp {fp:4d}  " +
                  " ".join(f"{h[6:]} {100 * v / max(cur, 1):.0f}%" for h, v in top) + f"   | {src[:50]}")
        cur, cr, n0, fp = 0, {}, k + 1, 0
