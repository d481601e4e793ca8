#!/usr/bin/env python
"""Top stall-sample SASS lines of an `ncu --page source --csv` export, in program order with a running share."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))[2:]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
tot = sum(float(r[2] or 0) for r in rows)
top = set(id(r) for r in sorted(rows, key=lambda r: -float(r[2] or 0))[:n])
for i, r in enumerate(rows):
    if id(r) in top: print(f"{i:4d} {float(r[2])/tot*100:5.1f}%  exec={r[5]:>9s}  {r[1].strip()[:100]}")
