#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (cold-cache, serialised times: read shares only)."""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) >= 15 and r[0].isdigit()]
skip = sys.argv[2:]  # kernel-name substrings to leave out (e.g. the synthetic generator)
agg = collections.OrderedDict()
for r in rows:
    k = r[4].split('(')[0].replace('<unnamed>::', '').replace('dge::', '').replace('void ', '')[:60]
    if any(s in k for s in skip): continue
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += float(r[14].replace(',', ''))
tot = sum(v[1] for v in agg.values())
print(f"{'kernel':60s} {'n':>5s} {'total us':>10s} {'share':>6s}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:60s} {v[0]:5d} {v[1]/1e3:10.1f} {100*v[1]/tot:5.1f}%")
print(f"{'TOTAL':60s} {sum(v[0] for v in agg.values()):5d} {tot/1e3:10.1f}")
