#!/usr/bin/env python
"""profiles/traffic.json from an `ncu --set full --page raw --csv` export of one warm-up + one timed step:
dram__bytes_read.sum + dram__bytes_write.sum per launch for k_fill_compact, and per launch GROUP (the size classes + hash tail that
bench.py times as one unit) for k_sort_dedup.  Uses the LAST step in the capture."""
import csv, json, sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
ki = hdr.index("Kernel Name")
ri, wi, ti = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")


def to_bytes(v, u):
    return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


def to_ms(v, u):
    return float(v.replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}[u]


launches = [(r[ki], to_bytes(r[ri], units[ri]) + to_bytes(r[wi], units[wi]), to_ms(r[ti], units[ti])) for r in rows[2:]]
fills = [l for l in launches if "k_fill_compact" in l[0]]
dedups = [l for l in launches if "k_sort_dedup" in l[0] or "k_dedup_sort" in l[0]]
per_step = len(dedups) // max(1, len(fills))
last = dedups[-per_step:]
out = {
    "k_fill_compact": {"dram_bytes_per_launch": fills[-1][1], "ncu_ms": fills[-1][2], "launch": fills[-1][0][:80]},
    "k_sort_dedup": {"dram_bytes_per_launch": sum(l[1] for l in last), "ncu_ms": sum(l[2] for l in last),
                     "launches_in_group": [(l[0][:60], round(l[1] / 1e6, 1), round(l[2], 3)) for l in last]},
    "source": sys.argv[1], "note": "ncu --set full --clock-control none, BASELINE configs[1] (400M reads), last of two steps; cold-cache serialised times",
}
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out, indent=1))
