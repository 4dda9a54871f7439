#!/usr/bin/env python3
"""Average duration per kernel name from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import csv
import sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
seen = {}
for r in rows[1:]:
    seen.setdefault(r[ki][:90], []).append(float(r[vi].replace(",", "")))
for k, v in seen.items():
    print(f"{sum(v) / len(v) / 1000:10.2f} us  x{len(v):4d}  {k}")
