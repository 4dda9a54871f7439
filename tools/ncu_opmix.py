#!/usr/bin/env python3
"""Executed SASS opcode mix from an `ncu --page source --csv --print-source cuda,sass` export.
Usage: ncu_opmix.py src.csv n_units"""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hdr = None
ops = collections.Counter()
seen = set()
for r in rows:
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr) or r[0] != "":
        continue
    addr = r[2]
    if addr in seen:
        continue
    seen.add(addr)
    sass = r[3].strip()
    m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', sass)
    op = m.group(2) if m else sass
    try:
        ops[op] += int(r[7])
    except ValueError:
        pass
tot = sum(ops.values())
print("total warp-instructions", tot, "per unit", tot / units)
for k, v in ops.most_common(50):
    print(f"{k:34s} {v:10d} {v / units:8.1f}")
