#!/usr/bin/env python3
"""Summarise an `ncu --page source --csv --print-source cuda,sass` export: stall reasons in total and the
top source lines by sampled stalls.  Usage: ncu_stalls.py src.csv [top_n]"""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path)))
# the export holds one section per source file: find header rows
totals = defaultdict(float)
by_line = defaultdict(lambda: defaultdict(float))
text = {}
cur_file = None
hdr = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    # only C source rows carry 'Line No'; sass rows carry Address. Use rows with Line No digits and a Source
    ln = r[0]
    if not ln.isdigit():
        continue
    key = (cur_file, int(ln))
    text[key] = r[1].strip()[:110]
    for k, v in d.items():
        if k.startswith("stall_") and "Not Issued" not in k:
            try:
                f = float(v)
            except ValueError:
                continue
            totals[k] += f
            by_line[key][k] += f
    try:
        by_line[key]["_inst"] += float(d.get("Instructions Executed", 0) or 0)
        by_line[key]["_samples"] += float(d.get("# Samples", 0) or 0)
    except ValueError:
        pass
tot = sum(totals.values())
print("total samples", tot)
for k, v in sorted(totals.items(), key=lambda kv: -kv[1])[:12]:
    print(f"  {k:28s} {v:10.0f} {100 * v / max(tot, 1):5.1f}%")
print("top lines:")
for key, d in sorted(by_line.items(), key=lambda kv: -kv[1]["_samples"])[:top]:
    reasons = sorted(((k, v) for k, v in d.items() if k.startswith("stall_")), key=lambda kv: -kv[1])[:3]
    rs = " ".join(f"{k[6:]}={v:.0f}" for k, v in reasons)
    print(f"  {key[0]}:{key[1]:4d} samples={d['_samples']:7.0f} inst={d['_inst']:9.0f} [{rs}] | {text[key]}")
