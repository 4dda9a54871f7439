#!/usr/bin/env python3
"""Split the stall samples of an `ncu --page source --csv --print-source cuda,sass` export into the barrier-delimited
phases of the kernel (SASS in address order, cut at BAR.SYNC / EXIT).  Usage: ncu_phases.py src.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = None
ins = {}
for r in rows:
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr) or r[0] != "" or not r[2]:
        continue
    ia, isrc, isamp, iinst = 2, 3, hdr.index("# Samples"), hdr.index("Instructions Executed")
    try:
        ins[int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])] = (r[isrc], float(r[isamp] or 0), float(r[iinst] or 0))
    except ValueError:
        pass
seg_s = seg_i = n = 0
total = sum(v[1] for v in ins.values())
print(f"{len(ins)} SASS instructions, {total:.0f} samples")
for a in sorted(ins):
    src, s, i = ins[a]
    seg_s += s
    seg_i += i
    n += 1
    if "BAR.SYNC" in src or "EXIT" in src or "WARPSYNC" in src and False:
        print(f"  up to {a:#x} {src[:34]:34s} sass={n:5d} samples={seg_s:8.0f} ({100*seg_s/total:5.1f}%) warp-inst={seg_i/1e6:8.2f} M")
        seg_s = seg_i = n = 0
