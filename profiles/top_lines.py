#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export by CUDA source line.
usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass | python profiles/top_lines.py [N]"""
import csv
import sys

rows = list(csv.reader(sys.stdin))
n_top = int(sys.argv[1]) if len(sys.argv) > 1 else 40
cur_file, cur_line, cur_src = "", "", ""
agg = {}
hdr = None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if len(r) >= 3 and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < 6:
        continue
    if r[0] != "":
        cur_line, cur_src = r[0], r[1].strip()
        continue
    try:
        s = int(r[4])
    except ValueError:
        continue
    inst = 0
    try:
        inst = int(r[hdr.index("Instructions Executed")]) if "Instructions Executed" in hdr else 0
    except ValueError:
        pass
    k = (cur_file, cur_line, cur_src)
    a = agg.setdefault(k, [0, 0])
    a[0] += s
    a[1] += inst
total = sum(v[0] for v in agg.values())
print("total samples", total)
for (f, l, src), (s, inst) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:n_top]:
    print("%6d %5.1f%% inst=%9d %s:%s  %s" % (s, 100.0 * s / max(1, total), inst, f, l, src[:110]))
