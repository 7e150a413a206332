#!/usr/bin/env python
"""Warp-state samples of an `ncu --set full --import-source on` capture, aggregated by SOURCE FUNCTION.
    ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv
    python profiles/scripts/samples_by_function.py src.csv [csrc dir] > profiles/<run>_samples_by_function.txt
Every CUDA source line of the export is attributed to the function whose definition encloses it (parsed from the csrc
files: `HIVED_DEV ... name(...) {` / `__device__` / `__global__`), header lines of other files to the file name."""
import csv
import os
import re
import sys

src_csv = sys.argv[1]
csrc = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))),
                                                           "hivedscheduler_b200", "csrc")
DEF = re.compile(r"^\s*(?:template\s*<[^>]*>\s*)?(?:HIVED_DEV(?:_NOINLINE)?|__device__|__global__|static|inline|HIVED_HD)\b[^;=]*?\b(\w+)\s*\([^;]*\)\s*(?:const\s*)?\{")
func_of = {}
for fn in os.listdir(csrc):
    if not fn.endswith((".h", ".inc", ".cu", ".hpp")):
        continue
    cur, table = None, {}
    lines = open(os.path.join(csrc, fn), errors="replace").read().splitlines()
    i = 0
    while i < len(lines):
        joined = lines[i]
        # a definition's signature may span two or three lines
        for extra in (1, 2):
            if "{" not in joined and i + extra < len(lines):
                joined += " " + lines[i + extra].strip()
        m = DEF.match(joined)
        if m and lines[i].startswith("  ") and not lines[i].startswith("      "):
            cur = m.group(1)
        elif m and not lines[i].startswith(" "):
            cur = m.group(1)
        table[i + 1] = cur
        i += 1
    func_of[fn] = table

csv.field_size_limit(1 << 30)
agg = {}
cur_file, hdr, cur_func = "", None, None
with open(src_csv, newline="") as f:
    for r in csv.reader(f):
        if len(r) >= 2 and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if len(r) >= 3 and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr) or r[0] in ("", "Function Name"):
            continue
        try:
            line = int(r[0])
            samples = int(r[hdr.index("# Samples")])
            inst = int(r[hdr.index("Instructions Executed")])
        except ValueError:
            continue
        name = func_of.get(cur_file, {}).get(line) or cur_file
        a = agg.setdefault(name, {"s": 0, "i": 0, "st": {}})
        a["s"] += samples
        a["i"] += inst
        for k, col in enumerate(hdr):
            if col.startswith("stall_") and "(" not in col:
                try:
                    a["st"][col[6:]] = a["st"].get(col[6:], 0) + int(r[k])
                except ValueError:
                    pass
total = sum(a["s"] for a in agg.values())
print("total samples %d instructions %d" % (total, sum(a["i"] for a in agg.values())))
for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["s"])[:45]:
    top = sorted(a["st"].items(), key=lambda kv: -kv[1])[:4]
    print("%8d %5.1f%% inst %10d  %-28s %s" % (a["s"], 100.0 * a["s"] / max(1, total), a["i"], name,
                                                " ".join("%s=%d" % kv for kv in top)))
