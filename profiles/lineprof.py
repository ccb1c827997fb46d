#!/usr/bin/env python
"""Executed warp instructions per source line of a kernel, per tile.
usage: ncu -i rep --page source --csv --print-source cuda,sass > src.csv; lineprof.py src.csv n_tiles [top]"""
import csv
import os
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
n = float(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "slamdunk_b200", "csrc")
src = {}
inst = defaultdict(int)
cur = hdr = ln = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur, hdr = r[1].split("/")[-1], None
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None:
        continue
    if r[0] != "":
        ln = int(r[0])
    try:
        ii = hdr.index("Instructions Executed")
        if r[ii] != "":
            inst[(cur, ln)] += int(r[ii])
    except (ValueError, IndexError):
        pass
tot = sum(inst.values()) / 2.0          # the cuda + sass view lists every instruction twice
print("warp instructions per tile: %.1f" % (tot / n))
for (f, l), v in sorted(inst.items(), key=lambda kv: -kv[1])[:top]:
    if f not in src and os.path.exists(os.path.join(here, f or "")):
        src[f] = open(os.path.join(here, f)).read().split("\n")
    text = src[f][l - 1].strip()[:110] if f in src and l - 1 < len(src[f]) else ""
    print("%6.1f  %-24s %4d  %s" % (v / 2.0 / n, f, l, text))
