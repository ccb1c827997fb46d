#!/usr/bin/env python
"""Executed SASS of the count kernel in address order with per-tile execution counts.
usage: ncu -i rep --page source --csv --print-source sass > sass.csv; hotpath.py sass.csv [n_tiles] [min_per_tile]
columns: executions per tile, average active threads, stall samples, instruction"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n_tiles = float(sys.argv[2]) if len(sys.argv) > 2 else 1562500.0
floor = float(sys.argv[3]) if len(sys.argv) > 3 else 0.25
hdr = rows[1]
ii, si, ti = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
tot = 0.0
for r in rows[2:]:
    if len(r) <= ii or not r[ii].isdigit():
        continue
    e = int(r[ii]) / n_tiles
    tot += e
    if e < floor:
        continue
    print("%5.2f %4.0f %5s  %s" % (e, int(r[ti]) / max(1, int(r[ii])), r[si], r[1].strip()[:100]))
print("total warp instructions per tile: %.1f" % tot)
