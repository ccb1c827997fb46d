#!/usr/bin/env python
"""Instruction / stall-sample share per source region of sd_count_kernel.cuh (the count kernel's device code).  Regions are delimited by
comment markers `// @region <name>` in the source; input is an
`ncu --page source --csv --print-source cuda,sass` dump."""
import csv
import re
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
src = open(sys.argv[2] if len(sys.argv) > 2 else "slamdunk_b200/csrc/sd_count_kernel.cuh").read().split("\n")
region_of = {}
cur = "preamble"
for i, line in enumerate(src, 1):
    m = re.search(r"//\s*@region\s+(\S+)", line)
    if m:
        cur = m.group(1)
    region_of[i] = cur
inst = defaultdict(int)
smp = defaultdict(int)
thr = defaultdict(int)
hdr = None
curfile = None
ln = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        curfile = r[1].split("/")[-1]
        hdr = None
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None:
        continue
    if r[0] != "":
        ln = int(r[0])
    try:
        ii, si, ti = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
        if r[ii] == "":
            continue
        key = region_of.get(ln, "?") if curfile in ("sd_count_kernel.cuh", "sd_count.cu") else "hdr:" + curfile
        inst[key] += int(r[ii])
        smp[key] += int(r[si] or 0)
        thr[key] += int(r[ti] or 0)
    except (ValueError, IndexError):
        pass
ti, ts = sum(inst.values()), sum(smp.values())
print("%-34s %7s %7s %6s" % ("region", "inst%", "smpl%", "thr/wi"))
for k, v in sorted(inst.items(), key=lambda kv: -kv[1]):
    print("%-34s %6.2f%% %6.2f%% %6.1f" % (k, 100.0 * v / ti, 100.0 * smp[k] / max(1, ts), thr[k] / max(1, v)))
