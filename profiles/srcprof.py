#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line.
usage: ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass > both.csv; srcprof.py both.csv [top]"""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
per = defaultdict(lambda: [0, 0, 0, ""])   # (file,line) -> [inst, samples, thread_inst, text]
cur_file = None
hdr = None
line_no = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        hdr = None
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None:
        continue
    if r[0] != "":
        line_no = r[0]
        per[(cur_file, line_no)][3] = r[1]
    try:
        i_inst = hdr.index("Instructions Executed")
        i_smp = hdr.index("# Samples")
        i_ti = hdr.index("Thread Instructions Executed")
    except ValueError:
        continue
    if len(r) > i_ti and r[i_inst] not in ("", None):
        try:
            per[(cur_file, line_no)][0] += int(r[i_inst])
            per[(cur_file, line_no)][1] += int(r[i_smp] or 0)
            per[(cur_file, line_no)][2] += int(r[i_ti] or 0)
        except ValueError:
            pass
tot_i = sum(v[0] for v in per.values())
tot_s = sum(v[1] for v in per.values())
print("total warp instructions %d, samples %d" % (tot_i, tot_s))
print("%-6s %-22s %7s %7s %6s  %s" % ("line", "file", "inst%", "smpl%", "thr/wi", "source"))
for (f, ln), v in sorted(per.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%-6s %-22s %6.2f%% %6.2f%% %6.1f  %s" % (ln, f[:22], 100.0 * v[0] / max(1, tot_i), 100.0 * v[1] / max(1, tot_s),
                                                   v[2] / max(1, v[0]), v[3].strip()[:110]))
