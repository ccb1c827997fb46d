#!/usr/bin/env python
"""ALU-pipe instruction count per CUDA source line (the count kernel is bound by the half-rate ALU pipe:
LOP3/SHF/IADD3/ISETP/SEL/...).  Input: `ncu --page source --csv --print-source cuda,sass` dump.
usage: aluprof.py both.csv [n_tiles] [top]"""
import csv
import re
import sys
from collections import defaultdict

ALU = {"LOP3", "SHF", "IADD3", "ISETP", "SEL", "VIADD", "PLOP3", "VIMNMX", "LEA", "PRMT", "MOV", "VIADDMNMX", "IABS", "ICMP", "FLO", "POPC", "BREV", "FSETP", "FSEL", "VOTE", "P2R", "R2P", "CS2R"}
rows = list(csv.reader(open(sys.argv[1])))
n_tiles = float(sys.argv[2]) if len(sys.argv) > 2 else 1562500.0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
per = defaultdict(lambda: [0, 0, ""])
hdr = None
cur = None
key = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        hdr = None
        continue
    if r[0] == "Line No":
        hdr = r
        ii = hdr.index("Instructions Executed")
        continue
    if hdr is None:
        continue
    if r[0] != "":
        key = (cur, int(r[0]))
        per[key][2] = r[1]
        continue
    if len(r) <= ii or not r[ii].isdigit():
        continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[3])
    if not m:
        continue
    n = int(r[ii])
    per[key][1] += n
    if m.group(2) in ALU:
        per[key][0] += n
ta = sum(v[0] for v in per.values())
tt = sum(v[1] for v in per.values())
print("ALU-pipe warp instructions per tile: %.1f of %.1f total" % (ta / n_tiles, tt / n_tiles))
for k, v in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%-22s %5d  alu/tile %6.1f  all/tile %6.1f | %s" % (k[0][:22], k[1], v[0] / n_tiles, v[1] / n_tiles, v[2].strip()[:100]))
