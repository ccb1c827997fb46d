#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list, and the share of one bench step
(sd_lean_kernel + sd_count_kernel + sd_finalize_kernel + sd_tcontent_kernel) taken by each of its kernels.
usage: launchsum.py launches.csv"""
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if r and not r[0].startswith("==")]
hdr = rows[0]
kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    if len(r) <= mv:
        continue
    v = float(r[mv].replace(",", ""))
    v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(r[mu], v)
    name = r[kn].split("(")[0][:60]
    tot[name][0] += 1
    tot[name][1] += v
total = sum(v[1] for v in tot.values())
print("ncu --metrics gpu__time_duration.sum --clock-control none; per-launch times are cold-cache and serialised: compare shares\n")
for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%-62s n=%3d %11.1f us  %5.1f%%" % (k, n, us, 100 * us / total))
step = {k: v for k, v in tot.items() if any(s in k for s in ("sd_lean_kernel", "sd_count_kernel", "sd_finalize_kernel", "sd_tcontent_kernel"))}
st = sum(v[1] / v[0] for v in step.values())
print("\none step = sd_lean_kernel + sd_count_kernel (staged, plain) + sd_finalize_kernel + sd_tcontent_kernel (+ memset fills of the outputs):")
for k, (n, us) in sorted(step.items(), key=lambda kv: -kv[1][1]):
    print("  %-42s %8.1f us per launch  share of step %5.1f%%" % (k, us / n, 100 * (us / n) / st))
