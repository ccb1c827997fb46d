#!/usr/bin/env python
"""The iterator-based alleyoop statistics end to end on a synthetic coordinate-sorted BAM (needs a GPU): wall time per
tool, the sd_read_stats kernel inside it, and beside them the oracle restatement of the reference (test infrastructure: that is why this script lives under tests/) on a small sample of
the same kind.   python tests/tools/stats_bench.py [n_reads] [genome_bp]"""
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import statsrun                                        # noqa: E402
import synth_files                                     # noqa: E402
from slamdunk_b200.dunks import stats                  # noqa: E402
from slamdunk_b200.engine import CountEngine           # noqa: E402
from slamdunk_b200.synth import SynthWorkload          # noqa: E402

TOOLS = ["rates", "tcperreadpos", "utrrates", "tcperutrpos", "snpeval"]


def make(d, n, genome_bp, n_utr):
    eng = CountEngine(0)
    wl = SynthWorkload(eng, seed=1, genome_bp=genome_bp, n_utr=n_utr)
    db, triples = wl.generate(wl.params(seed=3, first_read=0, n_reads=n), want_triples=True, coordinate_sorted=True, group=False)
    ref, bed, bam, n = synth_files.write_workload_files(d, wl, db, triples)
    eng.close()
    return ref, bed, bam, n


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    genome_bp = int(sys.argv[2]) if len(sys.argv) > 2 else 20_000_000
    d = tempfile.mkdtemp()
    ref, bed, bam, n = make(d, n, genome_bp, 4000)
    out = {"reads": n, "genome_bp": genome_bp, "n_utr": 4000, "host_threads": os.cpu_count(), "gpu": {}}
    for tool in TOOLS:
        for rep in range(2):                              # second run: FASTA cached, CUDA context warm
            csv = os.path.join(d, tool + ".csv")
            if os.path.exists(csv):
                os.remove(csv)
            t0 = time.time()
            statsrun.run_gpu(tool, csv, ref, bam, bed, None, 27, 100, False)
            wall = time.time() - t0
        tm = dict(stats.TIMINGS)
        out["gpu"][tool] = {"wall_s": wall, "kernel_ms": tm["kernel_ms"], "bam_inflate_s": tm["bam_inflate_s"], "bam_decode_s": tm["bam_decode_s"],
                            "reads_per_s_wall": n / wall, "reads_per_s_kernel": n / (tm["kernel_ms"] * 1e-3),
                            "kernel_GBps_of_resident_columns": tm["device_bytes"] / (tm["kernel_ms"] * 1e-3) / 1e9}
    d2 = tempfile.mkdtemp()
    ref2, bed2, bam2, n2 = make(d2, 20_000, 1_000_000, 200)
    out["cpu_oracle_small"] = {"reads": n2, "cores": 1}
    for tool in TOOLS:
        o, g = os.path.join(d2, tool + "_o.csv"), os.path.join(d2, tool + "_g.csv")
        t0 = time.time()
        statsrun.run_oracle(tool, o, ref2, bam2, bed2, None, 27, 100, False)
        cpu = time.time() - t0
        statsrun.run_gpu(tool, g, ref2, bam2, bed2, None, 27, 100, False)
        out["cpu_oracle_small"][tool] = {"seconds": cpu, "reads_per_s": n2 / cpu, "identical_output": open(o).read() == open(g).read()}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
