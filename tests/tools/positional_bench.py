#!/usr/bin/env python
"""`alleyoop positional-tracks` end to end on a synthetic coordinate-sorted BAM (needs a GPU): where the time goes
(host decode, device scan, device run-length compaction + text) and, beside it, the oracle restatement of the reference (test infrastructure: that is why this script lives under tests/)
on a small sample of the same kind.   python tests/tools/positional_bench.py [n_reads] [genome_bp]"""
import io
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import synth_files                                    # noqa: E402
from oracle import oracle                              # noqa: E402
from slamdunk_b200.dunks import tcounter               # noqa: E402
from slamdunk_b200.engine import CountEngine           # noqa: E402
from slamdunk_b200.synth import SynthWorkload          # noqa: E402


def make(d, n, genome_bp, n_utr):
    eng = CountEngine(0)
    wl = SynthWorkload(eng, seed=1, genome_bp=genome_bp, n_utr=n_utr)
    db, triples = wl.generate(wl.params(seed=3, first_read=0, n_reads=n), want_triples=True, coordinate_sorted=True, group=False)
    ref, bed, bam, n = synth_files.write_workload_files(d, wl, db, triples)
    eng.close()
    return ref, bam, n


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    genome_bp = int(sys.argv[2]) if len(sys.argv) > 2 else 20_000_000
    d = tempfile.mkdtemp()
    ref, bam, n = make(d, n, genome_bp, 4000)
    out = {"reads": n, "genome_bp": genome_bp, "host_threads": os.cpu_count()}
    devnull = open(os.devnull, "w")
    old = sys.stdout
    for rep in range(2):                                  # second run: FASTA cached, CUDA context warm
        sys.stdout = devnull
        t0 = time.time()
        tcounter.genomewideConversionRates(ref, None, bam, 27, os.path.join(d, "g_slamdunk_mapped"), 1, 1, io.StringIO())
        wall = time.time() - t0
        sys.stdout = old
    tm = dict(tcounter.OPTIONS["timings"])
    out["gpu"] = {"wall_s": wall, **{k: v for k, v in tm.items() if k != "stats"}}
    out["gpu"]["reads_per_s_wall"] = n / wall
    # CPU: the oracle (C arrays + the reference's per-position Python loop) on a small workload of the same kind
    d2 = tempfile.mkdtemp()
    ref2, bam2, n2 = make(d2, 50_000, 1_000_000, 200)
    t0 = time.time()
    oracle.positional_files(ref2, None, bam2, 27, os.path.join(d2, "o_slamdunk_mapped"), 1, 1, io.StringIO(), out=io.StringIO())
    cpu = time.time() - t0
    sys.stdout = devnull
    t0 = time.time()
    tcounter.genomewideConversionRates(ref2, None, bam2, 27, os.path.join(d2, "g_slamdunk_mapped"), 1, 1, io.StringIO())
    gpu_small = time.time() - t0
    sys.stdout = old
    same = all(open(os.path.join(d2, "o_slamdunk_mapped" + s)).read().split("\n", 1)[1] == open(os.path.join(d2, "g_slamdunk_mapped" + s)).read().split("\n", 1)[1]
               for s, _n, _d in oracle.TRACKS)
    out["cpu_oracle_small"] = {"reads": n2, "genome_bp": 1_000_000, "seconds": cpu, "positions_per_s": 1_000_000 / cpu,
                               "gpu_seconds_same_input": gpu_small, "identical_output": same, "cores": 1}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
