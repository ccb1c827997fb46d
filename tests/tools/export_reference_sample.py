#!/usr/bin/env python
"""GPU box: write a cfg2-shaped sample as the FILES the reference reads (FASTA, BED, filtered sorted BAM) into
gpurun_out/refsample/, together with what computeTconversions of this package makes of them, so that the unmodified
reference can be timed on the same files in the build container (tests/tools/reference_cpu_baseline.py; /root/reference
does not exist on the GPU box).  The sample is cfg2 scaled 1 : 5 -- 20 Mb genome, 4 000 UTRs of the same length
distribution, 200 000 reads: the reads per UTR of the 1 M-read prefix SURVEY.md 8(d) names -- to fit the 64 MiB that
travel back.
   python tests/tools/export_reference_sample.py [reads] [genome_bp] [utrs]"""
import gzip
import io
import json
import os
import shutil
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import synth_files                                        # noqa: E402
from slamdunk_b200.dunks import tcounter                  # noqa: E402
from slamdunk_b200.engine import CountEngine              # noqa: E402
from slamdunk_b200.synth import SynthWorkload             # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
    genome_bp = int(sys.argv[2]) if len(sys.argv) > 2 else 20_000_000
    utrs = int(sys.argv[3]) if len(sys.argv) > 3 else 4000
    out = os.path.join(ROOT, "gpurun_out", "refsample")
    os.makedirs(out, exist_ok=True)
    eng = CountEngine(0)
    wl = SynthWorkload(eng, seed=1, genome_bp=genome_bp, n_utr=utrs)
    wl.install()
    db, triples = wl.generate(wl.params(seed=3, first_read=0, n_reads=n, read_len=100), want_triples=True, coordinate_sorted=True, group=False)
    ref, bed, bam, n_written = synth_files.write_workload_files(out, wl, db, triples, name="cfg2s")
    eng.close()
    log = io.StringIO()
    t0 = time.perf_counter()
    tcounter.computeTconversions(ref, bed, None, bam, 100, 27, os.path.join(out, "gpu_tcount.tsv"), os.path.join(out, "gpu_plus.bedgraph"),
                                 os.path.join(out, "gpu_mins.bedgraph"), 1, log)
    wall = time.perf_counter() - t0
    with open(ref, "rb") as f, gzip.open(ref + ".gz", "wb", compresslevel=6) as g:
        shutil.copyfileobj(f, g)
    os.remove(ref)
    for ext in (".fai",):
        if os.path.exists(ref + ext):
            os.remove(ref + ext)
    json.dump({"reads": n_written, "genome_bp": genome_bp, "utrs": utrs, "gpu_computeTconversions_wall_s": wall,
               "gpu_timings": {k: v for k, v in tcounter.OPTIONS["timings"].items() if isinstance(v, (int, float))}},
              open(os.path.join(out, "meta.json"), "w"), indent=1)
    print(json.dumps({"out": out, "files": {f: os.path.getsize(os.path.join(out, f)) for f in sorted(os.listdir(out))}}))


if __name__ == "__main__":
    main()
