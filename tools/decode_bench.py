#!/usr/bin/env python
"""Host decode throughput (reported separately from the counting kernels):
write a synthetic coordinate-sorted BAM of N reads, then time sd_decode_bam / sd_load_fasta and the whole
file-level computeTconversions.  Needs a GPU for the generator.   python tools/decode_bench.py [N]"""
import io
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import synth_files                                    # noqa: E402
from slamdunk_b200.dunks import tcounter               # noqa: E402
from slamdunk_b200.engine import CountEngine, HostBatch, HostGenome   # noqa: E402
from slamdunk_b200.synth import SynthWorkload          # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    d = tempfile.mkdtemp()
    eng = CountEngine(0)
    wl = SynthWorkload(eng, seed=1, genome_bp=20_000_000, n_utr=4000)
    db, triples = wl.generate(wl.params(seed=3, first_read=0, n_reads=n), want_triples=True, coordinate_sorted=True, group=False)
    t0 = time.time()
    ref, bed, bam, n = synth_files.write_workload_files(d, wl, db, triples)
    t_write = time.time() - t0
    eng.close()
    out = {"reads": n, "bam_bytes": os.path.getsize(bam), "fasta_bytes": os.path.getsize(ref), "python_bam_write_s": t_write,
           "host_threads": os.cpu_count()}
    g = HostGenome(ref)
    out["fasta_decode_s"] = g.seconds
    for threads in (1, 4, 0):
        hb = HostBatch(bam, g.names, want_mp=True, threads=threads, group=True)
        out["decode_threads_%s" % (threads or "all")] = {
            "inflate_s": hb.seconds_inflate, "records_s": hb.seconds_decode, "inflated_bytes": hb.inflated_bytes,
            "reads_per_s": n / (hb.seconds_inflate + hb.seconds_decode),
            "compressed_MBps": hb.compressed_bytes / 1e6 / (hb.seconds_inflate + hb.seconds_decode)}
        hb.close()
    g.close()
    o = os.path.join(d, "o")
    t0 = time.time()
    tcounter.computeTconversions(ref, bed, None, bam, 100, 27, o + ".tsv", o + ".p", o + ".m", 1, io.StringIO())
    out["computeTconversions_s"] = time.time() - t0
    out["computeTconversions_timings"] = {k: v for k, v in tcounter.OPTIONS["timings"].items() if k != "stats"}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
