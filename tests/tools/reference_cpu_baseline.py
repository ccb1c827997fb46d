#!/usr/bin/env python
"""Build container (where /root/reference exists): time the UNMODIFIED reference -- slamdunk.dunks.tcounter.
computeTconversions imported verbatim, pysam / pybedtools / intervaltree replaced by the I/O-only shims of tests/shims --
on the cfg2-shaped sample tests/tools/export_reference_sample.py wrote on the GPU box, next to the oracle port on the same
files and cores, and compare all outputs byte for byte (reference == port == GPU).  SURVEY.md 8(d) "CPU reference timed
beside it"; the result is kept in profiles/.
   python tests/tools/reference_cpu_baseline.py [gpurun_out/refsample] [out.json]"""
import gzip
import io
import json
import os
import shutil
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import refrun                     # noqa: E402
from oracle import oracle         # noqa: E402


def main():
    src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "refsample")
    dst = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "profiles", "r2_reference_cpu_baseline.json")
    meta = json.load(open(os.path.join(src, "meta.json")))
    work = tempfile.mkdtemp(prefix="refsample_")
    ref = os.path.join(work, "cfg2s.fa")
    with gzip.open(os.path.join(src, "cfg2s.fa.gz"), "rb") as f, open(ref, "wb") as g:
        shutil.copyfileobj(f, g)
    bed, bam = os.path.join(src, "cfg2s.bed"), os.path.join(src, "cfg2s.bam")
    n = meta["reads"]
    # the reference, verbatim.  The shim's AlignmentFile decodes the BAM when it is opened (inside the call), so the
    # time includes BAM decode like a pysam run does; a second run with the process warm is reported too.
    runs = []
    for k in range(2):
        t0 = time.perf_counter()
        tsv, plus, minus, _ = refrun.run_count(ref, bed, None, bam, os.path.join(work, "ref%d" % k), 100, 27, 1)
        runs.append(time.perf_counter() - t0)
    from slamdunk_b200.io import bam as bamio
    t0 = time.perf_counter()
    bamio.read_bam(bam)
    t_read = time.perf_counter() - t0      # the Python BAM reader both the shim and the port's file-level entry go through
    port = {}
    for threads in (1, oracle.max_threads()):
        o = os.path.join(work, "orc%d" % threads)
        t0 = time.perf_counter()
        oracle.count_files(ref, bed, None, bam, 100, 27, o + ".tsv", o + ".p", o + ".m", 1, io.StringIO(), threads=threads)
        port[threads] = time.perf_counter() - t0
    same = {}
    o = os.path.join(work, "orc1")
    for name, a, b, c in (("tsv", tsv, o + ".tsv", os.path.join(src, "gpu_tcount.tsv")), ("plus", plus, o + ".p", os.path.join(src, "gpu_plus.bedgraph")),
                          ("minus", minus, o + ".m", os.path.join(src, "gpu_mins.bedgraph"))):
        ra, rb, rc = open(a).read(), open(b).read(), open(c).read()
        same[name] = {"reference_eq_port": ra == rb, "reference_eq_gpu": ra == rc, "bytes": len(ra)}
    best = min(runs)
    res = {
        "what": "unmodified reference computeTconversions (tcounter.py:125) through tests/shims on the cfg2-shaped sample written by "
                "tests/tools/export_reference_sample.py on the GPU box; build container, Python %s" % sys.version.split()[0],
        "sample": meta, "host_cpus": oracle.max_threads(),
        "reference_seconds": runs, "reference_reads_per_s_per_core": n / best,
        "reference_parallelism": "one process per BAM (slamdunk.py:516): a single sample runs on 1 core",
        "python_bam_read_seconds_included_in_each": t_read,
        "oracle_port_seconds": {str(k): v for k, v in port.items()},
        "oracle_port_reads_per_s": {str(k): n / v for k, v in port.items()},
        "port_over_reference_1core": best / port[1],
        "outputs_identical": same,
        "extrapolation_to_cfg2": "50 M reads at the measured per-core rate: %.0f s on one core (the reference cannot use more for one BAM)" % (50e6 / (n / best)),
    }
    json.dump(res, open(dst, "w"), indent=1)
    print(json.dumps(res, indent=1))
    shutil.rmtree(work, ignore_errors=True)
    assert all(v["reference_eq_port"] and v["reference_eq_gpu"] for v in same.values()), same


if __name__ == "__main__":
    main()
