#!/usr/bin/env python
"""Host-side decode / write timings that need no GPU (test infrastructure; the numbers under profiles/r2_host_*.json):
sd_decode_bam with the own block inflate and with zlib alone, sd_load_fasta, sd_rewrite_bam_indexed with the own deflate
and with zlib alone -- on ONE thread (comparable between machines whose cores differ in number) and on all threads.

    python tests/tools/host_bench.py [reads.bam ref.fa]      # without arguments: simulates 300 000 x 100 bp reads first
"""
import ctypes
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np                                     # noqa: E402

from slamdunk_b200 import _lib as L                    # noqa: E402
from slamdunk_b200.engine import HostBatch, HostGenome  # noqa: E402


def best(f, reps=3):
    out = None
    for _ in range(reps):
        r = f()
        if out is None or r[0] < out[0]:
            out = r
    return out


def decode(bam, names, threads):
    hb = HostBatch(bam, names, want_mp=True, threads=threads, group=True, seq2=True, wire=True)
    r = (hb.seconds_inflate + hb.seconds_decode, hb.seconds_inflate, hb.seconds_decode, hb.n_reads, hb.inflated_bytes, hb.compressed_bytes)
    hb.close()
    return r


def rewrite(bam, out, threads, level=6):
    hb = HostBatch(bam, [], want_mp=False, threads=threads)
    n, text = hb.n_reads, hb.header_text
    hb.close()
    keep = np.ones(max(1, n), dtype=np.uint8)
    err = ctypes.create_string_buffer(512)
    nw = ctypes.c_uint64(0)
    t = time.perf_counter()
    rc = L.load().sd_rewrite_bam_indexed(bam.encode(), out.encode(), (out + ".bai").encode(), text.encode(), keep.ctypes.data, None, n, 1, level,
                                         threads, ctypes.byref(nw), err, len(err))
    assert rc == L.SD_OK, err.value
    return (time.perf_counter() - t, os.path.getsize(out))


def main():
    d = tempfile.mkdtemp(prefix="sd_host_bench_")
    if len(sys.argv) > 2:
        bam, fasta = sys.argv[1], sys.argv[2]
    else:
        import simdata
        c = simdata.make_case(d, 5, n_reads=300000, read_len=(100, 100), n_utr=2000,
                              chroms=(("chr1", 3000000), ("chr2", 2000000), ("chrX", 1000000)), complexity=0.3, snp_density=0.0)
        bam, fasta = c.bam, c.ref
    out = {"bam": os.path.basename(bam), "bam_bytes": os.path.getsize(bam), "fasta_bytes": os.path.getsize(fasta), "host_threads": os.cpu_count()}
    for threads, tag in ((1, "1_thread"), (0, "all_threads")):
        g = best(lambda: (HostGenome(fasta, threads).seconds,))
        names = HostGenome(fasta).names
        res = {"fasta_s": g[0]}
        for mode in ("own", "zlib"):
            os.environ["SD_INFLATE"] = mode
            t, ti, tr, n, inflated, comp = best(lambda: decode(bam, names, threads))
            res["decode_inflate_" + mode] = {"total_s": t, "inflate_s": ti, "records_s": tr, "reads_per_s": n / t, "inflate_MBps": inflated / 1e6 / ti}
            out["reads"], out["inflated_bytes"] = n, inflated
        os.environ["SD_INFLATE"] = "own"
        for mode in ("own", "zlib"):
            os.environ["SD_DEFLATE"] = mode
            t, size = best(lambda: rewrite(bam, os.path.join(d, "rw.bam"), threads), reps=2)
            res["rewrite_level6_deflate_" + mode] = {"total_s": t, "out_bytes": size, "reads_per_s": out["reads"] / t}
        os.environ["SD_DEFLATE"] = "own"
        out[tag] = res
    print(json.dumps(out))


if __name__ == "__main__":
    main()
