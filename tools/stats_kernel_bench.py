#!/usr/bin/env python
"""sd_read_stats (the kernel behind alleyoop rates / tccontext / tcperreadpos / utrrates / tcperutrpos / snpeval) on a
device-resident synthetic batch of cfg2's shape: reads/s and GB/s per output set, CUDA events on the engine's stream.
The batch carries the complete MP column (synth.to_mp_all); bytes per read = the columns the kernel streams: record 20,
SEQ planes 16 per 32 bases, CIGAR ops, 8 per MP entry (quality bytes are touched only at the mismatching bases).
   python tools/stats_kernel_bench.py [n_reads] [read_len]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np                                      # noqa: E402

from slamdunk_b200.dunks.stats import read_stats_on_device   # noqa: E402
from slamdunk_b200.engine import CountEngine            # noqa: E402
from slamdunk_b200.synth import SynthWorkload, to_mp_all   # noqa: E402

SETS = [("rates", ("rates",)), ("tccontext", ("context",)), ("tcperreadpos", ("readpos",)), ("utrrates", ("utr_rates",)),
        ("tcperutrpos", ("utrpos",)), ("snpeval", ("utr_snp",)), ("all six", ("rates", "context", "readpos", "utr_rates", "utrpos", "utr_snp"))]


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
    read_len = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    eng = CountEngine(0)
    wl = SynthWorkload(eng, seed=1, genome_bp=100_000_000, n_utr=20000)
    wl.install()
    db, triples = wl.generate(wl.params(seed=3, first_read=0, n_reads=n, read_len=read_len), want_triples=True, coordinate_sorted=True)
    dbm = to_mp_all(db, triples, eng)
    del triples
    tiles = dbm.tensors["tiles"].cpu().numpy().view(np.uint64).reshape(-1, 4)
    bytes_per_read = 20 + 16 * ((read_len + 31) // 32) + (float(tiles[-1, 2]) + float(tiles[-1, 3])) / n
    peak = 6537.6
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    out = {"workload": "cfg2 shape: %d x %d bp reads, 20000 UTRs, coordinate-sorted, complete MP column" % (n, read_len),
           "bytes_per_read": bytes_per_read, "mp_entries_per_read": float(tiles[-1, 3]) / 8 / n, "hbm_peak_gbs": peak, "sets": {}}
    for name, want in SETS:
        _res, ms = read_stats_on_device(eng, dbm, len(wl.ann), 27, want, max_read_length=read_len, repeat=4)
        out["sets"][name] = {"kernel_ms": ms, "reads_per_s": n / (ms * 1e-3), "GBps": n * bytes_per_read / (ms * 1e-3) / 1e9,
                             "frac_of_hbm_peak": n * bytes_per_read / (ms * 1e-3) / 1e9 / peak}
    print(json.dumps(out))
    eng.close()


if __name__ == "__main__":
    main()
