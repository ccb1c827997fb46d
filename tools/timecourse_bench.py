#!/usr/bin/env python
"""BASELINE.json config 3 on one GPU: a 4sU time course of 12 samples x 30 M reads x 150 bp, generated on the device
(per-UTR half-lives, labelled fraction 1 - exp(-ln2 t / halflife)) and counted in one batched run by one engine --
genome and annotation installed once, the samples resident in HBM side by side, 12 count + finalize passes back to back.
   python tools/timecourse_bench.py [reads_per_sample] [read_len]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np                                      # noqa: E402
import torch                                            # noqa: E402

from slamdunk_b200 import _lib as L                     # noqa: E402
from slamdunk_b200.engine import CountEngine, make_params   # noqa: E402
from slamdunk_b200.synth import SynthWorkload, algorithmic_bytes_per_read, to_seq2   # noqa: E402

TIMES = [0, 15, 30, 45, 60, 90, 120, 180, 240, 360, 480, 720]


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 30_000_000
    read_len = int(sys.argv[2]) if len(sys.argv) > 2 else 150
    eng = CountEngine(0)
    wl = SynthWorkload(eng, seed=1, genome_bp=100_000_000, n_utr=20000)
    wl.install()
    params = make_params(27, 1, L.SD_MISMATCH_CIGAR, filter_on=True, filter_mq=2, filter_min_identity=0.95, filter_max_nm=-1)
    samples = []
    for k, t in enumerate(TIMES):
        wl.set_time_point(t)
        db, _ = wl.generate(wl.params(seed=10 + k, first_read=0, n_reads=n, read_len=read_len), coordinate_sorted=True)
        db2 = to_seq2(db, eng)
        db.tensors["seq"] = None
        del db
        torch.cuda.empty_cache()
        samples.append(db2)
    resident = sum(s.nbytes() for s in samples)
    counters = []

    def run():
        del counters[:]
        kms = []
        for s in samples:
            eng.reset_outputs()
            eng.count(s, params)
            eng.finalize()
            kms.append(None)
            with torch.cuda.stream(eng.stream):          # the outputs live on the engine's stream
                counters.append(eng.counters.clone())
        return kms

    for _ in range(2):
        run()
    eng.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    ev0.record(eng.stream)
    for _ in range(reps):
        run()
    ev1.record(eng.stream)
    eng.synchronize()
    ms = ev0.elapsed_time(ev1) / reps
    tc = [int(c.view(-1, L.SD_NCOUNTER)[:, L.SD_C_TCREADS].sum().item()) for c in counters]
    rc = [int(c.view(-1, L.SD_NCOUNTER)[:, L.SD_C_READS].sum().item()) for c in counters]
    tiles = samples[0].tensors["tiles"].cpu().numpy().view(np.uint64).reshape(-1, 4)
    bpr = algorithmic_bytes_per_read(read_len, float(tiles[-1, 2]) / 4.0 / n)
    print(json.dumps({"workload": "cfg3: %d samples x %d reads x %d bp, one batched run on one GPU" % (len(TIMES), n, read_len),
                      "time_points_min": TIMES, "ms_per_run": ms, "reads_per_s": len(TIMES) * n / (ms * 1e-3),
                      "algorithmic_GBps": len(TIMES) * n * bpr / (ms * 1e-3) / 1e9, "resident_GB": resident / 1e9,
                      "tc_read_fraction_per_sample": [round(a / max(1, b), 4) for a, b in zip(tc, rc)]}))
    eng.close()


if __name__ == "__main__":
    main()
