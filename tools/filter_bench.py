#!/usr/bin/env python
"""sd_filter (the decisions of `slamdunk filter`, dunks/filter.py:72-262) on a device-resident synthetic batch in mapper
order: default branch (MQ / identity / NM predicates) and multimapper-retention branch (UTR interval tree + query-name
groups).  Times the whole synchronous call -- kernels, its scratch allocations and the stats read-back -- with the
decisions left on the device.   python tools/filter_bench.py [n_reads]"""
import ctypes
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np                                      # noqa: E402
import torch                                            # noqa: E402

from slamdunk_b200 import _lib as L                     # noqa: E402
from slamdunk_b200.engine import CountEngine, make_params   # noqa: E402
from slamdunk_b200.synth import SynthWorkload           # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
    eng = CountEngine(0)
    wl = SynthWorkload(eng, seed=1, genome_bp=100_000_000, n_utr=20000)
    wl.install()
    db, _ = wl.generate(wl.params(seed=3, first_read=0, n_reads=n, frac_multimap=0.15), coordinate_sorted=False)   # mapper order
    params = make_params(27, 1, 0, filter_on=True, filter_mq=2, filter_min_identity=0.95, filter_max_nm=-1)
    dev = eng.device
    with torch.cuda.stream(eng.stream):
        keep = torch.zeros(n, dtype=torch.uint8, device=dev)
        state = torch.zeros(n, dtype=torch.uint8, device=dev)
        stats = torch.zeros(L.SD_NSTAT, dtype=torch.int64, device=dev)
        name_hash = (torch.arange(n, device=dev, dtype=torch.int64) // 3) * 0x9E3779B97F4A7C15 % (1 << 62)   # three alignments per name
    eng.stream.synchronize()
    a = wl.ann
    utr = [np.ascontiguousarray(x) for x in (a.chrom.astype(np.uint32), a.start.astype(np.int64), a.stop.astype(np.int64),
                                             np.arange(len(a), dtype=np.uint32))]
    ptr = lambda x: ctypes.c_void_p(x.ctypes.data)      # noqa: E731
    tiles = db.tensors["tiles"].cpu().numpy().view(np.uint64).reshape(-1, 4)
    bytes_per_read = 20 + float(tiles[-1, 2]) / n + 1 + 1      # record + CIGAR ops in, keep + state out
    out = {"workload": "%d x 100 bp reads in mapper order, 15 %% mapq 0, 20000 UTRs" % n, "bytes_per_read": bytes_per_read}
    for mode in ("default", "multimapper"):
        args = (None, None, None, None, None, 0) if mode == "default" else (name_hash.data_ptr(), ptr(utr[0]), ptr(utr[1]), ptr(utr[2]), ptr(utr[3]), len(a))
        best = None
        for _ in range(5):
            stats.zero_()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            eng._chk(eng.lib.sd_filter(eng.ctx, ctypes.byref(db.batch), ctypes.byref(params), args[0], args[1], args[2], args[3], args[4],
                                       args[5], keep.data_ptr(), stats.data_ptr(), state.data_ptr()), "sd_filter")
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        st = stats.cpu().numpy()
        out[mode] = {"ms_per_call": best * 1e3, "reads_per_s": n / best, "GBps": n * bytes_per_read / best / 1e9,
                     "kept": int((keep != 0).sum().item()), "filtered_stat": int(st[L.SD_S_FILTERED]), "mqfiltered": int(st[L.SD_S_MQFILTERED])}
    print(json.dumps(out))
    eng.close()


if __name__ == "__main__":
    main()
