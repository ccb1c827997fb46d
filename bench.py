#!/usr/bin/env python
"""bench.py -- reads/s of the `slamdunk count` hot path (filter + T>C scan + UTR reduce) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--reads R]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Workload: BASELINE.json configs[1] ("cfg2"): synthetic 50 M single-end 100 bp SLAM-seq reads over
20 000 UTRs on a ~100 Mb, 21-contig genome, 2.5 % T>C per reference T,
generated on the device from a counter-based RNG.  A step = one pass of the hot path over the
whole batch: zero the outputs, sd_count (filter predicates + conversion scan + interval join +
per-UTR / per-position reduction), sd_finalize (coverage scan) + sd_tcontent and, for N > 1, one
NCCL all-reduce of the int64 per-UTR counter table.  Multi-GPU runs are weak-scaling: every rank
owns a contiguous genome region (UTR-owner sharding) and counts R reads placed there.

One JSON line on stdout (rank 0).  `value` is device-resident throughput, `e2e` is the same work
through sd_count_host from pinned HOST buffers (H2D inside the timed region, counters read back),
`roofline` relates the count kernel's algorithmic bytes to the measured HBM copy peak, and
`cpu_baseline` is the oracle port of the reference's algorithm timed on this box's host cores
on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

FILTER = dict(filter_on=True, filter_mq=2, filter_min_identity=0.95, filter_max_nm=-1)   # slamdunk.py:371-373 defaults
MIN_QUAL, CONV_THRESHOLD = 27, 1                                                         # slamdunk.py:395,397 defaults


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.remove(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), samples=len(sm))
        out["reasons"] = sorted(reasons)
        return out


def build_workload(local_rank, rank, world, n_reads, read_len, genome_bp, n_utr, want_triples=False, first_read=None,
                   want_mp=False, coordinate_sorted=True):
    from slamdunk_b200.engine import CountEngine
    from slamdunk_b200.synth import SynthWorkload
    eng = CountEngine(local_rank)
    wl = SynthWorkload(eng, seed=1, genome_bp=genome_bp, n_utr=n_utr)
    wl.install()
    first, count = (0, 0) if world == 1 else wl.region_split(world)[rank]
    fr = rank * n_reads if first_read is None else first_read
    p = wl.params(seed=3, first_read=fr, n_reads=n_reads, read_len=read_len, utr_first=first, utr_count=count)
    db, triples = wl.generate(p, want_mp=want_mp, want_triples=want_triples, coordinate_sorted=coordinate_sorted)
    return eng, wl, db, triples


def oracle_inputs(wl, db, triples):
    import oracle_bridge as ob
    g = ob.genome_from_planes(wl.names, wl.chrom_off, wl.chrom_len, wl.lo.cpu().numpy().view(np.uint32),
                              wl.hi.cpu().numpy().view(np.uint32), wl.nmask.cpu().numpy().view(np.uint32))
    bn = ob.device_batch_to_numpy(db)
    reads, qual, mp = ob.batch_to_oracle(bn, triples.cpu().numpy())
    return g, bn, reads, qual, mp, ob.annotation_to_oracle(wl.ann, len(wl.names))


def oracle_pass(g, bn, reads, qual, mp, utrs, threads):
    """The reference's algorithm (oracle port): filter.Filter default branch, then computeTconversions."""
    from oracle import oracle
    from slamdunk_b200 import _lib as L
    n = bn["n_reads"]
    rec = bn["rec"][:n]
    fr = np.zeros(n, dtype=oracle.FREAD_DT)
    flags = rec[:, 1] >> 24
    fr["flag"] = np.where(flags & L.SD_F_REVERSE, 16, 0)
    fr["mapq"] = (rec[:, 1] >> 16) & 0xFF
    fr["xi"] = rec[:, 2].view(np.float32)
    fr["nm"] = rec[:, 4] & 0xFFFF
    keep = np.zeros(n, dtype=np.uint8)
    stats = np.zeros(7, dtype=np.int64)
    t0 = time.perf_counter()
    oracle.lib().orc_filter_default(fr.ctypes.data, n, FILTER["filter_mq"], FILTER["filter_min_identity"],
                                    FILTER["filter_max_nm"], keep.ctypes.data, stats.ctypes.data)
    kept = reads[keep.astype(bool)]
    rows, bg = oracle.count_arrays(g, kept, qual, utrs, MIN_QUAL, CONV_THRESHOLD, None, mp_triples=mp, threads=threads)
    return time.perf_counter() - t0, rows, bg, stats


def run_reference(args, rank, world, local_rank):
    """--impl reference: the reference's CPU algorithm (oracle port; the Python reference cannot travel
    to the GPU box) on the host cores, each step a bounded sample of the same workload."""
    if rank != 0:
        return
    from oracle import oracle
    threads = oracle.max_threads()
    sample = min(args.reads, args.cpu_sample)
    eng, wl, db, triples = build_workload(local_rank, 0, 1, sample, args.read_len, args.genome_bp, args.utrs, want_triples=True)
    inp = oracle_inputs(wl, db, triples)
    eng.close()
    times = []
    for i in range(args.warmup + args.steps):
        dt, _rows, _bg, _st = oracle_pass(*inp, threads=threads)
        if i >= args.warmup:
            times.append(dt)
    t = float(np.mean(times))
    v = sample / t
    line = {
        "impl": "reference", "metric": "reads/sec counted (filter+T>C scan+UTR reduce)", "value": v, "unit": "reads/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8/int64", "data": "synthetic",
        "config": workload_config(args, sample_reads=sample),
        "cpu_baseline": {"value": v, "unit": "reads/s", "cores": threads, "kind": "port",
                         "sample": "first %d reads of the cfg2 read stream per step, inputs pre-decoded in RAM" % sample},
        "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def workload_config(args, sample_reads=None):
    c = {"workload": "cfg2: synthetic %d x %d bp single-end SLAM-seq reads per GPU, %d UTRs, %.0f Mb genome on 21 contigs, "
                     "2.5%% T>C per reference T" % (args.reads, args.read_len, args.utrs, args.genome_bp / 1e6),
         "reads_per_gpu": args.reads, "read_len": args.read_len, "n_utr": args.utrs,
         "filter": "MQ>=2, identity>=0.95 (slamdunk defaults)", "min_base_qual": MIN_QUAL, "conversion_threshold": CONV_THRESHOLD,
         "mismatch_source": "cigar_walk", "read_order": args.order, "seq_bits": getattr(args, "seq_bits", 2),
         "qual_form": getattr(args, "qual_form", "planes") if getattr(args, "seq_bits", 2) == 2 else "bytes", "sharding": "utr_owner_regions" if args.gpus > 1 else "none",
         "l2_policy": "inputs (~%.1f GB per GPU) far exceed the 126 MB L2; no explicit flush" % (args.reads * (getattr(args, "seq_bits", 2) * 4 * ((args.read_len + 31) // 32) +
                                                                                                         (8 * ((args.read_len + 31) // 32) if getattr(args, "seq_bits", 2) == 2 and getattr(args, "qual_form", "planes") == "planes" else args.read_len) + 24.4) / 1e9)}
    if sample_reads is not None:
        c["sample_reads"] = sample_reads
    return c


_RESULT_FD = None


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries print there too (NCCL writes its version banner to stdout at
    communicator creation), so stdout is pointed at stderr for the whole run and the result line goes to the saved
    descriptor."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    os.write(_RESULT_FD if _RESULT_FD is not None else 1, data)


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--reads", type=int, default=50_000_000, help="reads per GPU")
    ap.add_argument("--read-len", type=int, default=100, dest="read_len")
    ap.add_argument("--utrs", type=int, default=20000)
    ap.add_argument("--genome-bp", type=int, default=100_000_000, dest="genome_bp")
    ap.add_argument("--cpu-sample", type=int, default=4_000_000, dest="cpu_sample")
    ap.add_argument("--order", default="sorted", choices=["sorted", "mapper"],
                    help="read order inside the batch: coordinate-sorted (what `slamdunk count` reads: the sorted BAM "
                         "written by `slamdunk filter`) or mapper order (random placement, the filter's input)")
    ap.add_argument("--seq-bits", type=int, default=2, choices=[2, 4], dest="seq_bits",
                    help="form of the SEQ column: 2 = 2-bit base codes, what sd_decode_bam hands the count path (default); "
                         "4 = the four BAM-nibble bit-planes")
    ap.add_argument("--qual-form", default="planes", choices=["planes", "bytes"], dest="qual_form",
                    help="device form of the base qualities: planes = bit-planes of the dictionary code staged with the tile "
                         "(what sd_count_host gives the kernel when the file has <= 16 distinct qualities and the SEQ column is "
                         "2-bit; default), bytes = one byte per base fetched at candidates")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        args.gpus = world

    if args.impl == "reference":
        run_reference(args, rank, world, local_rank)
        return

    import torch
    import torch.distributed as dist
    from slamdunk_b200 import _lib as L
    from slamdunk_b200.engine import make_params
    from slamdunk_b200.synth import (algorithmic_bytes_per_read, batch_to_wire, layout_bytes_per_read, to_qual_planes, to_seq2)

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    eng, wl, db, _ = build_workload(local_rank, rank, world, args.reads, args.read_len, args.genome_bp, args.utrs,
                                    coordinate_sorted=(args.order == "sorted"))
    params = make_params(MIN_QUAL, CONV_THRESHOLD, L.SD_MISMATCH_CIGAR, **FILTER)
    n_reads = args.reads
    if args.seq_bits == 2:
        db4 = db
        db = to_seq2(db4, eng)
        db4.tensors["seq"] = None      # free the 4-plane column
        del db4
        torch.cuda.empty_cache()
    db_rows = db                       # qualities as one byte per base: what the host batch of the e2e leg is made from
    qual_planes = 0
    if args.seq_bits == 2 and args.qual_form == "planes":
        dbq = to_qual_planes(db_rows, eng)
        if dbq is not None:
            db, qual_planes = dbq, int(dbq.batch.qual_bits)

    def step():
        eng.reset_outputs()
        eng.count(db, params)
        eng.finalize()
        if world > 1:
            with torch.cuda.stream(eng.stream):
                dist.all_reduce(eng.counters)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = eng.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms = []
    barrier()
    ev0.record(eng.stream)
    for _ in range(args.steps):
        step()
        kernel_ms.append(eng.last_kernel_ms())      # CUDA events around the count kernel, on its stream
    ev1.record(eng.stream)
    barrier()
    total_ms = ev0.elapsed_time(ev1)
    launches = eng.launch_count() - launches0
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=eng.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = total_ms / args.steps
    value = world * n_reads / (ms_per_step * 1e-3)
    stats = eng.stats.cpu().numpy()

    # roofline of the dominant kernel (sd_count_kernel)
    tiles = db.tensors["tiles"].cpu().numpy().view(np.uint64).reshape(-1, 4)
    mean_cig = float(tiles[-1, 2]) / 4.0 / n_reads
    bytes_per_read = algorithmic_bytes_per_read(args.read_len, mean_cig)
    layout_bytes = layout_bytes_per_read(args.read_len, mean_cig, args.seq_bits, qual_planes)
    k_ms = float(np.mean(kernel_ms))
    achieved = n_reads * bytes_per_read / (k_ms * 1e-3) / 1e9
    peak, peak_src = measured_peaks()
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": load_ncu_traffic(args), "kernel": "sd_count_kernel", "kernel_ms": k_ms,
                "algorithmic_bytes_per_read": bytes_per_read, "algorithmic_bytes_source": "SURVEY.md 8(d): ceil(L/2) + L + 4c + 17",
                "layout_bytes_per_read": layout_bytes, "achieved_layout_bytes": n_reads * layout_bytes / (k_ms * 1e-3) / 1e9,
                "frac_layout_bytes": n_reads * layout_bytes / (k_ms * 1e-3) / 1e9 / peak,
                "note": "achieved / frac follow the contract: SURVEY 8(d)'s bytes of the path's INPUT per read over the kernel time. "
                        "The columns are held coded (2-bit SEQ, quality bit-planes), so the kernel moves fewer bytes "
                        "(layout_bytes_per_read; traffic = ncu DRAM bytes per launch) and at this point is bound by instruction "
                        "issue (about 70 % of the issue slots, profiles/), not by HBM: frac can exceed 1 at longer reads.",
                "peak_source": peak_src,
                "frac_of_nominal_8TBs": achieved / 8000.0, "kernel_share_of_step": k_ms / ms_per_step}

    # end to end: pinned host buffers -> sd_count_host -> counters back on the host
    e2e = None
    if not args.no_e2e:
        hw = batch_to_wire(db_rows, eng)           # the packed host form sd_decode_bam produces for `count` (sd_wire)
        h2d = hw.h2d_bytes(params)
        d2h = eng.counters.numel() * 8 + eng.stats.numel() * 8

        def e2e_step():
            eng.reset_outputs()
            eng.count_wire(hw.wire, params)
            eng.finalize()
            if world > 1:
                with torch.cuda.stream(eng.stream):
                    dist.all_reduce(eng.counters)
            eng.synchronize()
            return eng.counters.cpu(), eng.stats.cpu()

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            out = e2e_step()
        barrier()
        dt = (time.perf_counter() - t0) / args.steps
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=eng.device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": world * n_reads / dt, "unit": "reads/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": dt * 1e3, "h2d_gbs": h2d / dt / 1e9, "h2d_bytes_per_read": h2d / n_reads,
               "host_form": "sd_wire: per-field columns, 2-bit SEQ codes, %d-bit quality codes, plain-match tiles without l_seq / CIGAR" % int(hw.wire.qual_bits),
               "counted_reads": int(out[1][L.SD_S_COUNTED])}
        del hw

    # CPU baseline + parity of the GPU counters with the oracle on a bounded sample (rank 0, N = 1)
    cpu = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle
        import oracle_bridge as ob
        sample = min(n_reads, args.cpu_sample)
        p = wl.params(seed=3, first_read=0, n_reads=sample, read_len=args.read_len)
        sdb, triples = wl.generate(p, want_triples=True, coordinate_sorted=(args.order == "sorted"))
        inp = oracle_inputs(wl, sdb, triples)
        threads = oracle.max_threads()
        dt, rows, bg, ostats = oracle_pass(*inp, threads=threads)
        eng.reset_outputs()
        eng.count(sdb, params)
        eng.finalize()
        res = eng.results()
        problems = ob.compare(res, rows, bg, wl.ann, eng.pos_off, wl.chrom_off)
        if res["stats"][:6].tolist() != ostats[:6].tolist():
            problems.append("filter stats differ")
        parity = "bit-exact vs oracle on %d reads" % sample if not problems else "MISMATCH: " + "; ".join(problems)
        cpu = {"value": sample / dt, "unit": "reads/s", "cores": threads, "kind": "port",
               "sample": "first %d reads of the same read stream, inputs pre-decoded in RAM, filter + count" % sample,
               "seconds": dt}

    if rank == 0:
        line = {
            "metric": "reads/sec counted (filter+T>C scan+UTR reduce)", "value": value, "unit": "reads/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/int64 (integer bit ops)",
            "data": "synthetic", "config": workload_config(args),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "parity": parity,
            "counted_reads_last_step": int(stats[L.SD_S_COUNTED]), "filtered_reads_last_step": int(stats[L.SD_S_FILTERED]),
            "slow_path_reads": int(stats[L.SD_S_SLOWPATH]),
        }
        emit(line)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def load_ncu_traffic(args):
    """dram bytes per launch of sd_count_kernel from the committed ncu summary (profiles/); None when this run is not
    the workload that capture was taken on (cfg2: 50 M x 100 bp per GPU, coordinate order)."""
    path = os.path.join(ROOT, "profiles", "count_kernel_dram.json")
    try:
        d = json.load(open(path))
        if args.reads != d.get("reads_per_launch") or args.read_len != 100 or args.order != "sorted" or args.seq_bits != d.get("seq_bits", 4) or args.qual_form != d.get("qual_form", "bytes"):
            return None
        return d.get("dram_bytes_per_launch")
    except Exception:
        return None


if __name__ == "__main__":
    main()
