"""World-size-2 gloo run on CPU of the multi-rank host logic: tile sharding + integer all-reduce.
Each rank counts its slice of the reads with the CPU oracle; the reduced table must equal the
unsharded count bit for bit (the property the GPU path relies on)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import simdata

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _count_slice(case_dir, lo, hi):
    """oracle counters for records [lo, hi) of the case's BAM -> int64 [n_utr, 5] + per-position dict arrays"""
    from oracle import oracle
    from slamdunk_b200.io import bam as bamio
    ref, bed, bam = (os.path.join(case_dir, "c" + e) for e in (".fa", ".bed", ".bam"))
    text, refs, records = bamio.read_bam(bam)
    fasta = bamio.read_fasta(ref)
    genome = oracle.Genome(list(fasta.keys()), [s.encode() for s in fasta.values()])
    names = [n for n, _ in refs]
    reads = np.zeros(hi - lo, dtype=oracle.READ_DT)
    quals, mps = [], []
    qo = mo = 0
    for k, r in enumerate(records[lo:hi]):
        reads[k]["chrom"] = genome.index.get(names[r.ref_id], -1)
        reads[k]["pos"], reads[k]["end"] = r.pos, r.reference_end
        reads[k]["l_qual"], reads[k]["flag"], reads[k]["mapq"] = len(r.qual), r.flag, r.mapq
        reads[k]["qual_off"] = qo
        quals.append(bytes(r.qual)); qo += len(r.qual)
        if r.has_tag("MP"):
            mp_ = r.get_tag("MP").encode()
            reads[k]["has_mp"], reads[k]["mp_off"], reads[k]["mp_len"] = 1, mo, len(mp_)
            mps.append(mp_); mo += len(mp_)
    reads = reads[reads["chrom"] >= 0]
    rows = oracle.read_bed(bed)
    utrs = np.zeros(len(rows), dtype=oracle.UTR_DT)
    ids = {}
    for i, (ch, s, e, _n, st) in enumerate(rows):
        utrs[i] = (genome.index.get(ch, -1), 1 if ch in names else 0, s, e, 1 if st == "-" else 0, ids.setdefault(ch, len(ids)))
    out, bg = oracle.count_arrays(genome, reads, np.frombuffer(b"".join(quals), dtype=np.uint8), utrs, 27, 1, None,
                                  mp_text=b"".join(mps))
    table = np.stack([out[k] for k in ("read_count", "tc_read_count", "multimap_count", "coverage_on_ts", "conversions_on_ts")], axis=1)
    return table.astype(np.int64), len(records)


def _worker(rank, world, port, case_dir, n_records, result_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from slamdunk_b200 import parallel
    r, w, _lr = parallel.init("gloo")
    assert (r, w) == (rank, world)
    # shard by 32-read tiles like the GPU path
    n_tiles = (n_records + 31) // 32
    t0, t1 = parallel.shard_range(n_tiles, rank, world)
    lo, hi = t0 * 32, min(n_records, t1 * 32)
    table, _ = _count_slice(case_dir, lo, hi)
    t = torch.from_numpy(table.copy())
    parallel.allreduce_sum_([t])
    if rank == 0:
        np.save(result_path, t.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partitions_exactly():
    from slamdunk_b200 import parallel
    for n in (0, 1, 7, 32, 1000, 1001):
        for w in (1, 2, 3, 8):
            parts = [parallel.shard_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo_reduce_equals_unsharded(tmp_path):
    c = simdata.make_case(str(tmp_path), seed=909, n_reads=1500, name="c")
    whole, n_records = _count_slice(str(tmp_path), 0, 1500)
    port = _free_port()
    result = str(tmp_path / "reduced.npy")
    mp.spawn(_worker, args=(2, port, str(tmp_path), n_records, result), nprocs=2, join=True)
    reduced = np.load(result)
    assert whole[:, 0].sum() > 0
    assert np.array_equal(reduced, whole)


def test_sub_batch_views(tmp_path):
    """sub_batch() slices a decoded host batch by tiles without copying; slices tile the whole."""
    from slamdunk_b200 import parallel
    from slamdunk_b200.engine import HostBatch
    c = simdata.make_case(str(tmp_path), seed=910, n_reads=300, name="c")
    hb = HostBatch(c.bam, list(c.genome.keys()))
    nt = hb.batch.n_tiles
    total = 0
    for r in range(3):
        t0, t1 = parallel.shard_range(nt, r, 3)
        sb = parallel.sub_batch(hb.batch, t0, t1)
        assert sb.n_tiles == t1 - t0 and sb.tile_reads == 32
        assert sb.rec == hb.batch.rec + t0 * 32 * 20 and sb.tiles == hb.batch.tiles + t0 * 32
        assert sb.seq == hb.batch.seq and sb.qual == hb.batch.qual
        total += sb.n_reads
    assert total == hb.n_reads
    hb.close()


def test_region_bounds_cut_between_utrs_no_read_can_cross():
    """bench.Bench.region_bounds (BASELINE.json config 5, chromosome-region sharding): the cuts tile the linear genome, each cut
    lies a read length before a UTR that no earlier UTR of its chromosome reaches, so a read -- placed within a read length of
    its UTR -- starts in the stretch that owns the UTR and never touches a UTR of another stretch."""
    import argparse
    import bench
    from slamdunk_b200.synth import genome_layout, make_annotation

    class FakeWorkload(object):
        def __init__(self, seed):
            self.chrom_len = np.array([300_000, 250_000, 180_000, 90_000], dtype=np.uint32)
            self.chrom_off, self.n_words = genome_layout(self.chrom_len)
            self.ann = make_annotation(seed, self.chrom_len, 400, overlap_frac=0.2)
            self.cum_weight = np.cumsum(self.ann.weight)
        region_split = __import__("slamdunk_b200.synth", fromlist=["SynthWorkload"]).SynthWorkload.region_split

    for seed in (1, 2, 3):
        wl = FakeWorkload(seed)
        b = bench.Bench.__new__(bench.Bench)
        b.wl, b.args = wl, argparse.Namespace(read_len=100)
        a = wl.ann
        gs = wl.chrom_off[a.chrom].astype(np.int64) + a.start
        ge = wl.chrom_off[a.chrom].astype(np.int64) + a.stop
        for n_parts in (2, 3, 8):
            bounds = b.region_bounds(n_parts)
            assert len(bounds) == n_parts and bounds[0][0] == 0 and bounds[-1][1] >= (1 << 32)
            assert all(bounds[k][1] == bounds[k + 1][0] and bounds[k][0] <= bounds[k][1] for k in range(n_parts - 1))
            for lo, hi in bounds[:-1]:
                cut = hi
                # reads of a UTR start in [start - 103, stop) and end before stop + 103: no UTR may straddle the cut with that margin
                left = gs - 103 < cut
                assert np.all((ge[left] + 103 <= cut) | ~left[left]) and np.all(gs[~left] - 103 >= cut)


def _scatter_worker(rank, world, port, bam_path, names, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import pickle
    from slamdunk_b200 import parallel
    from slamdunk_b200.engine import HostBatch, wire_arrays
    from wireutil import unpack_wire
    parallel.init("gloo")
    hb = HostBatch(bam_path, names, want_mp=True, group=True, seq2=True, wire=True) if rank == 0 else None
    mine, keep = parallel.scatter_wire(hb.wire if hb is not None else None, rank, world)
    rows, arrays = unpack_wire(wire_arrays(mine))
    with open(os.path.join(out_dir, "rank%d.pkl" % rank), "wb") as fh:
        pickle.dump({"rows": rows, "n_tiles": int(mine.n_tiles), "n_reads": int(mine.n_reads), "mp": arrays["mp"][arrays["first"][2]:].tolist(),
                     "scalars": (int(mine.qual_bits), list(mine.qual_dict), int(mine.max_lseq), int(mine.max_span), int(mine.max_tile_cig),
                                 int(mine.max_tile_mp), int(mine.n_xi_dict))}, fh)
    if rank == 0:
        whole, warr = unpack_wire(hb)
        with open(os.path.join(out_dir, "whole.pkl"), "wb") as fh:
            pickle.dump({"rows": whole, "n_tiles": int(hb.wire.n_tiles), "mp": warr["mp"].tolist()}, fh)
    dist.barrier()
    del keep
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_wire_scatter_gives_every_rank_its_tile_slice(tmp_path, world):
    """computeTconversions under torchrun: rank 0 decodes the BAM, parallel.scatter_wire hands every rank the tile slice
    shard_range() assigns it (gloo, CPU).  The slices, read back with the numpy unpacker, are the whole wire cut at the tile
    boundaries -- records, base codes, qualities, CIGAR ops, MP blocks -- and the scalars travel too."""
    import pickle
    from slamdunk_b200 import parallel
    c = simdata.make_case(str(tmp_path), seed=77, n_reads=900, complexity=1.5, name="c")
    names = list(c.genome.keys())
    port = _free_port()
    mp.spawn(_scatter_worker, args=(world, port, c.bam, names, str(tmp_path)), nprocs=world, join=True)
    whole = pickle.load(open(str(tmp_path / "whole.pkl"), "rb"))
    nt = whole["n_tiles"]
    got_rows, got_mp, scalars = [], [], None
    for r in range(world):
        d = pickle.load(open(str(tmp_path / ("rank%d.pkl" % r)), "rb"))
        t0, t1 = parallel.shard_range(nt, r, world)
        assert d["n_tiles"] == t1 - t0 and d["n_reads"] == min(900, t1 * 32) - t0 * 32
        assert scalars is None or scalars == d["scalars"]
        scalars = d["scalars"]
        got_rows += d["rows"]
        got_mp += d["mp"]
    assert len(got_rows) == len(whole["rows"]) == nt * 32
    assert got_rows == whole["rows"]
    assert got_mp == whole["mp"]


def test_parallel_select_sort_counts_like_the_lexsort_path(tmp_path):
    """oracle.select_sort (C, OpenMP: the bench's CPU arm uses it so that the coordinate sort standing in for the indexed BAM
    runs on all threads) gives count_arrays the same rows and bedgraph entries as numpy's lexsort of the kept reads."""
    from oracle import oracle
    from slamdunk_b200.io import bam as bamio
    simdata.make_case(str(tmp_path), seed=31, n_reads=4000, name="c", complexity=1.5)
    ref, bed, bam = (os.path.join(str(tmp_path), "c" + e) for e in (".fa", ".bed", ".bam"))
    text, refs, records = bamio.read_bam(bam)
    fasta = bamio.read_fasta(ref)
    genome = oracle.Genome(list(fasta.keys()), [s.encode() for s in fasta.values()])
    names = [n for n, _ in refs]
    rng = np.random.RandomState(1)
    perm = rng.permutation(len(records))            # mapper order
    reads = np.zeros(len(records), dtype=oracle.READ_DT)
    quals, mps = [], []
    qo = mo = 0
    for k, i in enumerate(perm):
        r = records[i]
        reads[k]["chrom"] = genome.index.get(names[r.ref_id], -1) if r.ref_id >= 0 else -1
        reads[k]["pos"], reads[k]["end"] = r.pos, r.reference_end
        reads[k]["l_qual"], reads[k]["flag"], reads[k]["mapq"] = len(r.qual), r.flag, r.mapq
        reads[k]["qual_off"] = qo
        quals.append(bytes(r.qual)); qo += len(r.qual)
        if r.has_tag("MP"):
            t = r.get_tag("MP").encode()
            reads[k]["has_mp"], reads[k]["mp_off"], reads[k]["mp_len"] = 1, mo, len(t)
            mps.append(t); mo += len(t)
    rows_bed = oracle.read_bed(bed)
    utrs = np.zeros(len(rows_bed), dtype=oracle.UTR_DT)
    ids = {}
    for i, (ch, s, e, _n, st) in enumerate(rows_bed):
        utrs[i] = (genome.index.get(ch, -1), 1 if ch in names else 0, s, e, 1 if st == "-" else 0, ids.setdefault(ch, len(ids)))
    keep = (rng.rand(len(reads)) < 0.8).astype(np.uint8)
    qual_pool, mp_text = np.frombuffer(b"".join(quals), dtype=np.uint8), b"".join(mps)
    want_rows, want_bg = oracle.count_arrays(genome, reads[keep.astype(bool)], qual_pool, utrs, 27, 1, None, mp_text=mp_text)
    for threads in (1, 4):
        srt, begin, span = oracle.select_sort(reads, keep, len(genome.names), threads=threads)
        assert len(srt) == int(keep.sum()) and begin[-1] == len(srt)
        key = srt["chrom"].astype(np.int64) * (1 << 32) + srt["pos"]
        assert np.all(np.diff(key) >= 0)
        rows, bg = oracle.count_arrays(genome, srt, qual_pool, utrs, 27, 1, None, mp_text=mp_text, threads=threads, presorted=(begin, span))
        assert rows.tobytes() == want_rows.tobytes() and bg.tobytes() == want_bg.tobytes()
    assert want_rows["read_count"].sum() > 1000
