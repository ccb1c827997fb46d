"""The device-side synthetic generator and the count path on its output, against the oracle."""
import numpy as np
import pytest

import oracle_bridge as ob
from oracle import oracle

pytestmark = pytest.mark.gpu


def _setup(genome_bp=3_000_000, n_utr=600, seed=5):
    from slamdunk_b200.engine import CountEngine
    from slamdunk_b200.synth import SynthWorkload
    eng = CountEngine(0)
    wl = SynthWorkload(eng, seed=seed, genome_bp=genome_bp, n_utr=n_utr, n_frac=0.01)
    wl.install()
    return eng, wl


def _oracle(wl, db, triples, min_qual, thr, threads=4):
    g = ob.genome_from_planes(wl.names, wl.chrom_off, wl.chrom_len, wl.lo.cpu().numpy().view(np.uint32),
                              wl.hi.cpu().numpy().view(np.uint32), wl.nmask.cpu().numpy().view(np.uint32))
    reads, qual, mp = ob.batch_to_oracle(ob.device_batch_to_numpy(db), triples.cpu().numpy())
    return oracle.count_arrays(g, reads, qual, ob.annotation_to_oracle(wl.ann, len(wl.names)), min_qual, thr,
                               None, mp_triples=mp, threads=threads)


@pytest.mark.parametrize("read_len,n_reads,mode", [(100, 200_000, "cigar"), (100, 100_000, "verify"), (150, 100_000, "mp"),
                                                   (250, 60_000, "cigar"), (36, 50_000, "verify")])
def test_synthetic_reads_count_like_the_oracle(read_len, n_reads, mode):
    from slamdunk_b200 import _lib as L
    from slamdunk_b200.engine import make_params
    eng, wl = _setup()
    p = wl.params(seed=9, first_read=0, n_reads=n_reads, read_len=read_len, frac_softclip=0.1, frac_indel=0.08,
                  frac_antisense=0.05, frac_labelled=0.8)
    db, triples = wl.generate(p, want_mp=True, want_triples=True)
    eng.count(db, make_params(27, 1, {"cigar": 0, "mp": 1, "verify": 2}[mode]))
    eng.finalize()
    res = eng.results()
    rows, bg = _oracle(wl, db, triples, 27, 1)
    assert res["stats"][L.SD_S_COUNTED] == n_reads
    assert res["stats"][L.SD_S_MP_DISAGREE] == 0
    assert rows["read_count"].sum() > n_reads * 0.9
    assert rows["conversions_on_ts"].sum() > 0
    assert ob.compare(res, rows, bg, wl.ann, eng.pos_off, wl.chrom_off) == []
    eng.close()


def test_generator_is_index_keyed_and_shards_add_up():
    """Any split of the read-index range gives the same reads, and per-shard counters sum to the
    unsharded result (the multi-GPU contract: integer sums are partition independent)."""
    import torch
    from slamdunk_b200.engine import make_params
    eng, wl = _setup()
    n = 90_000
    params = make_params(27, 1, 0)
    whole, _ = wl.generate(wl.params(seed=3, first_read=0, n_reads=n))
    eng.count(whole, params)
    eng.synchronize()
    ref_counters = eng.counters.clone()
    ref_cov = eng.pos_cov.clone()
    ref_tc = eng.pos_tc.clone()
    eng.reset_outputs()
    for first, cnt in ((0, 25_600), (25_600, 40_192), (65_792, n - 65_792)):   # multiples of the tile size
        part, _ = wl.generate(wl.params(seed=3, first_read=first, n_reads=cnt))
        if first == 0:
            a = whole.tensors["rec"][:cnt * 20]
            assert torch.equal(a, part.tensors["rec"][:cnt * 20])
            assert torch.equal(whole.tensors["qual"][:cnt * 100], part.tensors["qual"][:cnt * 100])
        eng.count(part, params)
    eng.synchronize()
    assert torch.equal(eng.counters, ref_counters)
    assert torch.equal(eng.pos_cov, ref_cov) and torch.equal(eng.pos_tc, ref_tc)
    eng.close()


def test_region_sharding_matches_unsharded():
    """UTR-owner region sharding: reads generated per region and counted separately
    sum to the same table as all of them counted together."""
    import torch
    from slamdunk_b200.engine import make_params
    eng, wl = _setup()
    params = make_params(27, 1, 0)
    parts = wl.region_split(4)
    assert sum(c for _f, c in parts) == len(wl.ann)
    batches = [wl.generate(wl.params(seed=11 + k, first_read=0, n_reads=30_000, utr_first=f, utr_count=c))[0]
               for k, (f, c) in enumerate(parts)]
    total = torch.zeros_like(eng.counters)
    for b in batches:
        eng.reset_outputs()
        eng.count(b, params)
        eng.synchronize()
        total += eng.counters
    eng.reset_outputs()
    for b in batches:
        eng.count(b, params)
    eng.synchronize()
    assert torch.equal(total, eng.counters)
    eng.close()


def test_count_host_equals_count_device():
    """The end-to-end entry (host buffers, chunked H2D overlapped with the kernel) == device-resident."""
    import torch
    from slamdunk_b200.engine import make_params
    from slamdunk_b200.synth import batch_to_host
    eng, wl = _setup()
    params = make_params(27, 1, 0)
    db, _ = wl.generate(wl.params(seed=21, first_read=0, n_reads=300_000))
    eng.count(db, params)
    eng.synchronize()
    want = (eng.counters.clone(), eng.pos_cov.clone(), eng.pos_tc.clone(), eng.stats.clone())
    for coded in (False, True):       # one byte per quality / the dictionary codes sd_decode_bam writes for binned qualities
        hb, keep = batch_to_host(db, eng if coded else None)
        assert hb.qual_bits == (2 if coded else 8)       # the generator draws from three quality values
        eng.reset_outputs()
        eng.count_host(hb, params)
        got = (eng.counters, eng.pos_cov, eng.pos_tc, eng.stats)
        for w, g in zip(want, got):
            assert torch.equal(w, g)
    with pytest.raises(Exception, match="codes"):         # device entry points take bytes only
        import ctypes
        eng._chk(eng.lib.sd_count(eng.ctx, ctypes.byref(hb), ctypes.byref(params)), "sd_count")
    eng.close()


def test_filter_predicates_fused_into_count():
    """filter_on: the default-branch predicates of filter.Filter (filter.py:241-251) run in front of the
    scan; stats equal the oracle's filter, counters equal counting only the survivors."""
    from slamdunk_b200 import _lib as L
    from slamdunk_b200.engine import make_params
    eng, wl = _setup()
    n = 80_000
    db, triples = wl.generate(wl.params(seed=31, first_read=0, n_reads=n, p_err=0.02, frac_indel=0.2, frac_multimap=0.2),
                              want_triples=True)
    eng.count(db, make_params(27, 1, 0, filter_on=True, filter_mq=2, filter_min_identity=0.95, filter_max_nm=3))
    eng.finalize()
    res = eng.results()
    bn = ob.device_batch_to_numpy(db)
    rec = bn["rec"][:n]
    fr = np.zeros(n, dtype=oracle.FREAD_DT)
    flags = rec[:, 1] >> 24
    fr["flag"] = np.where(flags & L.SD_F_REVERSE, 16, 0)
    fr["mapq"] = (rec[:, 1] >> 16) & 0xFF
    fr["xi"] = rec[:, 2].view(np.float32)
    fr["nm"] = rec[:, 4] & 0xFFFF
    keep = np.zeros(n, dtype=np.uint8)
    stats = np.zeros(7, dtype=np.int64)
    oracle.lib().orc_filter_default(fr.ctypes.data, n, 2, 0.95, 3, keep.ctypes.data, stats.ctypes.data)
    assert res["stats"][:6].tolist() == stats[:6].tolist()
    assert 0 < stats[2] < n and stats[3] > 0 and stats[4] > 0 and stats[5] > 0
    g = ob.genome_from_planes(wl.names, wl.chrom_off, wl.chrom_len, wl.lo.cpu().numpy().view(np.uint32),
                              wl.hi.cpu().numpy().view(np.uint32), wl.nmask.cpu().numpy().view(np.uint32))
    reads, qual, mp = ob.batch_to_oracle(bn, triples.cpu().numpy())
    rows, bg = oracle.count_arrays(g, reads[keep.astype(bool)], qual, ob.annotation_to_oracle(wl.ann, len(wl.names)),
                                   27, 1, None, mp_triples=mp, threads=4)
    assert ob.compare(res, rows, bg, wl.ann, eng.pos_off, wl.chrom_off) == []
    eng.close()


def test_snp_masks_at_scale_including_key_collisions():
    """cfg4-style: a dense VCF (1 SNP / 500 bp of annotated sequence) on chromosomes whose names collide under
    the reference's chrom+pos string key (chr1 / chr12 / chr19 ..., utils/SNPtools.py:33,54), 250 bp reads, 15 % mapq 0."""
    from slamdunk_b200.engine import make_params
    from slamdunk_b200.utils.snpmask import SNPMasks
    eng, wl = _setup(genome_bp=2_500_000, n_utr=500, seed=8)
    rng = np.random.RandomState(4)
    lo = wl.lo.cpu().numpy().view(np.uint32)
    hi = wl.hi.cpu().numpy().view(np.uint32)
    nm = wl.nmask.cpu().numpy().view(np.uint32)
    g = ob.genome_from_planes(wl.names, wl.chrom_off, wl.chrom_len, lo, hi, nm)
    records = []
    for u in range(len(wl.ann)):
        c, s, e = int(wl.ann.chrom[u]), int(wl.ann.start[u]), int(wl.ann.stop[u])
        seq = g.seqs[c]
        for p in rng.randint(s, e, size=max(1, (e - s) // 500)):
            b = chr(seq[p])
            if b == "T":
                records.append((wl.names[c], str(p + 1), "T", rng.choice(["C", "C", "G", "C,A"])))
            elif b == "A":
                records.append((wl.names[c], str(p + 1), "A", rng.choice(["G", "G", "g", "T"])))
    # positions on chr1x that alias positions on chr1 under the string key: 'chr1' + '2345' == 'chr12' + '345'
    for c, name in enumerate(wl.names):
        if name.startswith("chr1") and len(name) == 5:
            for p in rng.randint(1, min(int(wl.chrom_len[c]), 9999), size=300):
                records.append((name, str(p), "T", "C"))
                records.append((name, str(p), "A", "G"))
    sm = SNPMasks(wl.names, wl.chrom_off, wl.chrom_len, wl.n_words)
    for r in records:
        sm.add(*r)
    tcm, agm = sm.masks()
    assert sm.n_tc > len([r for r in records if r[2] == "T" and r[3] == "C"]) * 0.5
    eng.set_snps(tcm, agm)
    n = 60_000
    db, triples = wl.generate(wl.params(seed=77, first_read=0, n_reads=n, read_len=250, frac_multimap=0.15, p_conv=0.03),
                              want_mp=True, want_triples=True, coordinate_sorted=True)
    eng.count(db, make_params(27, 1, 2))
    eng.finalize()
    res = eng.results()
    reads, qual, mp = ob.batch_to_oracle(ob.device_batch_to_numpy(db), triples.cpu().numpy())
    h = oracle.make_snps(records)
    try:
        rows, bg = oracle.count_arrays(g, reads, qual, ob.annotation_to_oracle(wl.ann, len(wl.names)), 27, 1, h, mp_triples=mp, threads=4)
        rows0, _ = oracle.count_arrays(g, reads, qual, ob.annotation_to_oracle(wl.ann, len(wl.names)), 27, 1, None, mp_triples=mp, threads=4)
    finally:
        oracle.free_snps(h)
    assert rows0["conversions_on_ts"].sum() > rows["conversions_on_ts"].sum() > 0      # the mask removes conversions
    assert rows["multimap_count"].sum() > 0
    assert ob.compare(res, rows, bg, wl.ann, eng.pos_off, wl.chrom_off) == []
    eng.close()


def test_file_pipeline_on_a_synthetic_bam(tmp_path):
    """cfg2 in miniature through the FILE interface: generator -> FASTA/BED/BAM on disk -> computeTconversions
    (C++ decode with shape grouping, GPU count) versus the oracle's count_files on the same files."""
    import io
    import synth_files
    from slamdunk_b200.dunks import tcounter
    eng, wl = _setup(genome_bp=4_000_000, n_utr=800, seed=21)
    db, triples = wl.generate(wl.params(seed=5, first_read=0, n_reads=120_000, frac_multimap=0.05), want_triples=True,
                              coordinate_sorted=True, group=False)
    ref, bed, bam, n = synth_files.write_workload_files(str(tmp_path), wl, db, triples)
    eng.close()
    out, orc = str(tmp_path / "gpu"), str(tmp_path / "orc")
    glog, olog = io.StringIO(), io.StringIO()
    tcounter.computeTconversions(ref, bed, None, bam, 100, 27, out + ".tsv", out + ".p", out + ".m", 1, glog)
    tm = tcounter.OPTIONS["timings"]
    oracle.count_files(ref, bed, None, bam, 100, 27, orc + ".tsv", orc + ".p", orc + ".m", 1, olog, threads=8)
    for ext in (".tsv", ".p", ".m"):
        assert open(out + ext).read() == open(orc + ext).read(), ext
    assert glog.getvalue() == olog.getvalue() == ""
    assert tm["bam_reads"] == n and tm["stats"][8] == 0       # no MP / CIGAR-walk disagreement
    print("decode: inflate %.3fs records %.3fs; gpu count %.3fs; write %.3fs for %d reads"
          % (tm["bam_inflate_s"], tm["bam_decode_s"], tm["gpu_count_s"], tm["write_s"], n))


def test_full_size_cfg2_properties():
    """BASELINE.json configs[1] at full size (50 M x 100 bp, 20 000 UTRs, 100 Mb genome): too big for the oracle, so
    size-independent properties -- counting the tile range in three uneven pieces equals counting it whole, a second
    pass reproduces the first bit for bit (integer atomics: order independent), every counted read is attributed
    consistently, and the coverage integral equals the summed clipped read spans recorded in difference form."""
    import ctypes
    import torch
    from slamdunk_b200 import _lib as L
    from slamdunk_b200.engine import CountEngine, DeviceBatch, make_params
    from slamdunk_b200.synth import SynthWorkload
    eng = CountEngine(0)
    wl = SynthWorkload(eng, seed=1, genome_bp=100_000_000, n_utr=20000)
    wl.install()
    n = 50_000_000
    db, _ = wl.generate(wl.params(seed=3, first_read=0, n_reads=n), coordinate_sorted=True)
    params = make_params(27, 1, L.SD_MISMATCH_CIGAR, filter_on=True, filter_mq=2, filter_min_identity=0.95)
    eng.count(db, params)
    eng.synchronize()
    whole = (eng.counters.clone(), eng.pos_cov.clone(), eng.pos_tc.clone(), eng.stats.clone())
    # idempotence
    eng.reset_outputs()
    eng.count(db, params)
    eng.synchronize()
    for a, b in zip(whole, (eng.counters, eng.pos_cov, eng.pos_tc, eng.stats)):
        assert torch.equal(a, b)
    # three uneven tile ranges (device sub-batches: record and tile-table pointers advance, columns stay)
    nt = db.batch.n_tiles
    cuts = [0, nt // 7, nt // 7 + nt // 2, nt]
    eng.reset_outputs()
    for t0, t1 in zip(cuts[:-1], cuts[1:]):
        sb = L.Batch()
        ctypes.memmove(ctypes.byref(sb), ctypes.byref(db.batch), ctypes.sizeof(L.Batch))
        sb.n_tiles = t1 - t0
        sb.n_reads = min(n, t1 * 32) - t0 * 32
        sb.rec = db.batch.rec + t0 * 32 * 20
        sb.tiles = db.batch.tiles + t0 * 32
        eng.count(DeviceBatch(sb, db.tensors), params)
    eng.synchronize()
    for a, b in zip(whole, (eng.counters, eng.pos_cov, eng.pos_tc, eng.stats)):
        assert torch.equal(a, b)
    st = whole[3].cpu().numpy()
    c = whole[0].cpu().numpy().reshape(-1, L.SD_NCOUNTER)
    assert st[L.SD_S_MAPPED] == n and st[L.SD_S_FILTERED] == st[L.SD_S_COUNTED] < n     # mapq-0 reads (3 %) are filtered
    assert st[L.SD_S_MQFILTERED] + st[L.SD_S_IDFILTERED] + st[L.SD_S_FILTERED] == n
    assert c[:, L.SD_C_READS].sum() >= 0.9 * st[L.SD_S_COUNTED]                         # ~98 % sense-strand reads hit their UTR
    assert (c[:, L.SD_C_TCREADS] <= c[:, L.SD_C_READS]).all() and (c[:, L.SD_C_CONV_ON_T] <= c[:, L.SD_C_COV_ON_T]).all()
    assert c[:, L.SD_C_MULTIMAP].sum() == 0                                             # filtered out by MQ >= 2
    # difference form: +1 / -1 per (read, segment) cancel; after the scan no position is negative
    assert int(whole[1].sum()) == 0
    eng.finalize()
    eng.synchronize()
    assert int(eng.pos_cov.min()) >= 0 and int(eng.pos_tc.min()) >= 0
    assert int((eng.pos_tc > eng.pos_cov).sum()) == 0                                   # conversions only where covered
    eng.close()


def test_deep_coverage_tiles_aggregate_their_coverage_atomics():
    """Ten reads per annotated base: the 32 sorted reads of a tile start within a few positions, which switches the
    kernel to warp-aggregated coverage updates (the multi-GPU region shards look like this); counts stay exact."""
    from slamdunk_b200 import _lib as L
    from slamdunk_b200.engine import make_params
    eng, wl = _setup(genome_bp=400_000, n_utr=24, seed=11)
    p = wl.params(seed=21, first_read=0, n_reads=300_000, read_len=100, frac_softclip=0.05, frac_indel=0.02,
                  frac_antisense=0.02, frac_labelled=1.0)
    db, triples = wl.generate(p, want_mp=True, want_triples=True, coordinate_sorted=True)
    pos = db.tensors["rec"].cpu().numpy().view(np.int32).reshape(-1, 5)[: 300_000, 0].astype(np.int64)
    starts = pos.reshape(-1, 32)
    assert np.median(starts[:, -1] - starts[:, 0]) <= 8, "workload is not deep enough to reach the aggregated path"
    eng.count(db, make_params(27, 1, 0))
    eng.finalize()
    res = eng.results()
    rows, bg = _oracle(wl, db, triples, 27, 1)
    assert res["stats"][L.SD_S_COUNTED] == 300_000
    assert ob.compare(res, rows, bg, wl.ann, eng.pos_off, wl.chrom_off) == []
    eng.close()


@pytest.mark.parametrize("read_len,n_reads,mode,slow", [(100, 150_000, "cigar", False), (150, 80_000, "verify", False), (250, 40_000, "mp", False),
                                                        (36, 30_000, "cigar", False), (400, 8_000, "cigar", False), (100, 20_000, "cigar", True)])
def test_two_bit_seq_column_counts_like_the_plane_form(read_len, n_reads, mode, slow):
    """sd_batch.seq_bits = 2 (sd_seq_pack; what sd_decode_bam hands the count path): every output identical to the run on
    the four-plane column, on all window sizes, the per-base path, from the device and through sd_count_host."""
    import torch
    from slamdunk_b200 import _lib as L
    from slamdunk_b200.engine import make_params
    from slamdunk_b200.synth import batch_to_host, to_seq2
    eng, wl = _setup()
    p = wl.params(seed=21, first_read=0, n_reads=n_reads, read_len=read_len, frac_softclip=0.1, frac_indel=0.08,
                  frac_antisense=0.05, frac_labelled=0.8)
    db, _ = wl.generate(p, want_mp=True)
    params = make_params(27, 1, {"cigar": 0, "mp": 1, "verify": 2}[mode], force_slow_path=slow)
    read_tc4 = torch.zeros(int(db.batch.n_tiles) * 32, dtype=torch.uint8, device=eng.device)
    eng.count(db, params, read_tc=read_tc4)
    eng.finalize()
    want = eng.results()
    db2 = to_seq2(db, eng)
    assert db2.batch.seq_bits == 2 and db2.tensors["seq"].numel() * 2 == db.tensors["seq"].numel()
    eng.reset_outputs()
    read_tc2 = torch.zeros_like(read_tc4)
    eng.count(db2, params, read_tc=read_tc2)
    eng.finalize()
    got = eng.results()
    assert want["stats"][L.SD_S_COUNTED] == n_reads and want["counters"][:, L.SD_C_CONV_ON_T].sum() > 0
    for k in ("counters", "stats", "pos_cov", "pos_tc", "tcontent"):
        assert np.array_equal(want[k], got[k]), k
    assert torch.equal(read_tc4, read_tc2)
    hb, keep = batch_to_host(db2, eng)
    assert hb.seq_bits == 2
    eng.reset_outputs()
    eng.count_host(hb, params)
    eng.finalize()
    got = eng.results()
    for k in ("counters", "stats", "pos_cov", "pos_tc"):
        assert np.array_equal(want[k], got[k]), "host " + k
    eng.close()


def test_time_course_samples_counted_in_one_batched_run():
    """BASELINE.json config 3 in miniature: samples of a 4sU time course (labelled-read fraction 1 - exp(-ln2 t / halflife)
    per UTR, 150 bp reads) counted one after the other by one engine on one genome / annotation; every sample equals the
    oracle, and the conversion-carrying reads grow with the labelling time."""
    from slamdunk_b200 import _lib as L
    from slamdunk_b200.engine import make_params
    from slamdunk_b200.synth import to_seq2
    eng, wl = _setup()
    params = make_params(27, 1, 0)
    tc_reads = []
    for k, minutes in enumerate([0, 30, 120, 720]):
        frac = wl.set_time_point(minutes)
        assert frac.min() >= 0.0 and (minutes == 0 or 0.02 < frac.mean() < 1.0)
        p = wl.params(seed=10 + k, first_read=0, n_reads=40_000, read_len=150)
        db, triples = wl.generate(p, want_triples=True)
        eng.reset_outputs()
        eng.count(to_seq2(db, eng), params)
        eng.finalize()
        res = eng.results()
        rows, bg = _oracle(wl, db, triples, 27, 1)
        assert ob.compare(res, rows, bg, wl.ann, eng.pos_off, wl.chrom_off) == []
        tc_reads.append(int(res["counters"][:, L.SD_C_TCREADS].sum()))
    wl.set_time_point(None)
    assert tc_reads[0] < tc_reads[1] < tc_reads[2] < tc_reads[3]
    eng.close()


@pytest.mark.parametrize("read_len,n_reads,mode,slow,min_qual,n_qual", [
    (100, 150_000, "cigar", False, 27, 3), (100, 60_000, "cigar", False, 31, 3), (100, 60_000, "cigar", False, 0, 3), (100, 60_000, "cigar", False, 50, 3),
    (150, 80_000, "verify", False, 27, 3), (250, 40_000, "mp", False, 27, 3), (36, 30_000, "cigar", False, 20, 3), (400, 8_000, "cigar", False, 27, 3),
    (100, 20_000, "cigar", True, 27, 3), (100, 60_000, "cigar", False, 27, 11), (150, 40_000, "verify", False, 24, 11), (100, 30_000, "cigar", False, 27, 40)])
def test_plane_form_qualities_count_like_the_byte_column(read_len, n_reads, mode, slow, min_qual, n_qual):
    """sd_batch.qual_form = 1 (sd_qual_dict + sd_qual_to_planes): every output identical to the run on the byte column -- 2 and 4
    planes, thresholds below / between / above the dictionary values, all mismatch sources, the per-base path, from the device
    and through sd_count_host (which builds the planes itself); more than 16 distinct qualities: no plane form."""
    import torch
    from slamdunk_b200 import _lib as L
    from slamdunk_b200.engine import make_params
    from slamdunk_b200.synth import batch_to_host, to_qual_planes, to_seq2
    eng, wl = _setup()
    p = wl.params(seed=31, first_read=0, n_reads=n_reads, read_len=read_len, frac_softclip=0.1, frac_indel=0.08,
                  frac_antisense=0.05, frac_labelled=0.8)
    db, _ = wl.generate(p, want_mp=True)
    if n_qual != 3:      # spread the three generated values over n_qual distinct ones (pad bytes between tiles stay 0)
        q = db.tensors["qual"]
        idx = torch.arange(q.numel(), device=q.device, dtype=torch.int64)
        spread = (q.to(torch.int64) - 19 + (idx * 7919) % (n_qual // 3 + 1) * 3).clamp(2, 93).to(torch.uint8)
        q.copy_(torch.where(q > 0, spread, q))
    db2 = to_seq2(db, eng)
    params = make_params(min_qual, 1, {"cigar": 0, "mp": 1, "verify": 2}[mode], force_slow_path=slow)
    read_tc_b = torch.zeros(int(db.batch.n_tiles) * 32, dtype=torch.uint8, device=eng.device)
    eng.count(db2, params, read_tc=read_tc_b)
    eng.finalize()
    want = eng.results()
    dbq = to_qual_planes(db2, eng)
    if n_qual > 16:
        assert dbq is None
        eng.close()
        return
    assert dbq.batch.qual_form == 1 and dbq.batch.qual_bits == (2 if n_qual <= 4 else 4)
    eng.reset_outputs()
    read_tc_p = torch.zeros_like(read_tc_b)
    eng.count(dbq, params, read_tc=read_tc_p)
    eng.finalize()
    got = eng.results()
    assert want["stats"][L.SD_S_COUNTED] == n_reads
    if min_qual < 40:
        assert want["counters"][:, L.SD_C_CONV_ON_T].sum() > 0
    else:
        assert want["counters"][:, L.SD_C_CONV_ON_T].sum() == 0
    for k in ("counters", "stats", "pos_cov", "pos_tc", "tcontent"):
        assert np.array_equal(want[k], got[k]), k
    assert torch.equal(read_tc_b, read_tc_p)
    hb, keep = batch_to_host(db2, eng)
    assert hb.seq_bits == 2 and hb.qual_bits in (2, 4)
    eng.reset_outputs()
    eng.count_host(hb, params)
    eng.finalize()
    got = eng.results()
    for k in ("counters", "stats", "pos_cov", "pos_tc"):
        assert np.array_equal(want[k], got[k]), "host " + k
    eng.close()


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("sorted_", [True, False])
def test_staged_and_plain_kernels_agree(mode, sorted_):
    """The count kernel that stages each tile's reference window and annotation entries with the tile (batches with span
    descriptors, sd_tile_spans) gives the outputs of the kernel that looks them up per read -- on coordinate-sorted reads
    (nearly every tile staged) and in mapper order (most tiles handed on to the plain kernel), in every mismatch mode, with
    the filter predicates fused, and per read."""
    import torch
    from slamdunk_b200.engine import make_params
    from slamdunk_b200.synth import to_qual_planes, to_seq2
    eng, wl = _setup(genome_bp=2_000_000, n_utr=900, seed=11)
    p = wl.params(seed=4, first_read=0, n_reads=150_000, read_len=100, frac_softclip=0.1, frac_indel=0.08, frac_antisense=0.1)
    db, _ = wl.generate(p, want_mp=True, coordinate_sorted=sorted_)
    params = make_params(27, 1, mode, filter_on=True, filter_mq=2, filter_min_identity=0.95, filter_max_nm=-1)
    forms = [db, to_seq2(db, eng)]
    qp = to_qual_planes(forms[1], eng)
    if qp is not None:
        forms.append(qp)
    for form in forms:
        outs = []
        for spans in (False, True):
            eng.reset_outputs()
            rtc = torch.zeros(form.batch.n_tiles * 32, dtype=torch.uint8, device=eng.device)
            eng.count(form, params, rtc, spans=spans)
            eng.finalize()
            eng.synchronize()
            assert bool(form.batch.spans) == spans
            outs.append((eng.counters.clone(), eng.pos_cov.clone(), eng.pos_tc.clone(), eng.stats.clone(), rtc))
        for a, b in zip(*outs):
            assert torch.equal(a, b)
        assert int(outs[0][3][6]) > 100_000
    eng.close()


def test_stale_span_descriptors_are_ignored():
    """Descriptors made for another annotation must not be used: after sd_set_utrs the library ignores them (epoch), and the
    engine attaches fresh ones."""
    import torch
    from slamdunk_b200.engine import make_params
    eng, wl = _setup(genome_bp=1_000_000, n_utr=300, seed=3)
    db, _ = wl.generate(wl.params(seed=2, first_read=0, n_reads=40_000), coordinate_sorted=True)
    params = make_params(27, 1, 0)
    eng.count(db, params)
    a = wl.ann
    keep = np.arange(len(a)) % 2 == 0                       # another annotation: every second UTR
    eng.set_utrs(a.chrom[keep], a.start[keep], a.stop[keep], a.strand[keep], None)
    stale_key = db.spans_key
    db.spans_key = (eng.ctx.value, eng._epoch)               # pretend they are current: the library must still refuse them
    eng.count(db, params)
    eng.finalize()
    eng.synchronize()
    got = eng.counters.clone()
    db.spans_key = stale_key
    eng.reset_outputs()
    eng.count(db, params)                                    # fresh descriptors
    eng.finalize()
    eng.synchronize()
    assert torch.equal(got, eng.counters)
    eng.reset_outputs()
    eng.count(db, params, spans=False)
    eng.finalize()
    eng.synchronize()
    assert torch.equal(got, eng.counters)
    eng.close()


@pytest.mark.parametrize("mode,sorted_,qvals", [(0, True, 3), (2, True, 3), (1, False, 3), (0, False, 11), (2, True, 40)])
def test_wire_form_counts_like_the_device_batch(mode, sorted_, qvals):
    """sd_wire_plan / sd_wire_fill pack a device batch into the wire form, sd_count_wire expands it on the device and counts it:
    same outputs as counting the device batch directly -- plain-match tiles without l_seq / CIGAR, mixed tiles, tiles whose
    positions need 32 bits or that cross chromosomes (mapper order), 2-bit / 4-bit / byte qualities, with and without the
    filter columns and the MP column."""
    import torch
    from slamdunk_b200.engine import make_params
    from slamdunk_b200.synth import batch_to_wire, to_seq2
    eng, wl = _setup(genome_bp=2_000_000, n_utr=900, seed=21)
    p = wl.params(seed=8, first_read=0, n_reads=120_000 - 7, read_len=100 if qvals == 3 else 77, frac_softclip=0.1, frac_indel=0.08,
                  frac_antisense=0.1)
    db, _ = wl.generate(p, want_mp=True, coordinate_sorted=sorted_)
    if qvals != 3:      # more distinct qualities than the generator makes: 4-bit codes / plain bytes on the wire
        g = torch.Generator(device=eng.device).manual_seed(5)
        db.tensors["qual"].copy_(torch.randint(0, qvals, db.tensors["qual"].shape, generator=g, device=eng.device, dtype=torch.uint8) + 10)
    db2 = to_seq2(db, eng)
    hw = batch_to_wire(db2, eng)
    assert hw.wire.qual_bits == {3: 2, 11: 4, 40: 8}[qvals]
    for filt in (False, True):
        params = make_params(27, 1, mode, filter_on=filt, filter_mq=2, filter_min_identity=0.95, filter_max_nm=3 if filt else -1)
        eng.reset_outputs()
        eng.count(db2, params)
        eng.finalize()
        eng.synchronize()
        want = (eng.counters.clone(), eng.pos_cov.clone(), eng.pos_tc.clone(), eng.stats.clone())
        eng.reset_outputs()
        eng.count_wire(hw.wire, params)
        eng.finalize()
        eng.synchronize()
        for a, b in zip(want, (eng.counters, eng.pos_cov, eng.pos_tc, eng.stats)):
            assert torch.equal(a, b)
        assert int(want[3][6]) > 50_000
        assert hw.h2d_bytes(params) < (0.8 if qvals == 3 else 1.0) * sum(t.numel() * t.element_size() for t in db2.tensors.values() if t is not None)
    eng.close()
