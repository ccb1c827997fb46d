"""Parity of the CUDA count path (through the C ABI) with the oracle and the committed golden
vectors.  Needs a GPU: run with `pytest -m gpu` on the B200 box."""
import io
import os

import numpy as np
import pytest

import goldens
import simdata
from oracle import oracle

pytestmark = pytest.mark.gpu


def _read(p):
    with open(p) as fh:
        return fh.read()


def _count(case, tmp_path, mode="verify", tag="out"):
    from slamdunk_b200.dunks import tcounter
    out = str(tmp_path / tag)
    log = io.StringIO()
    tcounter.OPTIONS["mismatch_source"] = mode
    tcounter.computeTconversions(case.ref, case.bed, case.vcf, case.bam, case.max_read_length, case.min_qual,
                                 out + ".tsv", out + "_plus.bedgraph", out + "_mins.bedgraph", case.conv_threshold, log)
    tcounter.OPTIONS["mismatch_source"] = "verify"
    return out, log.getvalue(), tcounter.OPTIONS["timings"]


@pytest.mark.parametrize("mode", ["verify", "cigar", "mp"])
@pytest.mark.parametrize("name", goldens.COUNT_CASES)
def test_count_matches_reference_golden(name, mode, tmp_path):
    """Files written by the GPU path == files written by the reference's own code (tests/golden)."""
    c = goldens.CountCase(name)
    out, log, tm = _count(c, tmp_path, mode)
    assert _read(out + ".tsv") == _read(c.tsv)
    assert _read(out + "_plus.bedgraph") == _read(c.plus)
    assert _read(out + "_mins.bedgraph") == _read(c.minus)
    from slamdunk_b200 import _lib as L
    if mode == "verify":
        assert tm["stats"][L.SD_S_MP_DISAGREE] == 0
    if name != "cfg1":
        assert log == _read(os.path.join(c.dir, "expected.log"))
    assert tm["gpu_launches"] >= 4


@pytest.mark.parametrize("name", goldens.COUNT_CASES)
def test_count_matches_reference_golden_with_the_plane_seq_column(name, tmp_path):
    """the default hands sd_count the 2-bit SEQ column (OPTIONS["seq_bits"] = 2); the four-plane form gives the same files"""
    from slamdunk_b200.dunks import tcounter
    c = goldens.CountCase(name)
    tcounter.OPTIONS["seq_bits"] = 4
    try:
        out, _log, _tm = _count(c, tmp_path)
    finally:
        tcounter.OPTIONS["seq_bits"] = 2
    assert _read(out + ".tsv") == _read(c.tsv)
    assert _read(out + "_plus.bedgraph") == _read(c.plus)
    assert _read(out + "_mins.bedgraph") == _read(c.minus)


def _engine_run(case_files, min_qual, thr, mode, force_slow=False, tile_reads=32, want_read_tc=False, snps=True):
    """Engine-level run -> results dict (+ per-read tc)."""
    import torch
    from slamdunk_b200 import _lib as L
    from slamdunk_b200.engine import CountEngine, DeviceBatch, HostBatch, HostGenome, make_params
    from slamdunk_b200.utils.bed import BedIterator
    from slamdunk_b200.utils.snpmask import build_snp_masks
    ref, bed, vcf, bam = case_files
    g = HostGenome(ref)
    hb = HostBatch(bam, g.names, want_mp=True, tile_reads=tile_reads)
    utrs = list(BedIterator(bed))
    eng = CountEngine(0)
    eng.set_reference(g.lo, g.hi, g.nmask, g.chrom_off, g.chrom_len)
    if vcf and snps:
        tc, ag = build_snp_masks(vcf, g.names, g.chrom_off, g.chrom_len, g.n_words)
        if tc is not None:
            eng.set_snps(tc, ag)
    bam_names = set(hb.ref_names)
    eng.set_utrs([g.index.get(u.chromosome, L.SD_REF_NONE) for u in utrs], [u.start for u in utrs], [u.stop for u in utrs],
                 [1 if u.strand == "-" else 0 for u in utrs], [1 if u.chromosome in bam_names else 0 for u in utrs])
    db = DeviceBatch.from_host(hb.batch, eng.device)
    params = make_params(min_qual, thr, {"cigar": 0, "mp": 1, "verify": 2}[mode], force_slow_path=force_slow)
    rtc = torch.zeros(hb.batch.n_tiles * hb.batch.tile_reads, dtype=torch.uint8, device=eng.device) if want_read_tc else None
    eng.count(db, params, rtc)
    eng.finalize()
    res = eng.results()
    if rtc is not None:
        res["read_tc"] = rtc.cpu().numpy()[:hb.n_reads]
    res["kernel_ms"] = eng.last_kernel_ms()
    eng.close()
    hb.close()
    g.close()
    return res


@pytest.mark.parametrize("name", ["case_b", "case_d"])
@pytest.mark.parametrize("mode", ["cigar", "mp"])
def test_per_base_path_equals_word_parallel_path(name, mode):
    """Two independent device implementations (register-window bit ops vs plain per-base loops)."""
    c = goldens.CountCase(name)
    files = (c.ref, c.bed, c.vcf, c.bam)
    a = _engine_run(files, c.min_qual, c.conv_threshold, mode, force_slow=False)
    b = _engine_run(files, c.min_qual, c.conv_threshold, mode, force_slow=True)
    from slamdunk_b200 import _lib as L
    assert b["stats"][L.SD_S_SLOWPATH] == b["stats"][L.SD_S_COUNTED] > 0
    assert a["stats"][L.SD_S_SLOWPATH] < a["stats"][L.SD_S_COUNTED]
    for k in ("counters", "tcontent", "pos_cov", "pos_tc"):
        assert np.array_equal(a[k], b[k]), k
    assert a["stats"][L.SD_S_TC_TOTAL] == b["stats"][L.SD_S_TC_TOTAL]


def test_per_read_tc_counts_match_mp_tags():
    """read_tc[i] == what fillMismatchesNGM + the threshold rule give for read i (SlamSeqFile.py:313-348, :397)."""
    from slamdunk_b200.io import bam as bamio
    c = goldens.CountCase("case_c")     # no SNPs
    res = _engine_run((c.ref, c.bed, None, c.bam), c.min_qual, c.conv_threshold, "cigar", want_read_tc=True)
    _t, refs, recs = bamio.read_bam(c.bam)
    fasta = set(bamio.read_fasta(c.ref).keys())
    want = []
    for r in recs:
        n = 0
        if r.has_tag("MP") and refs[r.ref_id][0] in fasta:
            for ent in r.get_tag("MP").split(","):
                conv, rp, _fp = (int(x) for x in ent.split(":"))
                if conv == (2 if r.flag & 16 else 16) and r.qual[rp - 1] >= c.min_qual:
                    n += 1
        want.append(n if n >= c.conv_threshold else 0)
    assert res["read_tc"].tolist() == want


@pytest.mark.parametrize("seed,thr,minq,read_len,reads", [(301, 1, 27, (40, 100), 20000), (302, 2, 20, (100, 100), 20000),
                                                          (303, 1, 27, (150, 250), 8000), (304, 1, 27, (300, 400), 3000),
                                                          (305, 1, 27, (480, 700), 1500)])
def test_count_matches_oracle_on_random_data(seed, thr, minq, read_len, reads, tmp_path):
    """Seeded data sets too big for the Python reference but quick for the C oracle."""
    c = simdata.make_case(str(tmp_path), seed=seed, n_reads=reads, read_len=read_len, n_utr=120, complexity=1.5,
                          chroms=(("chr1", 90000), ("chr12", 20000), ("chr2", 30000), ("chrX", 8000)), snp_density=0.006)
    log = io.StringIO()
    o = str(tmp_path / "orc")
    oracle.count_files(c.ref, c.bed, c.vcf, c.bam, 100, minq, o + ".tsv", o + "_p.bg", o + "_m.bg", thr, log, threads=4)

    class Case:
        ref, bed, vcf, bam = c.ref, c.bed, c.vcf, c.bam
        max_read_length, min_qual, conv_threshold = 100, minq, thr
    out, glog, tm = _count(Case, tmp_path, "verify", "gpu")
    assert _read(out + ".tsv") == _read(o + ".tsv")
    assert _read(out + "_plus.bedgraph") == _read(o + "_p.bg")
    assert _read(out + "_mins.bedgraph") == _read(o + "_m.bg")
    assert glog == log.getvalue()


def test_mp_tags_that_disagree_with_the_cigar_walk_fall_back_to_mp_mode(tmp_path):
    """A BAM whose MP tags describe other mismatches than SEQ/CIGAR (e.g. aligned to another FASTA):
    the reference trusts MP; so must we, and the log says why."""
    from slamdunk_b200.io import bam as bamio
    c = simdata.make_case(str(tmp_path), seed=55, n_reads=1500, snp_density=0.0)
    text, refs, recs = bamio.read_bam(c.bam)
    changed = 0
    for r in recs:
        if r.has_tag("MP") and changed < 40:
            ents = r.get_tag("MP").split(",")
            conv, rp, fp = (int(x) for x in ents[0].split(":"))
            if conv in (2, 16):
                continue
            ents[0] = "%d:%d:%d" % (2 if r.flag & 16 else 16, rp, fp)   # claim a conversion SEQ does not show
            r.set_tag("MP", ",".join(ents), "Z")
            changed += 1
    assert changed > 5
    bam2 = str(tmp_path / "tampered.bam")
    bamio.write_bam(bam2, text, refs, recs)
    log = io.StringIO()
    o = str(tmp_path / "orc")
    oracle.count_files(c.ref, c.bed, None, bam2, 100, 27, o + ".tsv", o + "_p.bg", o + "_m.bg", 1, log)

    class Case:
        ref, bed, vcf, bam = c.ref, c.bed, None, bam2
        max_read_length, min_qual, conv_threshold = 100, 27, 1
    out, glog, tm = _count(Case, tmp_path, "verify", "gpu")
    assert "disagree" in glog
    assert _read(out + ".tsv") == _read(o + ".tsv")
    assert _read(out + "_plus.bedgraph") == _read(o + "_p.bg")
    assert _read(out + "_mins.bedgraph") == _read(o + "_m.bg")


def test_error_behaviour_matches_reference(tmp_path):
    from slamdunk_b200.dunks import tcounter
    from slamdunk_b200.io import bam as bamio
    c = goldens.CountCase("cfg1")
    text, refs, recs = bamio.read_bam(c.bam)
    # wrong @PG version -> RuntimeError (tcounter.py:156-157)
    bad = str(tmp_path / "v2.bam")
    bamio.write_bam(bad, text.replace("VN:3", "VN:2"), refs, recs)
    with pytest.raises(RuntimeError, match="Wrong filtered BAM file version detected \\(2\\)"):
        tcounter.computeTconversions(c.ref, c.bed, None, bad, 100, 27, str(tmp_path / "a.tsv"), str(tmp_path / "a.p"),
                                     str(tmp_path / "a.m"), 1, io.StringIO())
    # no @RG -> RuntimeError (misc.py:236)
    norg = str(tmp_path / "norg.bam")
    bamio.write_bam(norg, "\n".join(l for l in text.split("\n") if not l.startswith("@RG")), refs, recs)
    with pytest.raises(RuntimeError, match="RG is missing"):
        tcounter.computeTconversions(c.ref, c.bed, None, norg, 100, 27, str(tmp_path / "b.tsv"), str(tmp_path / "b.p"),
                                     str(tmp_path / "b.m"), 1, io.StringIO())
    # unstranded BED row -> RuntimeError after the rows before it were written (tcounter.py:169-170)
    bed = str(tmp_path / "u.bed")
    open(bed, "w").write(open(c.bed).read() + "chr5\t10\t20\tx\t0\t.\n")
    with pytest.raises(RuntimeError, match="does not contain stranded intervals"):
        tcounter.computeTconversions(c.ref, bed, None, c.bam, 100, 27, str(tmp_path / "c.tsv"), str(tmp_path / "c.p"),
                                     str(tmp_path / "c.m"), 1, io.StringIO())
    assert len(_read(str(tmp_path / "c.tsv")).strip().split("\n")) == 4
    # annotation md5 mismatch -> warning in the log, not an error (tcounter.py:159-161)
    log = io.StringIO()
    bed2 = str(tmp_path / "actb2.bed")
    open(bed2, "w").write(open(c.bed).read().replace("Actb", "ActB"))
    tcounter.computeTconversions(c.ref, bed2, None, c.bam, 100, 27, str(tmp_path / "d.tsv"), str(tmp_path / "d.p"),
                                 str(tmp_path / "d.m"), 1, log)
    assert log.getvalue().startswith("Warning: MD5 checksum of annotation")
    # missing VCF -> message, counts as without SNPs (SNPtools.py:50-51)
    tcounter.computeTconversions(c.ref, c.bed, str(tmp_path / "nope.vcf"), c.bam, 100, 27, str(tmp_path / "e.tsv"),
                                 str(tmp_path / "e.p"), str(tmp_path / "e.m"), 1, io.StringIO())
    assert "".join(l for l in open(str(tmp_path / "e.tsv")) if l[0] != "#") == _read(os.path.join(c.dir, "golden_body.tsv"))


def test_empty_inputs(tmp_path):
    """No reads at all / no UTR rows: default rows, empty bedgraphs."""
    from slamdunk_b200.dunks import tcounter
    from slamdunk_b200.io import bam as bamio
    c = goldens.CountCase("cfg1")
    text, refs, _recs = bamio.read_bam(c.bam)
    empty = str(tmp_path / "empty.bam")
    bamio.write_bam(empty, text, refs, [])
    o = str(tmp_path / "o")
    tcounter.computeTconversions(c.ref, c.bed, None, empty, 100, 27, o + ".tsv", o + ".p", o + ".m", 1, io.StringIO())
    assert _read(o + ".tsv").split("\n")[3] == "chr5\t120498\t122492\tActb\t1994\t+\t0\t0\t445\t0\t0\t0\t0\t0\t-1.0\t-1.0"
    assert _read(o + ".p") == "" and _read(o + ".m") == ""
    nobed = str(tmp_path / "none.bed")
    open(nobed, "w").write("")
    tcounter.computeTconversions(c.ref, nobed, None, c.bam, 100, 27, o + "2.tsv", o + "2.p", o + "2.m", 1, io.StringIO())
    assert len(_read(o + "2.tsv").strip().split("\n")) == 3


@pytest.mark.parametrize("mode", ["verify", "mp"])
def test_records_without_seq_are_counted_but_convert_nothing(tmp_path, mode):
    """Secondary alignments often carry SEQ '*': they still count as reads and coverage of their UTRs; their MP entries
    find no base quality (read position beyond the empty quality array -> 0, SlamSeqFile.py:326-329)."""
    from slamdunk_b200.io import bam as bamio
    c = simdata.make_case(str(tmp_path), seed=91, n_reads=1200, n_utr=30, snp_density=0.0)
    text, refs, recs = bamio.read_bam(c.bam)
    n = 0
    for i, r in enumerate(recs):
        if i % 7 == 3 or 320 <= i < 360:        # scattered, and one stretch covering a whole tile
            r.seq, r.qual = "", None
            r.flag |= 0x100
            n += 1
    assert n > 150
    bam2 = str(tmp_path / "noseq.bam")
    bamio.write_bam(bam2, text, refs, recs)
    log = io.StringIO()
    o = str(tmp_path / "orc")
    oracle.count_files(c.ref, c.bed, None, bam2, 100, 27, o + ".tsv", o + "_p.bg", o + "_m.bg", 1, log)

    class Case:
        ref, bed, vcf, bam = c.ref, c.bed, None, bam2
        max_read_length, min_qual, conv_threshold = 100, 27, 1
    out, glog, tm = _count(Case, tmp_path, mode, "gpu")
    assert _read(out + ".tsv") == _read(o + ".tsv")
    assert _read(out + "_plus.bedgraph") == _read(o + "_p.bg")
    assert _read(out + "_mins.bedgraph") == _read(o + "_m.bg")


@pytest.mark.parametrize("name", ["cfg1", "case_a", "case_b", "case_d", "case_e"])
@pytest.mark.parametrize("mode", [0, 2])
def test_decoder_wire_counts_like_the_columnar_batch(name, mode):
    """The product path sends the decoder's packed form (SD_DECODE_WIRE -> sd_count_wire); the same file through the columnar host
    batch (sd_count_host) must give identical outputs -- and so must rank-style slices of the wire (parallel.sub_wire) added up."""
    import torch
    from slamdunk_b200 import _lib as L, parallel
    from slamdunk_b200.engine import CountEngine, HostBatch, HostGenome, make_params
    from slamdunk_b200.utils.bed import BedIterator
    c = goldens.CountCase(name)
    genome = HostGenome(c.ref)
    hb = HostBatch(c.bam, genome.names, want_mp=True, group=True, seq2=True, wire=True)
    assert hb.wire is not None
    utrs = list(BedIterator(c.bed))
    eng = CountEngine(0)
    eng.set_reference(genome.lo, genome.hi, genome.nmask, genome.chrom_off, genome.chrom_len)
    eng.set_utrs(np.array([genome.index.get(u.chromosome, L.SD_REF_NONE) for u in utrs], dtype=np.uint32), np.array([u.start for u in utrs]),
                 np.array([u.stop for u in utrs]), np.array([1 if u.strand == "-" else 0 for u in utrs], dtype=np.uint8), None)
    params = make_params(27, 1, mode, filter_on=True, filter_mq=0, filter_min_identity=0.5)
    eng.count_host(hb.batch, params)
    eng.finalize()
    want = eng.results()
    eng.reset_outputs()
    eng.count_wire(hb.wire, params)
    eng.finalize()
    got = eng.results()
    for k in ("counters", "stats", "pos_cov", "pos_tc"):
        assert np.array_equal(want[k], got[k]), k
    assert want["stats"][L.SD_S_COUNTED] > 0
    nt = int(hb.batch.n_tiles)
    if nt >= 2:
        eng.reset_outputs()
        for r in range(3):
            t0, t1 = parallel.shard_range(nt, r, 3)
            if t1 > t0:
                eng.count_wire(parallel.sub_wire(hb.wire, t0, t1), params)
        eng.finalize()
        got = eng.results()
        for k in ("counters", "stats", "pos_cov", "pos_tc"):
            assert np.array_equal(want[k], got[k]), "slices " + k
    eng.close()
    hb.close()
    genome.close()
