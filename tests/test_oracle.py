"""The CPU oracle (oracle/) against the reference's golden vector, the committed
reference-through-shim vectors, and -- where /root/reference is present -- the reference's own
code run live on fresh random data.  CPU only."""
import ast
import io
import os

import pytest

import goldens
import refrun
import simdata
from oracle import oracle
from slamdunk_b200.io import bam as bamio


def _read(p):
    with open(p) as fh:
        return fh.read()


@pytest.mark.parametrize("name", goldens.COUNT_CASES)
def test_oracle_count_matches_golden(name, tmp_path):
    c = goldens.CountCase(name)
    out = str(tmp_path / "o")
    log = io.StringIO()
    oracle.count_files(c.ref, c.bed, c.vcf, c.bam, c.max_read_length, c.min_qual, out + ".tsv",
                       out + "_plus.bedgraph", out + "_mins.bedgraph", c.conv_threshold, log, threads=2)
    assert _read(out + ".tsv") == _read(c.tsv)
    assert _read(out + "_plus.bedgraph") == _read(c.plus)
    assert _read(out + "_mins.bedgraph") == _read(c.minus)


def test_cfg1_is_the_reference_golden_row():
    """slamdunk/test/data/reads_slamdunk_mapped_filtered_tcount.tsv:2"""
    c = goldens.CountCase("cfg1")
    body = "".join(l for l in open(c.tsv) if not l.startswith("#"))
    assert body == _read(os.path.join(c.dir, "golden_body.tsv"))
    assert body.split("\n")[1] == ("chr5\t120498\t122492\tActb\t1994\t+\t0.022222222222222223\t666666.6666666666\t"
                                   "445\t90\t2\t8\t4\t0\t-1.0\t-1.0")


def _check_filter(records, bam_names, bed_rows, mq, mi, nm, expected_ds, expected_records):
    keep, rd, stats = oracle.filter_records(records, bam_names, bed_rows, mq, mi, nm)
    ds = ast.literal_eval(expected_ds)
    assert stats["mapped"] == ds["mapped"]
    assert stats["mapped"] + stats["unmapped"] == ds["sequenced"]
    for k in ("filtered", "mqfiltered", "idfiltered", "nmfiltered", "multimapper"):
        assert stats[k] == ds[k], k
    written = []
    for i, r in enumerate(records):
        if keep[i] == 1:
            written.append(r.copy())
        elif keep[i] == 2:
            w = r.copy()
            w.flag &= ~0x900
            w.set_tag("RD", rd[i], "Z")
            written.append(w)
    got = [[r.qname, r.flag, r.ref_id, r.pos, r.get_tag("RD") if r.has_tag("RD") else None]
           for r in bamio.sort_records(written)]
    assert got == expected_records


@pytest.mark.parametrize("name", goldens.FILTER_CASES)
def test_oracle_filter_matches_golden(name):
    c = goldens.FilterCase(name)
    _text, refs, records = bamio.read_bam(c.bam)
    bed_rows = oracle.read_bed(c.bed) if c.bed else None
    _check_filter(records, [n for n, _ in refs], bed_rows, c.mq, c.min_identity, c.nm,
                  c.expected["DS"], c.expected["records"])


needs_ref = pytest.mark.skipif(not refrun.available(), reason="/root/reference not present")


@needs_ref
@pytest.mark.parametrize("seed,thr,minq,vcf", [(101, 1, 27, True), (102, 3, 13, True), (103, 1, 30, False)])
def test_oracle_vs_live_reference_count(seed, thr, minq, vcf, tmp_path):
    c = simdata.make_case(str(tmp_path), seed=seed, n_reads=1500, complexity=1.5, snp_density=0.01)
    rtsv, rplus, rminus, rlog = refrun.run_count(c.ref, c.bed, c.vcf if vcf else None, c.bam,
                                                 str(tmp_path / "ref"), 100, minq, thr)
    out = str(tmp_path / "orc")
    log = io.StringIO()
    oracle.count_files(c.ref, c.bed, c.vcf if vcf else None, c.bam, 100, minq, out + ".tsv", out + "_p.bg",
                       out + "_m.bg", thr, log)
    assert _read(out + ".tsv") == _read(rtsv)
    assert _read(out + "_p.bg") == _read(rplus)
    assert _read(out + "_m.bg") == _read(rminus)
    assert log.getvalue() == rlog


@needs_ref
@pytest.mark.parametrize("seed,bed,mq,mi,nm", [(201, False, 2, 0.95, -1), (202, True, 2, 0.95, -1), (203, True, 0, 0.85, 3)])
def test_oracle_vs_live_reference_filter(seed, bed, mq, mi, nm, tmp_path):
    c = simdata.make_filter_case(str(tmp_path), seed=seed, n_reads=800)
    out_bam = str(tmp_path / "f.bam")
    refrun.run_filter(c.bam, out_bam, c.bed if bed else None, mq, mi, nm)
    text, refs, recs = bamio.read_bam(out_bam)
    hdr = bamio.parse_header_text(text)
    expected = [[r.qname, r.flag, r.ref_id, r.pos, r.get_tag("RD") if r.has_tag("RD") else None] for r in recs]
    _t, refs0, records = bamio.read_bam(c.bam)
    _check_filter(records, [n for n, _ in refs0], oracle.read_bed(c.bed) if bed else None, mq, mi, nm,
                  hdr["RG"][0]["DS"], expected)


def test_bam_codec_roundtrip(tmp_path):
    c = simdata.make_case(str(tmp_path), seed=5, n_reads=300)
    text, refs, recs = bamio.read_bam(c.bam)
    assert refs == c.bam_refs and len(recs) == len(c.records)
    for a, b in zip(recs, c.records):
        assert (a.qname, a.flag, a.ref_id, a.pos, a.mapq, a.cigar, a.seq, a.qual) == \
               (b.qname, b.flag, b.ref_id, b.pos, b.mapq, b.cigar, b.seq, b.qual)
        assert [(t, v if ty != "f" else round(v, 5)) for t, ty, v in a.tags] == \
               [(t, v if ty != "f" else round(v, 5)) for t, ty, v in b.tags]


@pytest.mark.parametrize("name", goldens.POSITIONAL_CASES)
def test_oracle_positional_tracks_match_golden(name, tmp_path):
    """genomewideConversionRates (alleyoop positional-tracks): the oracle's eight run-length bedGraphs equal what the
    reference wrote for the same inputs (tests/golden/<case>/positional)."""
    c = goldens.PositionalCase(name)
    out = io.StringIO()
    oracle.positional_files(c.ref, c.vcf, c.bam, c.min_qual, str(tmp_path / c.prefix_name), c.conv_threshold,
                            c.coverage_cutoff, io.StringIO(), out=out)
    assert out.getvalue() == c.stdout
    for suffix, _n, _d in oracle.TRACKS:
        assert open(str(tmp_path / c.prefix_name) + suffix).read() == open(c.expected(suffix)).read(), suffix


@pytest.mark.skipif(not refrun.available(), reason="reference tree not present")
@pytest.mark.parametrize("seed,q,thr,cutoff", [(61, 27, 1, 1), (62, 10, 3, 4)])
def test_oracle_vs_live_reference_positional(seed, q, thr, cutoff, tmp_path):
    c = simdata.make_case(str(tmp_path), seed=seed, n_reads=700, n_utr=10, complexity=2.0, snp_density=0.02,
                          chroms=(("chr10", 5000), ("chr2", 4000), ("chr1", 2500), ("chrM", 600)))
    (tmp_path / "r").mkdir()
    (tmp_path / "o").mkdir()
    refrun.run_positional(c.ref, c.vcf, c.bam, str(tmp_path / "r" / "s_slamdunk_mapped_x"), q, thr, cutoff)
    oracle.positional_files(c.ref, c.vcf, c.bam, q, str(tmp_path / "o" / "s_slamdunk_mapped_x"), thr, cutoff, io.StringIO(), out=io.StringIO())
    for suffix, _n, _d in oracle.TRACKS:
        assert open(str(tmp_path / "o" / "s_slamdunk_mapped_x") + suffix).read() == open(str(tmp_path / "r" / "s_slamdunk_mapped_x") + suffix).read(), suffix


@pytest.mark.parametrize("name", goldens.READSEP_CASES)
def test_oracle_read_separation_matches_golden(name):
    """genomewideReadSeparation (alleyoop read-separator): the oracle's two record lists equal the reference's BAMs."""
    c = goldens.ReadSepCase(name)
    recs, tci, bgi = oracle.read_separation(c.ref, c.vcf, c.bam, c.min_qual, c.conv_threshold)
    key = lambda r: [r.qname, r.flag, r.ref_id, r.pos]   # noqa: E731
    assert [key(recs[i]) for i in tci] == c.expected["tc"]
    assert [key(recs[i]) for i in bgi] == c.expected["background"]
    names = [r.qname for r in recs]
    assert len(set(names)) < len(names)        # the case really has several alignments per name


# ---------------------------------------------------------------------------------------------------------------
# iterator-based alleyoop statistics + mle per-read lists (oracle/stats_oracle.py)
# ---------------------------------------------------------------------------------------------------------------
import statsrun  # noqa: E402
from oracle import stats_oracle  # noqa: E402


@pytest.mark.parametrize("name", goldens.STATS_CASES)
def test_stats_oracle_matches_reference_golden(name, tmp_path):
    c = goldens.StatsCase(name)
    ref_in_bam = statsrun.fasta_in_bam(c.ref, c.bam, str(tmp_path / "in_bam.fa"))
    for tool, strict in goldens.StatsCase.RUNS:
        out = str(tmp_path / (tool + str(strict) + ".csv"))
        statsrun.run_oracle(tool, out, ref_in_bam if tool == "tccontext" else c.ref, c.bam, c.bed, c.vcf, c.min_qual, c.max_read_length, strict)
        assert _read(out) == _read(c.expected(tool, strict)), (tool, strict)
    lists = stats_oracle.perread_lists(c.ref, c.bed, c.vcf, c.bam, c.max_read_length, c.min_qual, c.conv_threshold)
    cols = statsrun.perread_columns(c.perread)
    assert len(cols) == len(lists)
    for (n_txt, k_txt), (n_list, k_list) in zip(cols, lists):
        if n_list:
            assert n_txt == ",".join(map(str, n_list)) and k_txt == ",".join(map(str, k_list))
        else:
            assert (n_txt, k_txt) in (("0", "0"), ("", ""))     # default row / a row whose reads are all antisense


@pytest.mark.skipif(not refrun.available(), reason="reference tree not present")
@pytest.mark.parametrize("seed,q,maxlen,read_len", [(91, 27, 100, (40, 100)), (92, 0, 60, (30, 80))])
def test_stats_oracle_matches_live_reference(seed, q, maxlen, read_len, tmp_path):
    """fresh seed, rows without a strand (both read strands join them), reads longer than maxReadLength"""
    c = simdata.make_case(str(tmp_path), seed=seed, n_reads=1500, n_utr=30, complexity=2.0, snp_density=0.02, read_len=read_len)
    bed = str(tmp_path / "mixed.bed")
    statsrun.unstrand_some_rows(c.bed, bed)
    ref_in_bam = statsrun.fasta_in_bam(c.ref, c.bam, str(tmp_path / "in_bam.fa"))
    for tool, strict in goldens.StatsCase.RUNS:
        use_bed = c.bed if tool == "snpeval" else bed          # snpeval refuses rows without strand
        use_ref = ref_in_bam if tool == "tccontext" else c.ref
        r, o = str(tmp_path / "r.csv"), str(tmp_path / "o.csv")
        for f in (r, o):
            if os.path.exists(f):
                os.remove(f)
        refrun.run_stats(tool, r, use_ref, c.bam, bed=use_bed, vcf=c.vcf, min_qual=q, max_read_length=maxlen, strict=strict)
        statsrun.run_oracle(tool, o, use_ref, c.bam, use_bed, c.vcf, q, maxlen, strict)
        assert _read(r) == _read(o), (tool, strict)
    # a FASTA chromosome that the BAM does not know: tccontext dies in pysam's fetch (stats.py:259)
    for run in (lambda: refrun.run_stats("tccontext", str(tmp_path / "x.csv"), c.ref, c.bam),
                lambda: statsrun.run_oracle("tccontext", str(tmp_path / "y.csv"), c.ref, c.bam, None, None, q, maxlen, False)):
        with pytest.raises(ValueError, match="invalid contig"):
            run()


@pytest.mark.skipif(not refrun.available(), reason="reference tree not present")
@pytest.mark.parametrize("q,maxlen", [(0, 60), (27, 60), (36, 35)])
def test_stats_oracle_matches_live_reference_on_edge_cases(q, maxlen, tmp_path):
    """statsrun.make_edge_case: MP entries past the read end (quality 0, negative mirrored positions), N bases, clips,
    entries outside the UTR, a read over the chromosome end, SNPs under conversions ..."""
    ref, bed, vcf, bam = statsrun.make_edge_case(str(tmp_path))
    for tool, strict in goldens.StatsCase.RUNS:
        r, o = str(tmp_path / "r.csv"), str(tmp_path / "o.csv")
        for f in (r, o):
            if os.path.exists(f):
                os.remove(f)
        refrun.run_stats(tool, r, ref, bam, bed=bed, vcf=vcf, min_qual=q, max_read_length=maxlen, strict=strict)
        statsrun.run_oracle(tool, o, ref, bam, bed, vcf, q, maxlen, strict)
        assert _read(r) == _read(o), (tool, strict)
    perread = refrun.run_count_mle(ref, bed, vcf, bam, str(tmp_path / "m"), maxlen, q, 1)
    lists = stats_oracle.perread_lists(ref, bed, vcf, bam, maxlen, q, 1)
    cols = statsrun.perread_columns(perread)
    for (n_txt, k_txt), (n_list, k_list) in zip(cols, lists):
        if n_list:
            assert n_txt == ",".join(map(str, n_list)) and k_txt == ",".join(map(str, k_list))
