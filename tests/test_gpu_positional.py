"""`alleyoop positional-tracks` on the GPU (tcounter.genomewideConversionRates) against the reference's outputs
(tests/golden/<case>/positional) and the oracle."""
import io
import os

import pytest

import goldens
import simdata
from oracle import oracle

pytestmark = pytest.mark.gpu


def _run(ref, vcf, bam, q, thr, cutoff, prefix, mode="verify", capsys=None):
    from slamdunk_b200.dunks import tcounter
    log = io.StringIO()
    tcounter.OPTIONS["mismatch_source"] = mode
    try:
        tcounter.genomewideConversionRates(ref, vcf, bam, q, prefix, thr, cutoff, log)
    finally:
        tcounter.OPTIONS["mismatch_source"] = "verify"
    return log.getvalue()


@pytest.mark.parametrize("mode", ["verify", "mp", "cigar"])
@pytest.mark.parametrize("name", goldens.POSITIONAL_CASES)
def test_positional_tracks_match_reference_golden(name, mode, tmp_path, capsys):
    c = goldens.PositionalCase(name)
    prefix = str(tmp_path / c.prefix_name)
    _run(c.ref, c.vcf, c.bam, c.min_qual, c.conv_threshold, c.coverage_cutoff, prefix, mode)
    assert capsys.readouterr().out == c.stdout
    for suffix, _n, _d in oracle.TRACKS:
        assert open(prefix + suffix).read() == open(c.expected(suffix)).read(), suffix


@pytest.mark.parametrize("seed,q,thr,cutoff,read_len", [(71, 27, 1, 1, (40, 100)), (72, 15, 2, 6, (100, 100)), (73, 27, 1, 2, (200, 300))])
def test_positional_tracks_match_oracle_on_random_data(seed, q, thr, cutoff, read_len, tmp_path):
    """Reads reaching both chromosome ends, chromosomes without reads, a BAM chromosome missing from the FASTA order."""
    c = simdata.make_case(str(tmp_path), seed=seed, n_reads=6000, n_utr=40, complexity=1.5, snp_density=0.01, read_len=read_len,
                          chroms=(("chr10", 30000), ("chr2", 20000), ("chr1", 9000), ("chrM", 1200), ("chrUn_1", 700)))
    (tmp_path / "g").mkdir()
    (tmp_path / "o").mkdir()
    name = "s_slamdunk_mapped_filtered_positional_rates"
    _run(c.ref, c.vcf, c.bam, q, thr, cutoff, str(tmp_path / "g" / name))
    oracle.positional_files(c.ref, c.vcf, c.bam, q, str(tmp_path / "o" / name), thr, cutoff, io.StringIO(), out=io.StringIO())
    for suffix, _n, _d in oracle.TRACKS:
        assert open(str(tmp_path / "g" / name) + suffix).read() == open(str(tmp_path / "o" / name) + suffix).read(), suffix


def test_alleyoop_cli_writes_the_reference_file_names(tmp_path):
    from slamdunk_b200 import alleyoop
    c = goldens.PositionalCase("case_a")
    out = tmp_path / "tracks"
    alleyoop.run(["positional-tracks", "-o", str(out), "-r", c.ref, "-v", c.vcf, "-a", str(c.coverage_cutoff), c.bam])
    made = sorted(os.listdir(out))
    assert made == sorted(["case_a_positional_rates" + s for s, _n, _d in oracle.TRACKS] + ["case_a_positional_rates.log"])
    got = open(str(out / ("case_a_positional_rates" + oracle.TRACKS[2][0]))).read().split("\n", 1)[1]
    assert got == open(c.expected(oracle.TRACKS[2][0])).read().split("\n", 1)[1]


def _rec_key(r):
    return [r.qname, r.flag, r.ref_id, r.pos]


def _full(r):
    return (r.qname, r.flag, r.ref_id, r.pos, r.mapq, tuple(r.cigar), r.seq, None if r.qual is None else tuple(r.qual),
            tuple((t, ty, tuple(v) if isinstance(v, list) else v) for t, ty, v in r.tags))


@pytest.mark.parametrize("mode", ["verify", "mp"])
@pytest.mark.parametrize("name", goldens.READSEP_CASES)
def test_read_separator_matches_reference_golden(name, mode, tmp_path):
    """alleyoop read-separator (tcounter.genomewideReadSeparation): separation is by query NAME, input order kept."""
    from slamdunk_b200.dunks import tcounter
    from slamdunk_b200.io import bam as bamio
    c = goldens.ReadSepCase(name)
    prefix = str(tmp_path / "x")
    tcounter.OPTIONS["mismatch_source"] = mode
    try:
        tcounter.genomewideReadSeparation(c.ref, c.vcf, c.bam, c.min_qual, prefix, c.conv_threshold, io.StringIO())
    finally:
        tcounter.OPTIONS["mismatch_source"] = "verify"
    text, refs, src = bamio.read_bam(c.bam)
    by_key = {}
    for r in src:
        by_key.setdefault(tuple(_rec_key(r)), []).append(_full(r))
    for key, f in (("tc", "x_TCReads.bam"), ("background", "x_backgroundReads.bam")):
        t2, r2, recs = bamio.read_bam(str(tmp_path / f))
        assert t2 == text and r2 == refs
        assert [_rec_key(r) for r in recs] == c.expected[key]
        for r in recs:                                   # records are copied whole
            assert _full(r) in by_key[tuple(_rec_key(r))]


@pytest.mark.parametrize("seed,q,thr", [(81, 27, 1), (82, 27, 3), (83, 10, 0)])
def test_read_separator_matches_oracle_on_random_data(seed, q, thr, tmp_path):
    from slamdunk_b200.dunks import tcounter
    from slamdunk_b200.io import bam as bamio
    c = simdata.make_case(str(tmp_path), seed=seed, n_reads=5000, n_utr=30, complexity=1.5, snp_density=0.02, p_conv=0.04,
                          chroms=(("chr10", 30000), ("chr2", 20000), ("chr1", 9000), ("chrM", 1200)))
    tcounter.genomewideReadSeparation(c.ref, c.vcf, c.bam, q, str(tmp_path / "g"), thr, io.StringIO())
    recs, tci, bgi = oracle.read_separation(c.ref, c.vcf, c.bam, q, thr)
    assert [_rec_key(r) for r in bamio.read_bam(str(tmp_path / "g_TCReads.bam"))[2]] == [_rec_key(recs[i]) for i in tci]
    assert [_rec_key(r) for r in bamio.read_bam(str(tmp_path / "g_backgroundReads.bam"))[2]] == [_rec_key(recs[i]) for i in bgi]
    assert len(tci) > 0 and (thr == 0 or len(bgi) > 0)


def test_alleyoop_read_separator_cli(tmp_path):
    from slamdunk_b200 import alleyoop
    c = goldens.ReadSepCase("case_a")
    out = tmp_path / "sep"
    alleyoop.run(["read-separator", "-o", str(out), "-r", c.ref, "-v", c.vcf, c.bam])
    assert sorted(os.listdir(out)) == ["case_a_multi_TCReads.bam", "case_a_multi_TCReads.bam.bai", "case_a_multi_backgroundReads.bam",
                                       "case_a_multi_backgroundReads.bam.bai", "case_a_multi_read_separator.log"]
