"""The iterator-based alleyoop statistics (rates, utrrates, tcperreadpos, tcperutrpos, snpeval) and the mle per-read
lists of `count` on the GPU (slamdunk_b200/dunks/stats.py over sd_read_stats) against the reference's outputs
(tests/golden/<case>/stats) and the oracle restatement (oracle/stats_oracle.py).  Everything is integer: byte-identical
files."""
import io
import os

import pytest

import goldens
import simdata
import statsrun
from oracle import stats_oracle

pytestmark = pytest.mark.gpu


def _read(p):
    with open(p) as fh:
        return fh.read()


@pytest.mark.parametrize("name", goldens.STATS_CASES)
def test_stats_match_reference_golden(name, tmp_path):
    c = goldens.StatsCase(name)
    ref_in_bam = statsrun.fasta_in_bam(c.ref, c.bam, str(tmp_path / "in_bam.fa"))
    for tool, strict in goldens.StatsCase.RUNS:
        out = str(tmp_path / (tool + str(strict) + ".csv"))
        log = statsrun.run_gpu(tool, out, ref_in_bam if tool == "tccontext" else c.ref, c.bam, c.bed, c.vcf, c.min_qual, c.max_read_length, strict)
        assert _read(out) == _read(c.expected(tool, strict)), (tool, strict)
        assert "not written" in log and not os.path.exists(out + ".pdf")


@pytest.mark.parametrize("name", goldens.STATS_CASES)
def test_mle_perread_lists_match_reference_golden(name, tmp_path):
    """computeTconversions(mle=True): the _perread.tsv next to the tcount file (tcounter.py:137-143, 233-243, 305-309)"""
    from slamdunk_b200.dunks import tcounter
    c = goldens.StatsCase(name)
    cc = goldens.CountCase(name)
    out = str(tmp_path / "x_tcount.tsv")
    tcounter.computeTconversions(c.ref, c.bed, c.vcf, c.bam, c.max_read_length, c.min_qual, out, str(tmp_path / "p"), str(tmp_path / "m"),
                                 c.conv_threshold, io.StringIO(), True)
    assert _read(out) == _read(cc.tsv)
    assert _read(str(tmp_path / "x_tcount_perread.tsv")) == _read(c.perread)


@pytest.mark.parametrize("seed,q,maxlen,read_len,n_reads", [(95, 27, 100, (40, 100), 6000), (96, 0, 60, (30, 80), 4000),
                                                            (97, 20, 300, (200, 300), 3000)])
def test_stats_match_oracle_on_random_data(seed, q, maxlen, read_len, n_reads, tmp_path):
    """rows without a strand, UTRs past the chromosome end / on chromosomes missing from FASTA or BAM, reads longer than
    maxReadLength, SNPs incl. key collisions, several tiles"""
    c = simdata.make_case(str(tmp_path), seed=seed, n_reads=n_reads, n_utr=40, complexity=2.0, snp_density=0.02, read_len=read_len)
    bed = str(tmp_path / "mixed.bed")
    statsrun.unstrand_some_rows(c.bed, bed)
    ref_in_bam = statsrun.fasta_in_bam(c.ref, c.bam, str(tmp_path / "in_bam.fa"))
    for tool, strict in goldens.StatsCase.RUNS:
        use_bed = c.bed if tool == "snpeval" else bed
        use_ref = ref_in_bam if tool == "tccontext" else c.ref
        g, o = str(tmp_path / (tool + str(strict) + "_g.csv")), str(tmp_path / (tool + str(strict) + "_o.csv"))
        statsrun.run_gpu(tool, g, use_ref, c.bam, use_bed, c.vcf, q, maxlen, strict)
        statsrun.run_oracle(tool, o, use_ref, c.bam, use_bed, c.vcf, q, maxlen, strict)
        assert _read(g) == _read(o), (tool, strict)
    with pytest.raises(ValueError, match="invalid contig"):      # a FASTA chromosome the BAM does not know (stats.py:259)
        statsrun.run_gpu("tccontext", str(tmp_path / "x.csv"), c.ref, c.bam, None, None, q, maxlen, False)


@pytest.mark.parametrize("thr", [1, 2])
def test_mle_perread_lists_match_oracle_on_random_data(thr, tmp_path):
    from slamdunk_b200.dunks import tcounter
    c = simdata.make_case(str(tmp_path), seed=98, n_reads=5000, n_utr=40, complexity=2.0, snp_density=0.02, p_conv=0.05)
    out = str(tmp_path / "x_tcount.tsv")
    tcounter.computeTconversions(c.ref, c.bed, c.vcf, c.bam, 100, 27, out, str(tmp_path / "p"), str(tmp_path / "m"), thr, io.StringIO(), True)
    lists = stats_oracle.perread_lists(c.ref, c.bed, c.vcf, c.bam, 100, 27, thr)
    cols = statsrun.perread_columns(str(tmp_path / "x_tcount_perread.tsv"))
    assert len(cols) == len(lists) > 0
    for (n_txt, k_txt), (n_list, k_list) in zip(cols, lists):
        if n_list:
            assert n_txt == ",".join(map(str, n_list)) and k_txt == ",".join(map(str, k_list))
        else:
            assert (n_txt, k_txt) == ("0", "0")


def test_snpeval_stops_at_a_row_without_strand(tmp_path):
    """stats.py:791-792: rows before the bad one are written, then RuntimeError"""
    c = simdata.make_case(str(tmp_path), seed=99, n_reads=800, n_utr=12)
    bed = str(tmp_path / "mixed.bed")
    statsrun.unstrand_some_rows(c.bed, bed, every=6)     # row 1 loses its strand
    out = str(tmp_path / "s.csv")
    with pytest.raises(RuntimeError, match="stranded"):
        statsrun.run_gpu("snpeval", out, c.ref, c.bam, bed, c.vcf, 27, 100, False)
    assert len(_read(out).splitlines()) == 1


def test_stats_cli(tmp_path):
    """python -m slamdunk_b200.alleyoop rates / tcperreadpos / utrrates / tcperutrpos / snpeval: file names of alleyoop.py:147-262"""
    from slamdunk_b200 import alleyoop
    c = goldens.StatsCase("case_a")
    out = str(tmp_path / "out")
    snpdir = str(tmp_path / "snp")
    os.makedirs(snpdir)
    base = os.path.splitext(os.path.basename(c.bam))[0]
    os.symlink(c.vcf, os.path.join(snpdir, base + "_snp.vcf"))
    q, L = str(c.min_qual), str(c.max_read_length)
    alleyoop.run(["rates", "-o", out, "-r", c.ref, "-mq", q, c.bam])
    alleyoop.run(["tccontext", "-o", out, "-r", statsrun.fasta_in_bam(c.ref, c.bam, str(tmp_path / "in_bam.fa")), c.bam])
    alleyoop.run(["tcperreadpos", "-o", out, "-r", c.ref, "-s", snpdir, "-l", L, "-mq", q, c.bam])
    alleyoop.run(["utrrates", "-o", out, "-r", c.ref, "-b", c.bed, "-l", L, "-mq", q, c.bam])
    alleyoop.run(["tcperutrpos", "-o", out, "-r", c.ref, "-b", c.bed, "-s", snpdir, "-l", L, "-mq", q, c.bam])
    alleyoop.run(["snpeval", "-o", out, "-r", c.ref, "-b", c.bed, "-s", snpdir, "-l", L, "-q", q, c.bam])
    for suffix, tool in (("_overallrates", "rates"), ("_tccontext", "tccontext"), ("_tcperreadpos", "tcperreadpos"), ("_mutationrates_utr", "utrrates"),
                         ("_tcperutr", "tcperutrpos"), ("_SNPeval", "snpeval")):
        assert _read(os.path.join(out, base + suffix + ".csv")) == _read(c.expected(tool)), tool


def test_device_built_mp_column_gives_the_file_route_results(tmp_path):
    """synth.to_mp_all (the complete MP column built on the device from the generator's triples, used by
    tools/stats_kernel_bench.py) against the route every other test takes: the same reads written to a BAM, decoded with
    SD_DECODE_MP_ALL and scanned -- all six output sets identical."""
    import numpy as np
    import synth_files
    from slamdunk_b200.dunks import stats
    from slamdunk_b200.engine import CountEngine
    from slamdunk_b200.synth import SynthWorkload, to_mp_all
    eng = CountEngine(0)
    wl = SynthWorkload(eng, seed=5, genome_bp=3_000_000, n_utr=600, n_frac=0.01)
    wl.install()
    p = wl.params(seed=41, first_read=0, n_reads=30_000, frac_softclip=0.1, frac_indel=0.08, frac_antisense=0.05, frac_labelled=0.8)
    db, triples = wl.generate(p, want_triples=True, coordinate_sorted=True, group=False)
    want = ("rates", "context", "readpos", "utr_rates", "utrpos", "utr_snp")
    dev_res, ms = stats.read_stats_on_device(eng, to_mp_all(db, triples, eng), len(wl.ann), 27, want, max_read_length=100)
    assert ms > 0
    ref, bed, bam, n = synth_files.write_workload_files(str(tmp_path), wl, db, triples)
    eng.close()
    rs = stats.ReadStats(ref, bam, bed=bed)
    try:
        file_res = rs.run(27, want=want, max_read_length=100)
    finally:
        rs.close()
    for k in want:
        assert np.array_equal(dev_res[k], file_res[k]), k
    assert dev_res["rates"].sum() > 0 and dev_res["utr_snp"].reshape(-1, 3)[:, 0].sum() > 20_000


@pytest.mark.parametrize("q,maxlen", [(0, 60), (27, 60), (36, 35)])
def test_stats_match_oracle_on_edge_cases(q, maxlen, tmp_path):
    """statsrun.make_edge_case (pinned against the live reference in tests/test_oracle.py): MP entries past the read end
    (quality 0; negative mirrored positions on reverse reads), N bases, soft / hard clips, entries outside the UTR, a read
    over the chromosome end, SNPs under conversions, reads without MP tag, all through the GPU tools and the mle lists."""
    from slamdunk_b200.dunks import tcounter
    ref, bed, vcf, bam = statsrun.make_edge_case(str(tmp_path))
    for tool, strict in goldens.StatsCase.RUNS:
        g, o = str(tmp_path / (tool + str(strict) + "_g.csv")), str(tmp_path / (tool + str(strict) + "_o.csv"))
        statsrun.run_gpu(tool, g, ref, bam, bed, vcf, q, maxlen, strict)
        statsrun.run_oracle(tool, o, ref, bam, bed, vcf, q, maxlen, strict)
        assert _read(g) == _read(o), (tool, strict)
    out = str(tmp_path / "x_tcount.tsv")
    tcounter.OPTIONS["mismatch_source"] = "mp"        # the hand-made MP tags do not describe SEQ / CIGAR: count as slamdunk does
    try:
        tcounter.computeTconversions(ref, bed, vcf, bam, maxlen, q, out, str(tmp_path / "p"), str(tmp_path / "m"), 1, io.StringIO(), True)
    finally:
        tcounter.OPTIONS["mismatch_source"] = "verify"
    lists = stats_oracle.perread_lists(ref, bed, vcf, bam, maxlen, q, 1)
    cols = statsrun.perread_columns(str(tmp_path / "x_tcount_perread.tsv"))
    assert len(cols) == len(lists)
    for (n_txt, k_txt), (n_list, k_list) in zip(cols, lists):
        if n_list:
            assert n_txt == ",".join(map(str, n_list)) and k_txt == ",".join(map(str, k_list))
    from oracle import oracle
    oo = str(tmp_path / "o_tcount.tsv")
    oracle.count_files(ref, bed, vcf, bam, maxlen, q, oo, str(tmp_path / "op"), str(tmp_path / "om"), 1, io.StringIO())
    assert _read(out) == _read(oo) and _read(str(tmp_path / "p")) == _read(str(tmp_path / "op")) and _read(str(tmp_path / "m")) == _read(str(tmp_path / "om"))
