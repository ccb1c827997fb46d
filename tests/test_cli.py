"""The command line keeps the reference's argument surface (slamdunk/slamdunk.py:345-430)."""
import io
import os

import pytest

import goldens


def _ns(argv):
    from slamdunk_b200.cli import build_parser
    return build_parser().parse_args(argv)


def test_count_defaults_match_reference():
    a = _ns(["count", "-r", "ref.fa", "-b", "utr.bed", "-o", "out", "a.bam", "b.bam"])
    assert (a.command, a.bam, a.ref, a.bed, a.outputDir) == ("count", ["a.bam", "b.bam"], "ref.fa", "utr.bed", "out")
    assert (a.conversionThreshold, a.minQual, a.threads, a.maxLength) == (1, 27, 1, None)
    assert "snpDir" not in a and "vcfFile" not in a           # SUPPRESS defaults, tested with `in` (slamdunk.py:506-513)
    b = _ns(["count", "-r", "r", "-b", "b", "-o", "o", "-s", "snps", "-v", "x.vcf", "-c", "2", "-l", "120", "-q", "20", "-t", "4", "x.bam"])
    assert (b.snpDir, b.vcfFile, b.conversionThreshold, b.maxLength, b.minQual, b.threads) == ("snps", "x.vcf", 2, 120, 20, 4)


def test_filter_and_all_defaults_match_reference():
    f = _ns(["filter", "-o", "out", "x.bam"])
    assert (f.mq, f.identity, f.nm, f.bed, f.threads) == (2, 0.95, -1, None, 1)
    a = _ns(["all", "-r", "ref.fa", "-b", "utr.bed", "-o", "out", "reads.fq"])
    assert (a.trim5, a.maxPolyA, a.topn, a.mq, a.identity, a.nm, a.cov, a.var) == (12, 4, 1, 2, 0.95, -1, 10, 0.8)
    assert (a.conversionThreshold, a.minQual, a.sampleType, a.sampleTime, a.sampleIndex) == (1, 27, "pulse", 0, -1)
    assert a.multimap is False and a.filterbed is None and a.skipSAM is False and "vcfFile" not in a
    t = _ns(["all", "-r", "r", "-b", "b", "-o", "o", "-rl", "100", "-mbq", "27", "-5", "0", "reads.fq"])   # slamdunk/test/test_sample.sh
    assert (t.maxLength, t.minQual, t.trim5) == (100, 27, 0)


def test_estimate_max_read_length(tmp_path):
    """utils/misc.py:91-107"""
    from slamdunk_b200.io import bam as bamio
    from slamdunk_b200.cli import estimateMaxReadLength
    recs = [bamio.BamRecord("r%d" % i, 0, 0, 10 + i, 60, "%dM" % (50 + i % 5), "A" * (50 + i % 5), [30] * (50 + i % 5), [("XA", "i", 2)])
            for i in range(1500)]
    p = str(tmp_path / "a.bam")
    bamio.write_bam(p, "@HD\tVN:1.0\n@SQ\tSN:c\tLN:10000\n", [("c", 10000)], recs)
    assert len(bamio.head_records(p, 1000)) == 1000
    assert estimateMaxReadLength(p) == 54 + 2 + 10
    recs[3] = bamio.BamRecord("long", 0, 0, 13, 60, "90M", "A" * 90, [30] * 90, [("XA", "i", 0)])
    bamio.write_bam(p, "@HD\tVN:1.0\n@SQ\tSN:c\tLN:10000\n", [("c", 10000)], recs)
    assert estimateMaxReadLength(p) == -1


@pytest.mark.gpu
def test_count_cli_end_to_end(tmp_path):
    """`slamdunk count` file naming (slamdunk.py:180-183) and content on the reference's own sample."""
    from slamdunk_b200.cli import run
    c = goldens.CountCase("cfg1")
    out = str(tmp_path / "count")
    run(["count", "-r", c.ref, "-b", c.bed, "-o", out, "-l", "100", "-q", "27", c.bam])
    base = os.path.join(out, "reads_slamdunk_mapped_filtered")
    body = "".join(l for l in open(base + "_tcount.tsv") if not l.startswith("#"))
    assert body == open(os.path.join(c.dir, "golden_body.tsv")).read()
    assert open(base + "_tcount_plus.bedgraph").read() == open(c.plus).read()
    assert open(base + "_tcount_mins.bedgraph").read() == ""
    assert open(base + "_tcount.log").read().startswith("Using 100 as maximum read length.")
