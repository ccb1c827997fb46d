"""The command line keeps the reference's argument surface (slamdunk/slamdunk.py:345-430)."""
import io
import os

import pytest

import goldens

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


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


def test_collapse_matches_reference(tmp_path):
    """alleyoop collapse (tcounter.collapse, tcounter.py:39-111): rows of one gene name summed; against the reference when present,
    against a hand-made expectation otherwise."""
    import io
    import refrun
    from slamdunk_b200.dunks import tcounter
    hdr = "Chromosome\tStart\tEnd\tName\tLength\tStrand\tConversionRate\tReadsCPM\tTcontent\tCoverageOnTs\tConversionsOnTs\tReadCount\tTcReadCount\tmultimapCount\tConversionRateLower\tConversionRateUpper\n"
    rows = ["chr1\t10\t210\tgeneB\t200\t+\t0.02\t100.0\t50\t1000\t20\t30\t12\t1\t-1.0\t-1.0\n",
            "chr1\t900\t1000\tgeneA\t100\t-\t0\t0\t20\t0\t0\t0\t0\t0\t-1.0\t-1.0\n",
            "chr2\t10\t310\tgeneB\t300\t+\t0.5\t7.5\t70\t33\t7\t9\t3\t2\t-1.0\t-1.0\n",
            "chr2\tbroken\n"]
    src = tmp_path / "x_tcount.tsv"
    src.write_text("#slamdunk v0.4.3\t3\n#annotation:\ta.bed\tabc\n" + hdr + "".join(rows))
    log = io.StringIO()
    tcounter.collapse(str(src), str(tmp_path / "mine.csv"), log)
    got = (tmp_path / "mine.csv").read_text()
    assert "unexpected number of fields (2)" in log.getvalue()
    lines = got.split("\n")
    assert lines[0].split("\t")[:4] == ["gene_name", "length", "readsCPM", "conversionRate"]
    assert lines[1] == "geneA\t100\t0.0\t0\t20\t0\t0\t0\t0\t0"
    assert lines[2] == "geneB\t500\t1000000.0\t" + repr(27.0 / 1033) + "\t120\t1033\t27\t39\t15\t3"
    if refrun.available():
        ref_tc, _f = refrun.reference_modules()
        rlog = io.StringIO()
        ref_tc.collapse(str(src), str(tmp_path / "ref.csv"), rlog)
        assert got == (tmp_path / "ref.csv").read_text() and log.getvalue() == rlog.getvalue()
    from slamdunk_b200 import alleyoop
    alleyoop.run(["collapse", "-o", str(tmp_path / "c"), str(src)])
    assert (tmp_path / "c" / "x_tcount_collapsed.csv").read_text() == got


def test_summary_and_merge(tmp_path):
    """alleyoop summary (stats.readSummary, stats.py:531-589) and merge (plot/merge_rate_files.R): host-only table tools on
    the BAM header's DS field and tcount files.  summary is compared with the unmodified reference where it is present."""
    import io
    import refrun
    from slamdunk_b200 import alleyoop
    from slamdunk_b200.dunks import tables
    bam = os.path.join(GOLD, "cfg1", "reads_slamdunk_mapped_filtered.bam")
    tdir = tmp_path / "count"
    tdir.mkdir()
    body = open(os.path.join(GOLD, "cfg1", "expected_tcount.tsv")).read()
    (tdir / "reads_slamdunk_mapped_filtered_tcount.tsv").write_text(body)
    out = tmp_path / "summary.tsv"
    alleyoop.run(["summary", "-o", str(out), "-t", str(tdir), bam])
    lines = out.read_text().split("\n")
    assert lines[0].startswith("# slamdunk summary v")
    assert lines[1].split("\t")[-2:] == ["Counted", "Annotation"]
    row = lines[2].split("\t")
    assert row[0] == bam and int(row[12]) == tables.sumCounts(str(tdir / "reads_slamdunk_mapped_filtered_tcount.tsv"))
    alleyoop.run(["summary", "-o", str(tmp_path / "s2.tsv"), bam])
    assert (tmp_path / "s2.tsv").read_text().split("\n")[1].split("\t")[-2:] == ["Retained", "Annotation"]
    if refrun.available():
        refrun.reference_modules()
        from slamdunk.dunks import stats as rstats
        rstats.readSummary([bam], None, str(tmp_path / "ref.tsv"), io.StringIO())
        assert (tmp_path / "ref.tsv").read_text().split("\n")[1:] == (tmp_path / "s2.tsv").read_text().split("\n")[1:]
    # merge: two samples, second with doubled counts and a lower sample ID -> its column comes first
    hdr = body.split("\n")[:3]
    rows = [ln for ln in body.split("\n")[3:] if ln]
    a = tmp_path / "a_tcount.tsv"
    b = tmp_path / "b_tcount.tsv"
    h1 = hdr[0].split("\t")
    assert len(h1) == 7, h1
    a.write_text("\n".join(["\t".join(h1[:3] + ["sampleA", "2"] + h1[5:]), hdr[1], hdr[2]] + rows) + "\n")

    def doubled(r):
        f = r.split("\t")
        for k in (7, 9, 10, 11, 12, 13):
            f[k] = str(float(f[k]) * 2 if k == 7 else int(f[k]) * 2)
        return "\t".join(f)
    b.write_text("\n".join(["\t".join(h1[:3] + ["sampleB", "1"] + h1[5:]), hdr[1], hdr[2]] + [doubled(r) for r in rows]) + "\n")
    alleyoop.run(["merge", "-o", str(tmp_path / "m.tsv"), str(a), str(b)])
    m = (tmp_path / "m.tsv").read_text().split("\n")
    assert m[2] == "#Expression:\tTcReadCount / ReadCount"
    assert m[3].split("\t")[6:] == ["avgReadsCPM", "avgMultimapper", "avgTcontent", "avgCoverageOnTs", "sampleB", "sampleA"]
    f0, r0 = m[4].split("\t"), rows[0].split("\t")
    assert f0[:6] == r0[:6]
    assert abs(float(f0[9]) - 1.5 * int(r0[9])) < 1e-9
    want = int(r0[12]) / int(r0[11]) if int(r0[11]) else 0.0
    assert abs(float(f0[10]) - want) < 1e-12 and abs(float(f0[11]) - want) < 1e-12
    import pytest
    with pytest.raises(SystemExit) as e:
        alleyoop.run(["dedup", "-o", "x", "y.bam"])
    assert "not provided" in str(e.value)


def test_samples_are_dealt_out_over_ranks_or_tiles_are():
    """`torchrun -m slamdunk_b200 count`: at least one BAM per rank -> every rank takes its own samples (the reference's
    one-worker-per-BAM fan-out, no collective); fewer BAMs than ranks -> all ranks share every BAM (tile slices)."""
    from slamdunk_b200.cli import plan_samples
    assert plan_samples(3, 0, 1) == ("single", [0, 1, 2])
    for world in (2, 4, 8):
        seen = []
        for r in range(world):
            mode, mine = plan_samples(12, r, world)
            assert mode == "samples"
            seen += mine
        assert sorted(seen) == list(range(12))                 # every sample exactly once
        sizes = [len(plan_samples(12, r, world)[1]) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1
    for r in range(8):
        assert plan_samples(3, r, 8) == ("tiles", [0, 1, 2])
