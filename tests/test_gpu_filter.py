"""`slamdunk filter` decisions on the GPU against the reference's outputs (tests/golden/filt_*) and the oracle."""
import ast
import io

import numpy as np
import pytest

import goldens
import simdata
from oracle import oracle
from slamdunk_b200.io import bam as bamio

pytestmark = pytest.mark.gpu


def _run(bam, bed, mq, mi, nm, tmp_path, tag="out"):
    from slamdunk_b200.dunks import filter as gfilter
    out = str(tmp_path / (tag + ".bam"))
    log = io.StringIO()
    gfilter.Filter(bam, out, log, bed, mq, mi, nm, False, False, True)
    text, refs, recs = bamio.read_bam(out)
    import baiquery                                   # the filtered BAM comes with an index that answers region queries
    idx, _n = baiquery.read_bai(out + ".bai")
    assert len(idx) == len(refs)
    if recs:
        mid = recs[len(recs) // 2]
        lo, hi = max(0, mid.pos - 300), mid.pos + 300
        assert baiquery.fetch(out, idx, mid.ref_id, lo, hi) == baiquery.scan(out, mid.ref_id, lo, hi)
    hdr = bamio.parse_header_text(text)
    got = [[r.qname, r.flag, r.ref_id, r.pos, r.get_tag("RD") if r.has_tag("RD") else None] for r in recs]
    return hdr, got, log.getvalue()


@pytest.mark.parametrize("name", goldens.FILTER_CASES)
def test_filter_matches_reference_golden(name, tmp_path):
    c = goldens.FilterCase(name)
    hdr, got, log = _run(c.bam, c.bed, c.mq, c.min_identity, c.nm, tmp_path)
    assert hdr["RG"][0]["DS"] == c.expected["DS"]
    assert [dict(p) for p in hdr["PG"]] == c.expected["PG"]
    assert got == c.expected["records"]
    import re
    norm = lambda t: re.sub(r"on \S*%s\.bam\." % name, "on <bam>.", t)     # the log names the input path
    assert norm(log) == norm(c.expected["log"])


@pytest.mark.parametrize("seed,bed,mq,mi,nm,topn", [(401, True, 2, 0.95, -1, 3), (402, True, 2, 0.9, 5, 8), (403, False, 3, 0.97, 2, 3),
                                                    (404, True, 0, 0.0, -1, 12)])
def test_filter_matches_oracle_on_random_data(seed, bed, mq, mi, nm, topn, tmp_path):
    c = simdata.make_filter_case(str(tmp_path), seed=seed, n_reads=6000, topn=topn)
    _t, refs, records = bamio.read_bam(c.bam)
    bam_names = [n for n, _ in refs]
    keep, rd, stats = oracle.filter_records(records, bam_names, oracle.read_bed(c.bed) if bed else None, mq, mi, nm)
    written = []
    for i, r in enumerate(records):
        if keep[i] == 1:
            written.append(r.copy())
        elif keep[i] == 2:
            w = r.copy()
            w.flag &= ~0x900
            w.set_tag("RD", rd[i], "Z")
            written.append(w)
    want = [[r.qname, r.flag, r.ref_id, r.pos, r.get_tag("RD") if r.has_tag("RD") else None] for r in bamio.sort_records(written)]
    hdr, got, _log = _run(c.bam, c.bed if bed else None, mq, mi, nm, tmp_path)
    ds = ast.literal_eval(hdr["RG"][0]["DS"])
    for k in ("mapped", "filtered", "mqfiltered", "idfiltered", "nmfiltered", "multimapper"):
        assert ds[k] == stats[k], k
    assert ds["sequenced"] == stats["mapped"] + stats["unmapped"]
    assert got == want


def test_filtered_bam_feeds_count(tmp_path):
    """filter -> count chain: the FilteredReads written into @RG DS is the CPM denominator (tcounter.py:133,297)."""
    from slamdunk_b200.dunks import tcounter
    c = simdata.make_case(str(tmp_path), seed=77, n_reads=2000, filtered_header=False)
    hdr, got, _log = _run(c.bam, None, 2, 0.95, -1, tmp_path, "filtered")
    out = str(tmp_path / "c")
    tcounter.computeTconversions(c.ref, c.bed, c.vcf, str(tmp_path / "filtered.bam"), 100, 27, out + ".tsv", out + ".p", out + ".m", 1,
                                 io.StringIO())
    o = str(tmp_path / "o")
    oracle.count_files(c.ref, c.bed, c.vcf, str(tmp_path / "filtered.bam"), 100, 27, o + ".tsv", o + ".p", o + ".m", 1, io.StringIO())
    assert open(out + ".tsv").read() == open(o + ".tsv").read()
    assert ast.literal_eval(hdr["RG"][0]["DS"])["filtered"] == len([g for g in got if not g[1] & 0x900])
