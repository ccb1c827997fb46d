#!/usr/bin/env python
"""Regenerate tests/golden/* by running the UNMODIFIED reference (/root/reference) through
the I/O shims in tests/shims.  Only runs where /root/reference exists
(the build container); the GPU box uses the committed files.

    python tests/golden/make_golden.py

cfg1   the reference's own test sample (slamdunk/test/data): 12 reads,
       1 UTR.  Expected TSV body must equal the reference's golden file -- asserted here.
       The FASTA is stored as the chr5:120000-123000 window padded with N to full length
       (only that window is touched by the 12 reads and the UTR).
case_* randomized data sets from tests/simdata.py; expected files are whatever the reference
       code writes for them.
filt_* mapper-order BAMs through the reference's filter.Filter (default and multimapper-
       retention modes); expected = DS counters + the written records.
case_*/positional   the same inputs through tcounter.genomewideConversionRates (alleyoop positional-tracks):
       the eight run-length bedGraphs (python tests/golden/make_golden.py positional regenerates only these).
case_*/stats        dunks/stats.py's iterator-based tools (alleyoop rates, utrrates, tcperreadpos, tcperutrpos, snpeval; the two
       with -m also as *_strict) and the _perread.tsv of computeTconversions(mle=True) (`... make_golden.py stats`).
case_*/readsep      tcounter.genomewideReadSeparation (alleyoop read-separator) on the case's BAM plus same-name
       secondary copies; expected = the records of the two BAMs it writes (`... make_golden.py readsep`).
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
TESTS = os.path.dirname(HERE)
sys.path.insert(0, os.path.dirname(TESTS))
sys.path.insert(0, TESTS)

import refrun          # noqa: E402
import simdata         # noqa: E402
from slamdunk_b200.io import bam as bamio   # noqa: E402

REFDATA = os.path.join(refrun.REFERENCE_ROOT, "slamdunk", "test", "data")

# name, flag, pos, CIGAR, MP  
CFG1_ALIGNMENTS = [
    ("Read1_1_0", 0, 120494, "40M", "16:2:2"), ("Read2_0_0", 0, 120497, "37M", None),
    ("Read3_1_0", 0, 120498, "1S36M", None), ("Read4_1_1", 0, 120497, "37M", "16:2:2"),
    ("Read5_2_2", 0, 120500, "2S34M", None), ("Read6_1_0", 0, 120498, "36M", "16:16:16"),
    ("Read7_0_0", 0, 120443, "56M", None), ("Read8_1_1", 0, 120443, "55M1S", None),
    ("Read9_1_0", 0, 120443, "56M", "16:55:55"), ("Read10_0_0_rev", 16, 120496, "38M", None),
    ("Read11_1_0_rev", 16, 120496, "38M", "2:15:15"), ("Read12_1_0_rev", 16, 120496, "38M", "16:27:27"),
]

CASES = {
    # name: (make_case kwargs, count kwargs)
    "case_a": (dict(seed=11, n_reads=3000), dict(max_read_length=100, min_qual=27, conv_threshold=1, vcf=True)),
    "case_b": (dict(seed=12, n_reads=2500, complexity=2.5, p_conv=0.05), dict(max_read_length=100, min_qual=20, conv_threshold=2, vcf=True)),
    "case_c": (dict(seed=13, n_reads=2500, read_len=(100, 100), snp_density=0.0), dict(max_read_length=100, min_qual=27, conv_threshold=1, vcf=False)),
    "case_d": (dict(seed=14, n_reads=2000, read_len=(150, 250), snp_density=0.02, multimap_frac=0.15,
                    chroms=(("chr1", 40000), ("chr11", 9000), ("chr12", 7000), ("chrX", 8000))),
               dict(max_read_length=250, min_qual=27, conv_threshold=1, vcf=True)),
    "case_e": (dict(seed=15, n_reads=1200, read_len=(20, 60), n_utr=15), dict(max_read_length=60, min_qual=0, conv_threshold=0, vcf=True)),
}

FILTER_CASES = {
    "filt_default": (dict(seed=21, n_reads=1500), dict(bed=False, mq=2, min_identity=0.95, nm=-1)),
    "filt_nm": (dict(seed=22, n_reads=1200), dict(bed=False, mq=0, min_identity=0.8, nm=2)),
    "filt_multimap": (dict(seed=23, n_reads=1500), dict(bed=True, mq=2, min_identity=0.95, nm=-1)),
    "filt_multimap_nm": (dict(seed=24, n_reads=1200, topn=5), dict(bed=True, mq=2, min_identity=0.9, nm=4)),
}


def make_cfg1(out):
    os.makedirs(out, exist_ok=True)
    seqs = bamio.read_fasta(os.path.join(REFDATA, "ref.fa"))
    chr5 = seqs["chr5"]
    lo, hi = 120000, 123000
    padded = "N" * lo + chr5[lo:hi] + "N" * (len(chr5) - hi)
    assert len(padded) == len(chr5) == 237442
    bamio.write_fasta(os.path.join(out, "cfg1.fa"), {"chr5": padded}, width=len(padded))
    with open(os.path.join(REFDATA, "actb.bed")) as fh:
        bed_text = fh.read()
    # same bytes (the md5 goes into the TSV header) under the name the reference uses
    with open(os.path.join(out, "actb.bed"), "w") as fh:
        fh.write(bed_text)
    reads = {}
    with open(os.path.join(REFDATA, "reads.fq")) as fh:
        lines = [l.rstrip("\n") for l in fh]
    for i in range(0, len(lines) - 3, 4):
        reads[lines[i][1:]] = (lines[i + 1], lines[i + 3])
    comp = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N"}
    recs = []
    for name, flag, pos, cig, mp in CFG1_ALIGNMENTS:
        s, q = reads[name]
        qual = [ord(c) - 33 for c in q]
        if flag & 16:
            s = "".join(comp[c] for c in reversed(s))
            qual = qual[::-1]
        recs.append(bamio.BamRecord(name, flag, 0, pos, 60, cig, s, qual, [("MP", "Z", mp)] if mp else []))
    md5 = hashlib.md5(bed_text.encode()).hexdigest()
    ds = ("{'sequenced':12,'mapped':12,'filtered':12,'mqfiltered':0,'idfiltered':0,'nmfiltered':0,'multimapper':0,"
          "'dedup':0,'snps':0,'annotation':'actb.bed','annotationmd5':'%s'}" % md5)
    hdr = ("@HD\tVN:1.0\tSO:coordinate\n@SQ\tSN:chr5\tLN:237442\n@RG\tID:0\tSM:reads:pulse:0\tDS:%s\n"
           "@PG\tID:slamdunk\tPN:slamdunk filter v0.4.3\tVN:3\n" % ds)
    bam = os.path.join(out, "reads_slamdunk_mapped_filtered.bam")
    bamio.write_bam(bam, hdr, [("chr5", 237442)], bamio.sort_records(recs))
    tsv, plus, minus, _log = refrun.run_count(os.path.join(out, "cfg1.fa"), os.path.join(out, "actb.bed"), None, bam,
                                              os.path.join(out, "expected"), 100, 27, 1)
    body = "".join(l for l in open(tsv) if not l.startswith("#"))
    gold = open(os.path.join(REFDATA, "reads_slamdunk_mapped_filtered_tcount.tsv")).read()
    assert body == gold, "reference-through-shim does not reproduce the reference's golden TSV"
    # and on the unpadded FASTA too
    tsv2, _, _, _ = refrun.run_count(os.path.join(REFDATA, "ref.fa"), os.path.join(out, "actb.bed"), None, bam,
                                     "/tmp/_cfg1_full", 100, 27, 1)
    assert open(tsv2).read() == open(tsv).read()
    with open(os.path.join(out, "golden_body.tsv"), "w") as fh:
        fh.write(gold)
    print("cfg1 ok:", body.strip().split("\n")[1])


def make_count_case(name, mk, ck):
    out = os.path.join(HERE, name)
    if os.path.isdir(out):
        shutil.rmtree(out)
    c = simdata.make_case(out, name=name, **mk)
    vcf = c.vcf if ck["vcf"] else None
    if not ck["vcf"]:
        os.remove(c.vcf)
    tsv, plus, minus, log = refrun.run_count(c.ref, c.bed, vcf, c.bam, os.path.join(out, "expected"),
                                             ck["max_read_length"], ck["min_qual"], ck["conv_threshold"])
    with open(os.path.join(out, "params.json"), "w") as fh:
        json.dump(dict(make=mk, count=ck), fh, indent=1, sort_keys=True)
    with open(os.path.join(out, "expected.log"), "w") as fh:
        fh.write(log)
    n_rows = sum(1 for l in open(tsv)) - 3
    print(name, "rows", n_rows, "plus", sum(1 for _ in open(plus)), "minus", sum(1 for _ in open(minus)))


def make_filter_case(name, mk, fk):
    out = os.path.join(HERE, name)
    if os.path.isdir(out):
        shutil.rmtree(out)
    c = simdata.make_filter_case(out, name=name, **mk)
    out_bam = os.path.join(out, "expected_filtered.bam")
    log = refrun.run_filter(c.bam, out_bam, c.bed if fk["bed"] else None, fk["mq"], fk["min_identity"], fk["nm"])
    text, refs, recs = bamio.read_bam(out_bam)
    hdr = bamio.parse_header_text(text)
    with open(os.path.join(out, "expected.json"), "w") as fh:
        json.dump(dict(filter=fk, make=mk, DS=hdr["RG"][0]["DS"], PG=hdr["PG"],
                       records=[[r.qname, r.flag, r.ref_id, r.pos, r.get_tag("RD") if r.has_tag("RD") else None]
                                for r in recs], log=log), fh, indent=0)
    os.remove(out_bam)
    print(name, "kept", len(recs), hdr["RG"][0]["DS"])


# case -> (min_qual, conversion threshold, coverage cutoff) of the positional-tracks run on the count case's inputs
POSITIONAL = {"case_a": (27, 1, 1), "case_b": (20, 2, 3), "case_d": (27, 1, 2), "case_e": (0, 0, 1)}


def make_positional(name, min_qual, conv_threshold, cutoff):
    d = os.path.join(HERE, name)
    out = os.path.join(d, "positional")
    if os.path.isdir(out):
        shutil.rmtree(out)
    os.makedirs(out)
    vcf = os.path.join(d, name + ".vcf")
    prefix = os.path.join(out, name + "_slamdunk_mapped_filtered_positional_rates")
    stdout, _log = refrun.run_positional(os.path.join(d, name + ".fa"), vcf if os.path.exists(vcf) else None,
                                         os.path.join(d, name + ".bam"), prefix, min_qual, conv_threshold, cutoff)
    with open(os.path.join(out, "params.json"), "w") as fh:
        json.dump(dict(min_qual=min_qual, conv_threshold=conv_threshold, coverage_cutoff=cutoff, stdout=stdout), fh, indent=1, sort_keys=True)
    print(name, "positional lines", sum(sum(1 for _ in open(os.path.join(out, f))) for f in os.listdir(out) if f.endswith(".bedGraph")))


# case -> (min_qual, conversion threshold) of the read-separator run; the input is the count case's BAM plus secondary
# copies of every 9th record under the same query name (the separation is by NAME, tcounter.py:508-522)
READSEP = {"case_a": (27, 1), "case_b": (20, 2), "case_e": (0, 0)}


def make_readsep(name, min_qual, conv_threshold):
    d = os.path.join(HERE, name)
    out = os.path.join(d, "readsep")
    if os.path.isdir(out):
        shutil.rmtree(out)
    os.makedirs(out)
    text, refs, recs = bamio.read_bam(os.path.join(d, name + ".bam"))
    extra = []
    for i, r in enumerate(recs):
        if i % 9 == 4:
            c = r.copy()
            c.flag |= 0x100
            c.pos = max(0, min(refs[c.ref_id][1] - 400, c.pos + 57))
            c.tags = [t for t in c.tags if t[0] != "MP"]
            extra.append(c)
    bam = os.path.join(out, name + "_multi.bam")
    bamio.write_bam(bam, text, refs, bamio.sort_records(recs + extra))
    vcf = os.path.join(d, name + ".vcf")
    refrun.run_read_separator(os.path.join(d, name + ".fa"), vcf if os.path.exists(vcf) else None, bam, os.path.join(out, "x"),
                              min_qual, conv_threshold)
    exp = {"min_qual": min_qual, "conv_threshold": conv_threshold}
    for key, f in (("tc", "x_TCReads.bam"), ("background", "x_backgroundReads.bam")):
        t2, _r, rr = bamio.read_bam(os.path.join(out, f))
        assert t2 == text
        exp[key] = [[r.qname, r.flag, r.ref_id, r.pos] for r in rr]
        os.remove(os.path.join(out, f))
    with open(os.path.join(out, "expected.json"), "w") as fh:
        json.dump(exp, fh, indent=0)
    print(name, "readsep tc", len(exp["tc"]), "background", len(exp["background"]))


# the iterator-based alleyoop statistics (dunks/stats.py) and the mle per-read lists of `count` on the count cases' inputs,
# with each case's own min_qual / max_read_length / conversion threshold (tests/goldens.py StatsCase)
STATS = ["case_a", "case_b", "case_d", "case_e"]
STATS_RUNS = [("rates", False), ("tccontext", False), ("utrrates", False), ("utrrates", True), ("tcperreadpos", False), ("tcperutrpos", False),
              ("snpeval", False), ("snpeval", True)]


def make_stats(name):
    d = os.path.join(HERE, name)
    out = os.path.join(d, "stats")
    if os.path.isdir(out):
        shutil.rmtree(out)
    os.makedirs(out)
    p = json.load(open(os.path.join(d, "params.json")))["count"]
    ref, bam, bed = os.path.join(d, name + ".fa"), os.path.join(d, name + ".bam"), os.path.join(d, name + ".bed")
    vcf = os.path.join(d, name + ".vcf") if p["vcf"] else None
    import statsrun
    ref_in_bam = statsrun.fasta_in_bam(ref, bam, os.path.join(out, "tmp_in_bam.fa"))   # tccontext needs every FASTA chromosome in the BAM
    for tool, strict in STATS_RUNS:
        csv = os.path.join(out, tool + ("_strict" if strict else "") + ".csv")
        refrun.run_stats(tool, csv, ref_in_bam if tool == "tccontext" else ref, bam, bed=bed, vcf=vcf, min_qual=p["min_qual"], max_read_length=p["max_read_length"], strict=strict)
    tmp = os.path.join(out, "tmp")
    perread = refrun.run_count_mle(ref, bed, vcf, bam, tmp, p["max_read_length"], p["min_qual"], p["conv_threshold"])
    os.rename(perread, os.path.join(out, "tcount_perread.tsv"))
    for f in os.listdir(out):
        if f.startswith("tmp"):
            os.remove(os.path.join(out, f))
    print(name, "stats files", sorted(os.listdir(out)))


def main():
    if not refrun.available():
        sys.exit("reference tree not found; goldens can only be regenerated in the build container")
    if sys.argv[1:] == ["stats"]:
        for name in STATS:
            make_stats(name)
        return
    if sys.argv[1:] == ["positional"]:
        for name, args in POSITIONAL.items():
            make_positional(name, *args)
        return
    if sys.argv[1:] == ["readsep"]:
        for name, args in READSEP.items():
            make_readsep(name, *args)
        return
    make_cfg1(os.path.join(HERE, "cfg1"))
    for name, (mk, ck) in CASES.items():
        make_count_case(name, mk, ck)
    for name, (mk, fk) in FILTER_CASES.items():
        make_filter_case(name, mk, fk)
    for name, args in POSITIONAL.items():
        make_positional(name, *args)
    for name, args in READSEP.items():
        make_readsep(name, *args)
    for name in STATS:
        make_stats(name)


if __name__ == "__main__":
    main()
