"""Shared by the stats tests: run one iterator-based statistic through the oracle restatement (oracle/stats_oracle.py),
the GPU drop-in (slamdunk_b200/dunks/stats.py) or the unmodified reference (refrun), same arguments for all three."""
import io


def run_oracle(tool, out_csv, ref, bam, bed, vcf, min_qual, max_read_length, strict):
    from oracle import stats_oracle as so
    if tool == "rates":
        so.rates_file(ref, bam, min_qual, out_csv)
    elif tool == "tccontext":
        so.tccontext_file(ref, bam, out_csv)
    elif tool == "utrrates":
        so.utrrates_file(ref, bam, min_qual, strict, out_csv, bed, max_read_length)
    elif tool == "tcperreadpos":
        so.tcperreadpos_file(ref, bam, min_qual, max_read_length, out_csv, vcf)
    elif tool == "tcperutrpos":
        so.tcperutr_file(ref, bed, bam, min_qual, max_read_length, out_csv, vcf)
    elif tool == "snpeval":
        so.snpeval_file(ref, bed, vcf, bam, max_read_length, min_qual, out_csv, strict)
    else:
        raise KeyError(tool)


def run_gpu(tool, out_csv, ref, bam, bed, vcf, min_qual, max_read_length, strict):
    from slamdunk_b200.dunks import stats
    log = io.StringIO()
    pdf = out_csv + ".pdf"
    if tool == "rates":
        stats.statsComputeOverallRates(ref, bam, min_qual, out_csv, pdf, log)
    elif tool == "tccontext":
        stats.statsComputeTCContext(ref, bam, min_qual, out_csv, pdf, log)
    elif tool == "utrrates":
        stats.statsComputeOverallRatesPerUTR(ref, bam, min_qual, strict, out_csv, pdf, bed, max_read_length, log)
    elif tool == "tcperreadpos":
        stats.tcPerReadPos(ref, bam, min_qual, max_read_length, out_csv, pdf, vcf, log)
    elif tool == "tcperutrpos":
        stats.tcPerUtr(ref, bed, bam, min_qual, max_read_length, out_csv, pdf, vcf, log, False, True, True)
    elif tool == "snpeval":
        stats.computeSNPMaskedRates(ref, bed, vcf, bam, max_read_length, min_qual, 10, 0.8, out_csv, pdf, strict, log)
    else:
        raise KeyError(tool)
    return log.getvalue()


def perread_columns(path):
    """(tInReads text, tcInRead text) of every row of a _perread.tsv"""
    rows = [l.rstrip("\n").split("\t") for l in open(path) if not l.startswith("#") and not l.startswith("Chromosome")]
    return [(r[11], r[12]) for r in rows]


def unstrand_some_rows(bed_in, bed_out, every=4):
    """copy of a BED file in which every `every`-th row has no strand ('.'): the iterator then takes both read strands"""
    with open(bed_in) as fi, open(bed_out, "w") as fo:
        for i, line in enumerate(fi):
            f = line.rstrip("\n").split("\t")
            if i % every == 1 and len(f) > 5:
                f[5] = "."
            fo.write("\t".join(f) + "\n")


def fasta_in_bam(ref, bam, out_fa):
    """copy of a FASTA without the chromosomes missing from the BAM header: `alleyoop tccontext` fetches every FASTA
    chromosome from the BAM without asking whether it is there (stats.py:259) and dies on the first missing one"""
    from slamdunk_b200.io import bam as bamio
    names = set(n for n, _l in bamio.read_bam(bam)[1])
    seqs = bamio.read_fasta(ref)
    bamio.write_fasta(out_fa, dict((k, v) for k, v in seqs.items() if k in names))
    return out_fa


def make_edge_case(outdir):
    """A hand-made data set for the corners of the iterator that random data does not reach: MP entries pointing past the
    read (quality 0, SlamSeqFile.py:326-329; on a reverse read the mirrored position is negative and indexes the per-position
    lists from their end, stats.py:627), N bases and soft / hard clips in the reads (matrix diagonal over the un-clipped
    part only, SlamSeqFile.py:234-237), entries before the UTR start and past its end, a read reaching over the chromosome
    end, mapq 0, reads without MP tag, every conversion code incl. the N ones, a UTR shorter than utrNormFactor and one
    longer, SNPs under conversions on both strands.  -> (ref.fa, utrs.bed, snps.vcf, reads.bam); use with minQual 0 so that
    the out-of-read entries are kept."""
    import hashlib
    import os
    from slamdunk_b200.io import bam as bamio
    os.makedirs(outdir, exist_ok=True)
    import random
    rng = random.Random(4242)
    g = {"chrA": "".join(rng.choice("ACGT") for _ in range(1400)), "chrB": "".join(rng.choice("ACGTN") for _ in range(500))}
    ref = os.path.join(outdir, "edge.fa")
    bamio.write_fasta(ref, g)
    bed = os.path.join(outdir, "edge.bed")
    with open(bed, "w") as fh:
        fh.write("chrA\t100\t700\tlong_plus\t0\t+\n")       # longer than utrNormFactor
        fh.write("chrA\t650\t800\tshort_minus\t0\t-\n")     # shorter, overlaps the first
        fh.write("chrA\t1300\t1500\tover_the_end\t0\t+\n")  # past the chromosome end
        fh.write("chrB\t10\t200\tnmix\t0\t-\n")
        fh.write("chrA\t300\t320\ttiny\t0\t+\n")
    vcf = os.path.join(outdir, "edge.vcf")
    with open(vcf, "w") as fh:
        fh.write("##fileformat=VCFv4.1\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n")
        for chrom, pos, r, a in (("chrA", 153, "T", "C"), ("chrA", 706, "A", "G"), ("chrA", 310, "T", "C"), ("chrB", 60, "A", "G"),
                                 ("chrA", 500, "T", "C,G"), ("chrA", 505, "G", "A")):
            fh.write("%s\t%d\t.\t%s\t%s\t.\tPASS\t.\n" % (chrom, pos, r, a))
    q = lambda n, v=35: [v] * n            # noqa: E731
    R = bamio.BamRecord
    recs = [
        # forward reads on long_plus; MP is conv:readPos:refPos (1-based)
        R("f1", 0, 0, 150, 60, "40M", "ACGT" * 10, q(40), [("MP", "Z", "16:3:3,8:12:12,16:45:45"), ("NM", "i", 2), ("XI", "f", 0.95)]),   # 45 > 40: quality 0
        R("f2", 0, 0, 90, 60, "5S30M5S", "N" * 5 + "ACGTN" * 6 + "T" * 5, q(20, 30) + q(20, 10), [("MP", "Z", "16:8:3,19:10:5,24:12:7,4:20:15"), ("NM", "i", 3), ("XI", "f", 0.9)]),
        R("f3", 0, 0, 660, 0, "3H20M2D18M", "ACGTTGCATT" * 3 + "ACGTACGT", q(38), [("MP", "Z", "16:38:40,1:1:1"), ("NM", "i", 4), ("XI", "f", 0.9)]),   # mapq 0, deletion, near the UTR end
        R("f4", 0, 0, 305, 60, "20M", "T" * 20, q(20), []),                                                                                 # no MP tag at all
        R("f5", 0, 0, 1380, 60, "30M", "ACGTT" * 6, q(30), [("MP", "Z", "16:30:30"), ("NM", "i", 1), ("XI", "f", 0.97)]),                   # runs over the chromosome end
        R("f6", 0, 0, 305, 60, "10M1I9M", "TTTTTCCCCCGGGGGAAAAA", q(20), [("MP", "Z", "16:6:6,16:7:7,17:16:15"), ("NM", "i", 3), ("XI", "f", 0.85)]),   # conversion on a SNP (pos 310)
        # reverse reads
        R("r1", 16, 0, 700, 60, "30M", "AGCTA" * 6, q(30), [("MP", "Z", "2:7:7,2:35:35,13:2:2"), ("NM", "i", 3), ("XI", "f", 0.9)]),         # 35 > 30 on a reverse read: negative mirrored position
        R("r2", 16, 0, 640, 60, "4S26M", "GGGG" + "AGAGAGNNTTAACCGGAAGGTTCCAA", q(30), [("MP", "Z", "2:5:1,2:70:67,22:11:7"), ("NM", "i", 3), ("XI", "f", 0.9)]),
        R("r3", 16, 1, 50, 60, "25M", "NNAGAGAGTTAACCGGAAGGTTCCA", q(25), [("MP", "Z", "2:11:11,20:1:1,9:5:5"), ("NM", "i", 3), ("XI", "f", 0.9)]),
        R("r4", 16, 1, 5, 0, "10M", "AAAAAGGGGG", q(10), [("MP", "Z", "2:6:6,2:7:7"), ("NM", "i", 2), ("XI", "f", 0.8)]),
        R("x1", 0, 1, 100, 60, "20M", "ACGTACGTACGTACGTACGT", q(20), [("MP", "Z", "16:4:4"), ("NM", "i", 1), ("XI", "f", 0.95)]),            # antisense to the chrB UTR
    ]
    refs = [("chrA", 1400), ("chrB", 500)]
    md5 = hashlib.md5(open(bed, "rb").read()).hexdigest()
    ds = ("{'sequenced':11,'mapped':11,'filtered':11,'mqfiltered':0,'idfiltered':0,'nmfiltered':0,'multimapper':0,'dedup':0,'snps':6,"
          "'annotation':'edge.bed','annotationmd5':'%s'}" % md5)
    hdr = "@HD\tVN:1.0\tSO:coordinate\n" + "".join("@SQ\tSN:%s\tLN:%d\n" % r for r in refs)
    hdr += "@RG\tID:3\tSM:edge:pulse:45\tDS:%s\n@PG\tID:slamdunk\tPN:slamdunk filter v0.4.3\tVN:3\n" % ds
    bam = os.path.join(outdir, "edge.bam")
    bamio.write_bam(bam, hdr, refs, bamio.sort_records(recs))
    return ref, bed, vcf, bam
