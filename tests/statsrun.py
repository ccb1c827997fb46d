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
