"""Run the UNMODIFIED reference (/root/reference/slamdunk) through the I/O shims.

Test infrastructure only (oracle of last resort).
/root/reference exists in the build container, not on the GPU box: everything that
imports this module must skip when `available()` is False.
"""
import io
import os
import sys

REFERENCE_ROOT = os.environ.get("SLAMDUNK_REFERENCE", "/root/reference")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")
_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "slamdunk", "dunks"))


def _prepare():
    sys.dont_write_bytecode = True          # /root/reference is read-only
    for p in (REFERENCE_ROOT, _SHIMS, _REPO):
        if p not in sys.path:
            sys.path.insert(0, p)


def reference_modules():
    """-> (tcounter, filter) modules of the reference, imported verbatim."""
    if not available():
        raise RuntimeError("reference tree not present at " + REFERENCE_ROOT)
    _prepare()
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from slamdunk.dunks import tcounter, filter as rfilter
    return tcounter, rfilter


def run_count(ref, bed, vcf, bam, out_prefix, max_read_length=100, min_qual=27, conv_threshold=1):
    """reference tcounter.computeTconversions -> (tsv, plus, minus, log text)."""
    tcounter, _ = reference_modules()
    tsv = out_prefix + "_tcount.tsv"
    plus = out_prefix + "_tcount_plus.bedgraph"
    minus = out_prefix + "_tcount_mins.bedgraph"
    log = io.StringIO()
    tcounter.computeTconversions(ref, bed, vcf, bam, max_read_length, min_qual, tsv, plus, minus,
                                 conv_threshold, log)
    return tsv, plus, minus, log.getvalue()


def run_filter(in_bam, out_bam, bed=None, mq=2, min_identity=0.95, nm=-1):
    """reference filter.Filter with `samtools sort` replaced by an in-process coordinate sort
    (samtools is absent; the sort is I/O, not arithmetic). -> log text"""
    _, rfilter = reference_modules()
    from slamdunk_b200.io import bam as bamio

    def bam_sort(outputBAM, log, newHeader, verbose):
        text, refs, recs = bamio.read_bam(outputBAM)
        if newHeader is not None:
            hd = newHeader.to_dict() if hasattr(newHeader, "to_dict") else dict(newHeader)
            text = bamio.header_dict_to_text(hd)
        bamio.write_bam(outputBAM, text, refs, bamio.sort_records(recs))

    rfilter.bamSort = bam_sort
    log = io.StringIO()
    rfilter.Filter(in_bam, out_bam, log, bed, mq, min_identity, nm, False, False, True)
    return log.getvalue()


def run_positional(ref, vcf, bam, out_prefix, min_qual=27, conv_threshold=1, coverage_cutoff=1):
    """reference tcounter.genomewideConversionRates (alleyoop positional-tracks) -> (stdout text, log text);
    the eight bedGraphs are written next to out_prefix."""
    import contextlib
    tcounter, _ = reference_modules()
    log, out = io.StringIO(), io.StringIO()
    with contextlib.redirect_stdout(out):
        tcounter.genomewideConversionRates(ref, vcf, bam, min_qual, out_prefix, conv_threshold, coverage_cutoff, log)
    return out.getvalue(), log.getvalue()


def run_read_separator(ref, vcf, bam, out_prefix, min_qual=27, conv_threshold=1):
    """reference tcounter.genomewideReadSeparation (alleyoop read-separator): writes <prefix>_TCReads.bam and
    <prefix>_backgroundReads.bam."""
    tcounter, _ = reference_modules()
    log = io.StringIO()
    tcounter.genomewideReadSeparation(ref, vcf, bam, min_qual, out_prefix, conv_threshold, log)
    return log.getvalue()


def stats_module():
    """the reference's dunks/stats.py, imported verbatim; its R plot calls (callR) are switched off -- no Rscript here,
    and the plots are not arithmetic."""
    if not available():
        raise RuntimeError("reference tree not present at " + REFERENCE_ROOT)
    _prepare()
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from slamdunk.dunks import stats
    stats.callR = lambda *a, **k: None
    return stats


def run_stats(tool, out_csv, ref, bam, bed=None, vcf=None, min_qual=27, max_read_length=100, strict=False):
    """One of the reference's iterator-based alleyoop statistics -> log text; the CSV is written to out_csv.
    tool: rates | tccontext | utrrates | tcperreadpos | tcperutrpos | snpeval"""
    import warnings
    stats = stats_module()
    log = io.StringIO()
    pdf = out_csv + ".pdf"
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
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


def run_count_mle(ref, bed, vcf, bam, out_prefix, max_read_length=100, min_qual=27, conv_threshold=1):
    """reference computeTconversions(mle=True) with the R fit switched off -> path of the _perread.tsv it wrote"""
    tcounter, _ = reference_modules()
    tcounter.callR = lambda *a, **k: None
    tsv = out_prefix + "_tcount.tsv"
    log = io.StringIO()
    tcounter.computeTconversions(ref, bed, vcf, bam, max_read_length, min_qual, tsv, out_prefix + "_plus.bedgraph",
                                 out_prefix + "_mins.bedgraph", conv_threshold, log, True)
    return out_prefix + "_tcount_perread.tsv"
