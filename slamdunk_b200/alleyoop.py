#!/usr/bin/env python
"""`alleyoop` sub-commands that ride on the GPU count path: `positional-tracks`, `read-separator`, the iterator-based
statistics `rates`, `tccontext`, `utrrates`, `tcperreadpos`, `tcperutrpos`, `snpeval` (CSV output; the R plots are not part of this
package) and the table tool `collapse` (same options, defaults and output names as the reference,
slamdunk/alleyoop.py:115-262, 312-413, 450-575); the other alleyoop tools are plots and table utilities outside the
accelerated path.

    python -m slamdunk_b200.alleyoop positional-tracks -o out -r ref.fa [-s snpdir | -v vcf] [-c 1] [-a 1] [-q 27] a.bam ...
    python -m slamdunk_b200.alleyoop read-separator    -o out -r ref.fa [-s snpdir | -v vcf] [-c 1] [-q 27] a.bam ...
"""
from __future__ import print_function

import os
import sys
from argparse import SUPPRESS, ArgumentDefaultsHelpFormatter, ArgumentParser
from os.path import basename

from .cli import createDir, dunkFinished, getLogFile, message, stepFinished
from .utils.samheader import replaceExtension
from .version import __version__


def runPositionalRates(tid, bam, ref, minQual, conversionThreshold, coverageCutoff, outputDirectory, snpDirectory, vcfFile):
    """alleyoop.py:115-129 -> GPU tracks"""
    from .dunks import tcounter
    outputBedGraphPrefix = os.path.join(outputDirectory, replaceExtension(basename(bam), "", "_positional_rates"))
    outputLOG = os.path.join(outputDirectory, replaceExtension(basename(bam), ".log", "_positional_rates"))
    if vcfFile is not None:
        inputSNP = vcfFile
    elif snpDirectory is not None:
        inputSNP = os.path.join(snpDirectory, replaceExtension(basename(bam), ".vcf", "_snp"))
    else:
        inputSNP = None
    log = getLogFile(outputLOG)
    tcounter.genomewideConversionRates(ref, inputSNP, bam, minQual, outputBedGraphPrefix, conversionThreshold, coverageCutoff, log)
    log.close()
    stepFinished()


def runReadSeparator(tid, bam, ref, minQual, conversionThreshold, outputDirectory, snpDirectory, vcfFile):
    """alleyoop.py:131-146 -> GPU per-read conversion counts"""
    from .dunks import tcounter
    outputBAM = os.path.join(outputDirectory, replaceExtension(basename(bam), "", ""))
    outputLOG = os.path.join(outputDirectory, replaceExtension(basename(bam), ".log", "_read_separator"))
    if vcfFile is not None:
        inputSNP = vcfFile
    elif snpDirectory is not None:
        inputSNP = os.path.join(snpDirectory, replaceExtension(basename(bam), ".vcf", "_snp"))
    else:
        inputSNP = None
    log = getLogFile(outputLOG)
    tcounter.genomewideReadSeparation(ref, inputSNP, bam, minQual, outputBAM, conversionThreshold, log)
    log.close()
    stepFinished()


def runCollapse(tid, tcount, outputDirectory):
    """alleyoop.py:94-100"""
    from .dunks import tcounter
    outputTCOUNT = os.path.join(outputDirectory, replaceExtension(basename(tcount), ".csv", "_collapsed"))
    outputLOG = os.path.join(outputDirectory, replaceExtension(basename(tcount), ".log", "_collapsed"))
    log = getLogFile(outputLOG)
    tcounter.collapse(tcount, outputTCOUNT, log)
    log.close()
    stepFinished()


def _snp_input(bam, snpDirectory, vcfFile):
    if vcfFile is not None:
        return vcfFile
    if snpDirectory is not None:
        return os.path.join(snpDirectory, replaceExtension(basename(bam), ".vcf", "_snp"))
    return None


def _max_read_length(bam, maxReadLength):
    """alleyoop.py:170-174 and the like"""
    from .cli import estimateMaxReadLength
    if maxReadLength is None:
        maxReadLength = estimateMaxReadLength(bam)
    if maxReadLength < 0:
        print("Could not reliable estimate maximum read length. Please specify --max-read-length parameter.")
        sys.exit(0)
    return maxReadLength


def _stats_files(bam, outputDirectory, suffix):
    return tuple(os.path.join(outputDirectory, replaceExtension(basename(bam), ext, suffix)) for ext in (".csv", ".pdf", ".log"))


def runStatsRates(tid, bam, referenceFile, minMQ, outputDirectory):
    """alleyoop.py:147-154"""
    from .dunks import stats
    outputCSV, outputPDF, outputLOG = _stats_files(bam, outputDirectory, "_overallrates")
    log = getLogFile(outputLOG)
    stats.statsComputeOverallRates(referenceFile, bam, minMQ, outputCSV, outputPDF, log)
    log.close()
    stepFinished()


def runStatsTCContext(tid, bam, referenceFile, minMQ, outputDirectory):
    """alleyoop.py:156-163"""
    from .dunks import stats
    outputCSV, outputPDF, outputLOG = _stats_files(bam, outputDirectory, "_tccontext")
    log = getLogFile(outputLOG)
    stats.statsComputeTCContext(referenceFile, bam, minMQ, outputCSV, outputPDF, log)
    log.close()
    stepFinished()


def runStatsRatesUTR(tid, bam, referenceFile, minMQ, strictTCs, outputDirectory, utrFile, maxReadLength):
    """alleyoop.py:165-182"""
    from .dunks import stats
    outputCSV, outputPDF, outputLOG = _stats_files(bam, outputDirectory, "_mutationrates_utr")
    maxReadLength = _max_read_length(bam, maxReadLength)
    log = getLogFile(outputLOG)
    print("Using " + str(maxReadLength) + " as maximum read length.", file=log)
    stats.statsComputeOverallRatesPerUTR(referenceFile, bam, minMQ, strictTCs, outputCSV, outputPDF, utrFile, maxReadLength, log)
    log.close()
    stepFinished()


def runSNPeval(tid, bam, ref, bed, maxLength, minQual, coverageCutoff, variantFraction, strictTCs, outputDirectory, snpDirectory):
    """alleyoop.py:184-207"""
    from .dunks import stats
    outputCSV, outputPDF, outputLOG = _stats_files(bam, outputDirectory, "_SNPeval")
    if not os.path.isdir(snpDirectory):
        print("SNP directory does not exists. Abort.")
        sys.exit(0)
    inputSNP = os.path.join(snpDirectory, replaceExtension(basename(bam), ".vcf", "_snp"))
    maxLength = _max_read_length(bam, maxLength)
    log = getLogFile(outputLOG)
    print("Using " + str(maxLength) + " as maximum read length.", file=log)
    stats.computeSNPMaskedRates(ref, bed, inputSNP, bam, maxLength, minQual, coverageCutoff, variantFraction, outputCSV, outputPDF,
                                strictTCs, log)
    log.close()
    stepFinished()


def runTcPerReadPos(tid, bam, referenceFile, minMQ, maxReadLength, outputDirectory, snpDirectory, vcfFile):
    """alleyoop.py:210-235"""
    from .dunks import stats
    outputCSV, outputPDF, outputLOG = _stats_files(bam, outputDirectory, "_tcperreadpos")
    inputSNP = _snp_input(bam, snpDirectory, vcfFile)
    maxReadLength = _max_read_length(bam, maxReadLength)
    log = getLogFile(outputLOG)
    print("Using " + str(maxReadLength) + " as maximum read length.", file=log)
    stats.tcPerReadPos(referenceFile, bam, minMQ, maxReadLength, outputCSV, outputPDF, inputSNP, log)
    log.close()
    stepFinished()


def runTcPerUtr(tid, bam, referenceFile, bed, minMQ, maxReadLength, outputDirectory, snpDirectory, vcfFile):
    """alleyoop.py:237-262"""
    from .dunks import stats
    outputCSV, outputPDF, outputLOG = _stats_files(bam, outputDirectory, "_tcperutr")
    inputSNP = _snp_input(bam, snpDirectory, vcfFile)
    maxReadLength = _max_read_length(bam, maxReadLength)
    log = getLogFile(outputLOG)
    print("Using " + str(maxReadLength) + " as maximum read length.", file=log)
    stats.tcPerUtr(referenceFile, bed, bam, minMQ, maxReadLength, outputCSV, outputPDF, inputSNP, log, False, True, True)
    log.close()
    stepFinished()


STATS_COMMANDS = ("rates", "tccontext", "utrrates", "snpeval", "tcperreadpos", "tcperutrpos")


def build_parser():
    parser = ArgumentParser(description="alleyoop (GPU subset) v" + __version__, prog="alleyoop", formatter_class=ArgumentDefaultsHelpFormatter)
    parser.add_argument("--version", action="version", version="%(prog)s " + __version__)
    sub = parser.add_subparsers(help="", dest="command")
    p = sub.add_parser("positional-tracks", help="Genome-wide positional tracks as bedgraph", formatter_class=ArgumentDefaultsHelpFormatter)
    p.add_argument("bam", action="store", help="Bam file(s)", nargs="+")
    p.add_argument("-o", "--outputDir", type=str, required=True, dest="outputDir", default=SUPPRESS, help="Output directory for bedGraph files.")
    p.add_argument("-s", "--snp-directory", type=str, required=False, dest="snpDir", default=SUPPRESS, help="Directory containing SNP files.")
    p.add_argument("-v", "--vcf", type=str, required=False, dest="vcfFile", default=SUPPRESS, help="Skip SNP step and provide custom variant file.")
    p.add_argument("-r", "--reference", type=str, required=True, dest="ref", default=SUPPRESS, help="Reference fasta file")
    p.add_argument("-c", "--conversion-threshold", type=int, dest="conversionThreshold", required=False, default=1,
                   help="Number of T>C conversions required to count read as T>C read (default: %(default)d)")
    p.add_argument("-a", "--coverage-cutoff", type=int, dest="coverageCutoff", required=False, default=1,
                   help="Minimum coverage required to report nucleotide-conversion rate (default: %(default)d). Anything less than 1 will be set to 1 to avoid division by zero.")
    p.add_argument("-q", "--min-base-qual", type=int, default=27, required=False, dest="minQual", help="Min base quality for T -> C conversions (default: %(default)d)")
    p.add_argument("-t", "--threads", type=int, required=False, default=1, dest="threads", help="Thread number (default: %(default)d)")
    p.add_argument("--device", type=int, default=0, help="CUDA device")
    p = sub.add_parser("collapse", help="Collapse UTRs", formatter_class=ArgumentDefaultsHelpFormatter)
    p.add_argument("-o", "--outputDir", type=str, required=True, dest="outputDir", default=SUPPRESS, help="Output directory for mapped BAM files.")
    p.add_argument("-t", "--threads", type=int, required=False, default=1, dest="threads", help="Thread number")
    p.add_argument("tcount", action="store", help="Tcount file(s)", nargs="+")
    p = sub.add_parser("read-separator", help="Separate TC-reads from background reads genome-wide", formatter_class=ArgumentDefaultsHelpFormatter)
    p.add_argument("bam", action="store", help="Bam file(s)", nargs="+")
    p.add_argument("-o", "--outputDir", type=str, required=True, dest="outputDir", default=SUPPRESS, help="Output directory for bam files.")
    p.add_argument("-s", "--snp-directory", type=str, required=False, dest="snpDir", default=SUPPRESS, help="Directory containing SNP files.")
    p.add_argument("-v", "--vcf", type=str, required=False, dest="vcfFile", default=SUPPRESS, help="Skip SNP step and provide custom variant file.")
    p.add_argument("-r", "--reference", type=str, required=True, dest="ref", default=SUPPRESS, help="Reference fasta file")
    p.add_argument("-c", "--conversion-threshold", type=int, dest="conversionThreshold", required=False, default=1,
                   help="Number of T>C conversions required to count read as T>C read (default: %(default)d)")
    p.add_argument("-q", "--min-base-qual", type=int, default=27, required=False, dest="minQual", help="Min base quality for T -> C conversions (default: %(default)d)")
    p.add_argument("-t", "--threads", type=int, required=False, default=1, dest="threads", help="Thread number (default: %(default)d)")
    p.add_argument("--device", type=int, default=0, help="CUDA device")
    # the iterator-based statistics, options as in alleyoop.py:336-343, 354-378, 392-413
    p = sub.add_parser("rates", help="Calculate overall conversion rates on SLAM-seq datasets", formatter_class=ArgumentDefaultsHelpFormatter)
    p.add_argument("bam", action="store", help="Bam file(s)", nargs="+")
    p.add_argument("-o", "--outputDir", type=str, required=True, dest="outputDir", default=SUPPRESS, help="Output directory for mapped BAM files.")
    p.add_argument("-r", "--reference", type=str, required=True, dest="referenceFile", default=SUPPRESS, help="Reference fasta file")
    p.add_argument("-mq", "--min-basequality", type=int, required=False, default=27, dest="mq", help="Minimal base quality for SNPs")
    p.add_argument("-t", "--threads", type=int, required=False, default=1, dest="threads", help="Thread number")
    p = sub.add_parser("tccontext", help="Calculate T->C conversion context on SLAM-seq datasets", formatter_class=ArgumentDefaultsHelpFormatter)
    p.add_argument("bam", action="store", help="Bam file(s)", nargs="+")
    p.add_argument("-o", "--outputDir", type=str, required=True, dest="outputDir", default=SUPPRESS, help="Output directory for mapped BAM files.")
    p.add_argument("-r", "--reference", type=str, required=True, dest="referenceFile", default=SUPPRESS, help="Reference fasta file")
    p.add_argument("-mq", "--min-basequality", type=int, required=False, default=0, dest="mq", help="Minimal base quality for SNPs")
    p.add_argument("-t", "--threads", type=int, required=False, default=1, dest="threads", help="Thread number")
    p = sub.add_parser("utrrates", help="Calculate conversion rates per UTR on SLAM-seq datasets")
    p.add_argument("bam", action="store", help="Bam file(s)", nargs="+")
    p.add_argument("-o", "--outputDir", type=str, required=True, dest="outputDir", help="Output directory for mapped BAM files.")
    p.add_argument("-r", "--reference", type=str, required=True, dest="referenceFile", help="Reference fasta file")
    p.add_argument("-mq", "--min-basequality", type=int, required=False, default=27, dest="mq", help="Minimal base quality for SNPs (default: %(default)s)")
    p.add_argument("-m", "--multiTCStringency", dest="strictTCs", action="store_true", required=False, help="")
    p.add_argument("-t", "--threads", type=int, required=False, default=1, dest="threads", help="Thread number (default: %(default)s)")
    p.add_argument("-b", "--bed", type=str, required=True, dest="bed", help="BED file")
    p.add_argument("-l", "--max-read-length", type=int, required=False, dest="maxLength", help="Max read length in BAM file (default: %(default)s)")
    p = sub.add_parser("snpeval", help="Evaluate SNP calling")
    p.add_argument("bam", action="store", help="Bam file(s)", nargs="+")
    p.add_argument("-o", "--outputDir", type=str, required=True, dest="outputDir", help="Output directory for mapped BAM files.")
    p.add_argument("-s", "--snp-directory", type=str, required=True, dest="snpDir", help="Directory containing SNP files.")
    p.add_argument("-r", "--reference", type=str, required=True, dest="ref", help="Reference fasta file")
    p.add_argument("-b", "--bed", type=str, required=True, dest="bed", help="BED file")
    p.add_argument("-c", "--min-coverage", required=False, dest="cov", type=int, help="Minimum coverage to call variant (default: %(default)s)", default=10)
    p.add_argument("-f", "--var-fraction", required=False, dest="var", type=float, help="Minimum variant fraction to call variant (default: %(default)s)", default=0.8)
    p.add_argument("-m", "--multiTCStringency", dest="strictTCs", action="store_true", required=False, help="")
    p.add_argument("-l", "--max-read-length", type=int, required=False, dest="maxLength", help="Max read length in BAM file (default: %(default)s)")
    p.add_argument("-q", "--min-base-qual", type=int, default=27, required=False, dest="minQual", help="Min base quality for T -> C conversions (default: %(default)s)")
    p.add_argument("-t", "--threads", type=int, required=False, default=1, dest="threads", help="Thread number (default: %(default)s)")
    for name, text in (("tcperreadpos", "Calculate conversion rates per read position on SLAM-seq datasets"),
                       ("tcperutrpos", "Calculate conversion rates per UTR position on SLAM-seq datasets")):
        p = sub.add_parser(name, help=text)
        p.add_argument("bam", action="store", help="Bam file(s)", nargs="+")
        p.add_argument("-r", "--reference", type=str, required=True, dest="referenceFile", help="Reference fasta file")
        if name == "tcperutrpos":
            p.add_argument("-b", "--bed", type=str, required=True, dest="bed", help="BED file")
        p.add_argument("-s", "--snp-directory", type=str, required=False, dest="snpDir", help="Directory containing SNP files.")
        p.add_argument("-v", "--vcf", type=str, required=False, dest="vcfFile", default=SUPPRESS, help="Skip SNP step and provide custom variant file.")
        p.add_argument("-l", "--max-read-length", type=int, required=False, dest="maxLength", help="Max read length in BAM file")
        p.add_argument("-o", "--outputDir", type=str, required=True, dest="outputDir", help="Output directory for mapped BAM files.")
        p.add_argument("-mq", "--min-basequality", type=int, required=False, default=27, dest="mq", help="Minimal base quality for SNPs (default: %(default)s)")
        p.add_argument("-t", "--threads", type=int, required=False, dest="threads", default=1, help="Thread number (default: %(default)s)")
    for name in STATS_COMMANDS:
        sub.choices[name].add_argument("--device", type=int, default=0, help="CUDA device")
    # host-only table tools (alleyoop.py:379-390)
    p = sub.add_parser("summary", help="Display summary information and statistics on read numbers")
    p.add_argument("bam", action="store", help="Filtered BAM files (produced by slamdunk filter or all)", nargs="+")
    p.add_argument("-o", "--output", type=str, required=True, dest="outputFile", help="Output file")
    p.add_argument("-t", "--tcountDir", type=str, required=False, dest="countDirectory", help="Folder containing tcount files")
    p = sub.add_parser("merge", help="Merge T->C rates from multiple sample in one TSV file", formatter_class=ArgumentDefaultsHelpFormatter)
    p.add_argument("countFiles", action="store", help="tCount files", nargs="+")
    p.add_argument("-o", "--output", type=str, required=True, dest="outputFile", default=SUPPRESS, help="Output file")
    p.add_argument("-c", "--column", dest="column", type=str, required=False, default="TcReadCount / ReadCount", help="Column or expression used to summarize files.")
    p.add_argument("-n", "--columnname", dest="columnName", type=int, required=False, default=2, help="Index of meta data field to use as column name.")
    # tools of the reference that are NOT provided: registered so that the failure says why instead of "invalid choice"
    for name, what in UNSUPPORTED.items():
        p = sub.add_parser(name, help=what + " (not provided by slamdunk_b200)")
        p.add_argument("rest", nargs="*")
    return parser


UNSUPPORTED = {"dedup": "Deduplicate SLAM-seq aligned data", "dump": "Print all info available for slamdunk reads",
               "half-lifes": "Compute half-lifes"}


def _run_stats(args):
    """alleyoop.py:491-575: one sample after the other on the one GPU (the reference starts joblib workers)"""
    from .dunks import tcounter
    tcounter.OPTIONS["device"] = args.device
    tcounter.OPTIONS["threads"] = args.threads if args.threads > 1 else 0
    outputDirectory = args.outputDir
    createDir(outputDirectory)
    n = args.threads
    command = args.command
    label = {"rates": "rates", "snpeval": "SNPeval", "tccontext": "TC context"}.get(command, command)
    message("Running alleyoop " + label + " for " + str(len(args.bam)) + " files (" + str(n) + " threads)")
    snpDirectory = getattr(args, "snpDir", None)
    vcfFile = getattr(args, "vcfFile", None)
    for tid in range(0, len(args.bam)):
        bam = args.bam[tid]
        if command == "rates":
            runStatsRates(tid, bam, args.referenceFile, args.mq, outputDirectory)
        elif command == "tccontext":
            runStatsTCContext(tid, bam, args.referenceFile, args.mq, outputDirectory)
        elif command == "utrrates":
            runStatsRatesUTR(tid, bam, args.referenceFile, args.mq, args.strictTCs, outputDirectory, args.bed, args.maxLength)
        elif command == "snpeval":
            runSNPeval(tid, bam, args.ref, args.bed, args.maxLength, args.minQual, args.cov, args.var, args.strictTCs, outputDirectory,
                       snpDirectory)
        elif command == "tcperreadpos":
            runTcPerReadPos(tid, bam, args.referenceFile, args.mq, args.maxLength, outputDirectory, snpDirectory, vcfFile)
        else:
            runTcPerUtr(tid, bam, args.referenceFile, args.bed, args.mq, args.maxLength, outputDirectory, snpDirectory, vcfFile)
    dunkFinished()


def run(argv=None):
    parser = build_parser()
    if argv is None:
        argv = sys.argv[1:]
    if argv and argv[0] in UNSUPPORTED:
        sys.exit("alleyoop " + argv[0] + " is not provided by slamdunk_b200 (it is outside the accelerated count path); "
                 "use the reference's alleyoop for it.")
    args = parser.parse_args(argv)
    if args.command not in ("positional-tracks", "read-separator", "collapse", "summary", "merge") + STATS_COMMANDS:
        parser.print_help()
        sys.exit(0)
    if args.command == "summary":                        # alleyoop.py:533-536
        from .dunks import tables
        message("Running alleyoop summary for " + str(len(args.bam)) + " files")
        log = getLogFile(replaceExtension(args.outputFile, ".log"))
        tables.readSummary(args.bam, args.countDirectory, args.outputFile, log)
        log.close()
        dunkFinished()
        return
    if args.command == "merge":                          # alleyoop.py:538-542
        from .dunks import tables
        message("Running alleyoop merge for " + str(len(args.countFiles)) + " files")
        log = getLogFile(replaceExtension(args.outputFile, ".log"))
        tables.mergeRates(",".join(args.countFiles), args.outputFile, args.column, args.columnName, log)
        log.close()
        dunkFinished()
        return
    if args.command in STATS_COMMANDS:
        _run_stats(args)
        return
    if args.command == "collapse":
        createDir(args.outputDir)
        message("Running alleyoop collapse for " + str(len(args.tcount)) + " files (" + str(args.threads) + " threads)")
        for tid in range(0, len(args.tcount)):
            runCollapse(tid, args.tcount[tid], args.outputDir)
        dunkFinished()
        return
    from .dunks import tcounter
    tcounter.OPTIONS["device"] = args.device
    tcounter.OPTIONS["threads"] = args.threads if args.threads > 1 else 0
    outputDirectory = args.outputDir
    createDir(outputDirectory)
    snpDirectory = args.snpDir if "snpDir" in args else None
    vcfFile = args.vcfFile if "vcfFile" in args else None
    message("Running alleyoop " + args.command + " for " + str(len(args.bam)) + " files (" + str(args.threads) + " threads)")
    for tid in range(0, len(args.bam)):                  # one GPU: samples one after the other (alleyoop.py:463,479 use joblib workers)
        if args.command == "positional-tracks":
            runPositionalRates(tid, args.bam[tid], args.ref, args.minQual, args.conversionThreshold, args.coverageCutoff,
                               outputDirectory, snpDirectory, vcfFile)
        else:
            runReadSeparator(tid, args.bam[tid], args.ref, args.minQual, args.conversionThreshold, outputDirectory, snpDirectory, vcfFile)
    dunkFinished()


if __name__ == "__main__":
    run()
