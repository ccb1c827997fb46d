#!/usr/bin/env python
"""`alleyoop` sub-commands that ride on the GPU count path: `positional-tracks`, `read-separator` and the table tool
`collapse` (same options,
defaults and output names as the reference, slamdunk/alleyoop.py:115-144, 312-333, 450-481); the other alleyoop tools
are plots and table utilities outside the accelerated path.

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
    return parser


def run(argv=None):
    parser = build_parser()
    args = parser.parse_args(argv)
    if args.command not in ("positional-tracks", "read-separator", "collapse"):
        parser.print_help()
        sys.exit(0)
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
