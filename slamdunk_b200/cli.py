#!/usr/bin/env python
"""`slamdunk` command line with the count / filter steps on the GPU.

Same sub-commands, options, defaults, output file names and directory layout as the reference CLI
(slamdunk/slamdunk.py:332-524): `count`, `filter`, `all` (and `map` / `snp`, which -- exactly like
the reference -- only build the command lines of the external NextGenMap / samtools / VarScan
binaries and shell out; they are not part of the accelerated path).  `-t` is the number of host
decode threads: samples are processed one after the other on the GPU instead of one joblib worker
process per BAM (slamdunk.py:516).  Extra, opt-in: `--device`, `--mismatch-source`.

Several GPUs: `torchrun --nproc-per-node N -m slamdunk_b200 count ...` (one process per GPU).  With at least as many BAMs as
ranks the SAMPLES are dealt out over the ranks -- the reference's one-worker-per-BAM fan-out (slamdunk.py:516), no collective
(partitioning B); with fewer BAMs than ranks every BAM is counted by all ranks together (tile slices, one
all-reduce of the integer outputs, rank 0 writes).  `filter` deals out samples the same way.
"""
from __future__ import print_function

import os
import random
import subprocess
import sys
from argparse import SUPPRESS, ArgumentDefaultsHelpFormatter, ArgumentParser
from os.path import basename
from time import sleep

from .utils.samheader import checkStep, replaceExtension
from .version import __ngm_version__, __version__

mainOutput = sys.stderr
printOnly = False
verbose = False


def message(msg):
    print(msg, file=mainOutput)


def stepFinished():
    print(".", end="", file=mainOutput)


def dunkFinished():
    print("", file=mainOutput)


def createDir(directory):
    if not os.path.exists(directory):
        message("Creating output directory: " + directory)
        os.makedirs(directory)


def getLogFile(path):
    return open(path, "a")


def _run(cmd, log, dry=False):
    """utils/misc.py:184-196"""
    if verbose or dry:
        print(cmd, file=log)
    if not dry:
        p = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, shell=True)
        for line in iter(p.stdout.readline, b""):
            print(line.decode("utf-8", "replace"), end="", file=log)
        p.wait()
        if p.returncode != 0:
            raise RuntimeError("Error while executing command: \"" + cmd + "\"")


def readSampleFile(fileName):
    """slamdunk.py:82-101"""
    samples, infos = [], []
    with open(fileName, "r") as ins:
        for line in ins:
            line = line.strip()
            if len(line) > 1:
                if fileName.endswith(".tsv"):
                    cols = line.split("\t")
                elif fileName.endswith(".csv"):
                    cols = line.split(",")
                else:
                    raise RuntimeError("Unknown file extension found: " + fileName)
                if len(cols) < 4:
                    raise RuntimeError("Invalid sample file found: " + fileName)
                samples.append(cols[0])
                infos.append(cols[1] + ":" + cols[2] + ":" + cols[3])
    return samples, infos


def getSamples(bams, runOnly=-1):
    """slamdunk.py:103-123"""
    if len(bams) == 1 and (bams[0].endswith(".tsv") or bams[0].endswith(".csv")):
        samples, samplesInfos = readSampleFile(bams[0])
    else:
        samples = bams
        samplesInfos = [""] * len(samples)
    if runOnly > 0:
        if runOnly > len(samples):
            raise RuntimeError("Sample index out of range. " + str(runOnly) + " > " + str(len(samples)) + ". Check -i/--sample-index")
        message("Running only job " + str(runOnly))
        samples = [samples[runOnly - 1]]
        samplesInfos = [samplesInfos[runOnly - 1]]
    elif runOnly == 0:
        raise RuntimeError("Sample index (" + str(runOnly) + ") out of range. Starts with 1. Check -i/--sample-index")
    return samples, samplesInfos


def estimateMaxReadLength(bam):
    """utils/misc.py:91-107: first 1000 reads, query_length + XA tag; spread <= 10 -> max + 10 else -1."""
    from .io import bam as bamio
    minLength, maxLength = sys.maxsize, 0
    for read in bamio.head_records(bam, 1000):
        ln = len(read.seq) + read.get_tag("XA")
        minLength = min(minLength, ln)
        maxLength = max(maxLength, ln)
    return maxLength + 10 if (maxLength - minLength) <= 10 else -1


# ------------------------------------------------------------------ steps
def runMap(tid, inputBAM, referenceFile, threads, trim5p, maxPolyA, quantseqMapping, endtoendMapping, topn, sampleDescription,
           outputDirectory, skipSAM):
    """slamdunk.py:125-151 + dunks/mapper.py:77-109: NextGenMap command line (external binary)."""
    ext = ".bam" if skipSAM else ".sam"
    outputSAM = os.path.join(outputDirectory, replaceExtension(basename(inputBAM), ext, "_slamdunk_mapped"))
    outputLOG = os.path.join(outputDirectory, replaceExtension(basename(inputBAM), ".log", "_slamdunk_mapped"))
    sampleName = replaceExtension(basename(inputBAM), ".bam", "")
    sampleType, sampleTime = "NA", "-1"
    if sampleDescription != "":
        d = sampleDescription.split(":")
        if len(d) >= 1:
            sampleName = d[0]
        if len(d) >= 2:
            typeDict = {"p": "pulse", "c": "chase", "pulse": "pulse", "chase": "chase", "": "NA"}
            sampleType = typeDict.get(d[1], d[1])
        if len(d) >= 3:
            sampleTime = d[2]
    parameter = "--no-progress" if quantseqMapping is True else "--no-progress --slam-seq 2"
    if trim5p > 0:
        parameter += " -5 " + str(trim5p)
    if maxPolyA > -1:
        parameter += " --max-polya " + str(maxPolyA)
    parameter += " -e " if endtoendMapping is True else " -l "
    parameter += " --rg-id " + str(tid)
    if sampleName != "":
        parameter += " --rg-sm " + sampleName + ":" + sampleType + ":" + str(sampleTime)
    if topn > 1:
        parameter += " -n " + str(topn) + " --strata "
    log = getLogFile(outputLOG)
    if checkStep([referenceFile, inputBAM], [replaceExtension(outputSAM, ".bam")], False):
        flag = "" if outputSAM.endswith(".sam") else "-b "
        _run("ngm " + flag + "-r " + referenceFile + " -q " + inputBAM + " -t " + str(threads) + " " + parameter + " -o " + outputSAM,
             log, dry=printOnly)
    else:
        print("Skipped mapping for " + inputBAM, file=log)
    log.close()
    stepFinished()


def runSam2Bam(tid, bam, threads, outputDirectory):
    """slamdunk.py:153-157 + dunks/mapper.py:28-74 (samtools view, no sort)."""
    inputSAM = os.path.join(outputDirectory, replaceExtension(basename(bam), ".sam", "_slamdunk_mapped"))
    outputBAM = os.path.join(outputDirectory, replaceExtension(basename(bam), ".bam", "_slamdunk_mapped"))
    log = getLogFile(os.path.join(outputDirectory, replaceExtension(basename(bam), ".log", "_slamdunk_mapped")))
    if os.path.exists(inputSAM) and checkStep([inputSAM], [outputBAM + ".bai"]):
        _run(" ".join(["samtools view", "-@", str(threads), "-Sb", "-o", outputBAM, inputSAM]), log, dry=printOnly)
        if not printOnly and os.path.exists(inputSAM):
            os.remove(inputSAM)
    else:
        print("Skipped sorting for " + inputSAM, file=log)
    log.close()
    stepFinished()


def runFilter(tid, bam, bed, mq, minIdentity, maxNM, outputDirectory):
    """slamdunk.py:167-171 -> GPU filter"""
    from .dunks import filter as gfilter
    outputBAM = os.path.join(outputDirectory, replaceExtension(basename(bam), ".bam", "_filtered"))
    outputLOG = os.path.join(outputDirectory, replaceExtension(basename(bam), ".log", "_filtered"))
    log = getLogFile(outputLOG)
    gfilter.Filter(bam, outputBAM, log, bed, mq, minIdentity, maxNM, printOnly, verbose)
    log.close()
    stepFinished()


def runSnp(tid, referenceFile, minCov, minVarFreq, minQual, inputBAM, outputDirectory):
    """slamdunk.py:173-177 + dunks/snps.py:25-44: samtools mpileup | varscan mpileup2snp (external)."""
    outputSNP = os.path.join(outputDirectory, replaceExtension(basename(inputBAM), ".vcf", "_snp"))
    log = getLogFile(os.path.join(outputDirectory, replaceExtension(basename(inputBAM), ".log", "_snp")))
    if checkStep([inputBAM, referenceFile], [outputSNP], False):
        mpileupCmd = "samtools mpileup -B -A -f " + referenceFile + " " + inputBAM
        varscanCmd = ("varscan mpileup2snp  --strand-filter 0 --output-vcf --min-var-freq " + str(minVarFreq) +
                      " --min-coverage " + str(minCov) + " --variants 1")
        if verbose:
            print(mpileupCmd, file=log)
            print(varscanCmd, file=log)
        if not printOnly:
            with open(outputSNP, "w") as fileSNP:
                mpileup = subprocess.Popen(mpileupCmd, shell=True, stdout=subprocess.PIPE, stderr=log)
                varscan = subprocess.Popen(varscanCmd, shell=True, stdin=mpileup.stdout, stdout=fileSNP, stderr=log)
                varscan.wait()
    else:
        print("Skipping SNP calling", file=log)
    log.close()
    stepFinished()


def runCount(tid, bam, ref, bed, maxLength, minQual, conversionThreshold, outputDirectory, snpDirectory, vcfFile):
    """slamdunk.py:179-204 -> GPU count"""
    from .dunks import tcounter
    outputCSV = os.path.join(outputDirectory, replaceExtension(basename(bam), ".tsv", "_tcount"))
    outputBedgraphPlus = os.path.join(outputDirectory, replaceExtension(basename(bam), ".bedgraph", "_tcount_plus"))
    outputBedgraphMinus = os.path.join(outputDirectory, replaceExtension(basename(bam), ".bedgraph", "_tcount_mins"))
    outputLOG = os.path.join(outputDirectory, replaceExtension(basename(bam), ".log", "_tcount"))
    if vcfFile is not None:
        inputSNP = vcfFile
    elif snpDirectory is not None:
        inputSNP = os.path.join(snpDirectory, replaceExtension(basename(bam), ".vcf", "_snp"))
    else:
        inputSNP = None
    if maxLength is None:
        maxLength = estimateMaxReadLength(bam)
    if maxLength < 0:
        print("Difference between minimum and maximum read length is > 10. Please specify --max-read-length parameter.")
        sys.exit(0)
    log = getLogFile(outputLOG)
    print("Using " + str(maxLength) + " as maximum read length.", file=log)
    tcounter.computeTconversions(ref, bed, inputSNP, bam, maxLength, minQual, outputCSV, outputBedgraphPlus, outputBedgraphMinus,
                                 conversionThreshold, log)
    log.close()
    stepFinished()
    return outputCSV


def runAll(args):
    """slamdunk.py:206-330"""
    message("slamdunk all")
    if args.sampleIndex > -1:
        sec = random.randrange(200, 2000) / 1000.0
        message("Waiting " + str(sec) + " seconds")
        sleep(sec)
    outputDirectory = args.outputDir
    createDir(outputDirectory)
    n = args.threads
    referenceFile = args.referenceFile
    dunkPath = os.path.join(outputDirectory, "map")
    createDir(dunkPath)
    samples, samplesInfos = getSamples(args.files, runOnly=args.sampleIndex)
    message("Running slamDunk map for " + str(len(samples)) + " files (" + str(n) + " threads)")
    for i in range(0, len(samples)):
        bam = samples[i]
        if not args.sampleName or len(samples) > 1:
            sampleName = replaceExtension(basename(bam), "", "")
        else:
            sampleName = args.sampleName
        sampleInfo = samplesInfos[i]
        if sampleInfo == "":
            sampleInfo = sampleName + ":" + args.sampleType + ":" + str(args.sampleTime)
        tid = args.sampleIndex if args.sampleIndex > -1 else i
        runMap(tid, bam, referenceFile, n, args.trim5, args.maxPolyA, args.quantseq, args.endtoend, args.topn, sampleInfo, dunkPath, args.skipSAM)
    dunkFinished()
    if not args.skipSAM:
        message("Running slamDunk sam2bam for " + str(len(samples)) + " files (" + str(n) + " threads)")
        for tid in range(0, len(samples)):
            runSam2Bam(tid, samples[tid], n, dunkPath)
        dunkFinished()
    dunkbufferIn = [os.path.join(dunkPath, replaceExtension(basename(f), ".bam", "_slamdunk_mapped")) for f in samples]
    bed = args.bed
    if args.filterbed:
        bed = args.filterbed
        args.multimap = True
    if not args.multimap:
        bed = None
    dunkPath = os.path.join(outputDirectory, "filter")
    createDir(dunkPath)
    message("Running slamDunk filter for " + str(len(samples)) + " files (" + str(n) + " threads)")
    for tid in range(0, len(samples)):
        runFilter(tid, dunkbufferIn[tid], bed, args.mq, args.identity, args.nm, dunkPath)
    dunkFinished()
    dunkbufferIn = [os.path.join(dunkPath, replaceExtension(basename(f), ".bam", "_filtered")) for f in dunkbufferIn]
    snpDirectory, vcfFile = None, None
    if "vcfFile" not in args:
        dunkPath = os.path.join(outputDirectory, "snp")
        createDir(dunkPath)
        message("Running slamDunk SNP for " + str(len(samples)) + " files (" + str(max(1, int(n / 2)) if n > 1 else n) + " threads)")
        for tid in range(0, len(samples)):
            runSnp(tid, referenceFile, args.cov, args.var, args.minQual, dunkbufferIn[tid], dunkPath)
        snpDirectory = os.path.join(outputDirectory, "snp")
        dunkFinished()
    else:
        vcfFile = args.vcfFile
    dunkPath = os.path.join(outputDirectory, "count")
    createDir(dunkPath)
    message("Running slamDunk tcount for " + str(len(samples)) + " files (" + str(n) + " threads)")
    for tid in range(0, len(samples)):
        runCount(tid, dunkbufferIn[tid], referenceFile, args.bed, args.maxLength, args.minQual, args.conversionThreshold, dunkPath,
                 snpDirectory, vcfFile)
    dunkFinished()


# ---------------------------------------------------------------------------------------------
# Argument surface.  Flags, destinations, types and defaults are those of the reference CLI
# (slamdunk/slamdunk.py:345-430) so existing command lines keep working; they are kept as data.
# (flags, dest, type, default, required, help) ; type "flag" = store_true ; default SUPPRESS = absent unless given
# ---------------------------------------------------------------------------------------------
_S = SUPPRESS
_SAMPLE_OPTS = [
    (("-sn", "--sampleName"), "sampleName", str, None, False, "sample name used for every input"),
    (("-sy", "--sampleType"), "sampleType", str, "pulse", False, "sample type used for every input"),
    (("-st", "--sampleTime"), "sampleTime", int, 0, False, "sample time used for every input"),
    (("-i", "--sample-index"), "sampleIndex", int, -1, False, "process only sample <i> (1-based), for cluster arrays"),
    (("-ss", "--skip-sam"), "skipSAM", "flag", False, False, "let the mapper write BAM directly"),
]
_MAP_OPTS = [
    (("-5", "--trim-5p"), "trim5", int, 12, False, "bases trimmed from the 5' end of every read"),
    (("-n", "--topn"), "topn", int, 1, False, "alignments reported per read"),
    (("-a", "--max-polya"), "maxPolyA", int, 4, False, "As tolerated at the 3' end"),
    (("-t", "--threads"), "threads", int, 1, False, "threads"),
    (("-q", "--quantseq"), "quantseq", "flag", False, False, "plain QuantSeq alignment (no SLAM-seq scoring)"),
    (("-e", "--endtoend"), "endtoend", "flag", False, False, "end-to-end instead of local alignment"),
]
_FILTER_OPTS = [
    (("-mq", "--min-mq"), "mq", int, 2, False, "minimum mapping quality"),
    (("-mi", "--min-identity"), "identity", float, 0.95, False, "minimum alignment identity"),
    (("-nm", "--max-nm"), "nm", int, -1, False, "maximum NM per alignment (-1: off)"),
]
_COUNT_OPTS = [
    (("-c", "--conversion-threshold"), "conversionThreshold", int, 1, False, "T>C conversions that make a read a T>C read"),
    (("-q", "--min-base-qual"), "minQual", int, 27, False, "minimum base quality of a counted conversion"),
]
SUBCOMMANDS = {
    "map": dict(help="Map SLAM-seq read data", positional=("files", "sample sheet (csv/tsv) or read files"), options=[
        (("-r", "--reference"), "referenceFile", str, _S, True, "reference FASTA"),
        (("-o", "--outputDir"), "outputDir", str, _S, True, "output directory"),
    ] + _MAP_OPTS + _SAMPLE_OPTS),
    "filter": dict(help="Filter SLAM-seq aligned data", positional=("bam", "BAM file(s)"), options=[
        (("-o", "--outputDir"), "outputDir", str, None, True, "output directory"),
        (("-b", "--bed"), "bed", str, None, False, "BED file: switches to multimapper retention (no MQ cut)"),
    ] + _FILTER_OPTS + [(("-t", "--threads"), "threads", int, 1, False, "host decode threads")]),
    "snp": dict(help="Call SNPs on SLAM-seq aligned data", positional=("bam", "BAM file(s)"), options=[
        (("-o", "--outputDir"), "outputDir", str, _S, True, "output directory"),
        (("-r", "--reference"), "fasta", str, _S, True, "reference FASTA"),
        (("-c", "--min-coverage"), "cov", int, 10, False, "minimum coverage of a variant"),
        (("-f", "--var-fraction"), "var", float, 0.8, False, "minimum variant fraction"),
        (("-t", "--threads"), "threads", int, 1, False, "threads"),
    ]),
    "count": dict(help="Count T/C conversions in SLAM-seq aligned data", positional=("bam", "BAM file(s)"), options=[
        (("-o", "--outputDir"), "outputDir", str, _S, True, "output directory"),
        (("-s", "--snp-directory"), "snpDir", str, _S, False, "directory with <sample>_snp.vcf files"),
        (("-v", "--vcf"), "vcfFile", str, _S, False, "one variant file for all samples"),
        (("-r", "--reference"), "ref", str, _S, True, "reference FASTA"),
        (("-b", "--bed"), "bed", str, _S, True, "BED file of the 3' UTRs"),
        (("-l", "--max-read-length"), "maxLength", int, None, False, "maximum read length in the BAM"),
        (("-t", "--threads"), "threads", int, 1, False, "host decode threads"),
    ] + _COUNT_OPTS),
    "all": dict(help="Run entire SLAMdunk analysis", positional=("files", "sample sheet (csv/tsv) or read files"), options=[
        (("-r", "--reference"), "referenceFile", str, None, True, "reference FASTA"),
        (("-b", "--bed"), "bed", str, None, True, "BED file of the 3' UTRs"),
        (("-fb", "--filterbed"), "filterbed", str, None, False, "BED file for multimapper retention (implies -m)"),
        (("-v", "--vcf"), "vcfFile", str, _S, False, "skip SNP calling, use this variant file"),
        (("-o", "--outputDir"), "outputDir", str, None, True, "output directory"),
        (("-m", "--multimap"), "multimap", "flag", False, False, "resolve multimappers with the annotation (needs -n > 1)"),
        (("-mc", "--min-coverage"), "cov", int, 10, False, "minimum coverage of a variant"),
        (("-mv", "--var-fraction"), "var", float, 0.8, False, "minimum variant fraction"),
        (("-c", "--conversion-threshold"), "conversionThreshold", int, 1, False, "T>C conversions that make a read a T>C read"),
        (("-rl", "--max-read-length"), "maxLength", int, None, False, "maximum read length in the BAM"),
        (("-mbq", "--min-base-qual"), "minQual", int, 27, False, "minimum base quality of a counted conversion"),
    ] + _MAP_OPTS + _FILTER_OPTS + _SAMPLE_OPTS),
}


def build_parser():
    parser = ArgumentParser(description="SLAMdunk software for analyzing SLAM-seq data", formatter_class=ArgumentDefaultsHelpFormatter)
    parser.add_argument("--version", action="version", version="%(prog)s " + __version__)
    parser.add_argument("--device", type=int, default=0, help="CUDA device ordinal")
    parser.add_argument("--mismatch-source", default="verify", choices=["verify", "cigar", "mp"], dest="mismatchSource",
                        help="per-read mismatches from: CIGAR walk cross-checked against the MP tags, CIGAR walk only, MP tags only")
    sub = parser.add_subparsers(help="", dest="command")
    for name, spec in SUBCOMMANDS.items():
        sp = sub.add_parser(name, help=spec["help"], formatter_class=ArgumentDefaultsHelpFormatter)
        sp.add_argument(spec["positional"][0], action="store", nargs="+", help=spec["positional"][1])
        for flags, dest, typ, default, required, helptext in spec["options"]:
            if typ == "flag":
                sp.add_argument(*flags, dest=dest, action="store_true", required=False, help=helptext)
            else:
                sp.add_argument(*flags, dest=dest, type=typ, default=default, required=required, help=helptext)
    return parser


def plan_samples(n_samples, rank, world):
    """How `count` / `filter` use several ranks -> (mode, sample indices of this rank):
    "samples": at least one sample per rank -- this rank takes every world-th sample, alone, on its own GPU;
    "tiles":   fewer samples than ranks -- every rank takes part in every sample (tile slices + all-reduce);
    "single":  one process."""
    if world <= 1:
        return "single", list(range(n_samples))
    if n_samples >= world:
        return "samples", list(range(rank, n_samples, world))
    return "tiles", list(range(n_samples))


def run(argv=None):
    parser = build_parser()
    args = parser.parse_args(argv)
    command = args.command
    rank, world, local_rank = 0, 1, 0
    mode = "single"
    if command in ("count", "filter"):
        from . import parallel
        rank, world, local_rank = parallel.env()
        mode, mine = plan_samples(len(args.bam), rank, world)
        if world > 1:
            args.device = local_rank
            if mode == "tiles" and command == "count":
                parallel.init()                      # computeTconversions sees the process group and slices the tiles
            elif mode == "tiles":
                mine = mine if rank == 0 else []     # filter has no multi-rank form: rank 0 does the few samples
        args.bam = [args.bam[i] for i in mine]
        if world > 1 and rank != 0:
            global mainOutput
            mainOutput = open(os.devnull, "w")       # one voice on stderr
    if command in ("count", "filter", "all"):
        from .dunks import filter as gfilter
        from .dunks import tcounter
        tcounter.OPTIONS["device"] = gfilter.OPTIONS["device"] = args.device
        tcounter.OPTIONS["threads"] = gfilter.OPTIONS["threads"] = args.threads
        tcounter.OPTIONS["mismatch_source"] = args.mismatchSource

    if command == "map":
        createDir(args.outputDir)
        n = args.threads
        samples, samplesInfos = getSamples(args.files, runOnly=args.sampleIndex)
        message("Running slamDunk map for " + str(len(samples)) + " files (" + str(n) + " threads)")
        for i in range(0, len(samples)):
            bam = samples[i]
            sampleName = replaceExtension(basename(bam), "", "") if (not args.sampleName or len(samples) > 1) else args.sampleName
            sampleInfo = samplesInfos[i]
            if sampleInfo == "":
                sampleInfo = sampleName + ":" + args.sampleType + ":" + str(args.sampleTime)
            tid = args.sampleIndex if args.sampleIndex > -1 else i
            runMap(tid, bam, args.referenceFile, n, args.trim5, args.maxPolyA, args.quantseq, args.endtoend, args.topn, sampleInfo,
                   args.outputDir, args.skipSAM)
        dunkFinished()
        if not args.skipSAM:
            message("Running slamDunk sam2bam for " + str(len(samples)) + " files (" + str(n) + " threads)")
            for tid in range(0, len(samples)):
                runSam2Bam(tid, samples[tid], n, args.outputDir)
            dunkFinished()
    elif command == "filter":
        createDir(args.outputDir)
        message("Running slamDunk filter for " + str(len(args.bam)) + " files (" + str(args.threads) + " threads)")
        for tid in range(0, len(args.bam)):
            runFilter(tid, args.bam[tid], args.bed, args.mq, args.identity, args.nm, args.outputDir)
        dunkFinished()
    elif command == "snp":
        createDir(args.outputDir)
        n = args.threads
        if n > 1:
            n = int(n / 2)
        message("Running slamDunk SNP for " + str(len(args.bam)) + " files (" + str(n) + " threads)")
        for tid in range(0, len(args.bam)):
            runSnp(tid, args.fasta, args.cov, args.var, 15, args.bam[tid], args.outputDir)
        dunkFinished()
    elif command == "count":
        createDir(args.outputDir)
        snpDirectory = args.snpDir if "snpDir" in args else None
        vcfFile = args.vcfFile if "vcfFile" in args else None
        message("Running slamDunk tcount for " + str(len(args.bam)) + " files (" + str(args.threads) + " threads)")
        for tid in range(0, len(args.bam)):
            runCount(tid, args.bam[tid], args.ref, args.bed, args.maxLength, args.minQual, args.conversionThreshold, args.outputDir,
                     snpDirectory, vcfFile)
        dunkFinished()
    elif command == "all":
        runAll(args)
        dunkFinished()
    else:
        parser.error("Too few arguments.")


if __name__ == "__main__":
    run()
