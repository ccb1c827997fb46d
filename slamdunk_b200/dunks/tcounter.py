"""`slamdunk count` on the GPU: drop-in for the reference's tcounter.computeTconversions
(slamdunk/dunks/tcounter.py:125-330), same positional signature, same three output files.

What happens where
  host   BAM / FASTA decode (C++ threads, libslamdunk_b200.so), BED / VCF parsing with the
         reference's quirks, and every floating-point division + all text formatting, so the files
         are byte-identical to the reference's (str(float) of the same Python expressions);
  device filter predicates (optional), per-read T>C scan, interval join, per-UTR counters and the
         per-position coverage / conversion arrays (sd_count / sd_finalize / sd_tcontent).
"""
from __future__ import print_function

import os
import re
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from .. import _lib as L
from .. import parallel
from ..engine import CountEngine, HostBatch, HostGenome, make_params
from ..utils.bed import BedIterator
from ..utils.snpmask import build_snp_masks
from ..utils.samheader import SlamSeqInfo, getSampleInfo, md5, parse_sam_header, replaceExtension
from ..version import __bam_version__, __count_version__, __version__

HEADER = "\t".join(["Chromosome", "Start", "End", "Name", "Length", "Strand", "ConversionRate", "ReadsCPM", "Tcontent",
                    "CoverageOnTs", "ConversionsOnTs", "ReadCount", "TcReadCount", "multimapCount",
                    "ConversionRateLower", "ConversionRateUpper"])   # SlamSeqFile.py:80


class NegativeStartError(RuntimeError, TypeError):
    """tcounter.py:172-173 means to raise RuntimeError but concatenates str + BedEntry, so the
    reference actually dies with TypeError; this is both."""


def _row(utr, conversionRate, readsCPM, Tcontent, coverageOnTs, conversionsOnTs, readCount, tcReadCount, multimapCount):
    """SlamSeqInterval.__repr__ (slamseq/SlamSeqFile.py:99-100)."""
    return "\t".join([utr.chromosome, str(utr.start), str(utr.stop), utr.name, str(utr.stop - utr.start), utr.strand,
                      str(conversionRate), str(readsCPM), str(Tcontent), str(coverageOnTs), str(conversionsOnTs),
                      str(readCount), str(tcReadCount), str(multimapCount), str(-1.0), str(-1.0)])


class CountTimings(dict):
    pass


# knobs that are not part of the reference signature (set by the CLI / tests)
OPTIONS = {
    "device": 0,
    "threads": 0,                       # host decode threads (0 = all cores)
    "mismatch_source": "verify",        # "cigar" | "mp" | "verify" (CIGAR walk cross-checked against MP)
    "seq_bits": 2,                      # SEQ column handed to sd_count: 2 = 2-bit base codes (half the bytes), 4 = BAM nibble planes
    "wire": True,                       # send the reads over PCIe in the packed form (sd_wire / sd_count_wire); False: the columnar batch
    "decode_on_rank0": True,            # under torchrun: rank 0 decodes the BAM and scatters the wire slices (False: every rank decodes)
    "timings": None,                    # last run's CountTimings
}


_GENOMES = {}   # (path, mtime, size) -> HostGenome: a multi-sample run decodes the FASTA once


def _genome(ref):
    st = os.stat(ref)
    key = (os.path.abspath(ref), st.st_mtime_ns, st.st_size)
    g = _GENOMES.get(key)
    if g is not None:
        return g, True
    for old in list(_GENOMES):      # one genome at a time is kept resident
        _GENOMES.pop(old).close()
    g = HostGenome(ref, OPTIONS["threads"])
    _GENOMES[key] = g
    return g, False


def computeTconversions(ref, bed, snpsFile, bam, maxReadLength, minQual, outputCSV, outputBedgraphPlus,
                        outputBedgraphMinus, conversionThreshold, log, mle=False):
    tm = CountTimings()
    t0 = time.time()
    # multi-GPU: when the caller joined a process group (parallel.init / torchrun), every rank counts a slice of
    # the BAM's tiles, the integer outputs are all-reduced and rank 0 writes the files
    rank, world = 0, 1
    if parallel.dist.is_available() and parallel.dist.is_initialized():
        rank, world = parallel.dist.get_rank(), parallel.dist.get_world_size()

    genome, cached = _genome(ref)
    tm["fasta_decode_s"] = 0.0 if cached else genome.seconds
    mode = {"cigar": L.SD_MISMATCH_CIGAR, "mp": L.SD_MISMATCH_MP, "verify": L.SD_MISMATCH_VERIFY}[OPTIONS["mismatch_source"]]
    # the decoder also writes the packed form that crosses PCIe (sd_wire: 58 instead of 82 bytes per 100 bp read); sd_count_wire
    # expands it on the device.  Under torchrun ONE rank inflates and decodes the file and hands the others their tile slices
    # (parallel.scatter_wire): N ranks decoding the same BAM on the same cores is N times the host work for nothing.
    use_wire = OPTIONS["seq_bits"] == 2 and OPTIONS.get("wire", True)
    scatter = world > 1 and use_wire and OPTIONS.get("decode_on_rank0", True)
    hbatch = None
    if rank == 0 or not scatter:
        hbatch = HostBatch(bam, genome.names, want_mp=(mode != L.SD_MISMATCH_CIGAR), threads=OPTIONS["threads"], group=True,
                           seq2=OPTIONS["seq_bits"] == 2, wire=use_wire)
    info = [dict(header_text=hbatch.header_text, ref_names=hbatch.ref_names, n_reads=hbatch.n_reads, inflate=hbatch.seconds_inflate,
                 decode=hbatch.seconds_decode) if hbatch is not None else None]
    if scatter:
        parallel.dist.broadcast_object_list(info, src=0, group=parallel.host_group())
    info = info[0]
    tm["bam_inflate_s"] = info["inflate"]
    tm["bam_decode_s"] = info["decode"]
    tm["bam_reads"] = info["n_reads"]

    header = parse_sam_header(info["header_text"])
    sampleInfo = getSampleInfo(header)                       # tcounter.py:129
    slamseqInfo = SlamSeqInfo(header)                        # tcounter.py:131
    readNumber = slamseqInfo.FilteredReads                   # tcounter.py:133
    bedMD5 = md5(bed)

    if rank != 0:
        outputCSV = outputBedgraphPlus = outputBedgraphMinus = os.devnull
    fileCSV = open(outputCSV, "w")
    print("#slamdunk v" + __version__, __count_version__, "sample info:", sampleInfo.Name, sampleInfo.ID,
          sampleInfo.Type, sampleInfo.Time, sep="\t", file=fileCSV)
    print("#annotation:", os.path.basename(bed), bedMD5, sep="\t", file=fileCSV)
    print(HEADER, file=fileCSV)
    fileTest = None
    if mle:                                                  # tcounter.py:137-143
        fileTest = open(os.devnull if rank != 0 else replaceExtension(outputCSV, ".tsv", "_perread"), "w")
        print("#slamdunk v" + __version__, __count_version__, "sample info:", sampleInfo.Name, sampleInfo.ID,
              sampleInfo.Type, sampleInfo.Time, sep="\t", file=fileTest)
        print("#annotation:", os.path.basename(bed), bedMD5, sep="\t", file=fileTest)
        print(HEADER, file=fileTest)

    tc_mask, ag_mask = build_snp_masks(snpsFile, genome.names, genome.chrom_off, genome.chrom_len, genome.n_words,
                                       warn=print)          # SNPtools.py:40-51

    bamVersion = "None"                                      # SlamSeqFile.py:409-414
    for pg in header.get("PG", []):
        if pg.get("ID") == "slamdunk":
            bamVersion = pg.get("VN")
    if not bamVersion == __bam_version__:
        raise RuntimeError("Wrong filtered BAM file version detected (" + str(bamVersion) + "). Expected version " +
                           __bam_version__ + ". Please rerun slamdunk filter.")
    if slamseqInfo.AnnotationMD5 != bedMD5:
        print("Warning: MD5 checksum of annotation (" + bedMD5 + ") does not matched MD5 in filtered BAM files (" +
              str(slamseqInfo.AnnotationMD5) + "). Most probably the annotation filed changed after the filtered BAM files were created.",
              file=log)

    # BED rows in file order; a bad row stops the run where the reference would (rows before it are written)
    utrs = []
    pending_error = None
    for utr in BedIterator(bed):
        if not utr.hasStrand():
            pending_error = RuntimeError("Input BED file does not contain stranded intervals.")
            break
        if utr.start < 0:
            pending_error = NegativeStartError("Negativ start coordinate found. Please check the following entry in your BED file: " + repr(utr))
            break
        utrs.append(utr)
    bam_names = set(info["ref_names"])
    n = len(utrs)
    chrom_idx = np.array([genome.index.get(u.chromosome, L.SD_REF_NONE) for u in utrs], dtype=np.uint32)
    start = np.array([u.start for u in utrs], dtype=np.int64)
    stop = np.array([u.stop for u in utrs], dtype=np.int64)
    strand = np.array([1 if u.strand == "-" else 0 for u in utrs], dtype=np.uint8)
    in_bam = np.array([1 if u.chromosome in bam_names else 0 for u in utrs], dtype=np.uint8)
    tm["host_prepare_s"] = time.time() - t0 - tm["fasta_decode_s"] - tm["bam_inflate_s"] - tm["bam_decode_s"]

    t1 = time.time()
    eng = CountEngine(OPTIONS["device"] if world == 1 else parallel.env()[2])
    try:
        eng.set_reference(genome.lo, genome.hi, genome.nmask, genome.chrom_off, genome.chrom_len)
        if tc_mask is not None:
            eng.set_snps(tc_mask, ag_mask)
        eng.set_utrs(chrom_idx, start, stop, strand, in_bam)
        params = make_params(min_qual=minQual, conversion_threshold=conversionThreshold, mismatch_source=mode)
        eng.synchronize()
        t2 = time.time()
        t_sc = time.time()
        keep_alive = None
        if scatter:
            my_batch = None
            my_wire, keep_alive = parallel.scatter_wire(hbatch.wire if hbatch is not None else None, rank, world)
        else:
            my_batch, my_wire = hbatch.batch, hbatch.wire
            if world > 1:
                t0_, t1_ = parallel.shard_range(hbatch.batch.n_tiles, rank, world)
                my_batch = parallel.sub_batch(hbatch.batch, t0_, t1_)
                my_wire = parallel.sub_wire(hbatch.wire, t0_, t1_) if hbatch.wire is not None else None
        tm["scatter_s"] = time.time() - t_sc

        def count_mine(p):
            if my_wire is not None:
                eng.count_wire(my_wire, p)
            else:
                eng.count_host(my_batch, p)
        count_mine(params)
        if world > 1:
            parallel.reduce_engine_outputs(eng, positions=True)
        eng.synchronize()              # the all-reduce runs on eng.stream: every rank must branch on the REDUCED counter
        stats = eng.stats.cpu().numpy()
        if mode == L.SD_MISMATCH_VERIFY and stats[L.SD_S_MP_DISAGREE] > 0:
            # the BAM's MP tags do not describe the same mismatches as SEQ/CIGAR vs this FASTA: the reference
            # trusts MP, so do the same (still on the GPU) and say so
            print("Warning: %d reads carry MP tags that disagree with a CIGAR walk against %s; counting from the MP tags as slamdunk does."
                  % (int(stats[L.SD_S_MP_DISAGREE]), os.path.basename(ref)), file=log)
            eng.reset_outputs()
            params = make_params(min_qual=minQual, conversion_threshold=conversionThreshold, mismatch_source=L.SD_MISMATCH_MP)
            count_mine(params)
            if world > 1:
                parallel.reduce_engine_outputs(eng, positions=True)
        eng.finalize()
        res = eng.results()
        tm["gpu_setup_s"] = t2 - t1
        tm["gpu_count_s"] = time.time() - t2
        tm["gpu_launches"] = eng.launch_count()
        tm["stats"] = res["stats"].tolist()
        pos_off, pos_len = eng.pos_off, eng.pos_len
    finally:
        eng.close()

    t3 = time.time()
    counters = res["counters"]
    tcontent = res["tcontent"]
    pair_lo = pair_hi = pairs = None
    if mle:
        # the tInReads / tcInRead lists (tcounter.py:233-243) need every mismatch of a read, not only its conversions:
        # a second scan (sd_read_stats) over the BAM decoded with all MP entries, one (read, UTR) pair per list element
        from . import stats as _stats
        rs = _stats.ReadStats(ref, bam, snpsFile=snpsFile, bed=None)
        rs.utrs, rs.has_bed = utrs, True
        try:
            pairs = rs.run(minQual, conversion_threshold=conversionThreshold, strict=True, want=("pairs",),
                           pair_capacity=2 * info["n_reads"] + 1024)["pairs"]
        finally:
            rs.close()
        pair_lo = np.searchsorted(pairs["row"], np.arange(n), side="left")
        pair_hi = np.searchsorted(pairs["row"], np.arange(n), side="right")
    # the table as text in one native call (sd_format_tcount_rows: the same str() forms, tcounter.py:251-307); row by row in
    # Python for the mle lists and for a FilteredReads value that is not a plain int a double holds exactly
    native_rows = not mle and type(readNumber) is int and abs(readNumber) < 2 ** 53
    if native_rows:
        fileCSV.write(format_rows(utrs, counters, tcontent, readNumber))
    counter_rows = [] if native_rows else counters[:n].tolist()      # Python ints once, not one numpy scalar per cell
    tcontent_rows = [] if native_rows else tcontent[:n].tolist()
    for i, utr in enumerate([] if native_rows else utrs):
        row = counter_rows[i]
        readCount = row[L.SD_C_READS]
        Tcontent = tcontent_rows[i]
        if readCount > 0:                                    # tcounter.py:251
            coverageOnTs = row[L.SD_C_COV_ON_T]
            conversionsOnTs = row[L.SD_C_CONV_ON_T]
            readsCPM = 0
            if readNumber > 0:                               # tcounter.py:295-297
                readsCPM = readCount * 1000000.0 / readNumber
            conversionRate = 0
            if coverageOnTs > 0:                             # tcounter.py:301-303
                conversionRate = float(conversionsOnTs) / float(coverageOnTs)
            print(_row(utr, conversionRate, readsCPM, Tcontent, coverageOnTs, conversionsOnTs, readCount,
                       row[L.SD_C_TCREADS], row[L.SD_C_MULTIMAP]), file=fileCSV)
            if mle:                                          # tcounter.py:305: the two lists in the ReadCount / TcReadCount columns
                p = pairs[pair_lo[i]:pair_hi[i]]
                print(_row(utr, conversionRate, readsCPM, Tcontent, coverageOnTs, conversionsOnTs,
                           ",".join(str(int(x)) for x in p["t_count"]), ",".join(str(int(x)) for x in p["tc_count"]),
                           row[L.SD_C_MULTIMAP]), file=fileTest)
        else:
            print(_row(utr, 0, 0, Tcontent, 0, 0, 0, 0, 0), file=fileCSV)
            if mle:                                          # tcounter.py:168: the default MLE row never gets its Tcontent
                print(_row(utr, 0, 0, 0, 0, 0, 0, 0, 0), file=fileTest)
    fileCSV.close()
    if mle:
        fileTest.close()
    if pending_error is not None:
        raise pending_error

    _write_bedgraphs(utrs, chrom_idx, genome, res, pos_off, pos_len, counters, outputBedgraphPlus, outputBedgraphMinus)
    if mle:                                                  # tcounter.py:328-330 shells out to an R script that is not part of this package
        print("R plotters are not part of slamdunk_b200: the MLE fit over " + replaceExtension(outputCSV, ".tsv", "_perread") + " was not run.", file=log)
    tm["write_s"] = time.time() - t3
    tm["total_s"] = time.time() - t0
    OPTIONS["timings"] = tm
    if hbatch is not None:
        hbatch.close()


def format_rows(utrs, counters, tcontent, readNumber):
    """The per-UTR rows of the tcount table as one string: what `print(_row(...), file=fileCSV)` writes row by row."""
    import ctypes
    lib = L.load()
    n = len(utrs)
    if n == 0:
        return ""
    chroms = (ctypes.c_char_p * n)(*[u.chromosome.encode() for u in utrs])
    names = (ctypes.c_char_p * n)(*[u.name.encode() for u in utrs])
    start = np.array([u.start for u in utrs], dtype=np.int64)
    stop = np.array([u.stop for u in utrs], dtype=np.int64)
    strand = np.array([1 if u.strand == "-" else 0 for u in utrs], dtype=np.uint8)
    c = np.ascontiguousarray(np.asarray(counters)[:n], dtype=np.int64)
    t = np.ascontiguousarray(np.asarray(tcontent)[:n], dtype=np.int64)
    text = ctypes.c_void_p()
    length = ctypes.c_uint64(0)
    rc = lib.sd_format_tcount_rows(n, chroms, start.ctypes.data, stop.ctypes.data, names, strand.ctypes.data, c.ctypes.data, t.ctypes.data,
                                   int(readNumber), ctypes.byref(text), ctypes.byref(length))
    if rc != L.SD_OK:
        raise L.NativeError("sd_format_tcount_rows failed with code %d" % rc)
    try:
        return ctypes.string_at(text, length.value).decode()
    finally:
        lib.sd_free_text(text)


def _write_bedgraphs(utrs, chrom_idx, genome, res, pos_off, pos_len, counters, out_plus, out_minus):
    """tcounter.py:277-285 + 315-326: one line per covered T (A on '-') position of every UTR that had reads, in
    dict insertion order (BED order, ascending position, a position keeps the place of its first insertion).
    Written by the native writer (sd_write_bedgraphs), which prints rates like Python's repr(float)."""
    import ctypes
    lib = L.load()
    n = len(utrs)
    names = (ctypes.c_char_p * max(1, n))(*[u.chromosome.encode() for u in utrs])
    first = np.array([max(0, u.start) for u in utrs], dtype=np.int64)
    strand = np.array([1 if u.strand == "-" else 0 for u in utrs], dtype=np.uint8)
    has = np.ascontiguousarray(counters[:n, L.SD_C_READS] > 0, dtype=np.uint8) if n else np.zeros(0, np.uint8)
    known = np.asarray(chrom_idx[:n]) != L.SD_REF_NONE
    off = np.asarray(genome.chrom_off, dtype=np.int64)
    gstart = np.where(known, off[np.where(known, chrom_idx[:n], 0).astype(np.int64)] + first, 0).astype(np.uint64) if n else np.zeros(0, np.uint64)
    po = np.ascontiguousarray(pos_off, dtype=np.uint64)
    pl = np.ascontiguousarray(pos_len, dtype=np.uint32)
    planes = [np.ascontiguousarray(a, dtype=np.uint32) for a in (genome.lo, genome.hi, genome.nmask)]    # views, no copies
    cov = np.ascontiguousarray(res["pos_cov"], dtype=np.int32)
    tc = np.ascontiguousarray(res["pos_tc"], dtype=np.int32)
    ptr = lambda a: ctypes.c_void_p(a.ctypes.data)    # noqa: E731
    rc = lib.sd_write_bedgraphs(os.fsencode(out_plus), os.fsencode(out_minus), n, names, ptr(first), ptr(strand), ptr(has), ptr(po),
                                ptr(pl), ptr(gstart), ptr(planes[0]), ptr(planes[1]), ptr(planes[2]), ptr(cov), ptr(tc), len(cov), int(OPTIONS["threads"]))
    if rc != L.SD_OK:
        raise L.NativeError("sd_write_bedgraphs failed with code %d" % rc)


# ---------------------------------------------------------------------------------------------------------------
# alleyoop positional-tracks
# ---------------------------------------------------------------------------------------------------------------
TRACKS = [  # (file suffix, track name suffix, description, strand, SD_TRACK_*) in the order tcounter.py:349-365 opens / heads them
    ("_TC_rates_genomewide.bedGraph", " tc-conversions", "# T->C conversions / # reads on T per position genome-wide", 0, L.SD_TRACK_RATE),
    ("_AG_rates_genomewide.bedGraph", " ag-conversions", "# A->G conversions / # reads on A per position genome-wide", 1, L.SD_TRACK_RATE),
    ("_coverage_plus_genomewide.bedGraph", " plus-strand coverage", "# Reads on plus strand genome-wide", 0, L.SD_TRACK_COVERAGE),
    ("_coverage_minus_genomewide.bedGraph", " minus-strand coverage", "# Reads on minus strand genome-wide", 1, L.SD_TRACK_COVERAGE),
    ("_TC_conversions_genomewide.bedGraph", " T->C conversions", "# T->C conversions on plus strand genome-wide", 0, L.SD_TRACK_CONVERSIONS),
    ("_AG_conversions_genomewide.bedGraph", " A->G conversions", "# A->G conversions on minus strand genome-wide", 1, L.SD_TRACK_CONVERSIONS),
    ("_coverage_T_genomewide.bedGraph", " T-coverage", "# Plus-strand reads on Ts genome-wide", 0, L.SD_TRACK_COVERAGE_ON_BASE),
    ("_coverage_A_genomewide.bedGraph", " A-coverage", "# Minus-strand reads on As genome-wide", 1, L.SD_TRACK_COVERAGE_ON_BASE),
]


def _natural_keys(text):
    """SlamSeqBamFile.natural_keys (slamseq/SlamSeqFile.py:458-464): chr2 before chr10."""
    return [int(c) if c.isdigit() else c for c in re.split(r"(\d+)", text)]


def genomewideConversionRates(referenceFile, snpsFile, bam, minBaseQual, outputBedGraphPrefix, conversionThreshold,
                              coverageCutoff, log):
    """Drop-in for the reference's tcounter.genomewideConversionRates (dunks/tcounter.py:332-487, behind
    `alleyoop positional-tracks`): eight run-length bedGraphs per BAM -- coverage, coverage on T / A, conversions and
    conversion rates on both strands, genome-wide.

    Device: the same per-read scan as `count`, with one annotation row per chromosome strand that starts at position 1
    (the reference builds its iterator with startPosition 1, SlamSeqFile.py:451, so index p of its arrays is the 0-based
    genomic position p + 1 and the chromosome's first base is never counted), then sd_track_runs compacts every track
    into its change points.  Host: BAM / FASTA decode and the text of the change points (sd_write_track)."""
    import ctypes
    tm = CountTimings()
    t0 = time.time()
    genome, cached = _genome(referenceFile)
    tm["fasta_decode_s"] = 0.0 if cached else genome.seconds
    mode = {"cigar": L.SD_MISMATCH_CIGAR, "mp": L.SD_MISMATCH_MP, "verify": L.SD_MISMATCH_VERIFY}[OPTIONS["mismatch_source"]]
    # the decoder also writes the packed form that crosses PCIe (sd_wire: 58 instead of 82 bytes per 100 bp read); sd_count_wire
    # expands it on the device
    hbatch = HostBatch(bam, genome.names, want_mp=(mode != L.SD_MISMATCH_CIGAR), threads=OPTIONS["threads"], group=True,
                       seq2=OPTIONS["seq_bits"] == 2, wire=OPTIONS["seq_bits"] == 2 and OPTIONS.get("wire", True))
    tm["bam_inflate_s"], tm["bam_decode_s"], tm["bam_reads"] = hbatch.seconds_inflate, hbatch.seconds_decode, hbatch.n_reads
    tc_mask, ag_mask = build_snp_masks(snpsFile, genome.names, genome.chrom_off, genome.chrom_len, genome.n_words, warn=print)

    chromosomes = sorted(genome.names, key=_natural_keys)              # getChromosomes, SlamSeqFile.py:469-472
    bedGraphInfo = re.sub("_slamdunk_mapped.*", "", os.path.basename(outputBedGraphPrefix))
    print(bedGraphInfo)                                                # tcounter.py:347
    paths = []
    for suffix, name, desc, _strand, _kind in TRACKS:
        paths.append(outputBedGraphPrefix + suffix)
        with open(paths[-1], "w") as fh:
            print("track type=bedGraph name=\"" + bedGraphInfo + name + "\" description=\"" + desc + "\"", file=fh)

    # rows 2c (+) and 2c + 1 (-) for FASTA chromosome c: [1, chrLen + 1) clipped to the chromosome by sd_set_utrs
    chrom_idx, start, stop, strand, in_bam = _chromosome_rows(genome, hbatch)
    lengths = np.asarray(genome.chrom_len, dtype=np.int64)
    tm["host_prepare_s"] = time.time() - t0 - tm["fasta_decode_s"] - tm["bam_inflate_s"] - tm["bam_decode_s"]

    t1 = time.time()
    lib = L.load()
    eng = CountEngine(OPTIONS["device"])
    try:
        eng.set_reference(genome.lo, genome.hi, genome.nmask, genome.chrom_off, genome.chrom_len)
        if tc_mask is not None:
            eng.set_snps(tc_mask, ag_mask)
        eng.set_utrs(chrom_idx, start, stop, strand, in_bam)
        params = make_params(min_qual=minBaseQual, conversion_threshold=conversionThreshold, mismatch_source=mode,
                             flags=L.SD_PARAM_MP_ANY_BASE)
        eng.synchronize()
        t2 = time.time()
        eng.count_host(hbatch.batch, params)
        stats = eng.stats.cpu().numpy()
        if mode == L.SD_MISMATCH_VERIFY and stats[L.SD_S_MP_DISAGREE] > 0:
            print("Warning: %d reads carry MP tags that disagree with a CIGAR walk against %s; taking the conversions from the MP tags as slamdunk does."
                  % (int(stats[L.SD_S_MP_DISAGREE]), os.path.basename(referenceFile)), file=log)
            eng.reset_outputs()
            params = make_params(min_qual=minBaseQual, conversion_threshold=conversionThreshold, mismatch_source=L.SD_MISMATCH_MP,
                                 flags=L.SD_PARAM_MP_ANY_BASE)
            eng.count_host(hbatch.batch, params)
        eng.finalize()
        tm["gpu_setup_s"] = t2 - t1
        tm["gpu_count_s"] = time.time() - t2
        t3 = time.time()
        n_lines = 0

        def write_runs(path, chromosome, kind, pos, val, isf):     # the arrays stay referenced until the text is out
            rc = lib.sd_write_track(os.fsencode(path), chromosome.encode(), kind, len(pos), pos.ctypes.data, val.ctypes.data,
                                    isf.ctypes.data)
            if rc != L.SD_OK:
                raise IOError("sd_write_track failed with code %d on %s" % (rc, path))

        # one writer thread per track file: a file's chromosomes are appended in order, the eight files side by side
        # and under the next compaction on the device (the native writer runs without the GIL)
        writers = [ThreadPoolExecutor(max_workers=1) for _ in paths]
        pending = []
        try:
            for chromosome in chromosomes:
                c = genome.index[chromosome]
                for k, (path, (_suffix, _name, _desc, s, kind)) in enumerate(zip(paths, TRACKS)):
                    pos, val, isf = eng.track_runs(2 * c + s, kind, int(coverageCutoff), int(lengths[c]))
                    if len(pos) == 0:
                        continue
                    pending.append(writers[k].submit(write_runs, path, chromosome, kind, pos, val, isf))
                    n_lines += len(pos)
        finally:
            for w in writers:
                w.shutdown(wait=True)
        for f in pending:
            f.result()                                             # re-raises a writer's IOError
        tm["tracks_s"] = time.time() - t3
        tm["track_lines"] = n_lines
        tm["gpu_launches"] = eng.launch_count()
    finally:
        eng.close()
    tm["total_s"] = time.time() - t0
    OPTIONS["timings"] = tm
    if hbatch is not None:
        hbatch.close()


def _chromosome_rows(genome, hbatch):
    """Annotation rows 2c (+) and 2c + 1 (-) for FASTA chromosome c, [1, chrLen + 1) as the reference's whole-chromosome
    iterator sees it (readsInChromosome, SlamSeqFile.py:447-453)."""
    n_chrom = len(genome.names)
    bam_names = set(hbatch.ref_names)
    lengths = np.asarray(genome.chrom_len, dtype=np.int64)
    return (np.repeat(np.arange(n_chrom, dtype=np.uint32), 2), np.ones(2 * n_chrom, dtype=np.int64), np.repeat(lengths + 1, 2),
            np.tile(np.array([0, 1], dtype=np.uint8), n_chrom),
            np.repeat(np.array([1 if nm in bam_names else 0 for nm in genome.names], dtype=np.uint8), 2))


def genomewideReadSeparation(referenceFile, snpsFile, bam, minBaseQual, outputBAMPrefix, conversionThreshold, log):
    """Drop-in for the reference's tcounter.genomewideReadSeparation (dunks/tcounter.py:488-527, behind
    `alleyoop read-separator`): <prefix>_TCReads.bam holds every placed record whose query NAME belongs to a read with at
    least conversionThreshold T>C conversions in one of its alignments, <prefix>_backgroundReads.bam the other placed
    records; input order and header are kept.

    Device: the per-read conversion count of the count kernel (sd_count_reads).  Host: decode, the name set (64-bit name
    hashes from the decoder), the two BAMs and their BAI indexes (sd_rewrite_bam_indexed; pysamIndex, tcounter.py:526-527)."""
    import ctypes

    import torch
    from ..engine import DeviceBatch
    tm = CountTimings()
    t0 = time.time()
    genome, cached = _genome(referenceFile)
    mode = {"cigar": L.SD_MISMATCH_CIGAR, "mp": L.SD_MISMATCH_MP, "verify": L.SD_MISMATCH_VERIFY}[OPTIONS["mismatch_source"]]
    # the decoder also writes the packed form that crosses PCIe (sd_wire: 58 instead of 82 bytes per 100 bp read); sd_count_wire
    # expands it on the device
    hbatch = HostBatch(bam, genome.names, want_mp=(mode != L.SD_MISMATCH_CIGAR), threads=OPTIONS["threads"], group=True,
                       seq2=OPTIONS["seq_bits"] == 2, wire=OPTIONS["seq_bits"] == 2 and OPTIONS.get("wire", True))
    tm["bam_inflate_s"], tm["bam_decode_s"], tm["bam_reads"] = hbatch.seconds_inflate, hbatch.seconds_decode, hbatch.n_reads
    tc_mask, ag_mask = build_snp_masks(snpsFile, genome.names, genome.chrom_off, genome.chrom_len, genome.n_words, warn=print)
    n = hbatch.n_reads
    t1 = time.time()
    eng = CountEngine(OPTIONS["device"])
    try:
        eng.set_reference(genome.lo, genome.hi, genome.nmask, genome.chrom_off, genome.chrom_len)
        if tc_mask is not None:
            eng.set_snps(tc_mask, ag_mask)
        eng.set_utrs(*_chromosome_rows(genome, hbatch))
        db = DeviceBatch.from_host(hbatch.batch, eng.device, eng)
        with torch.cuda.stream(eng.stream):
            read_tc = torch.zeros(max(1, hbatch.batch.n_tiles * 32), dtype=torch.uint8, device=eng.device)
        eng.synchronize()
        # the threshold is applied here on the raw count (the kernel zeroes counts below its threshold, and 0 >= 0 must
        # still make a TC read for conversionThreshold 0): ask the kernel for every conversion
        params = make_params(min_qual=minBaseQual, conversion_threshold=0, mismatch_source=mode)
        eng.count(db, params, read_tc)
        eng.synchronize()              # sd_count_reads only queues the kernel on eng.stream; .cpu() copies on another stream
        stats = eng.stats.cpu().numpy()
        if mode == L.SD_MISMATCH_VERIFY and stats[L.SD_S_MP_DISAGREE] > 0:
            print("Warning: %d reads carry MP tags that disagree with a CIGAR walk against %s; taking the conversions from the MP tags as slamdunk does."
                  % (int(stats[L.SD_S_MP_DISAGREE]), os.path.basename(referenceFile)), file=log)
            eng.reset_outputs()
            eng.count(db, make_params(min_qual=minBaseQual, conversion_threshold=0, mismatch_source=L.SD_MISMATCH_MP), read_tc)
        eng.synchronize()
        tc_slot = read_tc.cpu().numpy()[:n].astype(np.int64)
        tm["gpu_s"] = time.time() - t1
    finally:
        eng.close()
    rec = hbatch.rec_array()[:n]
    flags = rec[:, 1] >> 24
    seen = ((rec[:, 1] & 0xFFFF) != L.SD_REF_NONE) & ((flags & (L.SD_F_PAD | L.SD_F_UNMAPPED)) == 0)   # chromosome in the FASTA
    placed_slot = rec[:, 0].view(np.int32) >= 0                                  # samFile.fetch(): records with a position
    is_tc_slot = seen & (tc_slot >= int(conversionThreshold))                    # isTcRead, SlamSeqFile.py:397-400
    order = hbatch.order().astype(np.int64)                                      # slot -> record index in the file
    name_hash = hbatch.name_hash()                                               # file order
    tc_names = np.unique(name_hash[order[is_tc_slot]])                           # tcReadDict, tcounter.py:508-516
    placed = np.zeros(n, dtype=bool)
    placed[order] = placed_slot
    in_tc = np.isin(name_hash, tc_names)
    text = hbatch.header_text
    hbatch.close()
    lib = L.load()
    err = ctypes.create_string_buffer(512)
    written = {}
    for suffix, keep in (("_backgroundReads.bam", placed & ~in_tc), ("_TCReads.bam", placed & in_tc)):
        k8 = np.ascontiguousarray(keep, dtype=np.uint8)
        nw = ctypes.c_uint64(0)
        rc = lib.sd_rewrite_bam_indexed(os.fsencode(bam), os.fsencode(outputBAMPrefix + suffix), os.fsencode(outputBAMPrefix + suffix + ".bai"),
                                        text.encode(), k8.ctypes.data, None, n, 0, 6,
                                OPTIONS["threads"], ctypes.byref(nw), err, len(err))
        if rc != L.SD_OK:
            raise IOError("sd_rewrite_bam_indexed: " + err.value.decode("utf-8", "replace"))
        written[suffix] = int(nw.value)
    tm["written"] = written
    tm["total_s"] = time.time() - t0
    OPTIONS["timings"] = tm


def collapse(expandedCSV, collapsedCSV, log):
    """Drop-in for the reference's tcounter.collapse (dunks/tcounter.py:39-111, `alleyoop collapse`): sum the tcount rows
    of every gene name (column 4) and recompute CPM and conversion rate.  Host-side table work on the count path's output;
    same column order, same float text (readsCPM is readCount / total reads * 1e6 as three float operations)."""
    genes = {}
    readNumber = 0
    with open(expandedCSV, "r") as f:
        for line in f:
            if line.startswith("#") or line.startswith("Chromosome"):
                continue
            fields = line.rstrip().split("\t")
            if len(fields) < 14:                                        # tcounter.py:86-87
                print("Error in TC file format - unexpected number of fields (" + str(len(fields)) + ") in the following line:\n" + line,
                      file=log)
                continue
            vals = [max(int(fields[k]), 0) for k in (4, 8, 9, 10, 11, 12, 13)]   # length, Tcontent, cov, conv, reads, tcReads, multimap
            acc = genes.get(fields[3])
            if acc is None:
                genes[fields[3]] = vals
            else:
                for k in range(7):
                    acc[k] += vals[k]
            readNumber += int(fields[11])                               # not clamped (tcounter.py:84)
    with open(collapsedCSV, "w") as out:
        print("gene_name", "length", "readsCPM", "conversionRate", "Tcontent", "coverageOnTs", "conversionsOnTs", "readCount",
              "tcReadCount", "multimapCount", sep="\t", file=out)
        for gene in sorted(genes.keys()):
            length, tcontent, cov, conv, reads, tcreads, multimap = genes[gene]
            conversionRate = 0
            if cov > 0:
                conversionRate = float(conv) / cov
            print(gene, length, float(reads) / float(readNumber) * 1000000, conversionRate, tcontent, cov, conv, reads, tcreads,
                  multimap, sep="\t", file=out)
