"""`alleyoop rates / tccontext / utrrates / tcperreadpos / tcperutrpos / snpeval` on the GPU: drop-ins for the reference's
dunks/stats.py functions that iterate SlamSeqBamIterator, same signatures, same CSV files.

    statsComputeOverallRates        stats.py:85-131
    statsComputeTCContext           stats.py:226-330
    statsComputeOverallRatesPerUTR  stats.py:331-528
    tcPerReadPos                    stats.py:591-657
    tcPerUtr                        stats.py:659-772
    computeSNPMaskedRates           stats.py:774-849

host    BAM / FASTA decode (C++ threads; the BAM with SD_DECODE_MP_ALL: every MP entry, not only T>C / A>G), BED / VCF
        parsing, the per-UTR percentages + medians of utrrates and all text;
device  sd_read_stats (csrc/sd_stats.cu): per-read mismatch list with base-quality and SNP tests, the 5 x 5 conversion
        matrix of every read, the interval join and all integer accumulation.
The PDF step of each tool shells out to the reference's R plotters, which are not part of this package: when the
output is due a line in the log says so and no PDF is written.  No CPU fallback: without the library or a GPU the
functions raise.
"""
from __future__ import print_function

import ctypes
import os
import warnings

import numpy as np
import torch

from .. import _lib as L
from ..engine import CountEngine, DeviceBatch, HostBatch
from ..utils.bed import BedIterator
from ..utils.samheader import SlamSeqInfo, getSampleInfo, parse_sam_header
from ..utils.snpmask import build_snp_masks
from ..version import __version__
from . import tcounter as _tc

utrNormFactor = 250                               # stats.py:35
baseNumber = 5
toBase = ["A", "C", "G", "T", "N"]

TIMINGS = {}                                      # last sd_read_stats run: kernel_ms, reads, device_bytes, decode seconds

PAIR_DT = np.dtype([("row", np.uint32), ("read", np.uint32), ("t_count", np.uint32), ("tc_count", np.uint32)])


def checkStep(inFiles, outFiles, force=False):
    """utils/misc.py:147-164"""
    if not all(os.path.exists(f) for f in inFiles):
        raise RuntimeError("One or more input files don't exist: " + str(inFiles))
    inFileDate = max(os.path.getmtime(x) for x in inFiles)
    if len(outFiles) > 0 and all(os.path.exists(f) for f in outFiles):
        outFileDate = min(os.path.getmtime(x) for x in outFiles)
        if outFileDate > inFileDate:
            return force is True
    return True


def _no_plot(what, outputPDF, log):
    print("R plotters are not part of slamdunk_b200: " + what + " plot " + str(outputPDF) + " not written.", file=log)


class ReadStats(object):
    """One sd_read_stats run over a BAM: decode, upload, scan, results as host numpy arrays."""

    def __init__(self, referenceFile, bam, snpsFile=None, bed=None, require_strand=False):
        self.genome, _cached = _tc._genome(referenceFile)
        self.hbatch = HostBatch(bam, self.genome.names, want_mp=True, threads=_tc.OPTIONS["threads"], mp_all=True)
        self.header = parse_sam_header(self.hbatch.header_text)
        self.utrs = []
        self.has_bed = bed is not None
        self.pending_error = None                  # a bad BED row stops the run where the reference would (rows before it are written)
        self.snp_masks = (None, None)
        if snpsFile is not None:
            g = self.genome
            self.snp_masks = build_snp_masks(snpsFile, g.names, g.chrom_off, g.chrom_len, g.n_words, warn=print)
        if bed is not None:
            for utr in BedIterator(bed):
                if require_strand and not utr.hasStrand():
                    self.pending_error = RuntimeError("Input BED file does not contain stranded intervals.")
                    break
                if utr.start < 0:
                    self.pending_error = _tc.NegativeStartError("Negativ start coordinate found. Please check the following entry in your BED file: " + repr(utr))
                    break
                self.utrs.append(utr)

    def run(self, min_qual, conversion_threshold=1, strict=False, want=(), max_read_length=0, utr_norm=utrNormFactor,
            pair_capacity=0):
        """want: subset of {"rates", "readpos", "utr_rates", "utr_snp", "utrpos", "context", "pairs"} -> dict of numpy arrays"""
        g, hb, utrs = self.genome, self.hbatch, self.utrs
        n = len(utrs)
        eng = CountEngine(_tc.OPTIONS["device"])
        try:
            eng.set_reference(g.lo, g.hi, g.nmask, g.chrom_off, g.chrom_len)
            if self.snp_masks[0] is not None:
                eng.set_snps(*self.snp_masks)
            if self.has_bed:
                bam_names = set(hb.ref_names)
                eng.set_utrs(np.array([g.index.get(u.chromosome, L.SD_REF_NONE) for u in utrs], dtype=np.uint32),
                             np.array([u.start for u in utrs], dtype=np.int64), np.array([u.stop for u in utrs], dtype=np.int64),
                             np.array([{"+": 0, "-": 1}.get(u.strand, 2) for u in utrs], dtype=np.uint8),
                             np.array([1 if u.chromosome in bam_names else 0 for u in utrs], dtype=np.uint8))
            dev = eng.device
            with torch.cuda.stream(eng.stream):
                dbatch = DeviceBatch.from_host(hb.batch, dev, eng)
                order = torch.from_numpy(hb.order().astype(np.int32)).to(dev) if hb.n_reads else None
                shapes = {"context": 20, "rates": 2 * L.SD_NRATE, "readpos": 6 * (max_read_length + 1), "utr_rates": max(1, n) * (L.SD_NRATE + 1),
                          "utr_snp": max(1, n) * 3, "utrpos": 6 * (utr_norm + 1)}
                out = dict((k, torch.zeros(shapes[k], dtype=torch.int64, device=dev)) for k in shapes if k in want)
                pairs = n_pairs = None
                if "pairs" in want:
                    pairs = torch.zeros(max(1, pair_capacity) * 4, dtype=torch.int32, device=dev)
                    n_pairs = torch.zeros(1, dtype=torch.int64, device=dev)
            eng.stream.synchronize()
            a = L.StatsArgs()
            a.min_qual, a.conversion_threshold, a.strict = int(min_qual), int(conversion_threshold), 1 if strict else 0
            a.max_read_length, a.utr_norm = int(max_read_length), int(utr_norm)
            a.d_order = order.data_ptr() if order is not None else None
            a.d_rates, a.d_readpos = _ptr(out.get("rates")), _ptr(out.get("readpos"))
            a.d_utr_rates, a.d_utr_snp, a.d_utrpos = _ptr(out.get("utr_rates")), _ptr(out.get("utr_snp")), _ptr(out.get("utrpos"))
            a.d_context = _ptr(out.get("context"))
            a.d_pairs, a.pair_capacity, a.d_n_pairs = _ptr(pairs), int(pair_capacity), _ptr(n_pairs)
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(eng.stream)
            eng._chk(eng.lib.sd_read_stats(eng.ctx, ctypes.byref(dbatch.batch), ctypes.byref(a)), "sd_read_stats")
            ev1.record(eng.stream)
            eng.synchronize()
            TIMINGS.update(kernel_ms=ev0.elapsed_time(ev1), reads=hb.n_reads, device_bytes=dbatch.nbytes(),
                           bam_inflate_s=hb.seconds_inflate, bam_decode_s=hb.seconds_decode)
            res = dict((k, v.cpu().numpy()) for k, v in out.items())
            if pairs is not None:
                found = int(n_pairs.item())
                if found > pair_capacity:
                    raise L.NativeError("sd_read_stats found %d (read, UTR) pairs, capacity was %d" % (found, pair_capacity))
                p = pairs.cpu().numpy().view(PAIR_DT)[:found]
                res["pairs"] = p[np.lexsort((p["read"], p["row"]))]
            res["gpu_launches"] = eng.launch_count()
            return res
        finally:
            eng.close()

    def close(self):
        self.hbatch.close()


def _ptr(t):
    return t.data_ptr() if t is not None else None


def read_stats_on_device(eng, dbatch, n_utr, min_qual, want, conversion_threshold=1, strict=False, max_read_length=0,
                         utr_norm=utrNormFactor, repeat=1):
    """sd_read_stats on a batch that is already in HBM (MP column in the SD_DECODE_MP_ALL form; reference, SNPs and
    annotation installed in `eng`) -> (dict of host numpy arrays, kernel ms of the last of `repeat` runs)."""
    dev = eng.device
    shapes = {"context": 20, "rates": 2 * L.SD_NRATE, "readpos": 6 * (max_read_length + 1), "utr_rates": max(1, n_utr) * (L.SD_NRATE + 1),
              "utr_snp": max(1, n_utr) * 3, "utrpos": 6 * (utr_norm + 1)}
    a = L.StatsArgs()
    a.min_qual, a.conversion_threshold, a.strict = int(min_qual), int(conversion_threshold), 1 if strict else 0
    a.max_read_length, a.utr_norm = int(max_read_length), int(utr_norm)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(repeat):
        with torch.cuda.stream(eng.stream):
            out = dict((k, torch.zeros(shapes[k], dtype=torch.int64, device=dev)) for k in shapes if k in want)
        a.d_rates, a.d_readpos = _ptr(out.get("rates")), _ptr(out.get("readpos"))
        a.d_utr_rates, a.d_utr_snp, a.d_utrpos = _ptr(out.get("utr_rates")), _ptr(out.get("utr_snp")), _ptr(out.get("utrpos"))
        a.d_context = _ptr(out.get("context"))
        ev0.record(eng.stream)
        eng._chk(eng.lib.sd_read_stats(eng.ctx, ctypes.byref(dbatch.batch), ctypes.byref(a)), "sd_read_stats")
        ev1.record(eng.stream)
        eng.synchronize()
    return dict((k, v.cpu().numpy()) for k, v in out.items()), ev0.elapsed_time(ev1)


def printRates(ratesFwd, ratesRev, f):
    """stats.py:74-83"""
    print("\t", end="", file=f)
    for i in range(0, 5):
        print(toBase[i] + "\t" + toBase[i].lower() + "\t", end="", file=f)
    print(file=f)
    for i in range(0, 5):
        print(toBase[i] + "\t", end="", file=f)
        for j in range(0, 5):
            print(str(ratesFwd[i * 5 + j]) + "\t" + str(ratesRev[i * 5 + j]) + "\t", end="", file=f)
        print(file=f)


def statsComputeOverallRates(referenceFile, bam, minBaseQual, outputCSV, outputPDF, log, printOnly=False, verbose=True, force=False):
    """stats.py:85-131 (alleyoop rates): summed 5 x 5 conversion matrices of the forward and of the reverse reads."""
    if not checkStep([bam, referenceFile], [outputCSV], force):
        print("Skipped computing overall rates for file " + bam, file=log)
    else:
        rs = ReadStats(referenceFile, bam)
        try:
            r = rs.run(minBaseQual, want=("rates",))["rates"].reshape(2, 25)
        finally:
            rs.close()
        fo = open(outputCSV, "w")
        print("# slamdunk rates v" + __version__, file=fo)
        printRates([int(x) for x in r[0]], [int(x) for x in r[1]], fo)
        fo.close()
    if not checkStep([bam, referenceFile], [outputPDF], force):
        print("Skipped computing overall rate pdfs for file " + bam, file=log)
    else:
        _no_plot("overall rates", outputPDF, log)


def statsComputeTCContext(referenceFile, bam, minQual, outputCSV, outputPDF, log, printOnly=False, verbose=True, force=False):
    """stats.py:226-330 (alleyoop tccontext): the bases next to every T of the forward reads / every A of the reverse reads."""
    if not checkStep([bam, referenceFile], [outputCSV], force):
        print("Skipped computing overall rates for file " + bam, file=log)
    else:
        frontCombinations = ["AT", "CT", "GT", "TT", "NT"]
        backCombinations = ["TA", "TC", "TG", "TT", "TN"]
        rs = ReadStats(referenceFile, bam)
        try:
            for chromosome in rs.genome.names:          # bamFile.fetch(region=chromosome) for every FASTA chromosome (stats.py:259)
                if chromosome not in rs.hbatch.ref_names:
                    raise ValueError("invalid contig `%s`" % chromosome)
            c = rs.run(minQual, want=("context",))["context"].reshape(2, 2, 5)
        finally:
            rs.close()
        fo = open(outputCSV, "w")
        print("\t".join(frontCombinations), file=fo)
        print("\t".join(str(int(x)) for x in c[0, 0]), file=fo)
        print("\t".join(str(int(x)) for x in c[0, 1]), file=fo)
        print("\t".join(backCombinations), file=fo)
        print("\t".join(str(int(x)) for x in c[1, 0]), file=fo)
        print("\t".join(str(int(x)) for x in c[1, 1]), file=fo)
        fo.close()
    if not checkStep([bam, referenceFile], [outputPDF], force):
        print("Skipped computing overall rate pdfs for file " + bam, file=log)
    else:
        _no_plot("T->C context", outputPDF, log)


PLOT_CONVERSIONS = ["A>T", "A>G", "A>C", "C>A", "C>G", "C>T", "G>A", "G>C", "G>T", "T>A", "T>G", "T>C"]   # stats.py:349-353


def _utr_percentages(totalRates, strand):
    """stats.py:386-486.  `conversionSum =+ x` assigns, so the row contributes iff its (un-swapped) T_T cell is > 0."""
    r = totalRates
    A_A, A_C, A_G, A_T = r[0], r[1], r[2], r[3]
    C_A, C_C, C_G, C_T = r[5], r[6], r[7], r[8]
    G_A, G_C, G_G, G_T = r[10], r[11], r[12], r[13]
    T_A, T_C, T_G, T_T = r[15], r[16], r[17], r[18]
    conversionSum = T_T
    if strand == "-":
        A_A, T_T = T_T, A_A
        G_G, C_C = C_C, G_G
        A_C, T_G = T_G, A_C
        A_G, T_C = T_C, A_G
        A_T, T_A = T_A, A_T
        C_A, G_T = G_T, C_A
        C_G, G_C = G_C, C_G
        C_T, G_A = G_A, C_T
    if not conversionSum > 0:
        return None
    Asum = A_A + A_C + A_G + A_T
    Csum = C_A + C_C + C_G + C_T
    Gsum = G_A + G_C + G_G + G_T
    Tsum = T_A + T_C + T_G + T_T

    def pct(x, s):
        return x / float(s) * 100 if s > 0 else 0
    return {"A>T": pct(A_T, Asum), "A>G": pct(A_G, Asum), "A>C": pct(A_C, Asum),
            "C>A": pct(C_A, Csum), "C>G": pct(C_G, Csum), "C>T": pct(C_T, Csum),
            "G>A": pct(G_A, Gsum), "G>C": pct(G_C, Gsum), "G>T": pct(G_T, Gsum),
            "T>A": pct(T_A, Tsum), "T>G": pct(T_G, Tsum), "T>C": pct(T_C, Tsum)}


def statsComputeOverallRatesPerUTR(referenceFile, bam, minBaseQual, strictTCs, outputCSV, outputPDF, utrBed, maxReadLength, log,
                                   printOnly=False, verbose=True, force=False):
    """stats.py:331-528 (alleyoop utrrates): per BED row the summed conversion matrices of its reads."""
    if not checkStep([bam, referenceFile], [outputCSV], force):
        print("Skipped computing overall rates for file " + bam, file=log)
    else:
        rs = ReadStats(referenceFile, bam, bed=utrBed)
        try:
            getSampleInfo(rs.header)                    # stats.py:333-335: raise what the reference raises on a bare BAM
            SlamSeqInfo(rs.header)
            utrs = rs.utrs
            t = rs.run(minBaseQual, strict=strictTCs, want=("utr_rates",))["utr_rates"].reshape(-1, 26)
        finally:
            rs.close()
        utrStats = dict((c, []) for c in PLOT_CONVERSIONS)
        for i, utr in enumerate(utrs):
            p = _utr_percentages([int(x) for x in t[i, :25]], utr.strand)
            if p is not None:
                for c in PLOT_CONVERSIONS:
                    utrStats[c].append(p[c])
        fo = open(outputCSV, "w")
        print("# slamdunk utrrates v" + __version__, file=fo)
        print("# Median-Conversions=", end="", file=fo)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")             # np.median([]) is nan (and warns), as in the reference
            print(",".join(c + ":" + str(np.median(utrStats[c])) for c in PLOT_CONVERSIONS), file=fo)
        print("Name", "Chr", "Start", "End", "Strand", "ReadCount", sep="\t", end="\t", file=fo)
        print("\t".join(a + "_" + b for a in toBase for b in toBase), file=fo)
        for i, utr in enumerate(utrs):
            print(utr.name, utr.chromosome, utr.start, utr.stop, utr.strand, int(t[i, 25]),
                  "\t".join(str(int(x)) for x in t[i, :25]), sep="\t", file=fo)
        fo.close()
    if not checkStep([bam, referenceFile], [outputPDF], force):
        print("Skipped computing global rate pdfs for file " + bam, file=log)
    else:
        _no_plot("global rate", outputPDF, log)


def tcPerReadPos(referenceFile, bam, minQual, maxReadLength, outputCSV, outputPDF, snpsFile, log, printOnly=False, verbose=True,
                 force=False):
    """stats.py:591-657 (alleyoop tcperreadpos): mismatches, T>C conversions and read coverage per read position."""
    if not checkStep([bam, referenceFile], [outputCSV], force):
        print("Skipped computing T->C per reads position for file " + bam, file=log)
    else:
        rs = ReadStats(referenceFile, bam, snpsFile=snpsFile)
        try:
            a = rs.run(minQual, want=("readpos",), max_read_length=maxReadLength)["readpos"].reshape(6, maxReadLength + 1)
        finally:
            rs.close()
        cov = np.cumsum(a[4:6], axis=1)                 # difference form -> reads covering each position
        foTC = open(outputCSV, "w")
        print("# slamdunk tcperreadpos v" + __version__, file=foTC)
        for i in range(0, maxReadLength):
            print(int(a[0, i]), int(a[1, i]), int(a[2, i]), int(a[3, i]), int(cov[0, i]), int(cov[1, i]), sep="\t", file=foTC)
        foTC.close()
    if not checkStep([outputCSV], [outputPDF], force):
        print("Skipped computing T->C per reads position plot for file " + bam, file=log)
    else:
        _no_plot("T->C per read position", outputPDF, log)


def tcPerUtr(referenceFile, utrBed, bam, minQual, maxReadLength, outputCSV, outputPDF, snpsFile, log, printOnly=False, verbose=True,
             force=False):
    """stats.py:659-772 (alleyoop tcperutrpos): the same over the last utrNormFactor positions of every UTR."""
    if not checkStep([bam, referenceFile], [outputCSV], force):
        print("Skipped computing T->C per UTR position for file " + bam, file=log)
    else:
        rs = ReadStats(referenceFile, bam, snpsFile=snpsFile, bed=utrBed)
        try:
            n_utr = len(rs.utrs)
            a = rs.run(minQual, want=("utrpos",))["utrpos"].reshape(6, utrNormFactor + 1)
        finally:
            rs.close()
        if verbose:
            for counter in range(10000, n_utr + 1, 10000):
                print("Handled " + str(counter) + " UTRs.", file=log)
        cov = np.cumsum(a[4:6], axis=1)
        N = utrNormFactor
        foTC = open(outputCSV, "w")
        print("# slamdunk tcperutr v" + __version__, file=foTC)
        for i in range(0, N):                           # the reverse-strand arrays are printed back to front (stats.py:760-765)
            print(int(a[0, i]), int(a[1, N - 1 - i]), int(a[2, i]), int(a[3, N - 1 - i]), int(cov[0, i]), int(cov[1, N - 1 - i]),
                  sep="\t", file=foTC)
        foTC.close()
    if not checkStep([outputCSV], [outputPDF], force):
        print("Skipped computing T->C per UTR position plot for file " + bam, file=log)
    else:
        _no_plot("T->C per UTR position", outputPDF, log)


def computeSNPMaskedRates(ref, bed, snpsFile, bam, maxReadLength, minQual, coverageCutoff, variantFraction, outputCSV, outputPDF,
                          strictTCs, log, printOnly=False, verbose=True, force=False):
    """stats.py:774-849 (alleyoop snpeval): per UTR the reads with a T>C conversion before and after SNP masking."""
    if not checkStep([bam, ref], [outputCSV], force):
        print("Skipped computing T->C per UTR with SNP masking for file " + bam, file=log)
    else:
        fileCSV = open(outputCSV, "w")                  # opened first, as the reference does (an error leaves it empty)
        try:
            rs = ReadStats(ref, bam, snpsFile=snpsFile, bed=bed, require_strand=True)
            try:
                utrs = rs.utrs
                t = rs.run(minQual, strict=strictTCs, want=("utr_snp",))["utr_snp"].reshape(-1, 3)
            finally:
                rs.close()
            for i, utr in enumerate(utrs):
                readCount, unmasked, masked = int(t[i, 0]), int(t[i, 1]), int(t[i, 2])
                containsSNP = 1 if unmasked != masked else 0
                print(utr.name + "\t" + str(readCount) + "\t" + str(unmasked) + "\t" + str(masked) + "\t" + str(containsSNP), file=fileCSV)
            if rs.pending_error is not None:
                raise rs.pending_error
        finally:
            fileCSV.close()
    if not checkStep([outputCSV], [outputPDF], force):
        print("Skipped computing T->C per UTR position plot for file " + bam, file=log)
    else:
        _no_plot("SNP evaluation", outputPDF, log)
