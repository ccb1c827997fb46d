"""oracle/stats_oracle.py -- CPU restatement of the reference's conversion statistics that share the count
path's read iterator.  TEST INFRASTRUCTURE ONLY: the checker the CUDA path
(slamdunk_b200/dunks/stats.py over sd_read_stats) is compared with; nothing in slamdunk_b200/ may import it.

Pure Python, loop for loop after the reference (paths relative to /root/reference/slamdunk):
    SlamSeqBamIterator / fillMismatchesNGM / computeRatesForRead   slamseq/SlamSeqFile.py:231-249, 313-402
    statsComputeOverallRates        dunks/stats.py:85-131      (alleyoop rates)
    statsComputeTCContext           dunks/stats.py:226-330     (alleyoop tccontext)
    statsComputeOverallRatesPerUTR  dunks/stats.py:331-528     (alleyoop utrrates)
    tcPerReadPos                    dunks/stats.py:591-657     (alleyoop tcperreadpos)
    tcPerUtr                        dunks/stats.py:659-772     (alleyoop tcperutrpos)
    computeSNPMaskedRates           dunks/stats.py:774-849     (alleyoop snpeval)
    computeTconversions(mle=True)   dunks/tcounter.py:137-143, 233-243, 305-309   (the _perread.tsv lists)

Parity status: PINNED by outputs of the unmodified reference run through the I/O shims -- the committed
tests/golden/case_*/stats files (tests/golden/make_golden.py stats) and, where /root/reference exists, live runs on
fresh seeds (tests/test_oracle.py).  The reference's own test suite holds no vector for these tools.
Meant for small cases (seconds for a few thousand reads).
"""
from __future__ import print_function

import ast
import re

import numpy as np

from slamdunk_b200.io import bam as bamio   # plain BAM / FASTA codec (I/O only)

from . import oracle as _o

VERSION = _o.VERSION
UTR_NORM = 250                                   # stats.py:35 utrNormFactor
TO_BASE = ["A", "C", "G", "T", "N"]              # stats.py:37


def _encode(base):
    """SlamSeqConversionRates.encodeBase, SlamSeqFile.py:47-57"""
    return {"A": 0, "C": 1, "G": 2, "T": 3}.get(base.upper(), 4)


class _Mismatch(object):
    __slots__ = ("readPosition", "referencePosition", "readBase", "referenceBase", "readBaseQlty", "isSnpPosition")

    def isTCMismatch(self, isReverse):                     # SlamSeqFile.py:137-141
        if isReverse:
            return self.referenceBase == "A" and self.readBase == "G" and not self.isSnpPosition
        return self.referenceBase == "T" and self.readBase == "C" and not self.isSnpPosition

    def isT(self, isReverse):                              # SlamSeqFile.py:143-147
        return self.referenceBase == ("A" if isReverse else "T")


class _Read(object):
    pass


class _Snps(object):
    """utils/SNPtools.py:23-60 incl. the chrom + str(pos) string keys"""

    def __init__(self, path, messages=None):
        self.tc, self.ag = set(), set()
        for chrom, pos, ref, alt in _o.read_vcf_records(path, messages):
            if ref.upper() == "T" and alt.upper() == "C":
                self.tc.add(chrom + pos)
            if ref.upper() == "A" and alt.upper() == "G":
                self.ag.add(chrom + pos)

    def isTCSnp(self, chrom, pos0):
        return (chrom + str(int(pos0) + 1)) in self.tc

    def isAGSnp(self, chrom, pos0):
        return (chrom + str(int(pos0) + 1)) in self.ag


def _query_alignment_sequence(r):
    """pysam: the read sequence without its soft-clipped ends"""
    seq, cig = r.seq, r.cigar
    start, end = 0, len(seq)
    k = 0
    while k < len(cig) and cig[k][0] in (4, 5):
        if cig[k][0] == 4:
            start += cig[k][1]
        k += 1
    k = len(cig) - 1
    while k >= 0 and cig[k][0] in (4, 5):
        if cig[k][0] == 4:
            end -= cig[k][1]
        k -= 1
    return seq[start:max(start, end)]


class _File(object):
    """SlamSeqBamFile (SlamSeqFile.py:404-476) over the plain codec."""

    def __init__(self, bam, ref, snps):
        self.text, self.bam_refs, records = bamio.read_bam(bam)
        self.bam_names = [n for n, _l in self.bam_refs]
        self.fasta = bamio.read_fasta(ref)
        self.snps = snps
        self.by_chrom = {}
        for r in records:
            if r.ref_id < 0 or r.reference_end is None:
                continue
            self.by_chrom.setdefault(self.bam_names[r.ref_id], []).append(r)
        for recs in self.by_chrom.values():
            recs.sort(key=lambda r: r.pos)              # coordinate order, file order among ties

    def chromosomes(self):                              # getChromosomes, SlamSeqFile.py:469-472
        refs = list(self.fasta.keys())
        refs.sort(key=lambda t: [int(c) if c.isdigit() else c for c in re.split(r"(\d+)", t)])
        return refs

    def _iterate(self, records, chromosome, startPosition, strand, minQual, conversionThreshold):
        for rec in records:                             # SlamSeqBamIterator.__next__, SlamSeqFile.py:364-402
            is_reverse = bool(rec.flag & 0x10)
            if (strand == "+" and is_reverse) or (strand == "-" and not is_reverse):
                continue
            read = _Read()
            read.name = rec.qname
            read.sequence = rec.seq
            read.isReverse = is_reverse
            read.isMultimapper = rec.mapq == 0
            # fillMismatchesNGM, SlamSeqFile.py:313-348
            tcCount, mismatches = 0, []
            if rec.has_tag("MP"):
                quals = rec.qual if rec.qual is not None else []
                for entry in rec.get_tag("MP").split(","):
                    conversion, readPos, refPos = entry.split(":")
                    conv = int(conversion)
                    refBase, readBase = TO_BASE[conv // 5], TO_BASE[conv % 5]
                    readPos = int(readPos) - 1
                    readQlty = 0 if readPos >= len(quals) else quals[readPos]
                    if readQlty >= minQual:
                        refPos = int(refPos) - 1
                        m = _Mismatch()
                        if is_reverse:
                            m.isSnpPosition = self.snps is not None and self.snps.isAGSnp(chromosome, rec.pos + refPos)
                            readPos = len(rec.seq) - readPos - 1
                        else:
                            m.isSnpPosition = self.snps is not None and self.snps.isTCSnp(chromosome, rec.pos + refPos)
                        m.readPosition, m.referencePosition = readPos, rec.pos - startPosition + refPos
                        m.readBase, m.referenceBase, m.readBaseQlty = readBase, refBase, readQlty
                        if m.isTCMismatch(is_reverse):
                            tcCount += 1
                        mismatches.append(m)
            read.mismatches, read.tcCount = mismatches, tcCount
            # computeRatesForRead, SlamSeqFile.py:231-249
            rates = [0] * 25
            aligned = _query_alignment_sequence(rec)
            for b in TO_BASE:
                rates[6 * _encode(b)] = aligned.count(b)
            for m in mismatches:
                if not m.isSnpPosition and m.readBaseQlty >= minQual:
                    rates[5 * _encode(m.referenceBase) + _encode(m.readBase)] += 1
                else:
                    rates[6 * _encode(m.referenceBase)] += 1
                rates[6 * _encode(m.readBase)] -= 1
            read.conversionRates = rates
            read.startRefPos = rec.pos - int(startPosition)
            read.endRefPos = rec.reference_end - int(startPosition)
            read.isTcRead = tcCount >= conversionThreshold
            yield read

    def readsInChromosome(self, chromosome, minQual=0, conversionThreshold=1):      # SlamSeqFile.py:447-453
        if chromosome in self.bam_names:
            return self._iterate(self.by_chrom.get(chromosome, []), chromosome, 1, ".", minQual, conversionThreshold)
        return iter([])

    def readInRegion(self, chromosome, start, stop, strand, maxReadLength, minQual=0, conversionThreshold=1):   # :419-445
        if chromosome not in self.fasta:
            return iter([])
        if chromosome not in self.bam_names:
            return iter([])
        clen = len(self.fasta[chromosome])
        fstart, fend = max(0, start), min(clen, stop)
        recs = [r for r in self.by_chrom.get(chromosome, []) if r.pos < fend and r.reference_end > fstart] if fstart <= fend else []
        return self._iterate(recs, chromosome, start, strand, minQual, conversionThreshold)


def _sum(a, b):
    return [int(x) + int(y) for x, y in zip(a, b)]       # stats.py:39-40 sumLists


def rates_file(referenceFile, bam, minBaseQual, outputCSV):
    """statsComputeOverallRates, stats.py:85-131"""
    fwd, rev = [0] * 25, [0] * 25
    f = _File(bam, referenceFile, None)
    for chromosome in f.chromosomes():
        for read in f.readsInChromosome(chromosome, minBaseQual):
            if read.isReverse:
                rev = _sum(rev, read.conversionRates)
            else:
                fwd = _sum(fwd, read.conversionRates)
    with open(outputCSV, "w") as fo:
        print("# slamdunk rates v" + VERSION, file=fo)
        print("\t", end="", file=fo)                       # printRates, stats.py:74-83
        for i in range(5):
            print(TO_BASE[i] + "\t" + TO_BASE[i].lower() + "\t", end="", file=fo)
        print(file=fo)
        for i in range(5):
            print(TO_BASE[i] + "\t", end="", file=fo)
            for j in range(5):
                print(str(fwd[i * 5 + j]) + "\t" + str(rev[i * 5 + j]) + "\t", end="", file=fo)
            print(file=fo)


def tccontext_file(referenceFile, bam, outputCSV):
    """statsComputeTCContext, stats.py:226-330: no iterator, no quality test -- the raw read sequences of every record
    placed on a FASTA chromosome"""
    front, back = ["AT", "CT", "GT", "TT", "NT"], ["TA", "TC", "TG", "TT", "TN"]
    comp = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N"}                  # utils/misc.py:276-280
    counts = {"5": {"fwd": dict((c, 0) for c in front), "rev": dict((c, 0) for c in front)},
              "3": {"fwd": dict((c, 0) for c in back), "rev": dict((c, 0) for c in back)}}
    _text, bam_refs, records = bamio.read_bam(bam)
    bam_names = [n for n, _l in bam_refs]
    fasta = bamio.read_fasta(referenceFile)
    for chromosome in fasta:
        if chromosome not in bam_names:
            raise ValueError("invalid contig `%s`" % chromosome)
        rid = bam_names.index(chromosome)
        for r in records:
            if r.ref_id != rid:
                continue
            seq, rev = r.seq, bool(r.flag & 0x10)
            for i in range(len(seq)):
                if seq[i] == "T" and not rev:
                    if i > 0:
                        counts["5"]["fwd"][seq[i - 1] + "T"] += 1
                    if i < len(seq) - 1:
                        counts["3"]["fwd"]["T" + seq[i + 1]] += 1
                if seq[i] == "A" and rev:
                    if i < len(seq) - 1:
                        counts["5"]["rev"][comp[seq[i + 1]] + "T"] += 1
                    if i > 0:
                        counts["3"]["rev"]["T" + comp[seq[i - 1]]] += 1
    with open(outputCSV, "w") as fo:
        print("\t".join(front), file=fo)
        print("\t".join(str(counts["5"]["fwd"][c]) for c in front), file=fo)
        print("\t".join(str(counts["5"]["rev"][c]) for c in front), file=fo)
        print("\t".join(back), file=fo)
        print("\t".join(str(counts["3"]["fwd"][c]) for c in back), file=fo)
        print("\t".join(str(counts["3"]["rev"][c]) for c in back), file=fo)


PLOT_CONVERSIONS = ["A>T", "A>G", "A>C", "C>A", "C>G", "C>T", "G>A", "G>C", "G>T", "T>A", "T>G", "T>C"]   # stats.py:349-353


def utr_percentages(totalRates, strand):
    """stats.py:386-486: the 12 per-UTR percentages, or None when the row adds nothing (conversionSum, assigned with
    `=+`, is just T_T of the un-swapped matrix)."""
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
    Asum, Csum = A_A + A_C + A_G + A_T, C_A + C_C + C_G + C_T
    Gsum, Tsum = G_A + G_C + G_G + G_T, T_A + T_C + T_G + T_T
    pct = lambda x, s: x / float(s) * 100 if s > 0 else 0    # noqa: E731
    return {"A>T": pct(A_T, Asum), "A>G": pct(A_G, Asum), "A>C": pct(A_C, Asum),
            "C>A": pct(C_A, Csum), "C>G": pct(C_G, Csum), "C>T": pct(C_T, Csum),
            "G>A": pct(G_A, Gsum), "G>C": pct(G_C, Gsum), "G>T": pct(G_T, Gsum),
            "T>A": pct(T_A, Tsum), "T>G": pct(T_G, Tsum), "T>C": pct(T_C, Tsum)}


def write_utrrates(outputCSV, rows):
    """stats.py:488-517. rows: (name, chromosome, start, stop, strand, readCount, totalRates[25]) in BED order."""
    import warnings
    stats = dict((c, []) for c in PLOT_CONVERSIONS)
    for _n, _c, _s, _e, strand, _rc, total in rows:
        p = utr_percentages(total, strand)
        if p is not None:
            for c in PLOT_CONVERSIONS:
                stats[c].append(p[c])
    with open(outputCSV, "w") as fo:
        print("# slamdunk utrrates v" + VERSION, file=fo)
        print("# Median-Conversions=", end="", file=fo)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")                  # np.median([]) -> nan, with a RuntimeWarning
            print(",".join(c + ":" + str(np.median(stats[c])) for c in PLOT_CONVERSIONS), file=fo)
        print("Name", "Chr", "Start", "End", "Strand", "ReadCount", sep="\t", end="\t", file=fo)
        print("\t".join(a + "_" + b for a in TO_BASE for b in TO_BASE), file=fo)
        for name, chrom, start, stop, strand, rc, total in rows:
            print(name, chrom, start, stop, strand, rc, "\t".join(str(x) for x in total), sep="\t", file=fo)


def utrrates_file(referenceFile, bam, minBaseQual, strictTCs, outputCSV, utrBed, maxReadLength):
    """statsComputeOverallRatesPerUTR, stats.py:331-528"""
    _check_header(bam)
    f = _File(bam, referenceFile, None)
    rows = []
    for chrom, start, stop, name, strand in _o.read_bed(utrBed):
        total, readCount = [0] * 25, 0
        for read in f.readInRegion(chrom, start, stop, strand, maxReadLength, minBaseQual):
            if not read.isTcRead and strictTCs and read.tcCount > 0:
                pass
            else:
                total = _sum(total, read.conversionRates)
                readCount += 1
        rows.append((name, chrom, start, stop, strand, readCount, total))
    write_utrrates(outputCSV, rows)


def _check_header(bam):
    """getSampleInfo + SlamSeqInfo (utils/misc.py:33-89, 229-241): raise what the reference raises"""
    text = bamio.read_bam(bam)[0]
    hdr = bamio.parse_header_text(text)
    if not ("RG" in hdr and len(hdr["RG"]) > 0):
        raise RuntimeError("Could not get mapped/unmapped/filtered read counts from BAM file. RG is missing. Please rerun slamdunk filter.")
    ast.literal_eval(hdr["RG"][0]["DS"])
    return hdr


def tcperreadpos_file(referenceFile, bam, minQual, maxReadLength, outputCSV, snpsFile):
    """tcPerReadPos, stats.py:591-657"""
    W = maxReadLength
    totF, totR, tcR, tcF, allR, allF = [0] * W, [0] * W, [0] * W, [0] * W, [0] * W, [0] * W
    f = _File(bam, referenceFile, _Snps(snpsFile))
    for chromosome in f.chromosomes():
        for read in f.readsInChromosome(chromosome, minQual):
            if len(read.sequence) <= W:
                tcCounts, mutCounts = [0] * W, [0] * W
                for m in read.mismatches:
                    if m.isTCMismatch(read.isReverse):
                        tcCounts[m.readPosition] += 1
                    else:
                        mutCounts[m.readPosition] += 1
                n = len(read.sequence)
                if read.isReverse:
                    tcR, allR = _sum(tcR, tcCounts), _sum(allR, mutCounts)
                    for i in range(n):
                        totR[i] += 1
                else:
                    tcF, allF = _sum(tcF, tcCounts), _sum(allF, mutCounts)
                    for i in range(n):
                        totF[i] += 1
    with open(outputCSV, "w") as fo:
        print("# slamdunk tcperreadpos v" + VERSION, file=fo)
        for i in range(W):
            print(allF[i], allR[i], tcF[i], tcR[i], totF[i], totR[i], sep="\t", file=fo)


def tcperutr_file(referenceFile, utrBed, bam, minQual, maxReadLength, outputCSV, snpsFile):
    """tcPerUtr, stats.py:659-772"""
    N = UTR_NORM
    totF, totR, tcR, tcF, allR, allF = [0] * N, [0] * N, [0] * N, [0] * N, [0] * N, [0] * N
    f = _File(bam, referenceFile, _Snps(snpsFile))
    for chrom, ustart, ustop, _name, strand in _o.read_bed(utrBed):
        L = ustop - ustart
        tcFc, mutFc, tcRc, mutRc = [0] * N, [0] * N, [0] * N, [0] * N
        for read in f.readInRegion(chrom, ustart, ustop, strand, maxReadLength, minQual):
            tcCounts, mutCounts = [0] * N, [0] * N
            for m in read.mismatches:
                p = m.referencePosition
                if strand == "+":
                    if p >= (L - N) and p < L:
                        p = N - (L - p)
                        if m.isTCMismatch(read.isReverse):
                            tcCounts[p] += 1
                        else:
                            mutCounts[p] += 1
                else:
                    if p >= 0 and p < min(L, N):
                        if m.isTCMismatch(read.isReverse):
                            tcCounts[p] += 1
                        else:
                            mutCounts[p] += 1
            if read.isReverse:
                tcRc, mutRc = _sum(tcRc, tcCounts), _sum(mutRc, mutCounts)
                start = max(0, min(min(L, N), read.startRefPos))
                end = max(0, min(min(L, N), read.endRefPos))
                for i in range(start, end):
                    totR[i] += 1
            else:
                tcFc, mutFc = _sum(tcFc, tcCounts), _sum(mutFc, mutCounts)
                start = min(L, max(L - N, read.startRefPos))
                end = min(L, max(L - N, read.endRefPos))
                for i in range(start, end):
                    totF[N - (L - i)] += 1
        tcF, allF = _sum(tcF, tcFc), _sum(allF, mutFc)
        tcR, allR = _sum(tcR, tcRc), _sum(allR, mutRc)
    with open(outputCSV, "w") as fo:
        print("# slamdunk tcperutr v" + VERSION, file=fo)
        rAll, rTc, rTot = allR[::-1], tcR[::-1], totR[::-1]
        for i in range(N):
            print(allF[i], rAll[i], tcF[i], rTc[i], totF[i], rTot[i], sep="\t", file=fo)


def snpeval_file(ref, bed, snpsFile, bam, maxReadLength, minQual, outputCSV, strictTCs):
    """computeSNPMaskedRates, stats.py:774-849"""
    f = _File(bam, ref, _Snps(snpsFile))
    with open(outputCSV, "w") as fo:
        for chrom, start, stop, name, strand in _o.read_bed(bed):
            if strand not in ("+", "-"):
                raise RuntimeError("Input BED file does not contain stranded intervals.")
            if start < 0:
                raise TypeError("Negativ start coordinate found.")     # str + BedEntry in the reference
            L = stop - start
            unmasked, masked, readCount = 0, 0, 0
            for read in f.readInRegion(chrom, start, stop, strand, maxReadLength, minQual):
                mismatches = read.mismatches
                if not read.isTcRead and strictTCs:
                    mismatches = []
                isTC = isTrueTC = False
                for m in mismatches:
                    if m.isTCMismatch(read.isReverse) and 0 <= m.referencePosition < L:
                        isTrueTC = True
                    un = (m.referenceBase == "A" and m.readBase == "G") if read.isReverse else (m.referenceBase == "T" and m.readBase == "C")
                    if un and 0 <= m.referencePosition < L:
                        isTC = True
                readCount += 1
                unmasked += 1 if isTC else 0
                masked += 1 if isTrueTC else 0
            print(name + "\t" + str(readCount) + "\t" + str(unmasked) + "\t" + str(masked) + "\t" + str(1 if unmasked != masked else 0), file=fo)


def perread_lists(ref, bed, snpsFile, bam, maxReadLength, minQual, conversionThreshold):
    """The tInReads / tcInRead lists of computeTconversions(mle=True), tcounter.py:196-243, one pair of lists per BED row."""
    f = _File(bam, ref, _Snps(snpsFile))
    out = []
    for chrom, start, stop, _name, strand in _o.read_bed(bed):
        L = stop - start
        tInReads, tcInRead = [], []
        for read in f.readInRegion(chrom, start, stop, strand, maxReadLength, minQual, conversionThreshold):
            mismatches = read.mismatches
            if read.tcCount < conversionThreshold:              # tcounter.py:210-214
                mismatches = []
            seq = read.sequence
            testN = (seq.count("a") + seq.count("A")) if read.isReverse else (seq.count("t") + seq.count("T"))
            testk = 0
            for m in mismatches:
                if 0 <= m.referencePosition < L:
                    if m.isT(read.isReverse):
                        testN += 1
                    if m.isTCMismatch(read.isReverse):
                        testk += 1
            tInReads.append(testN)
            tcInRead.append(testk)
        out.append((tInReads, tcInRead))
    return out
