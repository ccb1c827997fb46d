"""Small randomized SLAM-seq data sets for parity tests (test infrastructure).

Produces, from a seed, a consistent quadruple FASTA / BED6 / VCF / BAM the way the
reference's pipeline would see them after `slamdunk map` + `filter`:
NextGenMap-style tags (MP:Z mismatch list, XI:f identity, NM:i, XA:i), a @RG with the
`DS` read-count dict and the @PG slamdunk VN:3 line.

The MP tag is written from the simulation's own record of which aligned columns were
mutated -- NOT by re-walking the CIGAR -- so a CIGAR-walking implementation and an
MP-parsing one are checked against each other rather than against themselves.

Edge cases deliberately produced: both strands, overlapping / nested / duplicate UTRs,
UTRs past the chromosome end and on chromosomes missing from FASTA or BAM, soft clips,
insertions, deletions, N skips, '=' / 'X' ops, lower-case and N reference bases, base
qualities around the threshold, multimappers (mapq 0), reads that lost their MP tag,
SNPs incl. multi-allelic records and the reference's `chrom + str(pos)` key collision
(utils/SNPtools.py:33,54 -- 'chr1'+'2345' == 'chr12'+'345').
"""
from __future__ import annotations

import hashlib
import os
import random
from typing import Dict, List, Optional, Tuple

from slamdunk_b200.io import bam as bamio

BASES = "ACGT"
_CODE = {"A": 0, "C": 1, "G": 2, "T": 3}


def _code(b: str) -> int:
    return _CODE.get(b.upper(), 4)


class Case:
    """Paths + in-memory description of one generated data set."""

    def __init__(self):
        self.ref = self.bed = self.vcf = self.bam = self.raw_bam = None
        self.genome: Dict[str, str] = {}
        self.utrs: List[Tuple[str, int, int, str, str]] = []
        self.snps: List[Tuple[str, int, str, str]] = []
        self.records: List[bamio.BamRecord] = []
        self.bam_refs: List[Tuple[str, int]] = []
        self.params = {}


def _rand_seq(rng: random.Random, n: int) -> str:
    return "".join(rng.choice(BASES) for _ in range(n))


def make_genome(rng, chroms: List[Tuple[str, int]], n_frac=0.002, lower_frac=0.05) -> Dict[str, str]:
    genome = {}
    for name, length in chroms:
        s = list(_rand_seq(rng, length))
        # N runs
        n_runs = max(1, int(length * n_frac / 20))
        for _ in range(n_runs):
            a = rng.randrange(0, length)
            for i in range(a, min(length, a + rng.randint(1, 40))):
                s[i] = "N"
        # lower-case (soft-masked) stretches and the odd IUPAC code
        for _ in range(max(1, int(length * lower_frac / 50))):
            a = rng.randrange(0, length)
            for i in range(a, min(length, a + rng.randint(5, 100))):
                s[i] = s[i].lower()
        for _ in range(max(1, length // 5000)):
            s[rng.randrange(0, length)] = rng.choice("RYKMSWn")
        genome[name] = "".join(s)
    return genome


def make_utrs(rng, genome, n_utr, extra_chroms=()):
    utrs = []
    names = list(genome.keys())
    for i in range(n_utr):
        ch = rng.choice(names)
        L = len(genome[ch])
        ln = min(L // 2, max(30, int(rng.lognormvariate(6.0, 0.8))))
        start = rng.randrange(0, max(1, L - ln))
        strand = rng.choice("+-")
        utrs.append((ch, start, start + ln, "utr%d" % i, strand))
        r = rng.random()
        if r < 0.08:    # overlapping neighbour, same or other strand
            s2 = start + rng.randint(1, ln - 1)
            utrs.append((ch, s2, min(L, s2 + rng.randint(30, ln + 50)), "utr%d_ov" % i, rng.choice([strand, strand, "+", "-"])))
        elif r < 0.12:  # nested
            s2 = start + rng.randint(0, ln // 2)
            utrs.append((ch, s2, s2 + max(10, ln // 4), "utr%d_in" % i, strand))
        elif r < 0.14:  # exact duplicate under another name
            utrs.append((ch, start, start + ln, "utr%d_dup" % i, strand))
    # a UTR reaching past the chromosome end, one at position 0
    ch = names[0]
    L = len(genome[ch])
    utrs.append((ch, L - 150, L + 200, "utr_pastend", "+"))
    utrs.append((ch, 0, 120, "utr_zero", "-"))
    for ch in extra_chroms:  # chromosomes absent from FASTA and/or BAM
        utrs.append((ch, 100, 900, "utr_" + ch, "+"))
    rng.shuffle(utrs)
    return utrs


def make_snps(rng, genome, utrs, density, collide=True):
    """-> list of (chrom, pos1, ref, alt). ~half T>C / A>G, a few multi-allelic, some
    placed so that the concatenated key collides with another chromosome's position."""
    snps = {}
    for ch, start, stop, _n, _s in utrs:
        if ch not in genome:
            continue
        seq = genome[ch]
        for p in range(max(0, start - 60), min(len(seq), stop + 60)):
            if rng.random() < density:
                refb = seq[p].upper()
                if refb not in BASES:
                    continue
                r = rng.random()
                if refb == "T" and r < 0.7:
                    alt = "C"
                elif refb == "A" and r < 0.7:
                    alt = "G"
                else:
                    alt = rng.choice([b for b in BASES if b != refb])
                if rng.random() < 0.03:
                    alt = alt + "," + rng.choice([b for b in BASES if b != refb])
                snps[(ch, p + 1)] = (seq[p] if rng.random() < 0.5 else refb, alt if rng.random() < 0.9 else alt.lower())
    if collide:
        # chrA + digits == chrB + fewer digits  (e.g. chr1|2345 vs chr12|345)
        names = list(genome.keys())
        for a in names:
            for b in names:
                if a != b and b.startswith(a) and b[len(a):].isdigit():
                    extra = b[len(a):]
                    # a SNP recorded on chromosome b at pos q masks chromosome a at int(extra + str(q))
                    for ch, start, stop, _n, strand in utrs:
                        if ch != a:
                            continue
                        for p in range(start, min(stop, len(genome[a]))):
                            key = str(p + 1)
                            if key.startswith(extra) and len(key) > len(extra) and key[len(extra)] != "0":
                                q = int(key[len(extra):])
                                if q <= len(genome[b]) and rng.random() < 0.3:
                                    want = "T" if strand == "+" else "A"
                                    if genome[a][p].upper() == want:
                                        snps[(b, q)] = (want, "C" if want == "T" else "G")
    out = [(ch, pos, ra[0], ra[1]) for (ch, pos), ra in snps.items()]
    rng.shuffle(out)   # the reference does not need a sorted VCF
    return out


def write_vcf(path, snps):
    with open(path, "w") as fh:
        fh.write("##fileformat=VCFv4.1\n##source=VarScan2\n")
        fh.write("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tSample1\n")
        for ch, pos, ref, alt in snps:
            fh.write("%s\t%d\t.\t%s\t%s\t.\tPASS\tADP=20;WT=0;HET=0;HOM=1;NC=0\tGT:GQ\t1/1:255\n" % (ch, pos, ref, alt))


def write_bed(path, utrs):
    with open(path, "w") as fh:
        for ch, start, stop, name, strand in utrs:
            fh.write("%s\t%d\t%d\t%s\t0\t%s\n" % (ch, start, stop, name, strand))


def _pick_cigar(rng, L, complexity):
    """-> list of (op, len) with query length L. complexity scales the non-trivial share."""
    r = rng.random()
    c = complexity
    if r > 0.30 * c or L < 12:
        return [(0, L)]
    ops = []
    left = rng.randint(1, 5) if rng.random() < 0.45 else 0
    right = rng.randint(1, 5) if rng.random() < 0.45 else 0
    core = L - left - right
    if left:
        ops.append((4, left))
    kind = rng.random()
    if kind < 0.35 or core < 10:
        ops.append((0, core))
    elif kind < 0.60:   # insertion
        il = rng.randint(1, 3)
        a = rng.randint(2, core - il - 2)
        ops += [(0, a), (1, il), (0, core - a - il)]
    elif kind < 0.85:   # deletion
        a = rng.randint(2, core - 2)
        ops += [(0, a), (2, rng.randint(1, 4)), (0, core - a)]
    elif kind < 0.90:   # skipped region (spliced)
        a = rng.randint(2, core - 2)
        ops += [(0, a), (3, rng.randint(20, 400)), (0, core - a)]
    elif kind < 0.95:   # insertion + deletion
        a = rng.randint(2, max(2, core // 3))
        il = rng.randint(1, 2)
        b = rng.randint(2, max(2, core - a - il - 2))
        rest = core - a - il - b
        if rest >= 1:
            ops += [(0, a), (1, il), (0, b), (2, rng.randint(1, 3)), (0, rest)]
        else:
            ops.append((0, core))
    else:               # '=' / 'X' spelled alignment is handled by the caller (needs bases)
        ops.append((7, core))
    if right:
        ops.append((4, right))
    if rng.random() < 0.03:
        ops = [(5, rng.randint(1, 9))] + ops   # hard clip
    return ops


def simulate_read(rng, genome_up, ch, pos, L, reverse, snp_alleles, p_conv, p_err, labelled, complexity):
    """-> (cigar, seq, mismatches[(conv, readPos1, refPos1)]) or None if it does not fit."""
    ref = genome_up[ch]
    cigar = _pick_cigar(rng, L, complexity)
    ref_len = sum(l for op, l in cigar if op in (0, 2, 3, 7, 8))
    if pos < 0 or pos + ref_len > len(ref):
        return None
    seq = []
    mism = []
    q = 0
    r = 0
    conv_from, conv_to = ("A", "G") if reverse else ("T", "C")
    out_cigar = []
    for op, l in cigar:
        if op in (0, 7):
            run_op = None
            run_len = 0
            for i in range(l):
                rb = ref[pos + r]
                b = rb if rb in BASES else rng.choice(BASES)
                snp = snp_alleles.get((ch, pos + r + 1))
                if snp is not None and rng.random() < 0.9:
                    b = snp
                elif rb == conv_from and labelled and rng.random() < p_conv:
                    b = conv_to
                if rng.random() < p_err:
                    b = rng.choice(BASES + "N") if rng.random() < 0.1 else rng.choice(BASES)
                seq.append(b)
                rcode = _code(rb)
                if b != rb:
                    mism.append((5 * rcode + _code(b), q + 1, r + 1))
                if op == 7:
                    this = 7 if b == rb else 8
                    if this != run_op:
                        if run_len:
                            out_cigar.append((run_op, run_len))
                        run_op, run_len = this, 0
                    run_len += 1
                q += 1
                r += 1
            if op == 0:
                out_cigar.append((0, l))
            elif run_len:
                out_cigar.append((run_op, run_len))
        elif op == 1 or op == 4:
            seq.extend(rng.choice(BASES) for _ in range(l))
            q += l
            out_cigar.append((op, l))
        elif op in (2, 3):
            r += l
            out_cigar.append((op, l))
        else:
            out_cigar.append((op, l))
    return out_cigar, "".join(seq), mism


def make_case(outdir: str, seed: int, n_reads=2000, read_len=(30, 100), n_utr=40,
              chroms=(("chr1", 30000), ("chr12", 6000), ("chrX", 8000)),
              snp_density=0.004, p_conv=0.03, p_err=0.004, complexity=1.0,
              multimap_frac=0.05, drop_mp_frac=0.03, filtered_header=True,
              raw_for_filter=False, name="case") -> Case:
    rng = random.Random(seed)
    os.makedirs(outdir, exist_ok=True)
    c = Case()
    c.params = dict(seed=seed, n_reads=n_reads, read_len=read_len, n_utr=n_utr, chroms=list(chroms),
                    snp_density=snp_density, p_conv=p_conv, p_err=p_err, complexity=complexity)
    genome = make_genome(rng, list(chroms))
    c.genome = genome
    genome_up = {k: v.upper() for k, v in genome.items()}
    # BAM knows one chromosome the FASTA lacks ("chrUn") and lacks one the FASTA has (last one)
    fasta_names = list(genome.keys())
    bam_refs = [(n, len(genome[n])) for n in fasta_names[:-1]] if len(fasta_names) > 2 else [(n, len(genome[n])) for n in fasta_names]
    bam_refs.append(("chrUn", 5000))
    c.bam_refs = bam_refs
    bam_ref_ids = {n: i for i, (n, _l) in enumerate(bam_refs)}
    utrs = make_utrs(rng, genome, n_utr, extra_chroms=("chrUn", "chrNowhere"))
    c.utrs = utrs
    snps = make_snps(rng, genome, utrs, snp_density) if snp_density > 0 else []
    c.snps = snps
    snp_alleles = {}
    for ch, pos, ref, alt in snps:
        if "," not in alt and len(alt) == 1:
            snp_alleles[(ch, pos)] = alt.upper()

    c.ref = os.path.join(outdir, name + ".fa")
    c.bed = os.path.join(outdir, name + ".bed")
    c.vcf = os.path.join(outdir, name + ".vcf")
    c.bam = os.path.join(outdir, name + ".bam")
    bamio.write_fasta(c.ref, genome)
    write_bed(c.bed, utrs)
    write_vcf(c.vcf, snps)

    usable = [u for u in utrs if u[0] in genome and u[0] in bam_ref_ids]
    weights = [0.85 ** i + 0.01 for i in range(len(usable))]
    records = []
    lo, hi = read_len
    serial = 0
    while len(records) < n_reads:
        serial += 1
        L = rng.randint(lo, hi)
        if rng.random() < 0.06:      # intergenic read
            ch = rng.choice([n for n in genome if n in bam_ref_ids])
            pos = rng.randrange(0, max(1, len(genome[ch]) - L - 450))
            reverse = rng.random() < 0.5
        else:
            ch, start, stop, _n, strand = rng.choices(usable, weights)[0]
            pos = rng.randint(start - L + 1, max(start - L + 1, min(stop, len(genome[ch])) - 1))
            reverse = (strand == "-") if rng.random() < 0.93 else (strand == "+")
        labelled = rng.random() < 0.7
        sim = simulate_read(rng, genome_up, ch, pos, L, reverse, snp_alleles, p_conv, p_err, labelled, complexity)
        if sim is None:
            continue
        cigar, seq, mism = sim
        qual = [rng.choice((12, 20, 26, 27, 28, 30, 37, 37, 37, 41)) for _ in range(len(seq))]
        mapq = 0 if rng.random() < multimap_frac else rng.choice((1, 3, 20, 40, 60))
        aligned = sum(l for op, l in cigar if op in (0, 7, 8))
        nm = len(mism) + sum(l for op, l in cigar if op in (1, 2))
        xi = 1.0 - (nm / float(max(1, aligned)))
        tags = [("AS", "i", aligned * 10 - 15 * nm), ("NM", "i", nm), ("XI", "f", xi),
                ("XA", "i", rng.randint(0, 3))]
        if mism and rng.random() >= drop_mp_frac:
            tags.append(("MP", "Z", ",".join("%d:%d:%d" % m for m in mism)))
        if mapq == 0 and rng.random() < 0.5:
            tags.append(("RD", "Z", "%s:%d-%d" % (ch, pos, pos + 10)))
        flag = 16 if reverse else 0
        if rng.random() < 0.01:
            flag |= 0x100
        records.append(bamio.BamRecord("r%06d" % serial, flag, bam_ref_ids[ch], pos, mapq, cigar, seq, qual, tags))
    c.records = bamio.sort_records(records)
    md5 = hashlib.md5(open(c.bed, "rb").read()).hexdigest()
    n_primary = sum(1 for r in records if not (r.flag & 0x900))
    ds = ("{'sequenced':%d,'mapped':%d,'filtered':%d,'mqfiltered':0,'idfiltered':0,'nmfiltered':0,'multimapper':0,"
          "'dedup':0,'snps':%d,'annotation':'%s','annotationmd5':'%s'}"
          % (n_primary + 7, n_primary, n_primary, len(snps), os.path.basename(c.bed), md5))
    hdr = "@HD\tVN:1.0\tSO:coordinate\n"
    for n, l in bam_refs:
        hdr += "@SQ\tSN:%s\tLN:%d\n" % (n, l)
    hdr += "@RG\tID:%d\tSM:%s:pulse:%d\tDS:%s\n" % (seed % 7, name, seed % 90, ds)
    hdr += "@PG\tID:ngm\tPN:ngm\tVN:0.5.5\tCL:ngm --slam-seq 2\n"
    if filtered_header:
        hdr += "@PG\tID:slamdunk\tPN:slamdunk filter v0.4.3\tVN:3\n"
    bamio.write_bam(c.bam, hdr, bam_refs, c.records)
    return c


# --------------------------------------------------------------------------
# unsorted mapper-style output for the filter tests
# --------------------------------------------------------------------------
def make_filter_case(outdir: str, seed: int, n_reads=1500, name="raw", topn=3) -> Case:
    """Mapper-order BAM (as `ngm -n topn --strata` + `samtools view -b`, unsorted,
    slamdunk/dunks/mapper.py:31): alignments of one read adjacent, multimappers mapq 0
    with several alignments, unmapped reads, secondary flags, XI/NM tags."""
    rng = random.Random(seed)
    os.makedirs(outdir, exist_ok=True)
    c = Case()
    chroms = (("chr1", 20000), ("chr2", 9000))
    genome = make_genome(rng, list(chroms))
    genome_up = {k: v.upper() for k, v in genome.items()}
    c.genome = genome
    utrs = make_utrs(rng, genome, 25)
    utrs = [u for u in utrs if u[0] in genome and u[2] <= len(genome[u[0]])]
    c.utrs = utrs
    c.ref = os.path.join(outdir, name + ".fa")
    c.bed = os.path.join(outdir, name + ".bed")
    c.bam = os.path.join(outdir, name + ".bam")
    bamio.write_fasta(c.ref, genome)
    write_bed(c.bed, utrs)
    bam_refs = [(n, len(s)) for n, s in genome.items()]
    ids = {n: i for i, (n, _l) in enumerate(bam_refs)}
    records = []

    def aln(qname, mapq, flag_extra=0, near_utr=True):
        for _ in range(50):
            L = rng.randint(30, 80)
            if near_utr and rng.random() < 0.8:
                ch, start, stop, _n, strand = rng.choice(utrs)
                pos = rng.randint(start - L, stop)   # includes the [stop, stop+1) off-by-one zone
                reverse = strand == "-"
            else:
                ch = rng.choice(list(genome))
                pos = rng.randrange(0, len(genome[ch]) - L)
                reverse = rng.random() < 0.5
            sim = simulate_read(rng, genome_up, ch, pos, L, reverse, {}, 0.03, rng.choice((0.002, 0.02, 0.08)), True, 0.6)
            if sim is None:
                continue
            cigar, seq, mism = sim
            aligned = sum(l for op, l in cigar if op in (0, 7, 8))
            nm = len(mism) + sum(l for op, l in cigar if op in (1, 2))
            xi = 1.0 - (nm / float(max(1, aligned)))
            tags = [("NM", "i", nm), ("XI", "f", xi), ("XA", "i", 0)]
            if mism:
                tags.append(("MP", "Z", ",".join("%d:%d:%d" % m for m in mism)))
            qual = [rng.choice((20, 30, 37)) for _ in seq]
            return bamio.BamRecord(qname, (16 if reverse else 0) | flag_extra, ids[ch], pos, mapq, cigar, seq, qual, tags)
        raise RuntimeError("could not place read")

    for i in range(n_reads):
        qname = "q%05d" % i
        r = rng.random()
        if r < 0.05:
            records.append(bamio.BamRecord(qname, 4, -1, -1, 0, [], _rand_seq(rng, 40), [30] * 40, []))
        elif r < 0.30:
            k = rng.randint(1, topn)
            same_utr = rng.random() < 0.5
            first = aln(qname, 0)
            records.append(first)
            for j in range(1, k):
                if same_utr:
                    dup = first.copy()
                    dup.pos = max(0, first.pos + rng.randint(-3, 3))
                    dup.flag |= 0x100 if rng.random() < 0.5 else 0
                    records.append(dup)
                else:
                    records.append(aln(qname, 0, 0x100 if rng.random() < 0.5 else 0, near_utr=rng.random() < 0.7))
        else:
            records.append(aln(qname, rng.choice((1, 2, 3, 10, 60))))
    c.records = records
    hdr = "@HD\tVN:1.0\tSO:unsorted\n"
    for n, l in bam_refs:
        hdr += "@SQ\tSN:%s\tLN:%d\n" % (n, l)
    hdr += "@RG\tID:1\tSM:%s:chase:30\n" % name
    hdr += "@PG\tID:ngm\tPN:ngm\tVN:0.5.5\n"
    c.bam_refs = bam_refs
    bamio.write_bam(c.bam, hdr, bam_refs, records)
    return c
