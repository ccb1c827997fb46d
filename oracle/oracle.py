"""Python face of the CPU oracle (oracle/tcount_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of tcount_oracle.c.  Imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg; never by
slamdunk_b200/.

`count_files` / `filter_file` restate the file-level behaviour of the reference's
tcounter.computeTconversions (dunks/tcounter.py:125-330) and filter.Filter
(dunks/filter.py:213-315) on top of the C arithmetic: same header lines, same row
formatting (slamseq/SlamSeqFile.py:78-100), same bedgraph lines, same exceptions.
Parity: pinned against the reference's golden TSV and against the reference's own code
run through I/O shims (tests/test_oracle.py).
"""
from __future__ import annotations

import ast
import ctypes
import hashlib
import os
import subprocess
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "tcount_oracle.c")
_LIB = os.path.join(_HERE, "liboracle.so")

VERSION = "0.4.3"          # slamdunk/version.py:19
BAM_VERSION = "3"          # slamdunk/version.py:21
COUNT_VERSION = "3"        # slamdunk/version.py:23

READ_DT = np.dtype([("chrom", "<i4"), ("pos", "<i4"), ("end", "<i4"), ("l_qual", "<i4"),
                    ("flag", "<u2"), ("mapq", "u1"), ("has_mp", "u1"), ("mp_len", "<i4"),
                    ("qual_off", "<i8"), ("mp_off", "<i8")])
UTR_DT = np.dtype([("chrom", "<i4"), ("in_bam", "<i4"), ("start", "<i4"), ("stop", "<i4"),
                   ("strand", "<i4"), ("name_id", "<i4")])
ROW_DT = np.dtype([("tcontent", "<i8"), ("coverage_on_ts", "<i8"), ("conversions_on_ts", "<i8"),
                   ("read_count", "<i8"), ("tc_read_count", "<i8"), ("multimap_count", "<i8"),
                   ("has_reads", "<i8")])
BG_DT = np.dtype([("name_id", "<i4"), ("pos", "<i4"), ("strand", "<i4"), ("tc", "<i4"), ("cov", "<i4")])
FREAD_DT = np.dtype([("chrom_id", "<i4"), ("start", "<i4"), ("end", "<i4"), ("name_id", "<i4"),
                     ("flag", "<u2"), ("mapq", "u1"), ("pad", "u1"), ("xi", "<f4"), ("nm", "<i4")])
FUTR_DT = np.dtype([("chrom_id", "<i4"), ("start", "<i4"), ("stop", "<i4"), ("name_id", "<i4")])


class _CountArgs(ctypes.Structure):
    _fields_ = [("n_chrom", ctypes.c_int32),
                ("chrom_names", ctypes.POINTER(ctypes.c_char_p)),
                ("chrom_seq", ctypes.POINTER(ctypes.c_char_p)),
                ("chrom_len", ctypes.c_void_p),
                ("reads", ctypes.c_void_p),
                ("chrom_read_begin", ctypes.c_void_p),
                ("chrom_max_span", ctypes.c_void_p),
                ("qual_pool", ctypes.c_void_p),
                ("mp_text", ctypes.c_void_p),
                ("mp_triples", ctypes.c_void_p),
                ("n_utr", ctypes.c_int32),
                ("utrs", ctypes.c_void_p),
                ("min_qual", ctypes.c_int32),
                ("conversion_threshold", ctypes.c_int32),
                ("snps", ctypes.c_void_p)]


def build(force: bool = False) -> str:
    """gcc -O2 -fopenmp -shared -> oracle/liboracle.so (git-ignored, travels with gpurun)."""
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(_SRC):
        subprocess.check_call(["gcc", "-O2", "-fopenmp", "-shared", "-fPIC", "-o", _LIB, _SRC])
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        L = ctypes.CDLL(_LIB)
        L.orc_snps_new.restype = ctypes.c_void_p
        L.orc_snps_new.argtypes = [ctypes.c_int64]
        L.orc_snps_add.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p]
        L.orc_snps_free.argtypes = [ctypes.c_void_p]
        L.orc_count.argtypes = [ctypes.POINTER(_CountArgs), ctypes.c_void_p,
                                ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_int64), ctypes.c_int]
        L.orc_free.argtypes = [ctypes.c_void_p]
        L.orc_select_sort.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p,
                                      ctypes.c_void_p, ctypes.c_int]
        L.orc_select_sort.restype = ctypes.c_int64
        L.orc_positional_chrom.argtypes = [ctypes.POINTER(_CountArgs), ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p,
                                           ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.orc_positional_chrom.restype = ctypes.c_int
        L.orc_read_tc.argtypes = [ctypes.POINTER(_CountArgs), ctypes.c_int64, ctypes.c_void_p]
        L.orc_read_tc.restype = ctypes.c_int
        L.orc_filter_default.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_double,
                                         ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        L.orc_filter_default_mt.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_double,
                                            ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        L.orc_filter_multimap.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int32,
                                          ctypes.c_double, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        assert L.orc_sizeof_read() == READ_DT.itemsize and L.orc_sizeof_utr() == UTR_DT.itemsize
        assert L.orc_sizeof_row() == ROW_DT.itemsize and L.orc_sizeof_bg() == BG_DT.itemsize
        assert L.orc_sizeof_fread() == FREAD_DT.itemsize
        _lib = L
    return _lib


def max_threads() -> int:
    """Host threads available to this process: the scheduler's affinity mask, NOT OpenMP's default -- torchrun exports
    OMP_NUM_THREADS=1 to its workers, which is a launcher setting and not a property of the box."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        return max(1, os.cpu_count() or 1)


# --------------------------------------------------------------------------
# array-level entry (synthetic data, bench cpu_baseline)
# --------------------------------------------------------------------------
class Genome:
    def __init__(self, names: Sequence[str], seqs: Sequence[bytes]):
        self.names = list(names)
        self.seqs = [s if isinstance(s, bytes) else s.encode("ascii") for s in seqs]
        self.index = {n: i for i, n in enumerate(self.names)}
        self.lengths = np.array([len(s) for s in self.seqs], dtype=np.int32)


def make_snps(records: Sequence[Tuple[str, str, str, str]]):
    """records: (CHROM, POS text, REF, ALT) as pybedtools would index a VCF line
    (utils/SNPtools.py:43-49).  Returns an opaque handle (free with free_snps)."""
    h = lib().orc_snps_new(len(records))
    for ch, pos, ref, alt in records:
        lib().orc_snps_add(h, ch.encode(), str(pos).encode(), ref.encode(), alt.encode())
    return h


def free_snps(h) -> None:
    if h:
        lib().orc_snps_free(h)


def select_sort(reads: np.ndarray, keep: Optional[np.ndarray], n_chrom: int, threads: int = 1):
    """The kept reads in coordinate order (what `slamdunk filter` + samtools sort leave behind), bucketed and sorted on all
    threads by the C side.  -> (reads READ_DT sorted by (chrom, pos), begin int64[n_chrom + 1], max_span int32[n_chrom])"""
    reads = np.ascontiguousarray(reads, dtype=READ_DT)
    out = np.empty(len(reads), dtype=READ_DT)
    begin = np.zeros(n_chrom + 1, dtype=np.int64)
    span = np.ones(n_chrom, dtype=np.int32)
    k = np.ascontiguousarray(keep, dtype=np.uint8) if keep is not None else None
    n_out = lib().orc_select_sort(reads.ctypes.data, k.ctypes.data if k is not None else None, len(reads), n_chrom, out.ctypes.data,
                                  begin.ctypes.data, span.ctypes.data, int(threads))
    return out[:n_out], begin, span


def count_arrays(genome: Genome, reads: np.ndarray, qual_pool: np.ndarray, utrs: np.ndarray,
                 min_qual: int, conv_threshold: int, snps=None,
                 mp_text: Optional[bytes] = None, mp_triples: Optional[np.ndarray] = None,
                 threads: int = 1, presorted=None):
    """reads: READ_DT array (any order; sorted here by (chrom, pos) like an indexed BAM), or presorted = (begin, max_span) from
    select_sort() when they already are.
    -> (rows ROW_DT[n_utr], bedgraph BG_DT[...])"""
    L = lib()
    n_chrom = len(genome.names)
    if presorted is not None:
        reads = np.ascontiguousarray(reads, dtype=READ_DT)
        begin = np.ascontiguousarray(presorted[0], dtype=np.int64)
        span = np.ascontiguousarray(presorted[1], dtype=np.int32)
    else:
        reads = np.ascontiguousarray(reads, dtype=READ_DT)
        order = np.lexsort((reads["pos"], reads["chrom"]))
        reads = np.ascontiguousarray(reads[order])
        begin = np.searchsorted(reads["chrom"], np.arange(n_chrom + 1), side="left").astype(np.int64)
        span = np.ones(n_chrom, dtype=np.int32)
        for c in range(n_chrom):
            seg = reads[begin[c]:begin[c + 1]]
            if len(seg):
                span[c] = max(1, int((seg["end"] - seg["pos"]).max()))
    names = (ctypes.c_char_p * n_chrom)(*[n.encode() for n in genome.names])
    seqs = (ctypes.c_char_p * n_chrom)(*genome.seqs)
    utrs = np.ascontiguousarray(utrs, dtype=UTR_DT)
    qual_pool = np.ascontiguousarray(qual_pool, dtype=np.uint8)
    a = _CountArgs()
    a.n_chrom = n_chrom
    a.chrom_names = names
    a.chrom_seq = seqs
    a.chrom_len = genome.lengths.ctypes.data
    a.reads = reads.ctypes.data
    a.chrom_read_begin = begin.ctypes.data
    a.chrom_max_span = span.ctypes.data
    a.qual_pool = qual_pool.ctypes.data if len(qual_pool) else None
    keep = [names, seqs, begin, span, reads, qual_pool, utrs]
    if mp_text is not None:
        buf = ctypes.create_string_buffer(mp_text, len(mp_text) + 1)
        keep.append(buf)
        a.mp_text = ctypes.cast(buf, ctypes.c_void_p)
        a.mp_triples = None
    else:
        tr = np.ascontiguousarray(mp_triples if mp_triples is not None else np.zeros((0, 3)), dtype=np.int32)
        keep.append(tr)
        a.mp_text = None
        a.mp_triples = tr.ctypes.data if tr.size else None
    a.n_utr = len(utrs)
    a.utrs = utrs.ctypes.data
    a.min_qual = int(min_qual)
    a.conversion_threshold = int(conv_threshold)
    a.snps = snps
    rows = np.zeros(len(utrs), dtype=ROW_DT)
    bg_ptr = ctypes.c_void_p()
    n_bg = ctypes.c_int64()
    rc = L.orc_count(ctypes.byref(a), rows.ctypes.data, ctypes.byref(bg_ptr), ctypes.byref(n_bg), int(threads))
    if rc != 0:
        raise RuntimeError("orc_count failed: %d" % rc)
    n = n_bg.value
    bg = np.frombuffer(ctypes.string_at(bg_ptr.value, n * BG_DT.itemsize), dtype=BG_DT).copy() if n else np.zeros(0, BG_DT)
    L.orc_free(bg_ptr)
    del keep
    return rows, bg


# --------------------------------------------------------------------------
# file-level restatement
# --------------------------------------------------------------------------
def md5(fname: str) -> str:
    """utils/misc.py:84-89"""
    h = hashlib.md5()
    with open(fname, "rb") as f:
        for chunk in iter(lambda: f.read(4096), b""):
            h.update(chunk)
    return h.hexdigest()


def read_bed(path: str):
    """utils/BedReader.py:63-77 -- rstrip, split on TAB, int() on columns 1 and 2, strand optional."""
    out = []
    with open(path, "r") as fh:
        for line in fh:
            cols = line.rstrip().split("\t")
            strand = cols[5] if len(cols) > 5 else "."
            out.append((cols[0], int(cols[1]), int(cols[2]), cols[3], strand))
    return out


def read_vcf_records(path: Optional[str], messages: Optional[list] = None):
    """utils/SNPtools.py:40-51 (pybedtools iteration: header and blank lines skipped)."""
    recs = []
    if path is None:
        return recs
    if not os.path.exists(path):
        if messages is not None:
            messages.append("Warning: SNP file " + path + " not found.")
        return recs
    with open(path, "r") as fh:
        for line in fh:
            if line.startswith("#") or not line.strip():
                continue
            f = line.rstrip("\n").rstrip("\r").split("\t")
            recs.append((f[0], f[1], f[3], f[4]))
    return recs


def format_row(ch, start, stop, name, strand, conv_rate, cpm, tcontent, cov, conv, rc, tc, mm) -> str:
    """SlamSeqInterval.__repr__, slamseq/SlamSeqFile.py:99-100"""
    return "\t".join([ch, str(start), str(stop), name, str(stop - start), strand, str(conv_rate), str(cpm),
                      str(tcontent), str(cov), str(conv), str(rc), str(tc), str(mm), str(-1.0), str(-1.0)])


HEADER = "\t".join(["Chromosome", "Start", "End", "Name", "Length", "Strand", "ConversionRate", "ReadsCPM",
                    "Tcontent", "CoverageOnTs", "ConversionsOnTs", "ReadCount", "TcReadCount", "multimapCount",
                    "ConversionRateLower", "ConversionRateUpper"])


def count_files(ref, bed, snps_file, bam, max_read_length, min_qual, out_csv, out_plus, out_minus,
                conv_threshold, log, threads: int = 1):
    """File-level restatement of tcounter.computeTconversions (dunks/tcounter.py:125-330)."""
    from slamdunk_b200.io import bam as bamio   # plain BAM/FASTA codec (I/O only)
    text, bam_refs, records = bamio.read_bam(bam)
    hdr = bamio.parse_header_text(text)
    if not ("RG" in hdr and len(hdr["RG"]) > 0):                     # utils/misc.py:229-236
        raise RuntimeError("Could not get mapped/unmapped/filtered read counts from BAM file. RG is missing. Please rerun slamdunk filter.")
    rg = hdr["RG"][0]
    sm = rg["SM"].split(":")                                          # utils/misc.py:238-241
    ds = ast.literal_eval(rg["DS"])                                   # utils/misc.py:67
    read_number = ds.get("filtered", "NA")
    annotation_md5 = ds.get("annotationmd5", "NA")
    bed_md5 = md5(bed)
    fasta = bamio.read_fasta(ref)
    genome = Genome(list(fasta.keys()), [s.encode("ascii") for s in fasta.values()])

    with open(out_csv, "w") as csv:
        print("#slamdunk v" + VERSION, COUNT_VERSION, "sample info:", sm[0], rg["ID"], sm[1], sm[2], sep="\t", file=csv)
        print("#annotation:", os.path.basename(bed), bed_md5, sep="\t", file=csv)
        print(HEADER, file=csv)
        msgs = []
        snp_handle = make_snps(read_vcf_records(snps_file, msgs))
        for m in msgs:
            print(m)
        try:
            bam_version = "None"                                       # SlamSeqFile.py:409-414
            for pg in hdr.get("PG", []):
                if pg.get("ID") == "slamdunk":
                    bam_version = pg.get("VN")
            if bam_version != BAM_VERSION:
                raise RuntimeError("Wrong filtered BAM file version detected (" + str(bam_version) + "). Expected version " + BAM_VERSION + ". Please rerun slamdunk filter.")
            if annotation_md5 != bed_md5:
                print("Warning: MD5 checksum of annotation (" + bed_md5 + ") does not matched MD5 in filtered BAM files (" + annotation_md5 + "). Most probably the annotation filed changed after the filtered BAM files were created.", file=log)

            bam_names = [n for n, _l in bam_refs]
            # decode reads
            n = len(records)
            reads = np.zeros(n, dtype=READ_DT)
            quals = []
            mps = []
            qoff = 0
            moff = 0
            k = 0
            for r in records:
                if r.ref_id < 0:
                    continue
                end = r.reference_end
                if end is None:
                    continue
                rec = reads[k]
                rec["chrom"] = genome.index.get(bam_names[r.ref_id], -1)
                rec["pos"] = r.pos
                rec["end"] = end
                q = bytes(r.qual) if r.qual is not None else b""
                rec["l_qual"] = len(q)
                rec["flag"] = r.flag
                rec["mapq"] = r.mapq
                rec["qual_off"] = qoff
                quals.append(q)
                qoff += len(q)
                if r.has_tag("MP"):
                    mp = r.get_tag("MP").encode("ascii")
                    rec["has_mp"] = 1
                    rec["mp_off"] = moff
                    rec["mp_len"] = len(mp)
                    mps.append(mp)
                    moff += len(mp)
                k += 1
            reads = reads[:k]
            reads = reads[reads["chrom"] >= 0]
            qual_pool = np.frombuffer(b"".join(quals), dtype=np.uint8)
            mp_text = b"".join(mps)

            bed_rows = read_bed(bed)
            name_ids: Dict[str, int] = {}
            utrs = np.zeros(len(bed_rows), dtype=UTR_DT)
            bad_row = None
            for i, (ch, start, stop, _name, strand) in enumerate(bed_rows):
                if strand not in ("+", "-"):                          # tcounter.py:169-170
                    bad_row = (i, "strand")
                    break
                if start < 0:                                         # tcounter.py:172-173
                    bad_row = (i, "start")
                    break
                utrs[i] = (genome.index.get(ch, -1), 1 if ch in bam_names else 0, start, stop,
                           1 if strand == "-" else 0, name_ids.setdefault(ch, len(name_ids)))
            n_ok = len(bed_rows) if bad_row is None else bad_row[0]
            rows, bg = count_arrays(genome, reads, qual_pool, utrs[:n_ok], min_qual, conv_threshold,
                                    snp_handle, mp_text=mp_text, threads=threads)
            for i in range(n_ok):
                ch, start, stop, name, strand = bed_rows[i]
                row = rows[i]
                if row["has_reads"]:
                    cpm = 0
                    if read_number > 0:                               # tcounter.py:295-297
                        cpm = int(row["read_count"]) * 1000000.0 / read_number
                    rate = 0
                    if row["coverage_on_ts"] > 0:                     # tcounter.py:301-303
                        rate = float(int(row["conversions_on_ts"])) / float(int(row["coverage_on_ts"]))
                    line = format_row(ch, start, stop, name, strand, rate, cpm, int(row["tcontent"]),
                                      int(row["coverage_on_ts"]), int(row["conversions_on_ts"]),
                                      int(row["read_count"]), int(row["tc_read_count"]), int(row["multimap_count"]))
                else:
                    line = format_row(ch, start, stop, name, strand, 0, 0, int(row["tcontent"]), 0, 0, 0, 0, 0)
                print(line, file=csv)
            if bad_row is not None:
                if bad_row[1] == "strand":
                    raise RuntimeError("Input BED file does not contain stranded intervals.")
                # the reference concatenates str + BedEntry here and therefore dies with a TypeError
                raise TypeError("Negativ start coordinate found. Please check the following entry in your BED file: " + str(bed_rows[bad_row[0]]))
        finally:
            free_snps(snp_handle)

    id_to_name = {v: k for k, v in name_ids.items()}
    with open(out_plus, "w") as fp, open(out_minus, "w") as fm:      # tcounter.py:315-326
        for e in bg:
            rate = int(e["tc"]) * 100.0 / int(e["cov"])               # tcounter.py:252
            out = fp if e["strand"] == 0 else fm
            print(id_to_name[int(e["name_id"])], int(e["pos"]), int(e["pos"]) + 1, rate, file=out)
    return rows, bg


def _natural_keys(text):
    """SlamSeqBamFile.natural_keys, SlamSeqFile.py:458-464"""
    import re
    return [int(c) if c.isdigit() else c for c in re.split(r"(\d+)", text)]


TRACKS = [  # (file suffix, track name suffix, description) in the order tcounter.py:349-365 opens / heads them
    ("_TC_rates_genomewide.bedGraph", " tc-conversions", "# T->C conversions / # reads on T per position genome-wide"),
    ("_AG_rates_genomewide.bedGraph", " ag-conversions", "# A->G conversions / # reads on A per position genome-wide"),
    ("_coverage_plus_genomewide.bedGraph", " plus-strand coverage", "# Reads on plus strand genome-wide"),
    ("_coverage_minus_genomewide.bedGraph", " minus-strand coverage", "# Reads on minus strand genome-wide"),
    ("_TC_conversions_genomewide.bedGraph", " T->C conversions", "# T->C conversions on plus strand genome-wide"),
    ("_AG_conversions_genomewide.bedGraph", " A->G conversions", "# A->G conversions on minus strand genome-wide"),
    ("_coverage_T_genomewide.bedGraph", " T-coverage", "# Plus-strand reads on Ts genome-wide"),
    ("_coverage_A_genomewide.bedGraph", " A-coverage", "# Minus-strand reads on As genome-wide"),
]


def _rle_lines(chromosome, values, fh):
    """The reference's run-length writer (tcounter.py:417-476) for one track of one chromosome: `values` is a Python
    list holding exactly the objects the reference compares (ints, or int 0 / floats for the rate tracks).  A line is
    written at every position whose value differs from the running one; the run in progress at the end is never written."""
    prev, prev_pos = 0, 0
    for pos, v in enumerate(values):
        if prev != v:
            print(chromosome + "\t" + str(prev_pos + 1) + "\t" + str(pos + 1) + "\t" + str(prev), file=fh)
            prev, prev_pos = v, pos


def positional_files(ref, snps_file, bam, min_qual, prefix, conv_threshold, coverage_cutoff, log, out=None):
    """File-level restatement of tcounter.genomewideConversionRates (dunks/tcounter.py:332-487): eight run-length
    bedGraphs.  Small inputs only: the per-position part runs as plain Python."""
    import re
    from slamdunk_b200.io import bam as bamio
    text, bam_refs, records = bamio.read_bam(bam)
    fasta = bamio.read_fasta(ref)
    genome = Genome(list(fasta.keys()), [s.encode("ascii") for s in fasta.values()])
    msgs = []
    snp_handle = make_snps(read_vcf_records(snps_file, msgs))
    for m in msgs:
        print(m)
    bam_names = [n for n, _l in bam_refs]
    reads = np.zeros(len(records), dtype=READ_DT)
    quals, mps = [], []
    qoff = moff = k = 0
    for r in records:
        if r.ref_id < 0 or r.reference_end is None:
            continue
        rec = reads[k]
        rec["chrom"] = genome.index.get(bam_names[r.ref_id], -1)
        rec["pos"], rec["end"] = r.pos, r.reference_end
        q = bytes(r.qual) if r.qual is not None else b""
        rec["l_qual"], rec["flag"], rec["mapq"], rec["qual_off"] = len(q), r.flag, r.mapq, qoff
        quals.append(q)
        qoff += len(q)
        if r.has_tag("MP"):
            mp = r.get_tag("MP").encode("ascii")
            rec["has_mp"], rec["mp_off"], rec["mp_len"] = 1, moff, len(mp)
            mps.append(mp)
            moff += len(mp)
        k += 1
    reads = reads[:k]
    reads = reads[reads["chrom"] >= 0]
    order = np.lexsort((reads["pos"], reads["chrom"]))
    reads = np.ascontiguousarray(reads[order])
    n_chrom = len(genome.names)
    begin = np.searchsorted(reads["chrom"], np.arange(n_chrom + 1), side="left").astype(np.int64)
    span = np.ones(n_chrom, dtype=np.int32)
    qual_pool = np.frombuffer(b"".join(quals), dtype=np.uint8)
    mp_buf = ctypes.create_string_buffer(b"".join(mps), moff + 1)
    names = (ctypes.c_char_p * n_chrom)(*[n.encode() for n in genome.names])
    seqs = (ctypes.c_char_p * n_chrom)(*genome.seqs)
    a = _CountArgs()
    a.n_chrom, a.chrom_names, a.chrom_seq, a.chrom_len = n_chrom, names, seqs, genome.lengths.ctypes.data
    a.reads, a.chrom_read_begin, a.chrom_max_span = reads.ctypes.data, begin.ctypes.data, span.ctypes.data
    a.qual_pool = qual_pool.ctypes.data if len(qual_pool) else None
    a.mp_text, a.mp_triples = ctypes.cast(mp_buf, ctypes.c_void_p), None
    a.n_utr, a.utrs = 0, None
    a.min_qual, a.conversion_threshold, a.snps = int(min_qual), int(conv_threshold), snp_handle

    info = re.sub("_slamdunk_mapped.*", "", os.path.basename(prefix))       # tcounter.py:346
    print(info, file=out)                                                    # tcounter.py:347 (stdout)
    files = [open(prefix + suffix, "w") for suffix, _n, _d in TRACKS]
    try:
        for fh, (_s, name, desc) in zip(files, TRACKS):
            print("track type=bedGraph name=\"" + info + name + "\" description=\"" + desc + "\"", file=fh)
        f_rate_p, f_rate_m, f_cov_p, f_cov_m, f_tc, f_ag, f_t, f_a = files
        for chromosome in sorted(genome.names, key=_natural_keys):          # getChromosomes, SlamSeqFile.py:469-472
            c = genome.index[chromosome]
            n = int(genome.lengths[c])
            tc, ag, covp, covm = (np.zeros(max(1, n), dtype=np.int32) for _ in range(4))
            rc = lib().orc_positional_chrom(ctypes.byref(a), c, 1 if chromosome in bam_names else 0,
                                            tc.ctypes.data, ag.ctypes.data, covp.ctypes.data, covm.ctypes.data)
            if rc != 0:
                raise RuntimeError("orc_positional_chrom failed")
            seq = genome.seqs[c].decode("ascii")
            tcl, agl, cpl, cml = tc[:n].tolist(), ag[:n].tolist(), covp[:n].tolist(), covm[:n].tolist()
            # ref.fetch(start=pos + 1, end=pos + 2): the base right of index pos; past the end the fetch is empty
            base = [seq[p + 1].upper() if p + 1 < n else "" for p in range(n)]
            t_cov = [cpl[p] if cpl[p] > 0 and base[p] == "T" else 0 for p in range(n)]          # tcounter.py:427-439
            a_cov = [cml[p] if cml[p] > 0 and base[p] == "A" else 0 for p in range(n)]
            rate_p = [float(tcl[p]) / float(cpl[p]) if cpl[p] > 0 and cpl[p] >= coverage_cutoff else 0 for p in range(n)]
            rate_m = [float(agl[p]) / float(cml[p]) if cml[p] > 0 and cml[p] >= coverage_cutoff else 0 for p in range(n)]
            _rle_lines(chromosome, cpl, f_cov_p)
            _rle_lines(chromosome, cml, f_cov_m)
            _rle_lines(chromosome, t_cov, f_t)
            _rle_lines(chromosome, a_cov, f_a)
            _rle_lines(chromosome, tcl, f_tc)
            _rle_lines(chromosome, agl, f_ag)
            _rle_lines(chromosome, rate_p, f_rate_p)
            _rle_lines(chromosome, rate_m, f_rate_m)
    finally:
        for fh in files:
            fh.close()
        free_snps(snp_handle)


def read_separation(ref, snps_file, bam, min_qual, conv_threshold):
    """Restatement of tcounter.genomewideReadSeparation (dunks/tcounter.py:488-527) up to the two record lists:
    -> (records, tc_indices, background_indices) with indices into `records` (file order), as `samFile.fetch()` yields
    them (placed records only)."""
    from slamdunk_b200.io import bam as bamio
    _text, bam_refs, records = bamio.read_bam(bam)
    fasta = bamio.read_fasta(ref)
    genome = Genome(list(fasta.keys()), [s.encode("ascii") for s in fasta.values()])
    msgs = []
    snp_handle = make_snps(read_vcf_records(snps_file, msgs))
    bam_names = [n for n, _l in bam_refs]
    idx = [i for i, r in enumerate(records) if r.ref_id >= 0 and r.reference_end is not None]
    reads = np.zeros(len(idx), dtype=READ_DT)
    quals, mps = [], []
    qoff = moff = 0
    for k, i in enumerate(idx):
        r = records[i]
        rec = reads[k]
        rec["chrom"] = genome.index.get(bam_names[r.ref_id], -1)
        rec["pos"], rec["end"] = r.pos, r.reference_end
        q = bytes(r.qual) if r.qual is not None else b""
        rec["l_qual"], rec["flag"], rec["mapq"], rec["qual_off"] = len(q), r.flag, r.mapq, qoff
        quals.append(q)
        qoff += len(q)
        if r.has_tag("MP"):
            mp = r.get_tag("MP").encode("ascii")
            rec["has_mp"], rec["mp_off"], rec["mp_len"] = 1, moff, len(mp)
            mps.append(mp)
            moff += len(mp)
    n_chrom = len(genome.names)
    qual_pool = np.frombuffer(b"".join(quals), dtype=np.uint8)
    mp_buf = ctypes.create_string_buffer(b"".join(mps), moff + 1)
    names = (ctypes.c_char_p * n_chrom)(*[n.encode() for n in genome.names])
    seqs = (ctypes.c_char_p * n_chrom)(*genome.seqs)
    a = _CountArgs()
    a.n_chrom, a.chrom_names, a.chrom_seq, a.chrom_len = n_chrom, names, seqs, genome.lengths.ctypes.data
    a.reads = reads.ctypes.data
    a.qual_pool = qual_pool.ctypes.data if len(qual_pool) else None
    a.mp_text, a.mp_triples = ctypes.cast(mp_buf, ctypes.c_void_p), None
    a.min_qual, a.conversion_threshold, a.snps = int(min_qual), int(conv_threshold), snp_handle
    tc = np.zeros(max(1, len(idx)), dtype=np.int32)
    try:
        if lib().orc_read_tc(ctypes.byref(a), len(idx), tc.ctypes.data) != 0:
            raise RuntimeError("orc_read_tc failed")
    finally:
        free_snps(snp_handle)
    tc_names = set()                                                      # tcReadDict, tcounter.py:508-516
    for k, i in enumerate(idx):
        if tc[k] >= 0 and tc[k] >= conv_threshold:                        # isTcRead, SlamSeqFile.py:397-400
            tc_names.add(records[i].qname)
    placed = [i for i, r in enumerate(records) if r.ref_id >= 0]          # samFile.fetch(), tcounter.py:518
    return records, [i for i in placed if records[i].qname in tc_names], [i for i in placed if records[i].qname not in tc_names]


# --------------------------------------------------------------------------
# filter
# --------------------------------------------------------------------------
def filter_records(records, bam_names, bed_rows, mq=2, min_identity=0.8, nm=-1):
    """Decision part of filter.Filter (dunks/filter.py:225-271) on decoded records in file order.
    -> (keep uint8[n] (0 drop, 1 write, 2 write via dumpBufferToBam), rd_tags {index: str}, stats dict)"""
    L = lib()
    n = len(records)
    fr = np.zeros(n, dtype=FREAD_DT)
    chrom_ids: Dict[str, int] = {}
    name_ids: Dict[str, int] = {}
    for i, r in enumerate(records):
        unm = bool(r.flag & 4)
        fr[i]["chrom_id"] = chrom_ids.setdefault(bam_names[r.ref_id], len(chrom_ids)) if r.ref_id >= 0 else -1
        fr[i]["start"] = r.pos
        end = r.reference_end
        fr[i]["end"] = end if end is not None else r.pos + 1
        fr[i]["name_id"] = name_ids.setdefault(r.qname, len(name_ids))
        fr[i]["flag"] = r.flag
        fr[i]["mapq"] = r.mapq
        if not unm:
            fr[i]["xi"] = np.float32(r.get_tag("XI"))
            fr[i]["nm"] = int(r.get_tag("NM")) if (nm > -1 or r.has_tag("NM")) else 0
    keep = np.zeros(n, dtype=np.uint8)
    stats = np.zeros(7, dtype=np.int64)
    rd_tags = {}
    if bed_rows is None:
        L.orc_filter_default(fr.ctypes.data, n, int(mq), float(min_identity), int(nm), keep.ctypes.data, stats.ctypes.data)
    else:
        utr_names: Dict[str, int] = {}
        fu = np.zeros(len(bed_rows), dtype=FUTR_DT)
        for i, (ch, start, stop, name, _strand) in enumerate(bed_rows):
            fu[i] = (chrom_ids.setdefault(ch, len(chrom_ids)), start, stop, utr_names.setdefault(name, len(utr_names)))
        rb = np.zeros(n, dtype=np.int64)
        re_ = np.zeros(n, dtype=np.int64)
        rl = np.zeros(max(n, 1), dtype=np.int64)
        L.orc_filter_multimap(fr.ctypes.data, n, fu.ctypes.data, len(fu), float(min_identity), int(nm),
                              keep.ctypes.data, rb.ctypes.data, re_.ctypes.data, rl.ctypes.data, stats.ctypes.data)
        for i in np.nonzero(keep == 2)[0]:
            parts = []
            for j in rl[rb[i]:re_[i]]:
                r = records[int(j)]
                parts.append("%s:%d-%d" % (bam_names[r.ref_id], r.pos, r.reference_end))
            rd_tags[int(i)] = " ".join(parts)
    names = ("mapped", "unmapped", "filtered", "mqfiltered", "idfiltered", "nmfiltered", "multimapper")
    return keep, rd_tags, dict(zip(names, (int(x) for x in stats)))


def slamseq_info_repr(sequenced, mapped, filtered, mqf, idf, nmf, mm, dedup, snps, ann, annmd5) -> str:
    """SlamSeqInfo.__repr__, utils/misc.py:81-82"""
    return ("{'sequenced':%s,'mapped':%s,'filtered':%s,'mqfiltered':%s,'idfiltered':%s,'nmfiltered':%s,"
            "'multimapper':%s,'dedup':%s,'snps':%s,'annotation':'%s','annotationmd5':'%s'}"
            % (sequenced, mapped, filtered, mqf, idf, nmf, mm, dedup, snps, ann, annmd5))
