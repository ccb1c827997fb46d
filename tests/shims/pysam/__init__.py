"""I/O-only stand-in for pysam, used ONLY to run the unmodified reference
(/root/reference/slamdunk) as the parity oracle.

It implements the part of the pysam surface the reference's count/filter path
touches and nothing else; every arithmetic decision stays in the reference code.
htslib semantics mirrored here: indexed fetch returns records with
pos < end and bam_endpos > start in coordinate order without flag filtering;
reference_end = pos + max(1, sum of M/D/N/=/X), None when unmapped / no CIGAR;
query_sequence / query_qualities include soft clips.
"""
import bisect
import os

from slamdunk_b200.io import bam as _bam

__version__ = "0.0-shim"


class AlignmentHeader(dict):
    def to_dict(self):
        return dict(self)


class AlignedSegment(object):
    __slots__ = ("_r", "_af")

    def __init__(self, rec, af):
        self._r = rec
        self._af = af

    query_name = property(lambda s: s._r.qname)
    query_sequence = property(lambda s: s._r.seq)
    query_qualities = property(lambda s: s._r.qual)
    query_length = property(lambda s: len(s._r.seq))
    mapping_quality = property(lambda s: s._r.mapq)
    reference_start = property(lambda s: s._r.pos)
    reference_end = property(lambda s: s._r.reference_end)
    reference_id = property(lambda s: s._r.ref_id)
    flag = property(lambda s: s._r.flag)
    cigartuples = property(lambda s: list(s._r.cigar))
    is_reverse = property(lambda s: bool(s._r.flag & _bam.FLAG_REVERSE))
    is_unmapped = property(lambda s: bool(s._r.flag & _bam.FLAG_UNMAPPED))

    @property
    def reference_name(self):
        return self._af.getrname(self._r.ref_id) if self._r.ref_id >= 0 else None

    @property
    def query_alignment_sequence(self):
        cig = self._r.cigar
        seq = self._r.seq
        start, end = 0, len(seq)
        if cig and cig[0][0] == 4:
            start = cig[0][1]
        elif len(cig) > 1 and cig[0][0] == 5 and cig[1][0] == 4:
            start = cig[1][1]
        if cig and cig[-1][0] == 4:
            end -= cig[-1][1]
        elif len(cig) > 1 and cig[-1][0] == 5 and cig[-2][0] == 4:
            end -= cig[-2][1]
        return seq[start:end]

    def _getflag(self, bit):
        return bool(self._r.flag & bit)

    def _setflag(self, bit, v):
        if v:
            self._r.flag |= bit
        else:
            self._r.flag &= ~bit

    is_secondary = property(lambda s: s._getflag(_bam.FLAG_SECONDARY),
                            lambda s, v: s._setflag(_bam.FLAG_SECONDARY, v))
    is_supplementary = property(lambda s: s._getflag(_bam.FLAG_SUPPLEMENTARY),
                                lambda s, v: s._setflag(_bam.FLAG_SUPPLEMENTARY, v))

    def has_tag(self, tag):
        return self._r.has_tag(tag)

    def get_tag(self, tag):
        return self._r.get_tag(tag)

    def set_tag(self, tag, value, value_type=None, replace=True):
        self._r.set_tag(tag, value, value_type)


class AlignmentFile(object):
    def __init__(self, filename, mode="rb", header=None, template=None, **kw):
        self.filename = filename
        self.mode = mode
        self._records = []
        self._index = None
        if "w" in mode:
            if template is not None:
                self.header = AlignmentHeader(_copy_header(template.header))
                self._refs = list(template._refs)
            elif header is not None:
                hd = header.to_dict() if hasattr(header, "to_dict") else dict(header)
                self.header = AlignmentHeader(_copy_header(hd))
                self._refs = [(sq["SN"], int(sq["LN"])) for sq in self.header.get("SQ", [])]
            else:
                raise ValueError("need header or template for writing")
        else:
            text, refs, recs = _bam.read_bam(filename)
            self.header = AlignmentHeader(_bam.parse_header_text(text))
            self._refs = refs
            self._records = recs
        self.references = tuple(n for n, _ in self._refs)
        self.lengths = tuple(l for _, l in self._refs)
        self.nreferences = len(self._refs)

    # -- reading -----------------------------------------------------------
    def _build_index(self):
        idx = {}
        for r in self._records:
            if r.ref_id < 0:
                continue
            idx.setdefault(r.ref_id, []).append(r)
        self._index = {}
        for rid, recs in idx.items():
            recs.sort(key=lambda r: r.pos)   # stable: keeps file order among ties
            pos = [r.pos for r in recs]
            span = 1
            for r in recs:
                e = r.reference_end
                if e is not None and e - r.pos > span:
                    span = e - r.pos
            self._index[rid] = (pos, recs, span)

    def fetch(self, reference=None, start=None, end=None, region=None, until_eof=False, contig=None):
        if reference is None and contig is not None:
            reference = contig
        if reference is None and region is not None:       # "chr" or "chr:start-end" (1-based, inclusive), as samtools writes regions
            reference = region
            if ":" in region and region not in self.references:
                reference, span = region.rsplit(":", 1)
                a, _, b = span.replace(",", "").partition("-")
                start = int(a) - 1
                end = int(b) if b else None
        if reference is None:
            return iter([AlignedSegment(r, self) for r in self._records])
        if reference not in self.references:
            raise ValueError("invalid contig `%s`" % reference)
        rid = self.references.index(reference)
        if self._index is None:
            self._build_index()
        if rid not in self._index:
            return iter([])
        pos, recs, span = self._index[rid]
        if start is None:
            start = 0
        if end is None:
            end = self.lengths[rid]
        if start > end:
            raise ValueError("invalid coordinates: start (%d) > end (%d)" % (start, end))
        lo = bisect.bisect_left(pos, start - span)
        hi = bisect.bisect_left(pos, end)

        def gen():
            for i in range(lo, hi):
                r = recs[i]
                e = r.reference_end
                if e is None:
                    e = r.pos + 1
                if e > start:
                    yield AlignedSegment(r, self)
        return gen()

    def head(self, n):
        return iter([AlignedSegment(r, self) for r in self._records[:n]])

    def __iter__(self):
        return iter([AlignedSegment(r, self) for r in self._records])

    def getrname(self, rid):
        return self.references[rid]

    get_reference_name = getrname

    # -- writing -----------------------------------------------------------
    def write(self, read):
        self._records.append(read._r.copy())

    def close(self):
        if "w" in self.mode and self._records is not None:
            _bam.write_bam(self.filename, _bam.header_dict_to_text(self.header), self._refs, self._records)
            self._records = None


def _copy_header(h):
    out = {}
    for k, v in h.items():
        if isinstance(v, list):
            out[k] = [dict(x) if isinstance(x, dict) else x for x in v]
        elif isinstance(v, dict):
            out[k] = dict(v)
        else:
            out[k] = v
    return out


class FastaFile(object):
    def __init__(self, filename):
        self.filename = filename
        self._seqs = _bam.read_fasta(filename)
        self.references = tuple(self._seqs.keys())
        self.lengths = tuple(len(s) for s in self._seqs.values())
        self.nreferences = len(self._seqs)

    def fetch(self, reference=None, start=None, end=None, region=None):
        if reference not in self._seqs:
            raise KeyError("sequence '%s' not present" % reference)
        s = self._seqs[reference]
        if start is None:
            start = 0
        if end is None:
            end = len(s)
        if start < 0:
            raise ValueError("start out of range (%i)" % start)
        if start > end:
            raise ValueError("invalid coordinates: start (%d) > end (%d)" % (start, end))
        return s[start:end]

    def get_reference_length(self, reference):
        return len(self._seqs[reference])

    def close(self):
        pass


Fastafile = FastaFile
Samfile = AlignmentFile


def index(path, *args):
    """The reference indexes its output; the shim's fetch needs no index file."""
    return ""


def sort(*args):
    raise NotImplementedError("pysam.sort is not part of the shim")
