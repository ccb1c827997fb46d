"""Pure-Python BGZF/BAM codec (struct + zlib only).

Small, slow and complete: it writes the BAM files that the tests and the golden-vector
generator feed to both the reference (through tests/shims/pysam) and the native decoder
(slamdunk_b200/csrc/sd_decode.cpp), reads them back so the native decoder can be cross-checked
record by record, and is what `dunks/filter.py` uses to write the filtered BAM until the native
BAM writer exists.

The count path never imports this module: its BAM input goes through the C++ decoder.

Format follows the SAM/BAM specification (SAMv1 section 4): BGZF blocks are gzip
members carrying a 'BC' extra sub-field, the BAM stream is little-endian.
"""
from __future__ import annotations

import struct
import zlib
from typing import Dict, Iterable, Iterator, List, Optional, Sequence, Tuple

CIGAR_OPS = "MIDNSHP=X"
_CIGAR_CODE = {c: i for i, c in enumerate(CIGAR_OPS)}
SEQ_NIBBLES = "=ACMGRSVTWYHKDBN"
_SEQ_CODE = {c: i for i, c in enumerate(SEQ_NIBBLES)}
# which CIGAR ops consume reference / query
_CONSUMES_REF = (1, 0, 1, 1, 0, 0, 0, 1, 1)
_CONSUMES_QRY = (1, 1, 0, 0, 1, 0, 0, 1, 1)

FLAG_PAIRED = 0x1
FLAG_UNMAPPED = 0x4
FLAG_REVERSE = 0x10
FLAG_SECONDARY = 0x100
FLAG_SUPPLEMENTARY = 0x800

_BGZF_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")
_BGZF_MAX_PAYLOAD = 0xFF00


# --------------------------------------------------------------------------
# BGZF
# --------------------------------------------------------------------------
def bgzf_block(payload: bytes, level: int = 6) -> bytes:
    """One BGZF block (a gzip member with the BC extra field) for <=64 KiB of payload."""
    assert len(payload) <= 0x10000
    comp = zlib.compressobj(level, zlib.DEFLATED, -15)
    cdata = comp.compress(payload) + comp.flush()
    bsize = len(cdata) + 25  # total block size - 1
    assert bsize < 0x10000
    head = struct.pack("<BBBBIBBHBBHH", 0x1F, 0x8B, 8, 4, 0, 0, 0xFF, 6, 66, 67, 2, bsize)
    tail = struct.pack("<II", zlib.crc32(payload) & 0xFFFFFFFF, len(payload))
    return head + cdata + tail


def bgzf_compress(data: bytes, level: int = 6, block_payload: int = _BGZF_MAX_PAYLOAD) -> bytes:
    out = []
    for i in range(0, len(data), block_payload):
        out.append(bgzf_block(data[i:i + block_payload], level))
    out.append(_BGZF_EOF)
    return b"".join(out)


def bgzf_decompress(blob: bytes) -> bytes:
    """Inflate a whole BGZF file image."""
    out = []
    off = 0
    n = len(blob)
    while off < n:
        if blob[off:off + 4] != b"\x1f\x8b\x08\x04":
            raise ValueError("not a BGZF block at offset %d" % off)
        xlen = struct.unpack_from("<H", blob, off + 10)[0]
        # find the BC sub-field
        x = off + 12
        bsize = None
        while x < off + 12 + xlen:
            si1, si2, slen = struct.unpack_from("<BBH", blob, x)
            if si1 == 66 and si2 == 67 and slen == 2:
                bsize = struct.unpack_from("<H", blob, x + 4)[0]
            x += 4 + slen
        if bsize is None:
            raise ValueError("BGZF block without BC field")
        cstart = off + 12 + xlen
        cend = off + bsize + 1 - 8
        payload = zlib.decompress(blob[cstart:cend], -15)
        crc, isize = struct.unpack_from("<II", blob, cend)
        if len(payload) != isize or (zlib.crc32(payload) & 0xFFFFFFFF) != crc:
            raise ValueError("BGZF block checksum mismatch")
        out.append(payload)
        off += bsize + 1
    return b"".join(out)


# --------------------------------------------------------------------------
# records
# --------------------------------------------------------------------------
def parse_cigar_string(cigar: str) -> List[Tuple[int, int]]:
    """'2S34M' -> [(4, 2), (0, 34)]  (op code, length) like pysam cigartuples."""
    out = []
    num = ""
    for ch in cigar:
        if ch.isdigit():
            num += ch
        else:
            out.append((_CIGAR_CODE[ch], int(num)))
            num = ""
    if num:
        raise ValueError("dangling number in CIGAR " + cigar)
    return out


def cigar_to_string(cigar: Sequence[Tuple[int, int]]) -> str:
    return "".join("%d%s" % (l, CIGAR_OPS[op]) for op, l in cigar) if cigar else "*"


def reg2bin(beg: int, end: int) -> int:
    end -= 1
    if beg >> 14 == end >> 14:
        return ((1 << 15) - 1) // 7 + (beg >> 14)
    if beg >> 17 == end >> 17:
        return ((1 << 12) - 1) // 7 + (beg >> 17)
    if beg >> 20 == end >> 20:
        return ((1 << 9) - 1) // 7 + (beg >> 20)
    if beg >> 23 == end >> 23:
        return ((1 << 6) - 1) // 7 + (beg >> 23)
    if beg >> 26 == end >> 26:
        return ((1 << 3) - 1) // 7 + (beg >> 26)
    return 0


class BamRecord:
    """One alignment. Field names follow the BAM spec; helpers follow pysam's semantics."""

    __slots__ = ("qname", "flag", "ref_id", "pos", "mapq", "cigar", "next_ref_id", "next_pos",
                 "tlen", "seq", "qual", "tags")

    def __init__(self, qname: str, flag: int, ref_id: int, pos: int, mapq: int,
                 cigar, seq: str, qual: Optional[Sequence[int]], tags: Optional[list] = None,
                 next_ref_id: int = -1, next_pos: int = -1, tlen: int = 0):
        self.qname = qname
        self.flag = flag
        self.ref_id = ref_id
        self.pos = pos
        self.mapq = mapq
        self.cigar = parse_cigar_string(cigar) if isinstance(cigar, str) else list(cigar or [])
        self.seq = seq
        self.qual = None if qual is None else list(qual)
        # tags: list of (tag, type_char, value) preserving file order
        self.tags = list(tags or [])
        self.next_ref_id = next_ref_id
        self.next_pos = next_pos
        self.tlen = tlen

    # -- derived (htslib semantics) -------------------------------------
    @property
    def reference_length(self) -> int:
        return sum(l for op, l in self.cigar if _CONSUMES_REF[op])

    @property
    def reference_end(self) -> Optional[int]:
        """htslib bam_endpos as pysam exposes it: None when unmapped / no CIGAR,
        pos + max(1, reference length) otherwise."""
        if (self.flag & FLAG_UNMAPPED) or not self.cigar:
            return None
        rl = self.reference_length
        return self.pos + (rl if rl else 1)

    def get_tag(self, tag: str):
        for t, _ty, v in self.tags:
            if t == tag:
                return v
        raise KeyError("tag '%s' not present" % tag)

    def has_tag(self, tag: str) -> bool:
        return any(t == tag for t, _ty, _v in self.tags)

    def set_tag(self, tag: str, value, value_type: Optional[str] = None) -> None:
        self.tags = [x for x in self.tags if x[0] != tag]
        if value is None:
            return
        if value_type is None:
            value_type = "Z" if isinstance(value, str) else ("f" if isinstance(value, float) else "i")
        self.tags.append((tag, value_type, value))

    def copy(self) -> "BamRecord":
        return BamRecord(self.qname, self.flag, self.ref_id, self.pos, self.mapq, list(self.cigar),
                         self.seq, self.qual, list(self.tags), self.next_ref_id, self.next_pos, self.tlen)


def _pack_int_tag(tag: bytes, v: int) -> bytes:
    if v >= 0:
        if v <= 0xFF:
            return tag + b"C" + struct.pack("<B", v)
        if v <= 0xFFFF:
            return tag + b"S" + struct.pack("<H", v)
        return tag + b"I" + struct.pack("<I", v)
    if v >= -0x80:
        return tag + b"c" + struct.pack("<b", v)
    if v >= -0x8000:
        return tag + b"s" + struct.pack("<h", v)
    return tag + b"i" + struct.pack("<i", v)


_B_FMT = {"c": "b", "C": "B", "s": "h", "S": "H", "i": "i", "I": "I", "f": "f"}


def encode_tags(tags) -> bytes:
    out = []
    for tag, ty, v in tags:
        t = tag.encode("ascii")
        if ty == "Z":
            out.append(t + b"Z" + str(v).encode("ascii") + b"\0")
        elif ty == "A":
            out.append(t + b"A" + str(v).encode("ascii")[:1])
        elif ty == "f":
            out.append(t + b"f" + struct.pack("<f", float(v)))
        elif ty == "H":
            out.append(t + b"H" + str(v).encode("ascii") + b"\0")
        elif ty in "cCsSiI" and len(ty) == 1:
            if ty == "i":  # 'i' as the generic SAM integer type: pick the smallest BAM type
                out.append(_pack_int_tag(t, int(v)))
            else:
                out.append(t + ty.encode() + struct.pack("<" + _B_FMT[ty], int(v)))
        elif ty.startswith("B"):
            sub = ty[1] if len(ty) > 1 else "i"
            arr = list(v)
            out.append(t + b"B" + sub.encode() + struct.pack("<I", len(arr))
                       + struct.pack("<%d%s" % (len(arr), _B_FMT[sub]), *arr))
        else:
            raise ValueError("unsupported tag type %r" % (ty,))
    return b"".join(out)


def decode_tags(buf: bytes, off: int, end: int) -> list:
    tags = []
    while off < end:
        tag = buf[off:off + 2].decode("ascii")
        ty = chr(buf[off + 2])
        off += 3
        if ty == "Z" or ty == "H":
            z = buf.index(b"\0", off)
            tags.append((tag, ty, buf[off:z].decode("ascii")))
            off = z + 1
        elif ty == "A":
            tags.append((tag, "A", chr(buf[off])))
            off += 1
        elif ty == "f":
            tags.append((tag, "f", struct.unpack_from("<f", buf, off)[0]))
            off += 4
        elif ty in _B_FMT:
            fmt = _B_FMT[ty]
            size = struct.calcsize(fmt)
            tags.append((tag, ty, struct.unpack_from("<" + fmt, buf, off)[0]))
            off += size
        elif ty == "B":
            sub = chr(buf[off])
            n = struct.unpack_from("<I", buf, off + 1)[0]
            fmt = _B_FMT[sub]
            size = struct.calcsize(fmt)
            vals = list(struct.unpack_from("<%d%s" % (n, fmt), buf, off + 5))
            tags.append((tag, "B" + sub, vals))
            off += 5 + n * size
        else:
            raise ValueError("unknown tag type %r" % ty)
    return tags


def encode_record(r: BamRecord) -> bytes:
    name = r.qname.encode("ascii") + b"\0"
    l_seq = len(r.seq) if r.seq not in (None, "*") else 0
    seq = r.seq if l_seq else ""
    nib = [_SEQ_CODE.get(c.upper(), 15) for c in seq]
    if len(nib) & 1:
        nib.append(0)
    packed = bytes((nib[i] << 4) | nib[i + 1] for i in range(0, len(nib), 2))
    if r.qual is None:
        qual = b"\xff" * l_seq
    else:
        assert len(r.qual) == l_seq, "qual length != seq length"
        qual = bytes(r.qual)
    cig = b"".join(struct.pack("<I", (l << 4) | op) for op, l in r.cigar)
    end = r.reference_end
    b = reg2bin(r.pos, end if end is not None else r.pos + 1) if r.pos >= 0 else 4680
    core = struct.pack("<iiBBHHHiiii", r.ref_id, r.pos, len(name), r.mapq, b, len(r.cigar), r.flag,
                       l_seq, r.next_ref_id, r.next_pos, r.tlen)
    body = core + name + cig + packed + qual + encode_tags(r.tags)
    return struct.pack("<i", len(body)) + body


def decode_record(buf: bytes, off: int) -> Tuple[BamRecord, int]:
    (block_size,) = struct.unpack_from("<i", buf, off)
    s = off + 4
    ref_id, pos, l_name, mapq, _bin, n_cig, flag, l_seq, nref, npos, tlen = struct.unpack_from("<iiBBHHHiiii", buf, s)
    p = s + 32
    qname = buf[p:p + l_name - 1].decode("ascii")
    p += l_name
    cigar = []
    for i in range(n_cig):
        (v,) = struct.unpack_from("<I", buf, p + 4 * i)
        cigar.append((v & 0xF, v >> 4))
    p += 4 * n_cig
    nb = (l_seq + 1) // 2
    sb = buf[p:p + nb]
    seq = "".join(SEQ_NIBBLES[x >> 4] + SEQ_NIBBLES[x & 0xF] for x in sb)[:l_seq]
    p += nb
    qb = buf[p:p + l_seq]
    qual = None if (l_seq and qb[0] == 0xFF) else list(qb)
    p += l_seq
    tags = decode_tags(buf, p, s + block_size)
    return BamRecord(qname, flag, ref_id, pos, mapq, cigar, seq, qual, tags, nref, npos, tlen), s + block_size


# --------------------------------------------------------------------------
# SAM text header <-> dict (pysam AlignmentHeader.to_dict() layout)
# --------------------------------------------------------------------------
def parse_header_text(text: str) -> Dict[str, list]:
    hdr: Dict[str, list] = {}
    for line in text.split("\n"):
        if not line.startswith("@") or len(line) < 3:
            continue
        kind = line[1:3]
        if kind == "CO":
            hdr.setdefault("CO", []).append(line[4:])
            continue
        fields = {}
        for f in line.split("\t")[1:]:
            if len(f) >= 3 and f[2] == ":":
                fields[f[:2]] = f[3:]
        if kind == "HD":
            hdr["HD"] = fields
        else:
            if kind == "SQ" and "LN" in fields:
                fields["LN"] = int(fields["LN"])
            hdr.setdefault(kind, []).append(fields)
    return hdr


def header_dict_to_text(hdr: Dict[str, list]) -> str:
    lines = []
    if "HD" in hdr:
        lines.append("@HD\t" + "\t".join("%s:%s" % kv for kv in hdr["HD"].items()))
    for kind in ("SQ", "RG", "PG"):
        for rec in hdr.get(kind, []):
            lines.append("@" + kind + "\t" + "\t".join("%s:%s" % (k, v) for k, v in rec.items()))
    for kind, recs in hdr.items():
        if kind in ("HD", "SQ", "RG", "PG", "CO"):
            continue
        for rec in recs:
            lines.append("@" + kind + "\t" + "\t".join("%s:%s" % (k, v) for k, v in rec.items()))
    for co in hdr.get("CO", []):
        lines.append("@CO\t" + co)
    return "\n".join(lines) + "\n" if lines else ""


# --------------------------------------------------------------------------
# whole files
# --------------------------------------------------------------------------
def encode_bam(header_text: str, references: Sequence[Tuple[str, int]], records: Iterable[BamRecord]) -> bytes:
    """Uncompressed BAM stream."""
    text = header_text.encode("ascii")
    parts = [b"BAM\1", struct.pack("<i", len(text)), text, struct.pack("<i", len(references))]
    for name, length in references:
        nm = name.encode("ascii") + b"\0"
        parts.append(struct.pack("<i", len(nm)) + nm + struct.pack("<i", length))
    for r in records:
        parts.append(encode_record(r))
    return b"".join(parts)


def write_bam(path: str, header_text: str, references: Sequence[Tuple[str, int]],
              records: Iterable[BamRecord], level: int = 6) -> None:
    with open(path, "wb") as fh:
        fh.write(bgzf_compress(encode_bam(header_text, references, records), level))


def read_bam(path: str):
    """-> (header_text, [(name, length)], [BamRecord])"""
    with open(path, "rb") as fh:
        raw = bgzf_decompress(fh.read())
    if raw[:4] != b"BAM\1":
        raise ValueError("not a BAM file: " + path)
    (l_text,) = struct.unpack_from("<i", raw, 4)
    text = raw[8:8 + l_text].split(b"\0", 1)[0].decode("ascii")
    off = 8 + l_text
    (n_ref,) = struct.unpack_from("<i", raw, off)
    off += 4
    refs = []
    for _ in range(n_ref):
        (l_name,) = struct.unpack_from("<i", raw, off)
        name = raw[off + 4:off + 4 + l_name - 1].decode("ascii")
        (l_ref,) = struct.unpack_from("<i", raw, off + 4 + l_name)
        refs.append((name, l_ref))
        off += 8 + l_name
    records = []
    n = len(raw)
    while off < n:
        rec, off = decode_record(raw, off)
        records.append(rec)
    return text, refs, records


def head_records(path: str, n: int = 1000) -> List[BamRecord]:
    """First n alignment records, inflating only as many BGZF blocks as needed (pysam's AlignmentFile.head)."""
    out: List[BamRecord] = []
    buf = b""
    off = 0
    header_done = False
    with open(path, "rb") as fh:
        while len(out) < n:
            head = fh.read(18)
            if len(head) < 18:
                break
            xlen = struct.unpack_from("<H", head, 10)[0]
            extra = head[12:] + fh.read(xlen - 6)
            bsize = None
            x = 0
            while x + 4 <= len(extra):
                si1, si2, slen = struct.unpack_from("<BBH", extra, x)
                if si1 == 66 and si2 == 67:
                    bsize = struct.unpack_from("<H", extra, x + 4)[0]
                x += 4 + slen
            if bsize is None:
                raise ValueError("not a BGZF file: " + path)
            rest = fh.read(bsize + 1 - 12 - xlen)
            buf += zlib.decompress(rest[:-8], -15)
            if not header_done:
                if len(buf) < 12:
                    continue
                l_text = struct.unpack_from("<i", buf, 4)[0]
                if len(buf) < 12 + l_text:
                    continue
                p = 8 + l_text
                n_ref = struct.unpack_from("<i", buf, p)[0]
                p += 4
                ok = True
                for _ in range(n_ref):
                    if len(buf) < p + 4:
                        ok = False
                        break
                    l_name = struct.unpack_from("<i", buf, p)[0]
                    if len(buf) < p + 8 + l_name:
                        ok = False
                        break
                    p += 8 + l_name
                if not ok:
                    continue
                off = p
                header_done = True
            while len(out) < n and len(buf) >= off + 4:
                bs = struct.unpack_from("<i", buf, off)[0]
                if len(buf) < off + 4 + bs:
                    break
                rec, off = decode_record(buf, off)
                out.append(rec)
    return out


def sort_records(records: List[BamRecord]) -> List[BamRecord]:
    """Coordinate sort like `samtools sort` (unmapped ref_id=-1 last; ties keep input order)."""
    def key(r: BamRecord):
        rid = r.ref_id if r.ref_id >= 0 else 1 << 30
        return (rid, r.pos, 1 if r.flag & FLAG_REVERSE else 0)
    return sorted(records, key=key)


def read_fasta(path: str) -> Dict[str, str]:
    """FASTA -> ordered {name: sequence} (case preserved, name = first word of the header)."""
    seqs: Dict[str, List[str]] = {}
    cur = None
    with open(path, "r") as fh:
        for line in fh:
            if line.startswith(">"):
                cur = line[1:].split()[0] if line[1:].split() else ""
                seqs[cur] = []
            elif cur is not None:
                seqs[cur].append(line.strip())
    return {k: "".join(v) for k, v in seqs.items()}


def write_fasta(path: str, seqs: Dict[str, str], width: int = 60) -> None:
    with open(path, "w") as fh:
        for name, s in seqs.items():
            fh.write(">" + name + "\n")
            for i in range(0, len(s), width):
                fh.write(s[i:i + width] + "\n")
