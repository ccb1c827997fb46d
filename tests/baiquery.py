"""Independent reader of BAI indexes (SAM spec 5.2) and an indexed region fetch over a BGZF/BAM file, for testing the
native index writer: fetch(region) through the index must equal a linear scan."""
import struct
import zlib

from slamdunk_b200.io import bam as bamio


def read_bai(path):
    b = open(path, "rb").read()
    assert b[:4] == b"BAI\1"
    off = 4
    (n_ref,) = struct.unpack_from("<i", b, off)
    off += 4
    refs = []
    for _ in range(n_ref):
        (n_bin,) = struct.unpack_from("<i", b, off)
        off += 4
        bins = {}
        for _ in range(n_bin):
            bin_id, n_chunk = struct.unpack_from("<Ii", b, off)
            off += 8
            chunks = [struct.unpack_from("<QQ", b, off + 16 * k) for k in range(n_chunk)]
            off += 16 * n_chunk
            assert bin_id not in bins
            bins[bin_id] = chunks
        (n_intv,) = struct.unpack_from("<i", b, off)
        off += 4
        linear = list(struct.unpack_from("<%dQ" % n_intv, b, off))
        off += 8 * n_intv
        refs.append((bins, linear))
    n_no_coor = struct.unpack_from("<Q", b, off)[0] if off + 8 <= len(b) else None
    assert off + (8 if n_no_coor is not None else 0) == len(b)
    return refs, n_no_coor


def reg2bins(beg, end):
    """all bins that may hold alignments overlapping [beg, end) (spec 5.3)"""
    end -= 1
    out = [0]
    for shift, base in ((26, 1), (23, 9), (20, 73), (17, 585), (14, 4681)):
        out.extend(range(base + (beg >> shift), base + (end >> shift) + 1))
    return out


class Bgzf:
    def __init__(self, path):
        self.data = open(path, "rb").read()
        self.cache = {}

    def block(self, coff):
        if coff not in self.cache:
            d = self.data
            assert d[coff:coff + 4] == b"\x1f\x8b\x08\x04", "virtual offset does not point at a BGZF block"
            xlen = struct.unpack_from("<H", d, coff + 10)[0]
            bsize = struct.unpack_from("<H", d, coff + 16)[0] + 1
            self.cache[coff] = (zlib.decompress(d[coff + 12 + xlen:coff + bsize - 8], -15), bsize)
        return self.cache[coff]

    def read(self, voff, n):
        """n bytes from virtual offset -> (bytes, next virtual offset)"""
        coff, uoff = voff >> 16, voff & 0xFFFF
        out = b""
        while n:
            blk, bsize = self.block(coff)
            take = min(n, len(blk) - uoff)
            out += blk[uoff:uoff + take]
            n -= take
            uoff += take
            if uoff == len(blk) and n:
                coff, uoff = coff + bsize, 0
        if uoff == len(self.block(coff)[0]):
            coff, uoff = coff + self.block(coff)[1], 0
        return out, (coff << 16) | uoff


def fetch(bam_path, bai, ref_id, beg, end):
    """(qname, flag, pos) of the records on ref_id overlapping [beg, end), through the index"""
    bins, linear = bai[ref_id]
    min_off = linear[min(beg >> 14, len(linear) - 1)] if linear else 0
    chunks = []
    for bn in reg2bins(beg, end):
        for c0, c1 in bins.get(bn, ()):
            if c1 > min_off:
                chunks.append((c0, c1))
    chunks.sort()
    z = Bgzf(bam_path)
    seen, out = set(), []
    for c0, c1 in chunks:
        v = c0
        while v < c1:
            head, v2 = z.read(v, 4)
            (bs,) = struct.unpack("<i", head)
            body, v3 = z.read(v2, bs)
            rid, pos = struct.unpack_from("<ii", body, 0)
            l_name, n_cig, flag = body[8], struct.unpack_from("<H", body, 12)[0], struct.unpack_from("<H", body, 14)[0]
            reflen = 0
            if not flag & 4:
                for k in range(n_cig):
                    (op,) = struct.unpack_from("<I", body, 32 + l_name + 4 * k)
                    if (op & 15) in (0, 2, 3, 7, 8):
                        reflen += op >> 4
            rend = pos + max(1, reflen)
            if rid == ref_id and pos < end and rend > beg and v not in seen:
                seen.add(v)
                out.append((body[32:32 + l_name - 1].decode(), flag, pos))
            v = v3
    return sorted(out, key=lambda t: (t[2], t[0], t[1]))


def scan(bam_path, ref_id, beg, end):
    _t, _r, recs = bamio.read_bam(bam_path)
    out = []
    for r in recs:
        if r.ref_id != ref_id:
            continue
        rend = r.pos + (max(1, r.reference_length) if not r.flag & 4 else 1)
        if r.pos < end and rend > beg:
            out.append((r.qname, r.flag, r.pos))
    return sorted(out, key=lambda t: (t[2], t[0], t[1]))
