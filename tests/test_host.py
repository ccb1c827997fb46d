"""CPU-only checks of the host side: the C-ABI library loads and exports what include/*.h
declares, the native BAM / FASTA decoders agree with the plain Python codec, the SNP masks
reproduce the reference's chrom+pos string-key semantics."""
import math
import os
import re

import numpy as np
import pytest

import goldens
import simdata
from slamdunk_b200 import _lib as L
from slamdunk_b200.io import bam as bamio
from slamdunk_b200.utils import snpmask as SNPtools

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = L.load()
    text = open(os.path.join(ROOT, "include", "slamdunk_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    declared = set(re.findall(r"\b(sd_[a-z_0-9]+)\s*\(", text))
    assert declared, "no declarations found"
    for name in sorted(declared):
        assert hasattr(lib, name), "library does not export " + name
    assert declared == set(L.PROTOTYPES.keys())
    assert lib.sd_abi_version() == L.SD_ABI_VERSION == 9


def _decode(path, fasta_names, want_mp=True, tile_reads=32, threads=3):
    from slamdunk_b200.engine import HostBatch
    return HostBatch(path, fasta_names, want_mp=want_mp, tile_reads=tile_reads, threads=threads)


def _unslice(seq, tile_word_off, row, lseq, tile_reads=32):
    """bases of one read from the group-major plane-sliced tile block (include/slamdunk_b200.h, sd_batch.seq)"""
    out = []
    for i in range(lseq):
        g = tile_word_off + 4 * ((i >> 5) * tile_reads + row)
        nib = sum((((int(seq[g + p]) >> (i & 31)) & 1) << p) for p in range(4))
        out.append(bamio.SEQ_NIBBLES[nib])
    return "".join(out)


@pytest.mark.parametrize("tile_reads", [0, 32])
def test_bam_decoder_matches_python_codec(tmp_path, tile_reads):
    c = simdata.make_case(str(tmp_path), seed=77, n_reads=700, complexity=2.0)
    text, refs, recs = bamio.read_bam(c.bam)
    fasta_names = list(c.genome.keys())
    hb = _decode(c.bam, fasta_names, tile_reads=tile_reads)
    assert hb.n_reads == len(recs)
    assert hb.header_text == text
    assert hb.ref_names == [n for n, _ in refs] and hb.ref_lengths == [l for _, l in refs]
    rec = hb.rec_array()
    tiles = hb.tiles_array()
    seq, qual, cig, mp = hb.column("seq"), hb.column("qual"), hb.column("cigar"), hb.column("mp")
    assert (tiles % 16 == 0).all()
    T = 32
    prev_name = None
    for t in range(hb.batch.n_tiles):
        so, qo, co, mo = (int(x) for x in tiles[t])
        so //= 4; co //= 4; mo //= 4
        for k in range(T):
            i = t * T + k
            w = rec[i]
            flags = int(w[1]) >> 24
            if i >= len(recs):
                assert flags & L.SD_F_PAD
                continue
            r = recs[i]
            lseq, ncig = int(w[3]) & 0xFFFF, int(w[3]) >> 16
            nm, nmp = int(w[4]) & 0xFFFF, int(w[4]) >> 16
            assert np.int32(w[0]) == r.pos
            assert lseq == len(r.seq) and ncig == len(r.cigar)
            assert (int(w[1]) >> 16) & 0xFF == r.mapq
            ref = int(w[1]) & 0xFFFF
            bam_name = refs[r.ref_id][0]
            assert ref == (fasta_names.index(bam_name) if bam_name in fasta_names else L.SD_REF_NONE)
            assert bool(flags & L.SD_F_REVERSE) == bool(r.flag & 16)
            assert bool(flags & L.SD_F_SECONDARY) == bool(r.flag & 0x100)
            assert bool(flags & L.SD_F_HAS_MP) == r.has_tag("MP")
            assert bool(flags & L.SD_F_NEWNAME) == (r.qname != prev_name)
            assert not (flags & L.SD_F_PAD)
            prev_name = r.qname
            assert nm == r.get_tag("NM")
            assert np.float32(np.uint32(w[2]).view(np.float32)) == np.float32(r.get_tag("XI"))
            assert _unslice(seq, so, k, lseq) == r.seq.upper()
            assert list(qual[qo:qo + lseq]) == r.qual
            assert [(int(x) & 15, int(x) >> 4) for x in cig[co:co + ncig]] == r.cigar
            want = []
            if r.has_tag("MP"):
                for ent in r.get_tag("MP").split(","):
                    conv, rp, fp = (int(x) for x in ent.split(":"))
                    if conv == (2 if r.flag & 16 else 16):
                        want.append(((rp - 1) | ((fp - 1) << 16)))
            assert [int(x) for x in mp[mo:mo + nmp]] == want
            qo += lseq; co += ncig; mo += nmp
    hb.close()


def test_bam_decoder_unmapped_and_missing_tags(tmp_path):
    recs = [bamio.BamRecord("u1", 4, -1, -1, 0, [], "ACGTN", [30] * 5, []),
            bamio.BamRecord("m1", 0, 0, 10, 7, "5M", "ACGTA", None, [("NM", "i", 70000)]),
            bamio.BamRecord("m1", 16, 0, 12, 0, "2S3M", "ACGTA", [1, 2, 3, 4, 5], [("XI", "f", 0.5), ("MP", "Z", "2:3:1,16:4:2")])]
    p = str(tmp_path / "x.bam")
    bamio.write_bam(p, "@HD\tVN:1.0\n@SQ\tSN:c\tLN:100\n", [("c", 100)], recs)
    hb = _decode(p, ["c"])
    rec = hb.rec_array()
    f = [int(w[1]) >> 24 for w in rec[:3]]
    assert f[0] & L.SD_F_UNMAPPED and f[0] & L.SD_F_SKIP
    assert not (f[1] & L.SD_F_UNMAPPED) and math.isnan(float(np.uint32(rec[1][2]).view(np.float32)))
    assert int(rec[1][4]) & 0xFFFF == 65535          # NM saturates
    assert not (f[2] & L.SD_F_NEWNAME) and f[1] & L.SD_F_NEWNAME
    assert int(rec[2][4]) >> 16 == 1                 # only the A>G entry of a reverse read is kept
    assert list(hb.column("mp")[:1]) == [2 | (0 << 16)]
    hb.close()


def test_decoder_rejects_garbage(tmp_path):
    from slamdunk_b200.engine import HostBatch
    p = str(tmp_path / "bad.bam")
    open(p, "wb").write(b"this is not a bam file at all, not even close......")
    with pytest.raises(L.NativeError):
        HostBatch(p, ["c"])
    with pytest.raises(L.NativeError):
        HostBatch(str(tmp_path / "missing.bam"), ["c"])


def test_fasta_loader_matches_python(tmp_path):
    from slamdunk_b200.engine import HostGenome
    c = simdata.make_case(str(tmp_path), seed=3, n_reads=10)
    g = HostGenome(c.ref, threads=2)
    assert g.names == list(c.genome.keys())
    assert list(g.chrom_len) == [len(s) for s in c.genome.values()]
    for ci, (name, s) in enumerate(c.genome.items()):
        off = int(g.chrom_off[ci])
        assert off % 32 == 0
        up = s.upper()
        for plane, fn in ((g.is_base_mask(0), lambda ch: ch == "T"), (g.is_base_mask(1), lambda ch: ch == "A"),
                          (g.nmask, lambda ch: ch not in "ACGT")):
            bits = np.unpackbits(plane.view(np.uint8), bitorder="little")[off:off + len(s)]
            want = np.array([fn(ch) for ch in up], dtype=np.uint8)
            assert (bits == want).all()
        if ci + 1 < g.n_chrom:
            assert off + len(s) + 32 <= int(g.chrom_off[ci + 1])
    # pad positions are N
    bits = np.unpackbits(g.nmask.view(np.uint8), bitorder="little")
    assert bits[:int(g.chrom_off[0])].all()
    g.close()


def test_fasta_loader_multiline_and_description(tmp_path):
    from slamdunk_b200.engine import HostGenome
    p = str(tmp_path / "m.fa")
    open(p, "w").write(">chrA some description\nACGT\nacgtNN\r\n\n>chrB\nTTTT")
    g = HostGenome(p)
    assert g.names == ["chrA", "chrB"] and list(g.chrom_len) == [10, 4]
    t = np.unpackbits(g.is_base_mask(0).view(np.uint8), bitorder="little")
    a0, b0 = int(g.chrom_off[0]), int(g.chrom_off[1])
    assert list(t[a0:a0 + 10]) == [0, 0, 0, 1, 0, 0, 0, 1, 0, 0] and list(t[b0:b0 + 4]) == [1, 1, 1, 1]
    g.close()


def _reference_snp_sets(records):
    """utils/SNPtools.py:30-38 verbatim semantics"""
    tc, ag = {}, {}
    for ch, pos, ref, alt in records:
        if ref.upper() == "T" and alt.upper() == "C":
            tc[ch + pos] = True
        if ref.upper() == "A" and alt.upper() == "G":
            ag[ch + pos] = True
    return tc, ag


def test_snp_masks_reproduce_string_key_collisions():
    names = ["chr1", "chr12", "chr2", "2", "chrX"]
    lens = [30000, 4000, 500, 25000, 100]
    off, o = [], 32
    for l in lens:
        off.append(o)
        o = ((o + l + 31) // 32) * 32 + 32
    n_words = o // 32
    records = [("chr12", "345", "T", "C"), ("chr1", "2345", "A", "G"), ("chr1", "77", "t", "c"), ("chr1", "0077", "T", "C"),
               ("chr2", "12", "A", "G"), ("chr", "2100", "T", "C"), ("chrX", "5", "T", "C,G"), ("chrX", "6", "A", "g"),
               ("chr1", "23456", "T", "C"), ("chr12", "3999", "A", "G"), ("2", "17", "T", "C"), ("chr1", "99999", "T", "C"),
               ("chr10", "55", "T", "C")]
    sm = SNPtools.SNPMasks(names, off, lens, n_words)
    for r in records:
        sm.add(*r)
    tcm, agm = sm.masks()
    tc, ag = _reference_snp_sets(records)
    tb = np.unpackbits(tcm.view(np.uint8), bitorder="little")
    ab = np.unpackbits(agm.view(np.uint8), bitorder="little")
    for ci, name in enumerate(names):
        for p in range(lens[ci]):
            key = name + str(p + 1)                        # SNPtools.py:53-60
            assert bool(tb[off[ci] + p]) == (key in tc), (name, p)
            assert bool(ab[off[ci] + p]) == (key in ag), (name, p)
    assert tb.sum() == sum(1 for ci, n in enumerate(names) for p in range(lens[ci]) if n + str(p + 1) in tc)


def test_bed_reader_quirks(tmp_path):
    from slamdunk_b200.utils.bed import BedIterator
    p = str(tmp_path / "a.bed")
    open(p, "w").write("chr1\t5\t10\tn1\t0\t+\nchr1\t7\t9\tn2\nchr2\t1\t2\tn3\t.\t-\textra\n")
    rows = list(BedIterator(p))
    assert [(r.chromosome, r.start, r.stop, r.name, r.strand) for r in rows] == \
        [("chr1", 5, 10, "n1", "+"), ("chr1", 7, 9, "n2", "."), ("chr2", 1, 2, "n3", "-")]
    assert rows[0].hasStrand() and not rows[1].hasStrand() and rows[0].getLength() == 5


def test_decoder_grouping_is_a_stable_partition(tmp_path):
    """SD_DECODE_GROUP: plain single-match-block reads first (plus then minus strand), the rest after, each
    group in file order; `order` maps batch slots back to file records and the columns follow it."""
    c = simdata.make_case(str(tmp_path), seed=78, n_reads=600, complexity=2.0)
    _t, refs, recs = bamio.read_bam(c.bam)
    names = list(c.genome.keys())
    from slamdunk_b200.engine import HostBatch
    hb = HostBatch(c.bam, names, want_mp=True, group=True)
    order = hb.order().tolist()
    assert sorted(order) == list(range(len(recs)))

    def cls(r):
        simple = len(r.cigar) == 1 and r.cigar[0][0] in (0, 7, 8)
        return (0 if simple else 2) | (1 if r.flag & 16 else 0)
    keys = [(cls(recs[i]), i) for i in order]
    assert keys == sorted(keys)
    assert len(set(k for k, _ in keys)) >= 3
    rec = hb.rec_array()
    tiles = hb.tiles_array()
    seq, qual = hb.column("seq"), hb.column("qual")
    for t in range(hb.batch.n_tiles):
        so, qo = int(tiles[t][0]) // 4, int(tiles[t][1])
        for k in range(32):
            slot = t * 32 + k
            if slot >= len(recs):
                break
            r = recs[order[slot]]
            lseq = int(rec[slot][3]) & 0xFFFF
            assert lseq == len(r.seq) and np.int32(rec[slot][0]) == r.pos
            assert _unslice(seq, so, k, lseq) == r.seq.upper()
            assert list(qual[qo:qo + lseq]) == r.qual
            qo += lseq
    hb.close()
    plain = HostBatch(c.bam, names, want_mp=True, group=False)
    assert plain.order().tolist() == list(range(len(recs)))
    plain.close()


def test_parallel_record_index_equals_serial_walk(tmp_path, monkeypatch):
    """The chunked, verified record indexing gives the same batch as the serial walk (threads = 1)."""
    from slamdunk_b200.engine import HostBatch
    c = simdata.make_case(str(tmp_path), seed=79, n_reads=3000, complexity=2.0)
    names = list(c.genome.keys())
    serial = HostBatch(c.bam, names, want_mp=True, group=True, threads=1)
    monkeypatch.setenv("SD_DECODE_PAR_MIN", "0")
    for threads in (2, 5):
        par = HostBatch(c.bam, names, want_mp=True, group=True, threads=threads)
        assert par.n_reads == serial.n_reads == 3000
        assert np.array_equal(par.rec_array(), serial.rec_array()) and np.array_equal(par.order(), serial.order())
        for col in ("seq", "qual", "cigar", "mp"):
            assert np.array_equal(par.column(col), serial.column(col)), col
        assert np.array_equal(par.name_hash(), serial.name_hash())
        par.close()
    serial.close()


def test_native_float_repr_equals_python_repr():
    """sd_format_repr must print what Python's repr(float) prints: the bedgraph rates and nothing else go through it."""
    import ctypes
    import random
    import struct
    lib = L.load()
    buf = ctypes.create_string_buffer(64)

    def fmt(x):
        n = lib.sd_format_repr(ctypes.c_double(x), buf)
        return buf.value[:n].decode()
    rng = random.Random(7)
    samples = [0.0, 1.0, 100.0, 0.5, 1e16, 1e15, 9999999999999998.0, 1e-4, 1e-5, 0.0001234, 123456789.125, 1 / 3, 2 / 3, 1e22, 5e-324,
               1.7976931348623157e308, 14.285714285714286, 16.666666666666668, 0.022222222222222223, 666666.6666666666]
    for _ in range(20000):
        tc, cov = rng.randint(1, 5000), rng.randint(1, 200000)
        samples.append(tc * 100.0 / cov)
        samples.append(float(tc) / float(cov))
        samples.append(tc * 1000000.0 / cov)
    for _ in range(20000):
        x = struct.unpack("<d", struct.pack("<Q", rng.getrandbits(64)))[0]
        if x == x and abs(x) != float("inf"):
            samples.append(x)
    for x in samples:
        assert fmt(x) == repr(x), (x, fmt(x), repr(x))


def test_block_inflate_matches_zlib_and_refuses_what_it_cannot_decode(tmp_path, monkeypatch):
    """sd_inflate_raw (the decoder sd_decode_bam tries on every BGZF block before zlib): round trips over levels,
    strategies (stored, fixed, dynamic, Huffman-only, RLE), multi-block streams and sizes around the fast-loop margins;
    wrong sizes, truncated and corrupted streams come back as 0 or as bytes the block's CRC would reject -- never past
    the buffers.  Then a whole BAM decoded through it equals the same BAM decoded through zlib alone."""
    import ctypes
    import random
    import zlib
    lib = L.load()
    rng = random.Random(5)

    def own(comp, n):
        out = ctypes.create_string_buffer(n + 64)
        ctypes.memset(out, 0xAA, n + 64)
        ok = lib.sd_inflate_raw(comp, len(comp), ctypes.cast(out, ctypes.c_void_p), n)
        assert out.raw[n:] == b"\xAA" * 64, "wrote past the end"
        return ok, out.raw[:n]

    def deflate(data, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, memlevel=8):
        c = zlib.compressobj(level, zlib.DEFLATED, -15, memlevel, strategy)
        return c.compress(data) + c.flush()

    samples = [b"", b"a", bytes(1000), bytes(range(256)) * 250, os.urandom(65280),
               b"".join(rng.choice([b"ACGT", b"AAAA", b"FFFFFFFF", b"chr1", os.urandom(3)]) for _ in range(9000))[:65280],
               ("".join("%d:%d:%d," % (rng.randint(0, 24), rng.randint(1, 100), rng.randint(1, 100)) for _ in range(6000))).encode()]
    for n in (1, 2, 7, 8, 9, 16, 17, 257, 258, 259, 265, 266, 267, 273, 274, 275, 600):
        samples += [bytes(rng.randrange(4) for _ in range(n)), b"x" * n]
    decoded = 0
    for data in samples:
        for level in (0, 1, 6, 9):
            for strategy in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FIXED):
                comp = deflate(data, level, strategy, memlevel=rng.choice((1, 8)))
                ok, got = own(comp, len(data))
                if ok:                                   # a 0 would only mean "left to zlib"
                    decoded += 1
                    assert got == data
                assert own(comp, len(data) + 1)[0] == 0
                if data:
                    assert own(comp, len(data) - 1)[0] == 0
                    own(comp[:len(comp) // 2], len(data))
    assert decoded >= 0.9 * len(samples) * 16
    c = zlib.compressobj(6, zlib.DEFLATED, -15)          # several blocks, empty stored blocks in between
    parts = [os.urandom(500), b"A" * 3000, b"ACGT" * 500]
    comp = b"".join(c.compress(t) + c.flush(fl) for t, fl in zip(parts, (zlib.Z_SYNC_FLUSH, zlib.Z_FULL_FLUSH, zlib.Z_FINISH)))
    assert own(comp, 5500) == (1, b"".join(parts))
    base = os.urandom(1000) + b"ACGTTGCA" * 1500 + bytes(2000)
    comp = deflate(base)
    for _ in range(3000):
        b = bytearray(comp)
        for _k in range(rng.randint(1, 4)):
            b[rng.randrange(len(b))] ^= 1 << rng.randrange(8)
        own(bytes(b), len(base))
    for _ in range(1000):
        own(os.urandom(rng.randint(0, 300)), rng.randint(0, 1000))

    for _ in range(300):                                  # the block checksum: carry-less-multiply CRC-32 == zlib's
        blob = os.urandom(rng.choice((0, 1, 15, 16, 63, 64, 65, 79, 80, 127, 128, rng.randint(0, 70000))))
        assert lib.sd_crc32(blob, len(blob)) == zlib.crc32(blob)

    case = simdata.make_case(str(tmp_path), 77, n_reads=1500, read_len=(20, 120), name="inf")
    from slamdunk_b200.engine import HostBatch
    views = []
    for mode in ("own", "zlib"):
        monkeypatch.setenv("SD_INFLATE", mode)
        hb = HostBatch(case.bam, [n for n, _l in case.bam_refs], want_mp=True, group=True, seq2=True, wire=True)
        views.append((hb.rec_array().copy(), hb.column("seq").copy(), hb.column("qual").copy(), hb.column("cigar").copy(),
                      hb.column("mp").copy(), hb.inflated_bytes))
        hb.close()
    for a, b in zip(*views):
        assert np.array_equal(a, b)


def test_block_deflate_round_trips_through_zlib_and_the_own_inflate(tmp_path, monkeypatch):
    """sd_deflate_raw (what the BAM writer compresses its BGZF payloads with): every stream it writes must inflate to the
    input with zlib AND with sd_inflate_raw, at every level, on empty / tiny / incompressible (stored block) / repetitive
    inputs and on real BAM payloads, and must come out within a few percent of zlib's size at level 6; an output buffer
    that is too small is refused without a byte written past it.  Then the native BAM writer with it against the writer
    with zlib alone: same records."""
    import ctypes
    import random
    import zlib
    lib = L.load()
    rng = random.Random(3)

    def deflate(data, level=6, cap=None):
        cap = cap if cap is not None else len(data) + 1024
        out = ctypes.create_string_buffer(cap + 64)
        ctypes.memset(out, 0xAA, cap + 64)
        n = lib.sd_deflate_raw(data, len(data), ctypes.cast(out, ctypes.c_void_p), cap, level)
        assert out.raw[cap:] == b"\xAA" * 64
        return out.raw[:n] if n else None

    def inflate(comp, n):
        out = ctypes.create_string_buffer(n + 8)
        return out.raw[:n] if lib.sd_inflate_raw(comp, len(comp), ctypes.cast(out, ctypes.c_void_p), n) else None

    case = simdata.make_case(str(tmp_path), 91, n_reads=2500, read_len=(60, 100), name="defl")
    raw = bamio.bgzf_decompress(open(case.bam, "rb").read())
    samples = [b"", b"a", b"ab", b"abc", b"abcd", b"aaaa", b"abcabcabcabc", bytes(1000), bytes(65535), os.urandom(65535), os.urandom(100),
               bytes(range(256)) * 255, ("".join("%d:%d:%d," % (rng.randint(0, 24), rng.randint(1, 100), rng.randint(1, 100)) for _ in range(8000))).encode()[:65280]]
    samples += [bytes(rng.randrange(3) for _ in range(n)) for n in range(1, 40)]
    samples += [raw[k:k + 0xFF00] for k in range(0, len(raw), 0xFF00)]
    own6 = z6 = 0
    for data in samples:
        for level in (1, 3, 6, 9):
            comp = deflate(data, level)
            assert comp is not None
            d = zlib.decompressobj(-15)
            assert d.decompress(comp) + d.flush() == data and d.eof and d.unused_data == b""
            assert inflate(comp, len(data)) == data
        if len(data) > 20000:
            own6 += len(deflate(data, 6))
            c = zlib.compressobj(6, zlib.DEFLATED, -15)
            z6 += len(c.compress(data) + c.flush())
        assert deflate(data, 6, cap=16) is None or len(data) <= 11
    assert own6 <= 1.03 * z6
    assert lib.sd_deflate_raw(bytes(70000), 70000, ctypes.cast(ctypes.create_string_buffer(80000), ctypes.c_void_p), 80000, 6) == 0

    hb_n = len(bamio.read_bam(case.bam)[2])
    keep = np.ones(hb_n, dtype=np.uint8)
    outs = []
    for mode in ("own", "zlib"):
        monkeypatch.setenv("SD_DEFLATE", mode)
        out = str(tmp_path / ("rw_%s.bam" % mode))
        err = ctypes.create_string_buffer(512)
        nw = ctypes.c_uint64(0)
        text = bamio.read_bam(case.bam)[0]
        rc = lib.sd_rewrite_bam(case.bam.encode(), out.encode(), text.encode(), keep.ctypes.data, None, hb_n, 1, 6, 2, ctypes.byref(nw), err, len(err))
        assert rc == L.SD_OK, err.value
        outs.append(bamio.bgzf_decompress(open(out, "rb").read()))
    assert outs[0] == outs[1] and int(nw.value) == hb_n


def test_synthetic_workload_writer_fast_path_writes_the_same_bam(tmp_path):
    """tests/synth_files.py writes a device batch out as FASTA + BED + BAM (the files the file-level benchmarks read).  Its
    chunk-wise writer must produce the bytes of the record-by-record one: checked on batches decoded from simulated BAMs and
    held in CPU tensors, with reads of one length, of 1..70 and of 33..129 bases."""
    import types

    import torch

    import synth_files
    from slamdunk_b200.engine import DeviceBatch, HostBatch, HostGenome
    for seed, read_len, complexity in ((11, (100, 100), 0.3), (12, (1, 70), 1.0), (13, (33, 129), 1.0)):
        c = simdata.make_case(str(tmp_path), seed, n_reads=1200, read_len=read_len, complexity=complexity, name="w%d" % seed)
        g = HostGenome(c.ref)
        hb = HostBatch(c.bam, g.names, want_mp=True, mp_all=True, group=False, seq2=False, raw_qual=True)
        db = DeviceBatch.from_host(hb.batch, torch.device("cpu"))
        rng = np.random.default_rng(seed)
        tri = np.zeros((hb.n_reads, 19), dtype=np.int32)
        tri[:, 0] = rng.integers(0, 7, size=hb.n_reads)
        tri[:, 1:] = rng.integers(0, 300, size=(hb.n_reads, 18))

        class Ann:
            chrom, start, stop, strand = np.array([0, 1]), np.array([10, 20]), np.array([500, 700]), np.array([0, 1])

            def __len__(self):
                return 2
        wl = types.SimpleNamespace(lo=torch.from_numpy(g.lo.copy().view(np.int32)), hi=torch.from_numpy(g.hi.copy().view(np.int32)),
                                   nmask=torch.from_numpy(g.nmask.copy().view(np.int32)), names=g.names, chrom_off=g.chrom_off,
                                   chrom_len=g.chrom_len, ann=Ann())
        for sort in (True, False):
            a = synth_files.write_workload_files(str(tmp_path / ("slow%d%d" % (seed, sort))), wl, db, torch.from_numpy(tri), slow=True, sort=sort)
            b = synth_files.write_workload_files(str(tmp_path / ("fast%d%d" % (seed, sort))), wl, db, torch.from_numpy(tri), sort=sort)
            for x, y in zip(a[:3], b[:3]):
                assert open(x, "rb").read() == open(y, "rb").read()
        # and the FASTA it writes reads back as the genome it was given
        g2 = HostGenome(b[0])
        assert g2.names == g.names and np.array_equal(g2.lo, g.lo) and np.array_equal(g2.hi, g.hi)
        hb.close()


def test_native_tcount_rows_equal_the_python_rows():
    """sd_format_tcount_rows == SlamSeqInterval.__repr__ row by row (tcounter._row): int 0 against float forms of the two
    rates, the default row of a UTR without reads, FilteredReads 0, names outside ASCII, negative lengths."""
    import io

    from slamdunk_b200.dunks import tcounter

    class U(object):
        pass
    rng = np.random.default_rng(0)
    utrs = []
    for i in range(3000):
        u = U()
        u.chromosome, u.start, u.name, u.strand = "chr%d" % (i % 21), int(rng.integers(0, 10 ** 8)), ("g\u00e9ne%d" if i % 50 == 0 else "gene%d") % i, "+-"[i % 2]
        u.stop = u.start + int(rng.integers(-5, 5000))
        utrs.append(u)
    counters = rng.integers(0, 50000, size=(3000, 5)).astype(np.int64)
    counters[::7] = 0
    counters[1::11, L.SD_C_COV_ON_T] = 0
    counters[::13, L.SD_C_READS] = 1
    tcontent = rng.integers(0, 5000, size=3000).astype(np.int64)
    for readNumber in (50000000, 0, 1, 12345677):
        want = io.StringIO()
        for i, utr in enumerate(utrs):
            row = counters[i].tolist()
            if row[L.SD_C_READS] > 0:
                cov, conv = row[L.SD_C_COV_ON_T], row[L.SD_C_CONV_ON_T]
                cpm = row[L.SD_C_READS] * 1000000.0 / readNumber if readNumber > 0 else 0
                rate = float(conv) / float(cov) if cov > 0 else 0
                print(tcounter._row(utr, rate, cpm, int(tcontent[i]), cov, conv, row[L.SD_C_READS], row[L.SD_C_TCREADS], row[L.SD_C_MULTIMAP]), file=want)
            else:
                print(tcounter._row(utr, 0, 0, int(tcontent[i]), 0, 0, 0, 0, 0), file=want)
        assert tcounter.format_rows(utrs, counters, tcontent, readNumber) == want.getvalue()
    assert tcounter.format_rows([], counters, tcontent, 5) == ""


def test_compute_tconversions_host_side_with_a_stub_engine(tmp_path, monkeypatch):
    """The host half of computeTconversions (decode into the wire form, BED / VCF / header handling, the native row and
    bedgraph writers) run on the CPU with the device replaced by a stub that counts nothing: every UTR must come out as
    the reference's default row, the bedgraphs empty, the timings filled in.  (What the device adds is the GPU suite's job.)"""
    import io

    from slamdunk_b200.dunks import tcounter

    class Zeros(object):
        def __init__(self, n):
            self.a = np.zeros(n, dtype=np.int64)

        def cpu(self):
            return self

        def numpy(self):
            return self.a

    class StubEngine(object):
        def __init__(self, device):
            self.stats = Zeros(L.SD_NSTAT)
            self.calls = []

        def set_snps(self, *a):
            self.calls.append("snps")

        def set_reference(self, lo, hi, nmask, chrom_off, chrom_len):
            self.calls.append("ref")
            self.chrom_len = np.asarray(chrom_len, dtype=np.int64)

        def set_utrs(self, chrom_idx, start, stop, strand, in_bam=None):
            # a layout of its own: every row on a FASTA chromosome gets a private slice as long as its clipped interval
            self.n = len(chrom_idx)
            known = np.asarray(chrom_idx) != L.SD_REF_NONE
            clen = self.chrom_len[np.where(known, chrom_idx, 0).astype(np.int64)]
            length = np.where(known, np.maximum(0, np.minimum(np.asarray(stop), clen) - np.maximum(0, np.asarray(start))), 0)
            self.pos_len = length.astype(np.uint32)
            self.pos_off = (np.cumsum(length) - length).astype(np.uint64)
            self.pos_off[~known] = np.uint64(0xFFFFFFFFFFFFFFFF)
            self.n_pos = int(length.sum())

        def synchronize(self):
            pass

        def count_wire(self, wire, params):
            assert int(wire.n_tiles) > 0
            self.calls.append("count_wire")

        def count_host(self, batch, params):
            self.calls.append("count_host")

        def reset_outputs(self):
            pass

        def finalize(self):
            self.calls.append("finalize")

        def results(self):
            rng = np.random.default_rng(3)
            counters = np.zeros((self.n, L.SD_NCOUNTER), dtype=np.int64)
            counters[1::2, L.SD_C_READS] = 3                  # odd rows "had reads": only they reach the bedgraphs
            self.cov = rng.integers(0, 4, size=max(1, self.n_pos)).astype(np.int32)
            self.tc = np.minimum(self.cov, rng.integers(0, 3, size=max(1, self.n_pos))).astype(np.int32)
            return {"counters": counters, "tcontent": np.arange(self.n, dtype=np.int64), "stats": np.zeros(L.SD_NSTAT, dtype=np.int64),
                    "pos_cov": self.cov, "pos_tc": self.tc}

        def launch_count(self):
            return 0

        def close(self):
            self.calls.append("close")

    engines = []

    def make(device):
        engines.append(StubEngine(device))
        return engines[-1]
    monkeypatch.setattr(tcounter, "CountEngine", make)
    c = simdata.make_case(str(tmp_path), 5, n_reads=400, name="stub")
    out = str(tmp_path / "o")
    log = io.StringIO()
    tcounter.computeTconversions(c.ref, c.bed, c.vcf, c.bam, 100, 27, out + ".tsv", out + ".p", out + ".m", 1, log)
    assert engines[0].calls[-2:] == ["finalize", "close"] and "count_wire" in engines[0].calls
    lines = open(out + ".tsv").read().splitlines()
    assert lines[0].startswith("#slamdunk v") and lines[1].startswith("#annotation:") and lines[2] == tcounter.HEADER
    rows = [ln.split("\t") for ln in lines[3:]]
    utrs = [ln.rstrip("\n").split("\t") for ln in open(c.bed)]
    stranded = [u for u in utrs if len(u) > 5 and u[5] in "+-"]
    assert len(rows) == len(stranded) > 0
    filtered = tcounter.SlamSeqInfo(tcounter.parse_sam_header(bamio.read_bam(c.bam)[0])).FilteredReads      # the CPM denominator
    for k, (row, u) in enumerate(zip(rows, stranded)):
        assert row[:4] == u[:4] and row[5] == u[5]
        if k % 2 == 0:
            assert row[6:] == ["0", "0", str(k), "0", "0", "0", "0", "0", "-1.0", "-1.0"]
        else:       # reads but no coverage on Ts: int 0 rate, CPM from the header's FilteredReads
            assert row[6] == "0" and row[7] == repr(3 * 1000000.0 / filtered) and row[8:] == [str(k), "0", "0", "3", "0", "0", "-1.0", "-1.0"]
    # the bedgraphs: rows with reads, positions with coverage whose FASTA base is T ('+') / A ('-'), rate = tc * 100.0 / cov
    eng = engines[0]
    genome = {n_: s_.upper() for n_, s_ in bamio.read_fasta(c.ref).items()}
    want = {"+": [], "-": []}
    for k, u in enumerate(stranded):
        if k % 2 == 0 or u[0] not in genome:
            continue
        first = max(0, int(u[1]))
        for p_ in range(int(eng.pos_len[k])):
            idx = int(eng.pos_off[k]) + p_
            if eng.cov[idx] > 0 and genome[u[0]][first + p_] == ("T" if u[5] == "+" else "A"):
                rate = 0.0 if eng.tc[idx] == 0 else int(eng.tc[idx]) * 100.0 / int(eng.cov[idx])
                want[u[5]].append("%s %d %d %s\n" % (u[0], first + p_, first + p_ + 1, repr(rate)))
    assert len(want["+"]) > 50 and len(want["-"]) > 50
    assert open(out + ".p").read() == "".join(want["+"]) and open(out + ".m").read() == "".join(want["-"])
    tm = tcounter.OPTIONS["timings"]
    assert tm["bam_reads"] == 400 and tm["write_s"] >= 0 and tm["total_s"] > 0


def test_positional_tracks_host_side_with_a_stub_engine(tmp_path, monkeypatch):
    """genomewideConversionRates with the device replaced by a stub whose track_runs hands out made-up change points: the
    eight files, written by one thread per file, must hold the header line and then every chromosome's runs in the
    reference's chromosome order -- the same bytes as sequential sd_write_track calls."""
    import io

    from slamdunk_b200.dunks import tcounter

    def runs_of(row, kind):
        rng = np.random.default_rng(1000 * row + kind)
        n = int(rng.integers(0, 400))
        pos = np.sort(rng.choice(5000, size=n, replace=False)).astype(np.uint32)
        val = rng.integers(0, 50, size=n).astype(np.float64)
        isf = np.zeros(n, dtype=np.uint8)
        if kind == L.SD_TRACK_RATE:
            isf = (rng.random(n) < 0.7).astype(np.uint8)
            val = np.where(isf == 1, rng.random(n), 0.0)
        return pos, val, isf

    class Zeros(object):
        def cpu(self):
            return self

        def numpy(self):
            return np.zeros(L.SD_NSTAT, dtype=np.int64)

    class StubEngine(object):
        def __init__(self, device):
            self.stats = Zeros()

        def set_reference(self, *a):
            pass

        set_snps = set_utrs = count_host = set_reference

        def synchronize(self):
            pass

        reset_outputs = finalize = close = synchronize

        def track_runs(self, row, kind, cutoff, track_len):
            return runs_of(row, kind)

        def launch_count(self):
            return 0
    monkeypatch.setattr(tcounter, "CountEngine", StubEngine)
    c = simdata.make_case(str(tmp_path), 6, n_reads=300, name="pstub")
    prefix = str(tmp_path / "pstub_slamdunk_mapped_filtered")
    tcounter.genomewideConversionRates(c.ref, c.vcf, c.bam, 27, prefix, 1, 2, io.StringIO())
    genome = bamio.read_fasta(c.ref)
    names = list(genome.keys())
    lib = L.load()
    for suffix, tname, desc, strand, kind in tcounter.TRACKS:
        want = str(tmp_path / ("want" + suffix))
        with open(want, "w") as fh:
            print("track type=bedGraph name=\"pstub" + tname + "\" description=\"" + desc + "\"", file=fh)
        for chrom in sorted(names, key=tcounter._natural_keys):
            pos, val, isf = runs_of(2 * names.index(chrom) + strand, kind)
            if len(pos):
                assert lib.sd_write_track(want.encode(), chrom.encode(), kind, len(pos), pos.ctypes.data, val.ctypes.data, isf.ctypes.data) == L.SD_OK
        assert open(prefix + suffix).read() == open(want).read(), suffix
    assert tcounter.OPTIONS["timings"]["track_lines"] > 0


def test_bench_decode_object_runs_on_the_cpu_with_stubs(tmp_path, monkeypatch):
    """bench.py's `decode` leg (sample written out as files, sd_load_fasta, sd_decode_bam, the filter's rewrite, two file-level
    computeTconversions calls) end to end on the CPU: the workload generator hands out a batch decoded from a simulated BAM,
    the device is a stub.  The leg must fill in every field and never report an error."""
    import types

    import torch

    import bench
    from slamdunk_b200.dunks import tcounter
    from slamdunk_b200.engine import DeviceBatch, HostBatch, HostGenome
    c = simdata.make_case(str(tmp_path), 21, n_reads=1500, read_len=(100, 100), complexity=0.3, name="leg")
    g = HostGenome(c.ref)
    hb = HostBatch(c.bam, g.names, want_mp=True, mp_all=True, group=False, seq2=False, raw_qual=True)
    db = DeviceBatch.from_host(hb.batch, torch.device("cpu"))
    tri = torch.from_numpy(np.zeros((hb.n_reads, 19), dtype=np.int32))

    class Ann(object):
        chrom, start, stop, strand = np.array([0, 1]), np.array([10, 20]), np.array([500, 700]), np.array([0, 1])

        def __len__(self):
            return 2

    class Workload(object):
        lo, hi, nmask = (torch.from_numpy(a.copy().view(np.int32)) for a in (g.lo, g.hi, g.nmask))
        names, chrom_off, chrom_len, ann = g.names, g.chrom_off, g.chrom_len, Ann()

        def set_time_point(self, minutes):
            pass

        def params(self, **kw):
            assert set(kw) == {"seed", "first_read", "n_reads", "read_len", "frac_multimap"}
            return kw

        def generate(self, p, want_triples=False, coordinate_sorted=False, group=None):
            assert want_triples and coordinate_sorted and group is False
            return db, tri

    class Zeros(object):
        def cpu(self):
            return self

        def numpy(self):
            return np.zeros(L.SD_NSTAT, dtype=np.int64)

    class StubEngine(object):
        def __init__(self, device):
            self.stats = Zeros()

        def set_reference(self, *a):
            pass

        set_snps = count_wire = count_host = set_reference

        def set_utrs(self, chrom_idx, *a):
            self.n = len(chrom_idx)
            self.pos_len, self.pos_off = np.zeros(self.n, dtype=np.uint32), np.zeros(self.n, dtype=np.uint64)

        def synchronize(self):
            pass

        reset_outputs = finalize = close = synchronize

        def results(self):
            return {"counters": np.zeros((self.n, L.SD_NCOUNTER), dtype=np.int64), "tcontent": np.zeros(self.n, dtype=np.int64),
                    "stats": np.zeros(L.SD_NSTAT, dtype=np.int64), "pos_cov": np.zeros(1, dtype=np.int32), "pos_tc": np.zeros(1, dtype=np.int32)}

        def launch_count(self):
            return 0
    monkeypatch.setattr(tcounter, "CountEngine", StubEngine)
    fake = types.SimpleNamespace(wl=Workload(), args=types.SimpleNamespace(workload="cfg2", read_len=100))
    out = bench.decode_leg(fake, 1500)
    assert "error" not in out, out
    assert out["reads"] == 1500 and out["bam_reads_per_s"] > 0 and out["fasta_decode_s"] > 0 and out["bam_rewrite_s"] > 0
    assert out["file_level"]["second_run_stages"]["bam_reads"] == 1500 and out["file_level"]["output_bytes"] > 0
    assert set(out["bam_decode_one_thread"]) == {"inflate_own", "inflate_zlib"} and out["bam_rewrite_zlib_deflate"]["bytes"] > 0
    assert "SD_INFLATE" not in os.environ and "SD_DEFLATE" not in os.environ
    hb.close()


def _bedgraph_text_like_the_reference(names, first, strand, has, po, pl, gstart, isx, cov, tc):
    """The reference's loop (tcounter.py:277-285, 315-326): per strand one dict position -> rate filled UTR by UTR, printed
    in insertion order; a position already in the dict keeps its place (and gets the same value again)."""
    graphs = ({}, {})
    for i in range(len(names)):
        if not has[i] or pl[i] == 0:
            continue
        s = int(strand[i])
        for p in range(int(pl[i])):
            idx = int(po[i]) + p
            g = int(gstart[i]) + p
            if cov[idx] > 0 and (int(isx[s][g >> 5]) >> (g & 31)) & 1:
                rate = 0.0 if tc[idx] == 0 else int(tc[idx]) * 100.0 / int(cov[idx])
                graphs[s][(names[i], int(first[i]) + p, idx)] = rate
    return ["".join("%s %d %d %s\n" % (nm.decode(), pos, pos + 1, repr(rate)) for (nm, pos, _idx), rate in graph.items()) for graph in graphs]


@pytest.mark.parametrize("threads", [1, 3, 0])
@pytest.mark.parametrize("layout", ["disjoint", "overlapping"])
def test_native_bedgraph_writer_keeps_the_reference_order(tmp_path, threads, layout):
    """sd_write_bedgraphs on several threads == the reference's dict loop, also where same-strand rows share positions
    (nested, chained and identical UTRs: the first row in BED order prints a shared position)."""
    import ctypes
    lib = L.load()
    rng = np.random.default_rng(11)
    n, glen = 300, 600000
    chrom = np.arange(n) % 5
    strand = (rng.random(n) < 0.4).astype(np.uint8)
    has = (rng.random(n) < 0.85).astype(np.uint8)
    if layout == "disjoint":
        first = np.sort(rng.choice(glen // 1400, size=n, replace=False)).astype(np.int64) * 1400
        length = rng.integers(1, 1300, size=n)
        order = rng.permutation(n)                     # BED order is not position order
        first, length = first[order], length[order]
        po = np.zeros(n, dtype=np.uint64)
        po[np.argsort(first)] = np.concatenate([[0], np.cumsum(length[np.argsort(first)] + 1)[:-1]])
        pl = (length + 1).astype(np.uint32)
        n_pos = int((length + 1).sum())
    else:
        # rows of one strand share one position array indexed by genome position (what merged segments amount to)
        first = rng.integers(0, 5000, size=n).astype(np.int64)
        length = rng.integers(1, 1500, size=n)
        first[9], length[9] = first[4], length[4]      # identical rows (same chromosome: 9 % 5 == 4 % 5)
        strand[9] = strand[4]
        has[4] = has[9] = 1
        po = (first + (chrom * 2 + strand) * 8000).astype(np.uint64)
        pl = length.astype(np.uint32)
        n_pos = 80000
    pl[7] = 0
    gstart = (first + chrom * 100000 + 64).astype(np.uint64)
    words = (glen + 500000) // 32 + 8
    # random reference planes (2-bit base codes + N mask); "is T" / "is A" are what the reference's refSeq[position] test asks
    lo, hi, nm = (rng.integers(0, 2 ** 32, size=words, dtype=np.uint32) for _ in range(3))
    nm &= rng.integers(0, 2 ** 32, size=words, dtype=np.uint32)
    isx = [lo & hi & ~nm, ~lo & ~hi & ~nm]
    cov = rng.integers(-1, 40, size=n_pos).astype(np.int32)
    tc = np.where(rng.random(n_pos) < 0.3, rng.integers(0, 6, size=n_pos), 0).astype(np.int32)
    names = [("chr%d" % (c + 1)).encode() for c in chrom]
    want = _bedgraph_text_like_the_reference(names, first, strand, has, po, pl, gstart, isx, cov, tc)
    assert len(want[0]) > 1000 and len(want[1]) > 1000
    cn = (ctypes.c_char_p * n)(*names)
    ptr = lambda a: ctypes.c_void_p(a.ctypes.data)    # noqa: E731
    outs = [str(tmp_path / "plus.bedgraph"), str(tmp_path / "minus.bedgraph")]
    rc = lib.sd_write_bedgraphs(outs[0].encode(), outs[1].encode(), n, cn, ptr(first), ptr(strand), ptr(has), ptr(po), ptr(pl), ptr(gstart),
                                ptr(lo), ptr(hi), ptr(nm), ptr(cov), ptr(tc), n_pos, threads)
    assert rc == L.SD_OK
    for path, text in zip(outs, want):
        assert open(path).read() == text
    # an offset past the arrays is refused, not read
    po[0] = n_pos
    has[0], pl[0] = 1, 5
    assert lib.sd_write_bedgraphs(outs[0].encode(), outs[1].encode(), n, cn, ptr(first), ptr(strand), ptr(has), ptr(po), ptr(pl), ptr(gstart),
                                  ptr(lo), ptr(hi), ptr(nm), ptr(cov), ptr(tc), n_pos, threads) == -2   # SD_ERR_ARG


def _filter_states(records, bam_names, bed_rows, min_identity, nm):
    """What sd_filter reports next to keep[]: survivor class (0 / 1 unique / 2 multimapper) | 4 when a multimapper overlaps a
    UTR (filter.py:137) -- here derived on the host from the decoded records, for the writer test."""
    state = np.zeros(len(records), dtype=np.uint8)
    for i, r in enumerate(records):
        if r.flag & 4:
            continue
        if float(np.float32(r.get_tag("XI"))) < min_identity or (nm > -1 and int(r.get_tag("NM")) > nm):
            continue
        state[i] = 2 if r.mapq == 0 else 1
        if state[i] == 2:
            for ch, start, stop, _n, _s in bed_rows:
                if ch == bam_names[r.ref_id] and start < r.reference_end and stop + 1 > r.pos:
                    state[i] |= 4
                    break
    return state


def _rec_tuple(r):
    return (r.qname, r.flag, r.ref_id, r.pos, r.mapq, tuple(r.cigar), r.next_ref_id, r.next_pos, r.tlen, r.seq,
            None if r.qual is None else tuple(r.qual), tuple((t, ty, tuple(v) if isinstance(v, list) else v) for t, ty, v in r.tags))


@pytest.mark.parametrize("seed,use_bed,nm,threads", [(501, True, -1, 3), (502, True, 4, 1), (503, False, -1, 4)])
def test_native_bam_writer_matches_python_codec(tmp_path, seed, use_bed, nm, threads):
    """sd_rewrite_bam (host only): kept records byte-identical, RD tags as the oracle derives them from the reference's
    multimapList, samtools-sort order, and a BGZF stream the independent Python codec reads back."""
    import ctypes
    from oracle import oracle
    c = simdata.make_filter_case(str(tmp_path), seed=seed, n_reads=3000, topn=4)
    text, refs, records = bamio.read_bam(c.bam)
    bam_names = [n for n, _ in refs]
    bed_rows = oracle.read_bed(c.bed) if use_bed else None
    keep, rd, _stats = oracle.filter_records(records, bam_names, bed_rows, 2, 0.9, nm)
    state = _filter_states(records, bam_names, bed_rows or [], 0.9, nm)
    if use_bed:
        assert (keep == 2).any()
    want = []
    for i, r in enumerate(records):
        if keep[i] == 1:
            want.append(r.copy())
        elif keep[i] == 2:
            w = r.copy()
            w.flag &= ~0x900
            w.set_tag("RD", rd[i], "Z")
            want.append(w)
    want = bamio.sort_records(want)
    out = str(tmp_path / "w.bam")
    new_text = text + "@CO\trewritten\n"
    err = ctypes.create_string_buffer(256)
    nw = ctypes.c_uint64(0)
    lib = L.load()
    rc = lib.sd_rewrite_bam(c.bam.encode(), out.encode(), new_text.encode(), keep.ctypes.data, state.ctypes.data, len(records), 1, 6,
                            threads, ctypes.byref(nw), err, len(err))
    assert rc == 0, err.value
    assert nw.value == len(want)
    text2, refs2, got = bamio.read_bam(out)
    assert text2 == new_text and refs2 == refs
    assert [_rec_tuple(r) for r in got] == [_rec_tuple(r) for r in want]
    # wrong decision count is refused, not silently truncated
    rc = lib.sd_rewrite_bam(c.bam.encode(), out.encode(), new_text.encode(), keep.ctypes.data, state.ctypes.data, len(records) - 1, 1, 6,
                            threads, ctypes.byref(nw), err, len(err))
    assert rc != 0 and b"decisions" in err.value


@pytest.mark.parametrize("use_bed,mq,mi,nm", [(True, 2, 0.9, -1), (False, 3, 0.95, 4)])
def test_filter_host_side_with_the_decisions_of_the_oracle(tmp_path, monkeypatch, use_bed, mq, mi, nm):
    """filter.Filter's host half on the CPU: the device is replaced by a stub that hands back the ORACLE's decisions, so the
    test checks what the host does with them -- log lines, @RG DS / @PG rewrite, the sorted, compressed, indexed output with
    its RD tags -- against the same expectations the GPU suite uses."""
    import ast
    import io

    import torch
    from oracle import oracle
    from slamdunk_b200.dunks import filter as filter_mod
    c = simdata.make_filter_case(str(tmp_path), seed=611, n_reads=3000, topn=4)
    text, refs, records = bamio.read_bam(c.bam)
    bam_names = [n for n, _ in refs]
    bed_rows = oracle.read_bed(c.bed) if use_bed else None
    keep, rd, stats = oracle.filter_records(records, bam_names, bed_rows, mq, mi, nm)

    class StubEngine(object):
        def __init__(self, device):
            self.device = torch.device("cpu")

        def filter(self, db, params, name_hash=None, utrs=None):
            st = np.zeros(L.SD_NSTAT, dtype=np.int64)
            for k, name in ((L.SD_S_MAPPED, "mapped"), (L.SD_S_UNMAPPED, "unmapped"), (L.SD_S_FILTERED, "filtered"),
                            (L.SD_S_MQFILTERED, "mqfiltered"), (L.SD_S_IDFILTERED, "idfiltered"), (L.SD_S_NMFILTERED, "nmfiltered")):
                st[k] = stats[name]
            assert (utrs is None) == (not use_bed)
            return keep.copy(), st, _filter_states(records, bam_names, bed_rows or [], mi, nm)

        def close(self):
            pass

    class StubDeviceBatch(object):
        @staticmethod
        def from_host(batch, device, eng=None):
            return None
    monkeypatch.setattr(filter_mod, "CountEngine", StubEngine)
    monkeypatch.setattr(filter_mod, "DeviceBatch", StubDeviceBatch)
    out = str(tmp_path / "filtered.bam")
    log = io.StringIO()
    filter_mod.Filter(c.bam, out, log, c.bed if use_bed else None, mq, mi, nm)
    want = []
    for i, r in enumerate(records):
        if keep[i] == 1:
            want.append(r.copy())
        elif keep[i] == 2:
            w = r.copy()
            w.flag &= ~0x900
            w.set_tag("RD", rd[i], "Z")
            want.append(w)
    want = bamio.sort_records(want)
    text2, refs2, got = bamio.read_bam(out)
    assert refs2 == refs and [_rec_tuple(r) for r in got] == [_rec_tuple(r) for r in want]
    hdr = bamio.parse_header_text(text2)
    ds = ast.literal_eval(hdr["RG"][0]["DS"])
    for k in ("mapped", "filtered", "mqfiltered", "idfiltered", "nmfiltered", "multimapper"):
        assert ds[k] == stats[k], k
    assert ds["sequenced"] == stats["mapped"] + stats["unmapped"]
    assert hdr["PG"][-1]["ID"] == "slamdunk" and hdr["PG"][-1]["VN"] == "3"
    assert ("MM\t%d" % stats["multimapper"]) in log.getvalue()
    assert os.path.getsize(out + ".bai") > 8
    assert filter_mod.OPTIONS["stats"]["written"] == len(want)


def test_native_bam_writer_replaces_existing_rd_tag_and_handles_empty(tmp_path):
    import ctypes
    recs = [bamio.BamRecord("m1", 0x100, 0, 10, 0, "20M", "A" * 20, [30] * 20, [("RD", "Z", "old"), ("XI", "f", 1.0), ("NM", "C", 0),
                                                                                    ("ZB", "Bs", [1, -2, 3])]),
            bamio.BamRecord("u1", 0, 0, 5, 60, "20M", "C" * 20, [30] * 20, [("XI", "f", 1.0)])]
    src = str(tmp_path / "s.bam")
    bamio.write_bam(src, "@HD\tVN:1.0\n", [("chr1", 1000)], recs)
    lib = L.load()
    err = ctypes.create_string_buffer(256)
    nw = ctypes.c_uint64(0)
    keep = np.array([2, 1], dtype=np.uint8)
    state = np.array([2 | 4, 1], dtype=np.uint8)
    out = str(tmp_path / "o.bam")
    assert lib.sd_rewrite_bam(src.encode(), out.encode(), b"@HD\tVN:1.0\n", keep.ctypes.data, state.ctypes.data, 2, 1, 6, 2,
                              ctypes.byref(nw), err, len(err)) == 0, err.value
    _t, _r, got = bamio.read_bam(out)
    assert [r.qname for r in got] == ["u1", "m1"]
    assert got[1].flag == 0 and got[1].tags == [("XI", "f", 1.0), ("NM", "C", 0), ("ZB", "Bs", [1, -2, 3]), ("RD", "Z", "chr1:10-30")]
    keep[:] = 0
    assert lib.sd_rewrite_bam(src.encode(), out.encode(), b"@HD\tVN:1.0\n", keep.ctypes.data, None, 2, 1, 6, 2,
                              ctypes.byref(nw), err, len(err)) == 0, err.value
    assert nw.value == 0 and bamio.read_bam(out)[2] == []


def test_decoder_reports_the_widest_unspliced_alignment(tmp_path):
    """sd_batch.max_span sizes the count kernel's register window: the largest reference span, spliced alignments aside."""
    recs = [bamio.BamRecord("a", 0, 0, 10, 60, "50M", "A" * 50, [30] * 50, [("XI", "f", 1.0)]),
            bamio.BamRecord("b", 0, 0, 20, 60, "20M7D30M", "C" * 50, [30] * 50, [("XI", "f", 1.0)]),        # span 57
            bamio.BamRecord("c", 0, 0, 30, 60, "5S40M5I", "G" * 50, [30] * 50, [("XI", "f", 1.0)]),          # span 40
            bamio.BamRecord("d", 0, 0, 40, 60, "25M900N25M", "T" * 50, [30] * 50, [("XI", "f", 1.0)]),       # spliced: left out
            bamio.BamRecord("e", 4, -1, -1, 0, "", "T" * 80, [30] * 80, [])]
    p = str(tmp_path / "s.bam")
    bamio.write_bam(p, "@HD\tVN:1.0\n", [("chr1", 5000)], recs)
    hb = _decode(p, ["chr1"], want_mp=False)
    assert hb.batch.max_lseq == 80 and hb.batch.max_span == 57
    hb.close()


def test_decoder_codes_binned_qualities(tmp_path):
    """<= 4 distinct quality values -> 2-bit codes, <= 16 -> 4-bit, more -> bytes; column("qual") gives the phred bytes back."""
    import random
    rnd = random.Random(5)
    for levels, bits in (([2, 12, 23, 37], 2), (list(range(10, 40, 3)), 4), (list(range(2, 42)), 8)):
        recs = []
        for i in range(150):
            n = rnd.randint(20, 90)
            recs.append(bamio.BamRecord("r%d" % i, 0, 0, 10 + 3 * i, 60, "%dM" % n, "".join(rnd.choice("ACGT") for _ in range(n)),
                                        [rnd.choice(levels) for _ in range(n)], [("XI", "f", 1.0)]))
        p = str(tmp_path / ("q%d.bam" % bits))
        bamio.write_bam(p, "@HD\tVN:1.0\n", [("chr1", 5000)], recs)
        for raw in (False, True):
            from slamdunk_b200.engine import HostBatch
            hb = HostBatch(p, ["chr1"], want_mp=False, tile_reads=32, threads=2, raw_qual=raw)
            assert hb.batch.qual_bits == (8 if raw else bits)
            qual, tiles = hb.column("qual"), hb.tiles_array()
            for t in range(hb.batch.n_tiles):
                qo = int(tiles[t][1])
                for k in range(32):
                    i = t * 32 + k
                    if i >= len(recs):
                        break
                    n = len(recs[i].qual)
                    assert list(qual[qo:qo + n]) == recs[i].qual
                    qo += n
            hb.close()


def test_native_bai_index_answers_region_queries(tmp_path):
    """sd_rewrite_bam_indexed: fetching regions through the BAI (independent reader, tests/baiquery.py) equals a scan."""
    import ctypes
    import random
    import baiquery
    c = simdata.make_case(str(tmp_path), seed=611, n_reads=9000, n_utr=60, read_len=(30, 120), complexity=2.0,
                          chroms=(("chr1", 400000), ("chr2", 70000), ("chr3", 20000), ("chrEmpty", 5000)))
    text, refs, records = bamio.read_bam(c.bam)
    rnd = random.Random(3)
    keep = np.array([1 if rnd.random() < 0.8 else 0 for _ in records], dtype=np.uint8)
    out, bai = str(tmp_path / "o.bam"), str(tmp_path / "o.bam.bai")
    err = ctypes.create_string_buffer(256)
    nw = ctypes.c_uint64(0)
    lib = L.load()
    rc = lib.sd_rewrite_bam_indexed(c.bam.encode(), out.encode(), bai.encode(), text.encode(), keep.ctypes.data, None, len(records), 1, 6,
                                    3, ctypes.byref(nw), err, len(err))
    assert rc == 0, err.value
    idx, n_no_coor = baiquery.read_bai(bai)
    assert len(idx) == len(refs) and n_no_coor == 0
    kept = [r for r, k in zip(records, keep) if k]
    for rid in range(len(refs)):
        meta = idx[rid][0].get(37450)
        on_ref = [r for r in kept if r.ref_id == rid]
        if on_ref:
            assert meta is not None and meta[1] == (sum(1 for r in on_ref if not r.flag & 4), sum(1 for r in on_ref if r.flag & 4))
        else:
            assert idx[rid] == ({}, [])
    for _ in range(60):
        rid = rnd.randrange(3)
        ln = refs[rid][1]
        beg = rnd.randrange(ln)
        end = min(ln, beg + rnd.choice([1, 50, 2000, 40000, 300000]))
        assert baiquery.fetch(out, idx, rid, beg, end) == baiquery.scan(out, rid, beg, end), (rid, beg, end)
    # an unsorted output cannot be indexed
    rc = lib.sd_rewrite_bam_indexed(c.bam.encode(), out.encode(), bai.encode(), text.encode(), keep.ctypes.data, None, len(records), 0, 6,
                                    3, ctypes.byref(nw), err, len(err))
    srt = all((a.ref_id, a.pos) <= (b.ref_id, b.pos) for a, b in zip(kept, kept[1:]))
    assert (rc == 0) == srt


def test_bam_decoder_two_bit_seq_and_full_mp_columns(tmp_path):
    """SD_DECODE_SEQ2: 2-bit base codes (A=0 C=1 G=2 T=3, everything else A), half the plane column's bytes, same tile
    geometry; SD_DECODE_MP_ALL: every MP entry as {readPos-1 | refPos-1 << 16, conv}.  The two exclude each other."""
    from slamdunk_b200.engine import HostBatch
    c = simdata.make_case(str(tmp_path), seed=78, n_reads=500, complexity=2.0)
    text, refs, recs = bamio.read_bam(c.bam)
    recs[3].seq = "N" + recs[3].seq[1:-1] + "R"       # non-ACGT bases in a read
    recs[7].seq = "=" * len(recs[7].seq)
    path = str(tmp_path / "n.bam")
    bamio.write_bam(path, text, refs, recs)
    names = list(c.genome.keys())
    h4 = HostBatch(path, names, threads=3)
    h2 = HostBatch(path, names, threads=3, seq2=True)
    assert (h4.batch.seq_bits or 4) == 4 and h2.batch.seq_bits == 2
    t4, t2 = h4.tiles_array(), h2.tiles_array()
    assert (t2[:, 0] * 2 == t4[:, 0]).all() and (t2[:, 1:] == t4[:, 1:]).all()
    assert h2.batch.max_tile_seq * 2 == h4.batch.max_tile_seq
    s4, s2 = h4.column("seq"), h2.column("seq")
    assert len(s2) * 2 == len(s4)
    code = {"A": 0, "C": 1, "G": 2, "T": 3}
    rec = h4.rec_array()
    for i, r in enumerate(recs):
        t, k = divmod(i, 32)
        base2 = int(t2[t, 0]) // 4
        for j, ch in enumerate(r.seq.upper()):
            g = base2 + 2 * ((j >> 5) * 32 + k)
            got = ((int(s2[g]) >> (j & 31)) & 1) | (((int(s2[g + 1]) >> (j & 31)) & 1) << 1)
            assert got == code.get(ch, 0), (i, j, ch)
        lseq = int(rec[i, 3]) & 0xFFFF
        if lseq % 32:      # pad bits of the last group are zero
            g = base2 + 2 * ((lseq >> 5) * 32 + k)
            assert (int(s2[g]) | int(s2[g + 1])) >> (lseq & 31) == 0
    hm = HostBatch(path, names, threads=3, mp_all=True)
    mp, tm, recm = hm.column("mp"), hm.tiles_array(), hm.rec_array()
    for t in range(hm.batch.n_tiles):
        mo = int(tm[t, 3]) // 4
        for k in range(32):
            i = t * 32 + k
            if i >= len(recs):
                break
            nmp = int(recm[i, 4]) >> 16
            want = []
            if recs[i].has_tag("MP"):
                for ent in recs[i].get_tag("MP").split(","):
                    conv, rp, fp = (int(x) for x in ent.split(":"))
                    want += [(rp - 1) | ((fp - 1) << 16), conv]
            assert [int(x) for x in mp[mo:mo + 2 * nmp]] == want
            mo += 2 * nmp
    with pytest.raises(L.NativeError, match="exclude"):
        HostBatch(path, names, threads=3, mp_all=True, seq2=True)
    for h in (h4, h2, hm):
        h.close()


def test_utrrates_percentages_and_checkstep_host_logic(tmp_path):
    """host side of `alleyoop utrrates` (stats.py:386-486: the strand swap, the `conversionSum =+` quirk, division guards)
    against the oracle restatement on random matrices; checkStep's date rules (utils/misc.py:147-164)."""
    import random
    import time
    from oracle import stats_oracle
    from slamdunk_b200.dunks import stats
    rng = random.Random(7)
    seen_none = seen_some = 0
    for _ in range(400):
        m = [rng.choice([0, 0, 1, 3, 50, 1000]) for _ in range(25)]
        if rng.random() < 0.3:
            m[18] = 0                                   # T_T = 0: the row adds nothing, whatever else it holds
        for strand in ("+", "-", "."):
            a, b = stats._utr_percentages(m, strand), stats_oracle.utr_percentages(m, strand)
            assert a == b
            seen_none += a is None
            seen_some += a is not None
    assert seen_none and seen_some
    src, out = str(tmp_path / "in.txt"), str(tmp_path / "out.txt")
    open(src, "w").write("x")
    assert stats.checkStep([src], [out]) is True                       # output missing: run
    time.sleep(0.05)
    open(out, "w").write("y")
    assert stats.checkStep([src], [out]) is False                      # output newer: skip ...
    assert stats.checkStep([src], [out], force=True) is True           # ... unless forced
    with pytest.raises(RuntimeError, match="don't exist"):
        stats.checkStep([src, str(tmp_path / "missing")], [out])


from wireutil import unpack_wire as _unpack_wire      # noqa: E402


def _uniform_bam(path, n_reads, L=40, quals=(20, 30, 37)):
    """plain reads of one length on two chromosomes, positions up to 200 kb apart inside one tile, three distinct qualities:
    tiles without l_seq / CIGAR, with 32-bit positions, with per-read chromosomes"""
    from slamdunk_b200.io import bam as bamio
    rng = np.random.RandomState(3)
    refs = [("c1", 400000), ("c2", 300000)]
    recs = []
    for i in range(n_reads):
        ref = 0 if i < n_reads * 2 // 3 else 1
        pos = int(rng.randint(0, 1000)) + (200000 if (i // 8) % 5 == 4 and i % 2 else 0)
        seq = "".join(rng.choice(list("ACGTN"), p=[0.24, 0.24, 0.24, 0.24, 0.04]) for _ in range(L))
        recs.append(bamio.BamRecord("u%04d" % i, 16 if i % 3 == 0 else 0, ref, pos, 60 if i % 7 else 0, [(0, L)], seq,
                                    [int(rng.choice(list(quals))) for _ in range(L)],
                                    [("NM", "i", i % 4), ("XI", "f", 1.0 - (i % 4) / L), ("MP", "Z", "16:3:3" if i % 3 else "2:5:5")]))
    recs = bamio.sort_records(recs)
    hdr = "@HD\tVN:1.0\tSO:coordinate\n" + "".join("@SQ\tSN:%s\tLN:%d\n" % r for r in refs)
    bamio.write_bam(path, hdr, refs, recs)
    return [r[0] for r in refs]


@pytest.mark.parametrize("seed,n_reads,group,complexity", [(5, 700, True, 2.0), (6, 333, False, 1.0), (7, 64, True, 0.0), (0, 250, True, -1.0),
                                                           (0, 130, True, -2.0)])
def test_decoder_wire_form_holds_the_same_records_as_the_batch(tmp_path, seed, n_reads, group, complexity):
    """SD_DECODE_WIRE: the packed form sd_count_wire sends over PCIe, written by the decoder straight from the BAM records, read
    back with plain numpy and compared field by field with the columnar batch of the same call (and with the Python codec's
    records): positions incl. tiles on several chromosomes / wide position ranges, XI dictionary, CIGAR ops of the shaped
    tiles only, 2-bit base codes, quality codes, MP blocks."""
    from slamdunk_b200.engine import HostBatch
    from slamdunk_b200.io import bam as bamio
    if complexity < 0:
        bam_path = str(tmp_path / "uniform.bam")
        names = _uniform_bam(bam_path, n_reads, L=40 if complexity == -1.0 else 37,      # 37: rows end inside a byte of either column
                             quals=(20, 30, 37) if complexity == -1.0 else (2, 8, 11, 15, 20, 22, 25, 27, 30, 33, 37))
    else:
        c = simdata.make_case(str(tmp_path), seed=seed, n_reads=n_reads, complexity=complexity)
        names, bam_path = list(c.genome.keys()), c.bam
    hb = HostBatch(bam_path, names, want_mp=True, group=group, seq2=True, wire=True)
    assert hb.wire is not None and int(hb.wire.n_tiles) == int(hb.batch.n_tiles) and int(hb.wire.n_reads) == n_reads
    rows, w = _unpack_wire(hb)
    rec, tiles, order = hb.rec_array(), hb.tiles_array(), hb.order()
    _, _, recs = bamio.read_bam(bam_path)
    cig, qual = hb.column("cigar"), hb.column("qual")
    code_of = {"A": 0, "C": 1, "G": 2, "T": 3}
    T = w["tiles"]
    n_shape = 0
    for t in range(int(hb.batch.n_tiles)):
        assert int(T[t]["base_off"]) % 64 == 0 and int(T[t]["var_off"]) % 4 == 0
        n_shape += bool(int(T[t]["flags"]) & 4)
        co, qo = int(tiles[t][2]) // 4, int(tiles[t][1])
        assert int(T[t]["mp_off"]) == int(tiles[t][3]) // 4
        for k in range(32):
            s = t * 32 + k
            if s >= n_reads:
                assert rows[s] is None
                continue
            r, b = rows[s], rec[s]
            lseq, ncig = int(b[3]) & 0xFFFF, int(b[3]) >> 16
            assert (r["pos"], r["ref"], r["mapq"], r["flags"]) == (int(np.int32(b[0])), int(b[1]) & 0xFFFF, (int(b[1]) >> 16) & 0xFF, int(b[1]) >> 24)
            assert np.float32(r["xi"]).tobytes() == np.uint32(b[2]).tobytes() or (np.isnan(r["xi"]) and np.isnan(np.uint32(b[2]).view(np.float32)))
            assert (r["nm"], r["nmp"], r["lseq"]) == (int(b[4]) & 0xFFFF, int(b[4]) >> 16, lseq)
            assert r["ops"] == [int(x) for x in cig[co:co + ncig]]
            src = recs[order[s]]
            assert r["codes"] == [code_of.get(ch, 0) for ch in src.seq.upper()]
            assert r["qual"] == list(qual[qo:qo + lseq]) == src.qual
            co += ncig
            qo += lseq
    if complexity < 0:      # plain-match tiles of one read length carry neither l_seq nor CIGAR
        assert n_shape == 0 and w["qual_bits"] == (2 if complexity == -1.0 else 4) and len(w["xi_dict"]) <= 5
        assert any(int(x) & 1 for x in T["flags"][:-1]) and any(int(x) & 2 for x in T["flags"][:-1])      # 32-bit positions, per-read chromosomes
    else:
        assert n_shape > 0
    assert np.array_equal(w["mp"], hb.column("mp")[:len(w["mp"])])
    hb.close()
