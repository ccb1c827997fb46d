"""Write a synthetic workload (slamdunk_b200.synth) out as the FILES the reference pipeline reads --
FASTA, BED6 and a filtered, coordinate-sorted BAM with NextGenMap-style tags -- so that the file-level
entry points can be exercised on data sets far larger than the Python-simulated ones (test infrastructure)."""
import hashlib
import os

import numpy as np

import oracle_bridge as ob
from slamdunk_b200 import _lib as L
from slamdunk_b200.io import bam as bamio

_NIB = np.frombuffer(b"=ACMGRSVTWYHKDBN", dtype=np.uint8)


def write_workload_files(outdir, wl, db, triples, name="synth", sort=True, slow=False):
    """-> (ref.fa, utrs.bed, reads.bam, n_reads)"""
    os.makedirs(outdir, exist_ok=True)
    lo, hi, nm = (t.cpu().numpy().view(np.uint32) for t in (wl.lo, wl.hi, wl.nmask))
    g = ob.genome_from_planes(wl.names, wl.chrom_off, wl.chrom_len, lo, hi, nm)
    ref = os.path.join(outdir, name + ".fa")
    with open(ref, "wb") as fh:                                  # 70 bases per line
        for n_, s in zip(g.names, g.seqs):
            fh.write(b">" + n_.encode() + b"\n")
            a = np.frombuffer(s, dtype=np.uint8)
            full = len(a) // 70
            body = np.empty((full, 71), dtype=np.uint8)
            body[:, :70] = a[:full * 70].reshape(full, 70)
            body[:, 70] = 10
            fh.write(body.tobytes())
            if len(a) % 70:
                fh.write(a[full * 70:].tobytes() + b"\n")
    bed = os.path.join(outdir, name + ".bed")
    with open(bed, "w") as fh:
        for u in range(len(wl.ann)):
            fh.write("%s\t%d\t%d\tutr%05d\t0\t%s\n" % (wl.names[int(wl.ann.chrom[u])], wl.ann.start[u], wl.ann.stop[u], u,
                                                      "-" if wl.ann.strand[u] else "+"))
    bn = ob.device_batch_to_numpy(db)
    n = bn["n_reads"]
    tri = triples.cpu().numpy()
    seq_words = db.tensors["seq"].cpu().numpy().view(np.uint32)
    stream = (_record_stream_slow if slow else _record_stream)(bn, seq_words, tri, sort)
    md5 = hashlib.md5(open(bed, "rb").read()).hexdigest()
    ds = ("{'sequenced':%d,'mapped':%d,'filtered':%d,'mqfiltered':0,'idfiltered':0,'nmfiltered':0,'multimapper':0,'dedup':0,"
          "'snps':0,'annotation':'%s','annotationmd5':'%s'}" % (n, n, n, os.path.basename(bed), md5))
    hdr = "@HD\tVN:1.0\tSO:coordinate\n" + "".join("@SQ\tSN:%s\tLN:%d\n" % (a, b) for a, b in zip(wl.names, wl.chrom_len))
    hdr += "@RG\tID:7\tSM:%s:pulse:60\tDS:%s\n@PG\tID:slamdunk\tPN:slamdunk filter v0.4.3\tVN:3\n" % (name, ds)
    bam = os.path.join(outdir, name + ".bam")
    with open(bam, "wb") as fh:
        fh.write(bamio.bgzf_compress(bamio.encode_bam(hdr, [(a, int(b)) for a, b in zip(wl.names, wl.chrom_len)], []) + stream, 1))
    return ref, bed, bam, n


def _layout(bn):
    """per read: offsets of its rows in the qual / cigar columns and of its tile's group-major SEQ block"""
    n, T = bn["n_reads"], bn["tile_reads"]
    tiles = bn["tiles"]
    nt = len(tiles) - 1
    lseq_pad = (bn["rec"][:, 3] & 0xFFFF).astype(np.int64).reshape(nt, T)
    ncig_pad = (bn["rec"][:, 3] >> 16).astype(np.int64).reshape(nt, T)
    tile_seq = (tiles[:-1, 0] // 4).astype(np.int64)     # word offset of every tile's group-major seq block
    qual_off = (np.cumsum(lseq_pad, 1) - lseq_pad + tiles[:-1, 1].astype(np.int64)[:, None]).reshape(-1)[:n]
    cig_off = (np.cumsum(ncig_pad, 1) - ncig_pad + (tiles[:-1, 2] // 4).astype(np.int64)[:, None]).reshape(-1)[:n]
    return tile_seq, qual_off, cig_off


def _mp_text(tri_row, cnt):
    # the generator leaves mismatches over a reference N out of its triples; NextGenMap would list them,
    # they can never be T>C / A>G entries, so an (empty) tag keeps has_tag("MP") true
    return ",".join("%d:%d:%d" % tuple(e) for e in tri_row[1:1 + 3 * cnt].reshape(-1, 3)) if cnt else "24:1:1"


def _record_stream_slow(bn, seq, tri, sort):
    """BamRecord by BamRecord through the Python codec (the first writer; kept as the check of the fast one)."""
    n, T = bn["n_reads"], bn["tile_reads"]
    rec, qual, cig = bn["rec"][:n], bn["qual"], bn["cigar"]
    tile_seq, qual_off, cig_off = _layout(bn)
    K = (tri.shape[1] - 1) // 3
    records = []
    for i in range(n):
        w = rec[i]
        ls, nc = int(w[3]) & 0xFFFF, int(w[3]) >> 16
        flags = int(w[1]) >> 24
        gi = tile_seq[i // T] + 4 * (np.arange((ls + 31) // 32) * T + i % T)               # group g of row r at 16 * (g * T + r)
        words = seq[(gi[:, None] + np.arange(4)[None, :]).reshape(-1)]
        planes = np.unpackbits(words.view(np.uint8).reshape(-1, 4, 4), axis=2, bitorder="little")   # [group, plane, 32 bases]
        nib = (planes[:, 0, :] | (planes[:, 1, :] << 1) | (planes[:, 2, :] << 2) | (planes[:, 3, :] << 3)).reshape(-1)[:ls]
        cnt = int(tri[i, 0])
        assert cnt <= K
        tags = [("NM", "i", int(w[4]) & 0xFFFF), ("XI", "f", float(np.uint32(w[2]).view(np.float32))), ("XA", "i", 0)]
        if flags & L.SD_F_HAS_MP:
            tags.append(("MP", "Z", _mp_text(tri[i], cnt)))
        records.append(bamio.BamRecord("s%08d" % i, 16 if flags & L.SD_F_REVERSE else 0, int(w[1]) & 0xFFFF, int(np.int32(w[0])),
                                       (int(w[1]) >> 16) & 0xFF, [(int(c) & 15, int(c) >> 4) for c in cig[cig_off[i]:cig_off[i] + nc]],
                                       _NIB[nib].tobytes().decode(), qual[qual_off[i]:qual_off[i] + ls].tolist(), tags))
    if sort:
        records = bamio.sort_records(records)
    return b"".join(bamio.encode_record(r) for r in records)


_CONSUMES_REF = np.array([1, 0, 1, 1, 0, 0, 0, 1, 1, 0, 0, 0, 0, 0, 0, 0], dtype=np.int64)      # M I D N S H P = X


def _reg2bin(beg, end):
    """bamio.reg2bin on arrays (SAM spec 5.3)"""
    end = end - 1
    out = np.zeros(len(beg), dtype=np.int64)
    done = np.zeros(len(beg), dtype=bool)
    for shift, base in ((14, 4681), (17, 585), (20, 73), (23, 9), (26, 1)):
        hit = ~done & ((beg >> shift) == (end >> shift))
        out[hit] = base + (beg[hit] >> shift)
        done |= hit
    return out


def _record_stream(bn, seq, tri, sort, chunk=50000):
    """The same bytes with the per-base work done on whole chunks of reads: nibbles from the SEQ planes, quality rows, BAM
    bins and the fixed 36-byte core as arrays; Python touches a read only to join its pieces and print its MP tag."""
    n, T = bn["n_reads"], bn["tile_reads"]
    rec, qual, cig = bn["rec"][:n], bn["qual"], bn["cigar"].astype(np.int64)
    tile_seq, qual_off, cig_off = _layout(bn)
    K = (tri.shape[1] - 1) // 3
    assert n == 0 or int(tri[:n, 0].max()) <= K
    lseq = (rec[:, 3] & 0xFFFF).astype(np.int64)
    ncig = (rec[:, 3] >> 16).astype(np.int64)
    flags = (rec[:, 1] >> 24).astype(np.int64)
    ref_id = (rec[:, 1] & 0xFFFF).astype(np.int64)
    pos = rec[:, 0].view(np.int32).astype(np.int64)
    mapq = ((rec[:, 1] >> 16) & 0xFF).astype(np.int64)
    bam_flag = np.where(flags & L.SD_F_REVERSE, 16, 0)
    contrib = np.concatenate([[0], np.cumsum((cig >> 4) * _CONSUMES_REF[cig & 15])])
    reflen = contrib[cig_off + ncig] - contrib[cig_off]
    end = np.where((ncig > 0) & (reflen > 0), pos + reflen, pos + 1)         # BamRecord.reference_end, else pos + 1 (encode_record)
    bins = np.where(pos >= 0, _reg2bin(np.maximum(pos, 0), end), 4680)
    order = np.arange(n)
    if sort:                                                                 # bamio.sort_records: (ref, pos, strand), ties in input order
        order = np.lexsort((bam_flag, pos, np.where(ref_id >= 0, ref_id, 1 << 30)))
    nm = (rec[:, 4] & 0xFFFF).astype(np.int64)
    xi_bytes = np.ascontiguousarray(rec[:, 2]).view(np.uint8).reshape(n, 4)
    core_dt = np.dtype([("block_size", "<i4"), ("ref_id", "<i4"), ("pos", "<i4"), ("l_name", "u1"), ("mapq", "u1"), ("bin", "<u2"),
                        ("n_cig", "<u2"), ("flag", "<u2"), ("l_seq", "<i4"), ("nref", "<i4"), ("npos", "<i4"), ("tlen", "<i4")])
    pieces = [None] * n
    for c0 in range(0, n, chunk):
        c1 = min(n, c0 + chunk)
        m = c1 - c0
        ls = lseq[c0:c1]
        Lmax = int(ls.max()) if m else 0
        G = (Lmax + 31) // 32
        idx = np.arange(c0, c1)
        # group g of row r of a tile at word 4 * (g * T + r) of the tile's block; groups past a read's own are masked below
        gi = tile_seq[idx // T][:, None] + 4 * (np.arange(G)[None, :] * T + (idx % T)[:, None])
        words = seq[np.minimum(gi[:, :, None] + np.arange(4)[None, None, :], len(seq) - 1)]              # [m, G, 4 planes]
        bits = np.unpackbits(np.ascontiguousarray(words).view(np.uint8).reshape(m, G, 4, 4), axis=3, bitorder="little")
        nib = (bits[:, :, 0, :] | (bits[:, :, 1, :] << 1) | (bits[:, :, 2, :] << 2) | (bits[:, :, 3, :] << 3)).reshape(m, G * 32)
        nib[np.arange(G * 32)[None, :] >= ls[:, None]] = 0
        if (G * 32) & 1:
            nib = np.concatenate([nib, np.zeros((m, 1), np.uint8)], axis=1)
        packed = (nib[:, 0::2] << 4) | nib[:, 1::2]
        qrows = qual[np.minimum(qual_off[c0:c1][:, None] + np.arange(Lmax)[None, :], len(qual) - 1)] if Lmax else np.zeros((m, 0), np.uint8)
        core = np.zeros(m, dtype=core_dt)
        core["ref_id"], core["pos"], core["l_name"], core["mapq"] = ref_id[c0:c1], pos[c0:c1], 10, mapq[c0:c1]
        core["bin"], core["n_cig"], core["flag"], core["l_seq"] = bins[c0:c1], ncig[c0:c1], bam_flag[c0:c1], ls
        core["nref"], core["npos"], core["tlen"] = -1, -1, 0
        cig32 = cig.astype("<u4")
        for j in range(m):
            i = c0 + j
            l, nc = int(ls[j]), int(ncig[i])
            v = int(nm[i])
            tags = (b"NMC" + bytes((v,)) if v <= 0xFF else b"NMS" + v.to_bytes(2, "little")) + b"XIf" + xi_bytes[i].tobytes() + b"XAC\0"
            if flags[i] & L.SD_F_HAS_MP:
                tags += b"MPZ" + _mp_text(tri[i], int(tri[i, 0])).encode("ascii") + b"\0"
            body = (b"s%08d\0" % i) + cig32[cig_off[i]:cig_off[i] + nc].tobytes() + packed[j, :(l + 1) // 2].tobytes() + qrows[j, :l].tobytes() + tags
            core["block_size"][j] = 32 + len(body)
            pieces[i] = body
        cores = core.tobytes()
        for j in range(m):
            pieces[c0 + j] = cores[36 * j:36 * j + 36] + pieces[c0 + j]
    return b"".join(pieces[i] for i in order)
