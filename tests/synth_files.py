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


def write_workload_files(outdir, wl, db, triples, name="synth", sort=True):
    """-> (ref.fa, utrs.bed, reads.bam, n_reads)"""
    os.makedirs(outdir, exist_ok=True)
    lo, hi, nm = (t.cpu().numpy().view(np.uint32) for t in (wl.lo, wl.hi, wl.nmask))
    g = ob.genome_from_planes(wl.names, wl.chrom_off, wl.chrom_len, lo, hi, nm)
    ref = os.path.join(outdir, name + ".fa")
    with open(ref, "w") as fh:
        for n_, s in zip(g.names, g.seqs):
            fh.write(">" + n_ + "\n")
            s = s.decode()
            for i in range(0, len(s), 70):
                fh.write(s[i:i + 70] + "\n")
    bed = os.path.join(outdir, name + ".bed")
    with open(bed, "w") as fh:
        for u in range(len(wl.ann)):
            fh.write("%s\t%d\t%d\tutr%05d\t0\t%s\n" % (wl.names[int(wl.ann.chrom[u])], wl.ann.start[u], wl.ann.stop[u], u,
                                                      "-" if wl.ann.strand[u] else "+"))
    bn = ob.device_batch_to_numpy(db)
    n = bn["n_reads"]
    rec = bn["rec"][:n]
    tiles = bn["tiles"]
    seq = db.tensors["seq"].cpu().numpy().view(np.uint32)
    qual = bn["qual"]
    cig = bn["cigar"]
    tri = triples.cpu().numpy()
    T = bn["tile_reads"]
    nt = len(tiles) - 1
    lseq_pad = (bn["rec"][:, 3] & 0xFFFF).astype(np.int64).reshape(nt, T)
    ncig_pad = (bn["rec"][:, 3] >> 16).astype(np.int64).reshape(nt, T)
    tile_seq = (tiles[:-1, 0] // 4).astype(np.int64)     # word offset of every tile's group-major seq block
    qual_off = (np.cumsum(lseq_pad, 1) - lseq_pad + tiles[:-1, 1].astype(np.int64)[:, None]).reshape(-1)[:n]
    cig_off = (np.cumsum(ncig_pad, 1) - ncig_pad + (tiles[:-1, 2] // 4).astype(np.int64)[:, None]).reshape(-1)[:n]
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
        s = _NIB[nib].tobytes().decode()
        cnt = int(tri[i, 0])
        assert cnt <= K
        tags = [("NM", "i", int(w[4]) & 0xFFFF), ("XI", "f", float(np.uint32(w[2]).view(np.float32))), ("XA", "i", 0)]
        if flags & L.SD_F_HAS_MP:
            ents = tri[i, 1:1 + 3 * cnt].reshape(-1, 3)
            # the generator leaves mismatches over a reference N out of its triples; NextGenMap would list them,
            # they can never be T>C / A>G entries, so an (empty) tag keeps has_tag("MP") true
            tags.append(("MP", "Z", ",".join("%d:%d:%d" % tuple(e) for e in ents) if cnt else "24:1:1"))
        records.append(bamio.BamRecord("s%08d" % i, 16 if flags & L.SD_F_REVERSE else 0, int(w[1]) & 0xFFFF, int(np.int32(w[0])),
                                       (int(w[1]) >> 16) & 0xFF, [(int(c) & 15, int(c) >> 4) for c in cig[cig_off[i]:cig_off[i] + nc]],
                                       s, qual[qual_off[i]:qual_off[i] + ls].tolist(), tags))
    if sort:
        records = bamio.sort_records(records)
    md5 = hashlib.md5(open(bed, "rb").read()).hexdigest()
    ds = ("{'sequenced':%d,'mapped':%d,'filtered':%d,'mqfiltered':0,'idfiltered':0,'nmfiltered':0,'multimapper':0,'dedup':0,"
          "'snps':0,'annotation':'%s','annotationmd5':'%s'}" % (n, n, n, os.path.basename(bed), md5))
    hdr = "@HD\tVN:1.0\tSO:coordinate\n" + "".join("@SQ\tSN:%s\tLN:%d\n" % (a, b) for a, b in zip(wl.names, wl.chrom_len))
    hdr += "@RG\tID:7\tSM:%s:pulse:60\tDS:%s\n@PG\tID:slamdunk\tPN:slamdunk filter v0.4.3\tVN:3\n" % (name, ds)
    bam = os.path.join(outdir, name + ".bam")
    bamio.write_bam(bam, hdr, [(a, int(b)) for a, b in zip(wl.names, wl.chrom_len)], records, level=1)
    return ref, bed, bam, n
