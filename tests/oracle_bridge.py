"""Feed a synthetic columnar batch to the CPU oracle (test infrastructure; bench.py's cpu_baseline leg).

The oracle wants what the reference reads from pysam: reference_start / reference_end, flag, mapq,
qualities and the MP mismatch list.  The MP list comes from the generator's own triples
(sd_synth_fill), not from walking the encoded alignment again.
"""
import numpy as np

from oracle import oracle
from slamdunk_b200 import _lib as L

_CONSUMES_REF = np.array([1, 0, 1, 1, 0, 0, 0, 1, 1, 0, 0, 0, 0, 0, 0, 0], dtype=np.int64)


def genome_from_planes(names, chrom_off, chrom_len, lo, hi, nmask):
    """bit-planes -> oracle.Genome (upper-case ACGTN text per chromosome)."""
    lut = np.frombuffer(b"ACGTNNNN", dtype=np.uint8)
    seqs = []
    for off, ln in zip(chrom_off, chrom_len):
        off, ln = int(off), int(ln)
        w0, w1 = off >> 5, (off + ln + 31) >> 5
        bl = np.unpackbits(lo[w0:w1].view(np.uint8), bitorder="little")[:ln]
        bh = np.unpackbits(hi[w0:w1].view(np.uint8), bitorder="little")[:ln]
        bn = np.unpackbits(nmask[w0:w1].view(np.uint8), bitorder="little")[:ln]
        seqs.append(lut[bl + 2 * bh + 4 * bn].tobytes())
    return oracle.Genome(list(names), seqs)


def batch_to_oracle(batch_np, triples):
    """batch_np: dict of host numpy arrays rec [n_pad,5] u32, cigar u32, qual u8, tiles [nt+1,4] u64,
    n_reads, tile_reads.  triples: int32 [n, 1 + 3K] -> (reads READ_DT, qual_pool, mp_triples)"""
    n, T = batch_np["n_reads"], batch_np["tile_reads"]
    rec = batch_np["rec"][:n]
    tiles = batch_np["tiles"]
    nt = len(tiles) - 1
    lseq = (rec[:, 3] & 0xFFFF).astype(np.int64)
    ncig_pad = (batch_np["rec"][:, 3] >> 16).astype(np.int64).reshape(nt, T)
    lseq_pad = (batch_np["rec"][:, 3] & 0xFFFF).astype(np.int64).reshape(nt, T)
    cig_off = (np.cumsum(ncig_pad, axis=1) - ncig_pad + (tiles[:-1, 2] // 4).astype(np.int64)[:, None]).reshape(-1)[:n]
    qual_off = (np.cumsum(lseq_pad, axis=1) - lseq_pad + tiles[:-1, 1].astype(np.int64)[:, None]).reshape(-1)[:n]
    ncig = ncig_pad.reshape(-1)[:n]
    cig = batch_np["cigar"].astype(np.int64)
    contrib = (cig >> 4) * _CONSUMES_REF[cig & 15]
    C = np.concatenate([[0], np.cumsum(contrib)])
    reflen = C[cig_off + ncig] - C[cig_off]
    flags = rec[:, 1] >> 24
    reads = np.zeros(n, dtype=oracle.READ_DT)
    reads["chrom"] = (rec[:, 1] & 0xFFFF).astype(np.int32)
    reads["pos"] = rec[:, 0].view(np.int32)
    reads["end"] = reads["pos"] + np.maximum(1, reflen).astype(np.int32)
    reads["l_qual"] = lseq
    reads["flag"] = np.where(flags & L.SD_F_REVERSE, 16, 0) | np.where(flags & L.SD_F_SECONDARY, 0x100, 0)
    reads["mapq"] = (rec[:, 1] >> 16) & 0xFF
    reads["has_mp"] = (flags & L.SD_F_HAS_MP) != 0
    reads["qual_off"] = qual_off
    K = (triples.shape[1] - 1) // 3
    cnt = triples[:, 0].astype(np.int64)
    assert cnt.max(initial=0) <= K, "a read has more mismatches than SD_SYNTH_MAX_TRIPLES"
    reads["mp_len"] = cnt
    reads["mp_off"] = np.cumsum(cnt) - cnt
    tri = triples[:, 1:].reshape(n, K, 3)
    mask = np.arange(K)[None, :] < cnt[:, None]
    mp_triples = np.ascontiguousarray(tri[mask], dtype=np.int32)
    keep = ((flags & (L.SD_F_UNMAPPED | L.SD_F_SKIP | L.SD_F_PAD)) == 0)
    return reads[keep], batch_np["qual"], mp_triples


def device_batch_to_numpy(db):
    """DeviceBatch -> the dict batch_to_oracle wants."""
    b = db.batch
    t = db.tensors
    return {
        "n_reads": int(b.n_reads), "tile_reads": int(b.tile_reads),
        "rec": t["rec"].cpu().numpy().view(np.uint32).reshape(-1, 5),
        "cigar": t["cigar"].cpu().numpy().view(np.uint32),
        "qual": t["qual"].cpu().numpy(),
        "tiles": t["tiles"].cpu().numpy().view(np.uint64).reshape(-1, 4),
    }


def annotation_to_oracle(ann, n_chrom):
    utrs = np.zeros(len(ann), dtype=oracle.UTR_DT)
    utrs["chrom"] = ann.chrom
    utrs["in_bam"] = 1
    utrs["start"] = ann.start
    utrs["stop"] = ann.stop
    utrs["strand"] = ann.strand
    utrs["name_id"] = ann.chrom
    return utrs


def compare(res, rows, bg, ann, pos_off, chrom_off):
    """GPU engine results vs oracle rows / bedgraph entries; returns a list of problems (empty = parity)."""
    bad = []
    c = res["counters"]
    pairs = (("read_count", L.SD_C_READS), ("tc_read_count", L.SD_C_TCREADS), ("multimap_count", L.SD_C_MULTIMAP),
             ("coverage_on_ts", L.SD_C_COV_ON_T), ("conversions_on_ts", L.SD_C_CONV_ON_T))
    for name, k in pairs:
        if not np.array_equal(rows[name], c[:, k]):
            i = int(np.nonzero(rows[name] != c[:, k])[0][0])
            bad.append("%s differs at UTR %d: oracle %d gpu %d" % (name, i, rows[name][i], c[i, k]))
    if not np.array_equal(rows["tcontent"], res["tcontent"]):
        bad.append("tcontent differs")
    if len(bg):
        # every oracle bedgraph entry (chrom, pos, strand) -> compact offset through the first UTR containing it
        order = np.lexsort((ann.start, ann.chrom))
        assert (order == np.arange(len(order))).all()
        for s in (0, 1):
            e = bg[bg["strand"] == s]
            if not len(e):
                continue
            sel = np.nonzero(ann.strand == s)[0]
            key_u = ann.chrom[sel].astype(np.int64) * (1 << 32) + ann.start[sel]
            key_e = e["name_id"].astype(np.int64) * (1 << 32) + e["pos"]
            # last UTR (same strand) starting at or before the position; walk back until it contains it
            j = np.searchsorted(key_u, key_e, side="right") - 1
            for _ in range(64):
                miss = (ann.stop[sel[j]] <= e["pos"]) | (ann.chrom[sel[j]] != e["name_id"])
                if not miss.any():
                    break
                j = np.where(miss, j - 1, j)
            u = sel[j]
            off = pos_off[u].astype(np.int64) + (e["pos"] - ann.start[u])
            if not np.array_equal(res["pos_cov"][off], e["cov"]):
                bad.append("per-position coverage differs on strand %d" % s)
            if not np.array_equal(res["pos_tc"][off], e["tc"]):
                bad.append("per-position conversions differ on strand %d" % s)
    return bad
