"""Fingerprint of everything sd_decode_bam / sd_load_fasta produce, for before / after comparisons of decoder changes
(test infrastructure):   python tests/tools/decoder_fingerprint.py reads.bam ref.fa [reads2.bam ref2.fa ...] > fp.json
Every BAM is decoded under the six option sets the package uses, on 1 and 3 threads (which must agree); the md5 covers the
record table, every column, the tile table, the order / name-hash arrays, the wire form and the scalar fields."""
import sys, os, hashlib, ctypes, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from slamdunk_b200.engine import HostBatch, HostGenome
from slamdunk_b200 import _lib as L

def raw_qual(hb):
    b = hb.batch
    bits = int(b.qual_bits) or 8
    nq = int(hb.tiles_array()[-1][1])
    nbytes = nq * bits // 8
    if not nbytes: return b""
    return bytes((ctypes.c_uint8 * nbytes).from_address(b.qual))

def fp(path, fasta, **kw):
    g = HostGenome(fasta)
    out = {}
    for threads in (1, 3):
        hb = HostBatch(path, g.names, threads=threads, **kw)
        h = hashlib.md5()
        h.update(hb.rec_array().tobytes()); h.update(hb.tiles_array().tobytes())
        for c in ("seq", "cigar", "mp"):
            h.update(hb.column(c).tobytes())
        h.update(raw_qual(hb))
        b = hb.batch
        h.update(repr((int(b.qual_bits), list(b.qual_dict), int(b.seq_bits), int(b.max_lseq), int(b.max_tile_seq), int(b.max_tile_qual), int(b.max_tile_cig), int(b.max_tile_mp), int(b.max_span), hb.n_mapped_no_xi, hb.n_mapped_no_nm, hb.header_text, hb.ref_names)).encode())
        h.update(hb.order().tobytes()); h.update(hb.name_hash().tobytes())
        if hb.wire is not None:
            w = hb.wire_arrays()
            for k in sorted(w):
                v = w[k]
                h.update(k.encode())
                h.update(v.tobytes() if hasattr(v, "tobytes") else repr(v).encode())
            W = hb.wire
            h.update(repr((int(W.n_reads), int(W.n_tiles), int(W.max_lseq), int(W.max_span), int(W.max_tile_cig), int(W.max_tile_mp), int(W.n_xi_dict))).encode())
        out[threads] = h.hexdigest()
        hb.close()
    g.close()
    assert out[1] == out[3], out
    return out[1]

def fasta_fp(path):
    out = {}
    for threads in (1, 3):
        g = HostGenome(path, threads)
        h = hashlib.md5()
        for a in (g.lo, g.hi, g.nmask, g.chrom_len, g.chrom_off):
            h.update(np.ascontiguousarray(a).tobytes())
        h.update(repr((g.names, g.n_words, g.n_chrom)).encode())
        out[threads] = h.hexdigest()
        g.close()
    assert out[1] == out[3], path
    return out[1]


if __name__ == "__main__":
    res = {}
    args = sys.argv[1:]
    for bam, fa in zip(args[0::2], args[1::2]):
        res[fa] = fasta_fp(fa)
        for name, kw in (("count_verify", dict(want_mp=True, group=True, seq2=True, wire=True)),
                         ("count_cigar", dict(want_mp=False, group=True, seq2=True, wire=True)),
                         ("planes", dict(want_mp=True, group=True, seq2=False)),
                         ("stats", dict(want_mp=True, mp_all=True, group=False)),
                         ("rawq", dict(want_mp=False, raw_qual=True, seq2=True)),
                         ("filter", dict(want_mp=False))):
            res[os.path.basename(bam) + ":" + name] = fp(bam, fa, **kw)
    print(json.dumps(res, indent=0))
