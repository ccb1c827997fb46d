"""Synthetic SLAM-seq workloads (BASELINE.json configs) generated on the device.

Host side: genome layout, UTR annotation and expression profile (numpy, seeded).  Device side:
sd_synth_genome / sd_synth_plan / sd_synth_fill in libslamdunk_b200.so write the genome planes and
the columnar read batches.  SURVEY.md section 8(d) gives the recipe (cfg2: ~100 Mb genome on 21
contigs, 20 000 UTRs with log-normal lengths, geometric expression profile, 100 bp reads, 2.5 %
T>C per reference T, 0.2 % errors, 5 % soft clips, 2 % indels, 3 % mapq 0, 98 % sense strand).
"""
from __future__ import annotations

import ctypes
from typing import Optional

import numpy as np
import torch

from . import _lib as L
from .engine import CountEngine, DeviceBatch

MOUSE_LIKE_CHROMS = ["chr%d" % i for i in range(1, 20)] + ["chrX", "chrY"]


def genome_layout(chrom_len):
    """chromosome starts in the linear bit space: multiples of 32, >= 32 pad positions between."""
    off, bits = [], 32
    for l in chrom_len:
        off.append(bits)
        bits = ((bits + int(l) + 31) // 32) * 32 + 32
    return np.array(off, dtype=np.uint32), bits // 32


class Annotation(object):
    """UTRs sorted by genome position (BED order = genome order) with expression weights."""

    def __init__(self, chrom, start, stop, strand, weight):
        self.chrom = np.asarray(chrom, dtype=np.uint32)
        self.start = np.asarray(start, dtype=np.int64)
        self.stop = np.asarray(stop, dtype=np.int64)
        self.strand = np.asarray(strand, dtype=np.uint8)
        self.weight = np.asarray(weight, dtype=np.float64)

    def __len__(self):
        return len(self.chrom)


def make_annotation(seed, chrom_len, n_utr, overlap_frac=0.05):
    rng = np.random.RandomState(seed)
    chrom_len = np.asarray(chrom_len, dtype=np.int64)
    share = chrom_len / chrom_len.sum()
    per = np.floor(share * n_utr).astype(int)
    per[0] += n_utr - per.sum()
    chrom, start, stop, strand = [], [], [], []
    for c, k in enumerate(per):
        if k <= 0:
            continue
        ln = np.clip(np.exp(rng.normal(np.log(1000.0), 0.8, size=k)), 200, 10000).astype(np.int64)
        slot = chrom_len[c] // k
        ln = np.minimum(ln, max(50, slot - 10))
        base = np.arange(k, dtype=np.int64) * slot
        st = base + (rng.rand(k) * np.maximum(1, slot - ln)).astype(np.int64)
        ov = rng.rand(k) < overlap_frac                      # overlap the previous UTR
        ov[0] = False
        prev_stop = np.roll(st + ln, 1)
        st = np.where(ov, np.maximum(0, prev_stop - (ln // 3 + 1)), st)
        sp = np.minimum(st + ln, chrom_len[c])
        chrom.append(np.full(k, c)); start.append(st); stop.append(sp); strand.append(rng.randint(0, 2, size=k))
    chrom = np.concatenate(chrom); start = np.concatenate(start); stop = np.concatenate(stop); strand = np.concatenate(strand)
    order = np.lexsort((start, chrom))
    chrom, start, stop, strand = chrom[order], start[order], stop[order], strand[order]
    # geometric expression profile over a random ranking of the UTRs
    n = len(chrom)
    rank = rng.permutation(n)
    p = 8.0 / n
    weight = np.power(1.0 - p, rank.astype(np.float64)) + 1e-4
    return Annotation(chrom, start, stop, strand, weight)


class SynthWorkload(object):
    """Genome + annotation resident on one GPU; hands out read batches."""

    def __init__(self, engine: CountEngine, seed=1, genome_bp=100_000_000, n_utr=20000, chrom_names=None,
                 n_frac=0.001):
        self.eng = engine
        self.lib = engine.lib
        dev = engine.device
        self.names = list(chrom_names or MOUSE_LIKE_CHROMS)
        nc = len(self.names)
        # mouse-like decreasing chromosome sizes
        w = np.linspace(1.6, 0.5, nc)
        self.chrom_len = np.maximum(2000, (w / w.sum() * genome_bp)).astype(np.uint32)
        self.chrom_off, self.n_words = genome_layout(self.chrom_len)
        with torch.cuda.stream(engine.stream):
            self.lo = torch.empty(self.n_words, dtype=torch.int32, device=dev)
            self.hi = torch.empty(self.n_words, dtype=torch.int32, device=dev)
            self.nmask = torch.empty(self.n_words, dtype=torch.int32, device=dev)
            self.d_chrom_off = torch.from_numpy(self.chrom_off.view(np.int32)).to(dev)
            self.d_chrom_len = torch.from_numpy(self.chrom_len.view(np.int32)).to(dev)
        engine.stream.synchronize()
        engine._chk(self.lib.sd_synth_genome(engine.ctx, seed, self.lo.data_ptr(), self.hi.data_ptr(), self.nmask.data_ptr(),
                                             self.n_words, self.chrom_off.ctypes.data, self.chrom_len.ctypes.data, nc,
                                             ctypes.c_float(n_frac)), "sd_synth_genome")
        self.ann = make_annotation(seed + 1, self.chrom_len, n_utr)
        a = self.ann
        cum = np.cumsum(a.weight)
        with torch.cuda.stream(engine.stream):
            self.d_utr_chrom = torch.from_numpy(a.chrom.astype(np.int32)).to(dev)
            self.d_utr_start = torch.from_numpy(a.start.astype(np.int32)).to(dev)
            self.d_utr_end = torch.from_numpy(a.stop.astype(np.int32)).to(dev)
            self.d_utr_strand = torch.from_numpy(a.strand.astype(np.uint8)).to(dev)
            self.d_cum = torch.from_numpy(cum).to(dev)
        engine.stream.synchronize()
        self.cum_weight = cum
        t = L.SynthTables()
        t.lo, t.hi, t.nmask = self.lo.data_ptr(), self.hi.data_ptr(), self.nmask.data_ptr()
        t.chrom_off, t.chrom_len, t.n_chrom = self.d_chrom_off.data_ptr(), self.d_chrom_len.data_ptr(), nc
        t.n_utr = len(a)
        t.utr_chrom, t.utr_start, t.utr_end = self.d_utr_chrom.data_ptr(), self.d_utr_start.data_ptr(), self.d_utr_end.data_ptr()
        t.utr_strand, t.utr_cum_weight = self.d_utr_strand.data_ptr(), self.d_cum.data_ptr()
        self.tables = t

    def set_time_point(self, minutes=None, halflife_seed=7):
        """4sU time course (BASELINE.json config 3): every UTR gets a half-life drawn uniformly from [30, 720] min
        (splash.py:126-127) and, at labelling time `minutes`, the labelled-read fraction 1 - exp(-ln2 t / halflife)
        (simulator.py:291-302); None goes back to the global frac_labelled of params()."""
        if minutes is None:
            self.d_label_frac = None
            self.tables.utr_label_frac = None
            return None
        hl = np.random.default_rng(halflife_seed).uniform(30.0, 720.0, len(self.ann))
        frac = (1.0 - np.exp(-np.log(2.0) * float(minutes) / hl)).astype(np.float32)
        with torch.cuda.stream(self.eng.stream):
            self.d_label_frac = torch.from_numpy(frac).to(self.eng.device)
        self.eng.stream.synchronize()
        self.tables.utr_label_frac = self.d_label_frac.data_ptr()
        return frac

    def set_sample_snps(self, spacing=500, seed=30, p_allele=0.9, tc_ag_share=0.5, multi_allelic=0.02):
        """BASELINE.json config 4: the sample's own SNPs, one per `spacing` bp genome-wide.  `tc_ag_share` of them are T>C /
        A>G (placed on a reference T / A), the rest other substitutions; `multi_allelic` of the T>C / A>G ones are written
        with a second alternative allele ("C,G"), which the reference's SNPDictionary never matches (SNPtools.py:33-37).
        Reads of the generator then carry the alternative allele with probability p_allele.
        -> VCF records [(CHROM, POS text, REF, ALT)] as VarScan would list them, for utils.snpmask.SNPMasks / the oracle."""
        rng = np.random.default_rng(seed)
        lo, hi, nm = (t.cpu().numpy().view(np.uint32) for t in (self.lo, self.hi, self.nmask))
        tc = np.zeros(self.n_words, dtype=np.uint32)
        ag = np.zeros(self.n_words, dtype=np.uint32)
        other = np.zeros(self.n_words, dtype=np.uint32)
        records = []
        letters = "ACGT"
        for c, name in enumerate(self.names):
            clen, off = int(self.chrom_len[c]), int(self.chrom_off[c])
            k = clen // spacing
            if k == 0:
                continue
            pos = np.unique(rng.integers(0, clen, size=k))
            g = off + pos
            w, b = g >> 5, (g & 31).astype(np.uint32)
            code = ((lo[w] >> b) & 1) | (((hi[w] >> b) & 1) << 1)
            isn = ((nm[w] >> b) & 1).astype(bool)
            want_conv = rng.random(len(pos)) < tc_ag_share
            multi = rng.random(len(pos)) < multi_allelic
            for j in range(len(pos)):
                if isn[j]:
                    continue
                cj = int(code[j])
                bit = np.uint32(1) << b[j]
                if want_conv[j] and cj in (0, 3):
                    alt = "G" if cj == 0 else "C"
                    (ag if cj == 0 else tc)[w[j]] |= bit
                    records.append((name, str(int(pos[j]) + 1), letters[cj], alt + (",T" if multi[j] and cj == 0 else (",G" if multi[j] else ""))))
                else:
                    other[w[j]] |= bit
                    records.append((name, str(int(pos[j]) + 1), letters[cj], letters[(cj + 1) & 3]))
        with torch.cuda.stream(self.eng.stream):
            self.d_snp = [torch.from_numpy(x.view(np.int32)).to(self.eng.device) for x in (tc, ag, other)]
        self.eng.stream.synchronize()
        self.tables.snp_tc, self.tables.snp_ag, self.tables.snp_other = (t.data_ptr() for t in self.d_snp)
        self.tables.p_snp_allele = float(p_allele)
        return records

    def install(self):
        """Hand genome and annotation to the engine (sd_set_reference / sd_set_utrs)."""
        self.eng.set_reference(self.lo, self.hi, self.nmask, self.chrom_off, self.chrom_len)
        a = self.ann
        self.eng.set_utrs(a.chrom, a.start, a.stop, a.strand, None)

    def region_split(self, n_parts):
        """Contiguous UTR ranges (genome order) with equal expected read counts -> [(first, count)]."""
        cum = self.cum_weight
        cuts = [0]
        for k in range(1, n_parts):
            cuts.append(int(np.searchsorted(cum, cum[-1] * k / n_parts)))
        cuts.append(len(cum))
        for k in range(1, len(cuts)):
            cuts[k] = max(cuts[k], cuts[k - 1] + 1)
        cuts[-1] = len(cum)
        return [(cuts[k], cuts[k + 1] - cuts[k]) for k in range(n_parts)]

    def params(self, seed, first_read, n_reads, read_len=100, tile_reads=32, p_conv=0.025, p_err=0.002,
               frac_softclip=0.05, frac_indel=0.02, frac_multimap=0.03, frac_antisense=0.02, frac_labelled=1.0,
               utr_first=0, utr_count=0) -> L.SynthParams:
        p = L.SynthParams()
        p.seed, p.first_read, p.n_reads = int(seed), int(first_read), int(n_reads)
        p.read_len, p.tile_reads = int(read_len), int(tile_reads)
        p.p_conv, p.p_err = p_conv, p_err
        p.frac_softclip, p.frac_indel, p.frac_multimap = frac_softclip, frac_indel, frac_multimap
        p.frac_antisense, p.frac_labelled = frac_antisense, frac_labelled
        p.utr_first, p.utr_count = int(utr_first), int(utr_count)
        return p

    @staticmethod
    def _copy_params(p: L.SynthParams) -> L.SynthParams:
        q = L.SynthParams()
        ctypes.memmove(ctypes.byref(q), ctypes.byref(p), ctypes.sizeof(L.SynthParams))
        return q

    def generate(self, p: L.SynthParams, want_mp=False, want_triples=False, coordinate_sorted=False, group=None, gpos_range=None):
        """-> (DeviceBatch, triples tensor or None).
        coordinate_sorted: order the reads by alignment start like the sorted BAM `slamdunk count` reads;
        group (default: same as coordinate_sorted): additionally order them by alignment shape and strand the way
        sd_decode_bam's SD_DECODE_GROUP does (plain match blocks first).  Triples follow the batch order.
        gpos_range = (lo, hi): keep only the reads of the index range [first_read, first_read + n_reads) whose alignment
        starts at a linear genome position in [lo, hi) -- one rank's share of a read stream sharded by chromosome region
        (BASELINE.json config 5); returns (None, None) when no read falls there.  The reads themselves depend only on
        (seed, read index), so the shards of any split add up to the unsharded stream."""
        eng, dev = self.eng, self.eng.device
        n = int(p.n_reads)
        T = int(p.tile_reads)
        order = None
        if group is None:
            group = coordinate_sorted
        if coordinate_sorted or group or gpos_range is not None:
            with torch.cuda.stream(eng.stream):
                gpos = torch.empty(n, dtype=torch.int32, device=dev)
                cls = torch.empty(n, dtype=torch.uint8, device=dev)
            eng.stream.synchronize()
            p.order = None
            eng._chk(self.lib.sd_synth_positions(eng.ctx, ctypes.byref(p), ctypes.byref(self.tables), gpos.data_ptr(),
                                                 cls.data_ptr()), "sd_synth_positions")
            with torch.cuda.stream(eng.stream):
                g64 = gpos.to(torch.int64) & 0xFFFFFFFF
                key = g64 if coordinate_sorted else torch.arange(n, device=dev, dtype=torch.int64)
                if group:
                    key = key | (cls.to(torch.int64) << 40)
                if gpos_range is not None:
                    sel = torch.nonzero((g64 >= int(gpos_range[0])) & (g64 < int(gpos_range[1]))).view(-1)
                    order = sel[torch.argsort(key[sel], stable=True)].to(torch.int32)
                    del sel
                else:
                    order = torch.argsort(key, stable=True).to(torch.int32)
                del gpos, cls, key, g64
            eng.stream.synchronize()
            if gpos_range is not None:
                n = int(order.numel())
                if n == 0:
                    return None, None
                p = self._copy_params(p)
                p.n_reads = n
            p.order = order.data_ptr()
        n_tiles = (n + T - 1) // T
        with torch.cuda.stream(eng.stream):
            counts = torch.empty(n, dtype=torch.int32, device=dev)
            tiles = torch.empty((n_tiles + 1) * 4, dtype=torch.int64, device=dev)
        eng.stream.synchronize()
        totals = (ctypes.c_uint64 * 4)()
        eng._chk(self.lib.sd_synth_plan(eng.ctx, ctypes.byref(p), ctypes.byref(self.tables), counts.data_ptr(),
                                        tiles.data_ptr(), totals), "sd_synth_plan")
        with torch.cuda.stream(eng.stream):
            tens = {
                "rec": torch.empty(n_tiles * T * 20, dtype=torch.uint8, device=dev),
                "seq": torch.empty(max(16, totals[0]), dtype=torch.uint8, device=dev),
                "qual": torch.zeros(max(16, totals[1]), dtype=torch.uint8, device=dev),
                "cigar": torch.zeros(max(16, totals[2]), dtype=torch.uint8, device=dev),
                "mp": torch.zeros(max(16, totals[3]), dtype=torch.uint8, device=dev) if want_mp else None,
                "tiles": tiles,
            }
            triples = torch.zeros((n, 1 + 3 * L.SD_SYNTH_MAX_TRIPLES), dtype=torch.int32, device=dev) if want_triples else None
        eng.stream.synchronize()
        b = L.Batch()
        b.rec, b.seq, b.qual, b.cigar = tens["rec"].data_ptr(), tens["seq"].data_ptr(), tens["qual"].data_ptr(), tens["cigar"].data_ptr()
        b.mp = tens["mp"].data_ptr() if want_mp else None
        b.tiles = tiles.data_ptr()
        eng._chk(self.lib.sd_synth_fill(eng.ctx, ctypes.byref(p), ctypes.byref(self.tables), counts.data_ptr(),
                                        ctypes.byref(b), triples.data_ptr() if want_triples else None), "sd_synth_fill")
        p.order = None
        del order
        return DeviceBatch(b, tens), triples


def algorithmic_bytes_per_read(read_len, mean_cigar_ops):
    """compulsory bytes per read of the path's INPUT, each touched once, independent of how the
    columns are coded in HBM -- ceil(L/2) (SEQ as BAM stores it, 4 bits per base) + L (QUAL) + 4 c (CIGAR ops) +
    12 (pos, ref_id + flag + mapq + n_cigar, XI) + 4 (NM + pad) + 1 (per-read tc output): 171 B at L = 100, c = 1."""
    return (read_len + 1) // 2 + read_len + 4.0 * mean_cigar_ops + 17


def layout_bytes_per_read(read_len, mean_cigar_ops, seq_bits=4, qual_planes=0):
    """Bytes per read of the columns as THIS layout holds them in HBM: SEQ in 32-base groups of
    seq_bits planes, QUAL as bytes (of which the count kernel touches only those at candidates) or as qual_planes
    bit-planes per group, CIGAR ops, the 20-byte record."""
    groups = (read_len + 31) // 32
    return seq_bits * 4 * groups + (qual_planes * 4 * groups if qual_planes else read_len) + 4.0 * mean_cigar_ops + 20


def to_mp_all(db: DeviceBatch, triples: torch.Tensor, eng) -> DeviceBatch:
    """The same device batch with the COMPLETE MP column sd_read_stats takes (what sd_decode_bam writes with
    SD_DECODE_MP_ALL: per entry {readPos-1 | refPos-1 << 16, conv}), built from the generator's own record of the
    mismatching columns (`triples`, batch order; 1-based positions as NextGenMap lists them).  A read that carries an MP
    tag but no mismatch gets the entry 24:1:1 that tests/synth_files.py writes into its BAM, so that both routes agree."""
    dev = db.tensors["rec"].device
    nt, T, n = int(db.batch.n_tiles), int(db.batch.tile_reads), int(db.batch.n_reads)
    with torch.cuda.stream(eng.stream):
        rec = db.tensors["rec"].clone().view(torch.int32).view(-1, 5)
        has_mp = ((rec[:n, 1] >> 24) & L.SD_F_HAS_MP) != 0
        cnt = torch.where(has_mp, triples[:, 0].clamp(min=1), torch.zeros_like(triples[:, 0])).to(torch.int64)
        dummy = has_mp & (triples[:, 0] == 0)
        slots = torch.zeros(nt * T, dtype=torch.int64, device=dev)
        slots[:n] = cnt
        per_tile = slots.view(nt, T).sum(1)
        tile_bytes = (per_tile * 8 + 15) // 16 * 16
        mp_off = torch.zeros(nt + 1, dtype=torch.int64, device=dev)
        mp_off[1:] = torch.cumsum(tile_bytes, 0)
        within = torch.cumsum(slots.view(nt, T), 1) - slots.view(nt, T)
        first = (mp_off[:nt, None] // 8 + within).view(-1)[:n]                 # entry index of every read's first entry
        total = int(cnt.sum().item())
        col = torch.zeros(max(4, int(mp_off[-1].item()) // 4), dtype=torch.int32, device=dev)
        if total:
            read_of = torch.repeat_interleave(torch.arange(n, device=dev), cnt)
            k = torch.arange(total, device=dev) - torch.repeat_interleave(torch.cumsum(cnt, 0) - cnt, cnt)
            conv = triples[read_of, 1 + 3 * k].to(torch.int64)
            rp = triples[read_of, 2 + 3 * k].to(torch.int64) - 1
            fp = triples[read_of, 3 + 3 * k].to(torch.int64) - 1
            d = dummy[read_of]
            conv = torch.where(d, torch.full_like(conv, 24), conv)
            rp = torch.where(d, torch.zeros_like(rp), rp)
            fp = torch.where(d, torch.zeros_like(fp), fp)
            pos = first[read_of] + k
            col[2 * pos] = (rp | (fp << 16)).to(torch.int32)
            col[2 * pos + 1] = conv.to(torch.int32)
        rec[:n, 4] = (rec[:n, 4] & 0xFFFF) | (cnt << 16).to(torch.int32)
        tiles = db.tensors["tiles"].clone().view(torch.int64).view(-1, 4)
        tiles[:, 3] = mp_off
        max_tile_mp = int(tile_bytes.max().item()) if nt else 0
    eng.stream.synchronize()
    tens = dict(db.tensors)
    tens["rec"], tens["mp"], tens["tiles"] = rec.view(-1).view(torch.uint8), col.view(torch.uint8), tiles.view(-1).view(torch.uint8)
    b = L.Batch()
    ctypes.memmove(ctypes.byref(b), ctypes.byref(db.batch), ctypes.sizeof(L.Batch))
    b.rec, b.mp, b.tiles = tens["rec"].data_ptr(), tens["mp"].data_ptr(), tens["tiles"].data_ptr()
    b.max_tile_mp = max_tile_mp
    return DeviceBatch(b, tens)


def to_qual_planes(db: DeviceBatch, eng):
    """The same device batch (2-bit SEQ form) with its quality column as bit-planes of the dictionary code (sd_qual_dict +
    sd_qual_to_planes, sd_batch.qual_form = 1) -- what sd_count_host hands the kernel for a file with at most 16 distinct
    qualities.  Returns None when the column has more distinct values than that (the byte column stays)."""
    nt = int(db.batch.n_tiles)
    tl = db.tensors["tiles"].view(torch.int64)
    n_q = int(tl[4 * nt + 1].item()) if nt else 0
    n_seq = int(tl[4 * nt].item()) if nt else 0
    bits = ctypes.c_uint32(8)
    dict16 = (ctypes.c_uint8 * 16)()
    eng._chk(eng.lib.sd_qual_dict(eng.ctx, db.tensors["qual"].data_ptr(), n_q, ctypes.byref(bits), dict16), "sd_qual_dict")
    if bits.value == 8:
        return None
    assert int(db.batch.seq_bits) == 2
    with torch.cuda.stream(eng.stream):
        planes = torch.zeros(max(16, n_seq * bits.value // 2), dtype=torch.uint8, device=db.tensors["seq"].device)
        tiles2 = torch.empty_like(db.tensors["tiles"])
    eng.stream.synchronize()
    eng._chk(eng.lib.sd_qual_to_planes(eng.ctx, ctypes.byref(db.batch), bits.value, dict16, planes.data_ptr(), tiles2.data_ptr()),
             "sd_qual_to_planes")
    eng.synchronize()
    tens = dict(db.tensors)
    tens["qual"], tens["tiles"] = planes, tiles2
    b = L.Batch()
    ctypes.memmove(ctypes.byref(b), ctypes.byref(db.batch), ctypes.sizeof(L.Batch))
    b.qual, b.tiles = planes.data_ptr(), tiles2.data_ptr()
    b.qual_bits, b.qual_form = bits.value, 1
    for i in range(16):
        b.qual_dict[i] = dict16[i]
    b.max_tile_qual = int(db.batch.max_tile_seq) // 8 * 4 * bits.value
    return DeviceBatch(b, tens)


def to_seq2(db: DeviceBatch, eng) -> DeviceBatch:
    """The same device batch with its SEQ column in the 2-bit form (sd_seq_pack), the form sd_decode_bam hands the
    count path (SD_DECODE_SEQ2).  The other columns are shared with `db`."""
    nt = int(db.batch.n_tiles)
    n_seq = int(db.tensors["tiles"].view(torch.int64)[4 * nt].item()) if nt else 0
    with torch.cuda.stream(eng.stream):
        seq2 = torch.zeros(max(16, n_seq // 2), dtype=torch.uint8, device=db.tensors["seq"].device)
        tiles2 = torch.empty_like(db.tensors["tiles"])
    eng.stream.synchronize()
    eng._chk(eng.lib.sd_seq_pack(eng.ctx, ctypes.byref(db.batch), seq2.data_ptr(), tiles2.data_ptr()), "sd_seq_pack")
    eng.synchronize()
    tens = dict(db.tensors)
    tens["seq"], tens["tiles"] = seq2, tiles2
    b = L.Batch()
    ctypes.memmove(ctypes.byref(b), ctypes.byref(db.batch), ctypes.sizeof(L.Batch))
    b.seq, b.tiles = seq2.data_ptr(), tiles2.data_ptr()
    b.seq_bits = 2
    b.max_tile_seq = int(db.batch.max_tile_seq) // 2
    return DeviceBatch(b, tens)


def batch_to_host(db: DeviceBatch, eng=None):
    """Pinned host copy of a device batch -> (L.Batch with host pointers, keep-alive dict).  With an engine the quality
    column is stored the way sd_decode_bam stores it: as dictionary codes when it has at most 16 distinct values."""
    host = {}
    bits = 8
    dict16 = (ctypes.c_uint8 * 16)()
    for k, t in db.tensors.items():
        if t is None:
            host[k] = None
            continue
        if k == "qual" and eng is not None:
            nt = int(db.batch.n_tiles)
            n_q = int(db.tensors["tiles"].view(torch.int64)[4 * nt + 1].item()) if nt else 0
            if n_q:
                codes = torch.empty(n_q // 2, dtype=torch.uint8, device=t.device)
                cb = ctypes.c_uint32(8)
                eng._chk(eng.lib.sd_qual_pack(eng.ctx, t.data_ptr(), n_q, codes.data_ptr(), ctypes.byref(cb), dict16), "sd_qual_pack")
                bits = int(cb.value)
                if bits != 8:
                    t = codes[: n_q * bits // 8]
        h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        h.copy_(t)
        host[k] = h
    b = L.Batch()
    ctypes.memmove(ctypes.byref(b), ctypes.byref(db.batch), ctypes.sizeof(L.Batch))
    b.rec, b.seq, b.qual, b.cigar = host["rec"].data_ptr(), host["seq"].data_ptr(), host["qual"].data_ptr(), host["cigar"].data_ptr()
    b.mp = host["mp"].data_ptr() if host["mp"] is not None else None
    b.tiles = host["tiles"].data_ptr()
    b.qual_bits = bits
    for i in range(16):
        b.qual_dict[i] = dict16[i]
    return b, host


def host_batch_bytes(host: dict) -> int:
    """bytes sd_count_host copies to the device for this host batch"""
    return sum(t.numel() * t.element_size() for t in host.values() if t is not None)


class HostWire(object):
    """Pinned host copy of a batch in the wire form (sd_wire): what sd_decode_bam hands `count` and sd_count_wire sends over
    PCIe.  `wire` holds the host pointers, `host` keeps the buffers alive."""

    def __init__(self, wire: L.Wire, host: dict):
        self.wire = wire
        self.host = host

    def h2d_bytes(self, params: L.Params) -> int:
        """bytes sd_count_wire copies to the device for a call with these parameters"""
        h = self.host
        use_mp = params.mismatch_source != L.SD_MISMATCH_CIGAR
        cols = ["tiles", "pos16", "mqf", "seq", "qual", "var"]
        if params.filter_on:
            cols.append("xi16" if h.get("xi16") is not None else "xi32")
            if params.filter_max_nm > -1:
                cols.append("nm16")
        if use_mp:
            cols += ["nmp16", "mp"]
        n = sum(h[c].numel() * h[c].element_size() for c in cols if h.get(c) is not None)
        if params.filter_on and h.get("xi16") is not None:
            n += h["xi_dict"].numel() * 4
        return int(n)


def batch_to_wire(db: DeviceBatch, eng) -> HostWire:
    """Wire form of a device batch whose SEQ column is 2-bit (to_seq2) and whose qualities are bytes: packed on the device
    (sd_wire_plan / sd_wire_fill; XI dictionary with torch.unique), then copied to page-locked host memory."""
    dev = db.tensors["rec"].device
    nt = int(db.batch.n_tiles)
    tl = db.tensors["tiles"].view(torch.int64)
    n_q = int(tl[4 * nt + 1].item()) if nt else 0
    bits = ctypes.c_uint32(8)
    dict16 = (ctypes.c_uint8 * 16)()
    eng._chk(eng.lib.sd_qual_dict(eng.ctx, db.tensors["qual"].data_ptr(), n_q, ctypes.byref(bits), dict16), "sd_qual_dict")
    qb = int(bits.value)
    with torch.cuda.stream(eng.stream):
        wt = torch.zeros((nt + 1) * 32, dtype=torch.uint8, device=dev)
    eng.stream.synchronize()
    totals = (ctypes.c_uint64 * 4)()
    eng._chk(eng.lib.sd_wire_plan(eng.ctx, ctypes.byref(db.batch), wt.data_ptr(), totals), "sd_wire_plan")
    n_bases, n_var, n_mp, max_cig = (int(x) for x in totals)
    has_mp = bool(db.batch.mp)
    with torch.cuda.stream(eng.stream):
        xi_bits = db.tensors["rec"].view(torch.int32).view(-1, 5)[:, 2].contiguous()
        xdict, inv = torch.unique(xi_bits, return_inverse=True)
        small = xdict.numel() <= 65536
        dv = {
            "tiles": wt,
            "pos16": torch.zeros(nt * 32, dtype=torch.int16, device=dev),
            "mqf": torch.zeros(nt * 32, dtype=torch.int16, device=dev),
            "xi16": inv.to(torch.int16) if small else None,        # codes 0..65535 wrap into int16 bit patterns
            "xi32": None if small else torch.zeros(nt * 32, dtype=torch.int32, device=dev),
            "nm16": torch.zeros(nt * 32, dtype=torch.int16, device=dev),
            "nmp16": torch.zeros(nt * 32, dtype=torch.int16, device=dev) if has_mp else None,
            "seq": torch.zeros(n_bases // 4 + 16, dtype=torch.uint8, device=dev),
            "qual": torch.zeros(n_bases * qb // 8 + 16, dtype=torch.uint8, device=dev),
            "var": torch.zeros(n_var + 4, dtype=torch.int32, device=dev),
            "mp": torch.zeros(n_mp + 4, dtype=torch.int32, device=dev) if has_mp else None,
            "xi_dict": xdict.contiguous() if small else None,
        }
        if small and xdict.numel() > 32768:
            dv["xi16"] = (inv & 0xFFFF).to(torch.int32).to(torch.int16)
    eng.stream.synchronize()
    w = L.Wire()
    ptr = lambda t: t.data_ptr() if t is not None else None      # noqa: E731
    w.tiles, w.pos16, w.mqf, w.xi16, w.nm16, w.nmp16 = ptr(dv["tiles"]), ptr(dv["pos16"]), ptr(dv["mqf"]), ptr(dv["xi16"]), ptr(dv["nm16"]), ptr(dv["nmp16"])
    w.xi32, w.seq, w.qual, w.var, w.mp = ptr(dv["xi32"]), ptr(dv["seq"]), ptr(dv["qual"]), ptr(dv["var"]), ptr(dv["mp"])
    w.qual_bits = qb
    for i in range(16):
        w.qual_dict[i] = dict16[i]
    eng._chk(eng.lib.sd_wire_fill(eng.ctx, ctypes.byref(db.batch), ptr(dv["xi16"]), ctypes.byref(w)), "sd_wire_fill")
    w.max_tile_cig = max_cig
    host = {}
    for k, t in dv.items():
        if t is None:
            host[k] = None
            continue
        h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        h.copy_(t)
        host[k] = h
    hp = lambda k: host[k].data_ptr() if host[k] is not None else None      # noqa: E731
    w.tiles, w.pos16, w.mqf, w.xi16, w.nm16, w.nmp16 = hp("tiles"), hp("pos16"), hp("mqf"), hp("xi16"), hp("nm16"), hp("nmp16")
    w.xi32, w.seq, w.qual, w.var, w.mp = hp("xi32"), hp("seq"), hp("qual"), hp("var"), hp("mp")
    w.xi_dict = hp("xi_dict")
    w.n_xi_dict = int(host["xi_dict"].numel()) if host["xi_dict"] is not None else 0
    return HostWire(w, host)
