"""Host driver of the native count path: owns the device buffers (torch tensors) and calls the C ABI.

torch is plumbing here -- device memory, streams, torch.distributed for the multi-GPU reduce; every
kernel on the path lives in libslamdunk_b200.so (include/slamdunk_b200.h).
"""
from __future__ import annotations

import ctypes
import os
import threading
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib as L


def _np_ptr(a: np.ndarray):
    return ctypes.c_void_p(a.ctypes.data)


def _t_ptr(t: Optional[torch.Tensor]):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


class HostGenome(object):
    """FASTA decoded into the linear bit-plane space (sd_load_fasta)."""

    def __init__(self, path: str, threads: int = 0):
        lib = L.load()
        err = ctypes.create_string_buffer(512)
        out = ctypes.POINTER(L.HGenome)()
        rc = lib.sd_load_fasta(os.fsencode(path), int(threads), ctypes.byref(out), err, len(err))
        if rc != L.SD_OK:
            raise L.NativeError("sd_load_fasta(%s) failed (%d): %s" % (path, rc, err.value.decode("utf-8", "replace")))
        self._h = out
        g = out.contents
        self.n_chrom = int(g.n_chrom)
        self.names = [g.names[i].decode() for i in range(self.n_chrom)]
        self.index = {}
        for i, n in enumerate(self.names):
            self.index.setdefault(n, i)
        self.chrom_len = np.ctypeslib.as_array(g.chrom_len, shape=(self.n_chrom,)).copy()
        self.chrom_off = np.ctypeslib.as_array(g.chrom_off, shape=(self.n_chrom,)).copy()
        self.n_words = int(g.n_words)
        self.lo = np.ctypeslib.as_array(g.lo, shape=(self.n_words,))
        self.hi = np.ctypeslib.as_array(g.hi, shape=(self.n_words,))
        self.nmask = np.ctypeslib.as_array(g.nmask, shape=(self.n_words,))
        self.seconds = float(g.seconds)

    def close(self):
        if self._h is not None:
            self.lo = self.hi = self.nmask = None
            L.load().sd_hgenome_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def is_base_mask(self, strand: int) -> np.ndarray:
        """uint32 plane: bit set where the upper-cased FASTA base is T (strand 0) or A (strand 1)."""
        if strand == 0:
            return self.lo & self.hi & ~self.nmask
        return ~self.lo & ~self.hi & ~self.nmask


def wire_arrays(w: "L.Wire") -> dict:
    """numpy views of a HOST sd_wire's columns (tests, diagnostics, the multi-rank scatter)"""
    nt = int(w.n_tiles)

    def view(addr, count, dt):
        if not addr or not count:
            return np.zeros(0, dtype=dt)
        return np.frombuffer((ctypes.c_uint8 * (count * np.dtype(dt).itemsize)).from_address(addr), dtype=dt)
    tiles = view(w.tiles, (nt + 1) * 32, np.uint8).view(L.WIRE_TILE_DT)
    b0, v0, m0 = (int(tiles[0][k]) for k in ("base_off", "var_off", "mp_off")) if nt else (0, 0, 0)
    b1, v1, m1 = (int(tiles[nt][k]) for k in ("base_off", "var_off", "mp_off")) if nt else (0, 0, 0)
    qb = int(w.qual_bits)
    # the per-base, var and mp columns are addressed by the tiles' absolute offsets: the views start at the column bases
    return {"tiles": tiles, "pos16": view(w.pos16, nt * 32, np.uint16), "mqf": view(w.mqf, nt * 32, np.uint16),
            "xi16": view(w.xi16, nt * 32, np.uint16), "xi32": view(w.xi32, nt * 32, np.float32),
            "nm16": view(w.nm16, nt * 32, np.uint16), "nmp16": view(w.nmp16, nt * 32, np.uint16),
            "seq": view(w.seq, b1 // 4, np.uint8), "qual": view(w.qual, b1 * qb // 8, np.uint8),
            "var": view(w.var, v1, np.uint32), "mp": view(w.mp, m1, np.uint32),
            "xi_dict": view(w.xi_dict, int(w.n_xi_dict), np.float32), "qual_bits": qb, "qual_dict": list(w.qual_dict),
            "first": (b0, v0, m0)}


class HostBatch(object):
    """A BAM file decoded into the columnar batch (sd_decode_bam); buffers are page-locked."""

    def __init__(self, path: str, fasta_names: Sequence[str], want_mp: bool = True, tile_reads: int = 0,
                 threads: int = 0, group: bool = False, raw_qual: bool = False, mp_all: bool = False, seq2: bool = False,
                 wire: bool = False):
        lib = L.load()
        err = ctypes.create_string_buffer(512)
        out = ctypes.POINTER(L.HBatch)()
        names = (ctypes.c_char_p * max(1, len(fasta_names)))(*[n.encode() for n in fasta_names])
        rc = lib.sd_decode_bam(os.fsencode(path), names, len(fasta_names),
                               (L.SD_DECODE_MP if want_mp else 0) | (L.SD_DECODE_GROUP if group else 0) |
                               (L.SD_DECODE_RAW_QUAL if raw_qual else 0) | (L.SD_DECODE_MP_ALL if mp_all else 0) | (L.SD_DECODE_SEQ2 if seq2 else 0) |
                               (L.SD_DECODE_WIRE if wire else 0), int(tile_reads),
                               int(threads), ctypes.byref(out), err, len(err))
        if rc != L.SD_OK:
            raise L.NativeError("sd_decode_bam(%s) failed (%d): %s" % (path, rc, err.value.decode("utf-8", "replace")))
        self._h = out
        h = out.contents
        self.batch = h.batch
        self.n_reads = int(h.batch.n_reads)
        self.header_text = (h.header_text or b"").decode("utf-8", "replace")
        self.ref_names = [h.ref_names[i].decode() for i in range(h.n_ref)]
        self.ref_lengths = [int(h.ref_lengths[i]) for i in range(h.n_ref)]
        self.seconds_inflate = float(h.seconds_inflate)
        self.seconds_decode = float(h.seconds_decode)
        self.compressed_bytes = int(h.compressed_bytes)
        self.inflated_bytes = int(h.inflated_bytes)
        self.n_mapped_no_xi = int(h.n_mapped_no_xi)
        self.n_mapped_no_nm = int(h.n_mapped_no_nm)
        self.wire = h.wire if (wire and int(h.wire.n_tiles)) else None      # sd_wire with host pointers (SD_DECODE_WIRE)

    # numpy views (tests, diagnostics)
    def rec_array(self) -> np.ndarray:
        b = self.batch
        n = b.n_tiles * b.tile_reads
        buf = (ctypes.c_uint32 * (5 * n)).from_address(b.rec) if n else None
        return np.frombuffer(buf, dtype=np.uint32).reshape(n, 5) if n else np.zeros((0, 5), np.uint32)

    def tiles_array(self) -> np.ndarray:
        b = self.batch
        buf = (ctypes.c_uint64 * (4 * (b.n_tiles + 1))).from_address(b.tiles)
        return np.frombuffer(buf, dtype=np.uint64).reshape(b.n_tiles + 1, 4)

    def column(self, which: str) -> np.ndarray:
        b = self.batch
        t = self.tiles_array()[-1]
        if which == "seq":
            n, addr, dt = int(t[0]) // 4, b.seq, ctypes.c_uint32
        elif which == "qual":
            bits = int(b.qual_bits) or 8
            if bits != 8:        # dictionary codes -> phred bytes
                nq = int(t[1])
                if not nq or not b.qual:
                    return np.zeros(0, dtype=np.uint8)
                codes = np.frombuffer((ctypes.c_uint8 * (nq * bits // 8)).from_address(b.qual), dtype=np.uint8)
                idx = (codes[:, None] >> (np.arange(8 // bits, dtype=np.uint8) * bits)[None, :]) & ((1 << bits) - 1)
                return np.array(list(b.qual_dict), dtype=np.uint8)[idx.reshape(-1)]
            n, addr, dt = int(t[1]), b.qual, ctypes.c_uint8
        elif which == "cigar":
            n, addr, dt = int(t[2]) // 4, b.cigar, ctypes.c_uint32
        elif which == "mp":
            n, addr, dt = int(t[3]) // 4, b.mp, ctypes.c_uint32
        else:
            raise KeyError(which)
        if not n or not addr:
            return np.zeros(0, dtype=np.dtype(dt))
        return np.frombuffer((dt * n).from_address(addr), dtype=np.dtype(dt))

    def wire_arrays(self) -> dict:
        """numpy views of the wire form's columns (tests, diagnostics)"""
        return wire_arrays(self.wire)

    def order(self) -> np.ndarray:
        """batch slot -> record index in the file"""
        n = self.n_reads
        if not n:
            return np.zeros(0, np.uint32)
        return np.ctypeslib.as_array(self._h.contents.order, shape=(n,))

    def name_hash(self) -> np.ndarray:
        n = self.n_reads
        if not n:
            return np.zeros(0, np.uint64)
        return np.ctypeslib.as_array(self._h.contents.name_hash, shape=(n,))

    def close(self):
        if self._h is not None:
            L.load().sd_hbatch_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DeviceBatch(object):
    """Columnar batch resident in HBM (torch uint8 tensors) plus the sd_batch that points at it."""

    def __init__(self, batch: L.Batch, tensors: dict):
        self.batch = batch
        self.tensors = tensors

    @staticmethod
    def from_host(hb: L.Batch, device: torch.device, eng: Optional["CountEngine"] = None) -> "DeviceBatch":
        """Upload a host batch.  A coded quality column (hb.qual_bits 2 / 4) is expanded to bytes: on the device when an
        engine is given (sd_qual_unpack), else with numpy before the copy."""
        def up(addr, nbytes):
            if not addr or nbytes == 0:
                return torch.zeros(16, dtype=torch.uint8, device=device)
            src = np.frombuffer((ctypes.c_uint8 * nbytes).from_address(addr), dtype=np.uint8)
            return torch.from_numpy(src).to(device)
        nt = hb.n_tiles
        tiles = np.frombuffer((ctypes.c_uint64 * (4 * (nt + 1))).from_address(hb.tiles), dtype=np.uint64).reshape(nt + 1, 4)
        tot = tiles[-1]
        T = {
            "rec": up(hb.rec, nt * hb.tile_reads * 20),
            "seq": up(hb.seq, int(tot[0])),
            "qual": None,
            "cigar": up(hb.cigar, int(tot[2])),
            "mp": up(hb.mp, int(tot[3])) if hb.mp else None,
            "tiles": up(hb.tiles, (nt + 1) * 32),
        }
        bits, n_q = int(hb.qual_bits) or 8, int(tot[1])
        if bits == 8 or n_q == 0:
            T["qual"] = up(hb.qual, n_q)
        elif eng is not None:
            codes = up(hb.qual, n_q * bits // 8)
            T["qual"] = torch.empty(n_q, dtype=torch.uint8, device=device)
            dict16 = (ctypes.c_uint8 * 16)(*hb.qual_dict)
            eng._chk(eng.lib.sd_qual_unpack(eng.ctx, _t_ptr(codes), n_q, bits, dict16, _t_ptr(T["qual"])), "sd_qual_unpack")
        else:
            codes = np.frombuffer((ctypes.c_uint8 * (n_q * bits // 8)).from_address(hb.qual), dtype=np.uint8)
            per = 8 // bits
            idx = (codes[:, None] >> (np.arange(per, dtype=np.uint8) * bits)[None, :]) & ((1 << bits) - 1)
            T["qual"] = torch.from_numpy(np.array(list(hb.qual_dict), dtype=np.uint8)[idx.reshape(-1)]).to(device)
        b = L.Batch()
        ctypes.memmove(ctypes.byref(b), ctypes.byref(hb), ctypes.sizeof(L.Batch))
        b.rec, b.seq, b.qual, b.cigar = T["rec"].data_ptr(), T["seq"].data_ptr(), T["qual"].data_ptr(), T["cigar"].data_ptr()
        b.mp = T["mp"].data_ptr() if T["mp"] is not None else None
        b.tiles = T["tiles"].data_ptr()
        b.qual_bits = 8
        return DeviceBatch(b, T)

    def nbytes(self) -> int:
        return sum(t.numel() for t in self.tensors.values() if t is not None)


def make_params(min_qual=27, conversion_threshold=1, mismatch_source=L.SD_MISMATCH_CIGAR, filter_on=False,
                filter_mq=0, filter_min_identity=0.0, filter_max_nm=-1, force_slow_path=False, flags=0) -> L.Params:
    p = L.Params()
    p.min_qual = int(min_qual)
    p.conversion_threshold = int(conversion_threshold)
    p.mismatch_source = int(mismatch_source)
    p.filter_on = 1 if filter_on else 0
    p.filter_mq = int(filter_mq)
    p.filter_max_nm = int(filter_max_nm)
    p.filter_min_identity = float(filter_min_identity)
    p.force_slow_path = 1 if force_slow_path else 0
    p.flags = int(flags)
    return p


class CountEngine(object):
    """One context = one GPU.  Typical use:

        eng = CountEngine(device=0)
        eng.set_reference(lo, hi, nmask, chrom_off, chrom_len)      # numpy uint32 planes
        eng.set_snps(tc_mask, ag_mask)                              # optional
        eng.set_utrs(chrom_idx, start, stop, strand, in_bam)        # BED order
        eng.count(device_batch, params) / eng.count_host(host_batch, params)   # any number of batches
        (optionally all-reduce eng.counters / eng.pos_cov / eng.pos_tc across ranks)
        eng.finalize(); res = eng.results()
    """

    def __init__(self, device: int = 0):
        if not torch.cuda.is_available():
            raise RuntimeError("slamdunk_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.lib = L.load()
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        self.stream = torch.cuda.Stream(device=self.device)
        ctx = ctypes.c_void_p()
        L.check(self.lib.sd_create(device, ctypes.c_void_p(self.stream.cuda_stream), ctypes.byref(ctx)), None, "sd_create")
        self.ctx = ctx
        self.n_utr = 0
        self.n_utr_set = False
        self._epoch = 0
        self.n_positions = 0
        self.counters = self.pos_cov = self.pos_tc = self.stats = self.tcontent = None
        self._keep = {}

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.sd_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc, what):
        L.check(rc, self.ctx, what)

    def synchronize(self):
        self._chk(self.lib.sd_device_synchronize(self.ctx), "sd_device_synchronize")

    # -- inputs ------------------------------------------------------------------------
    def set_reference(self, lo, hi, nmask, chrom_off, chrom_len):
        """lo/hi/nmask: numpy uint32 planes (host) or torch int32 tensors already on the device."""
        def dev(x):
            if isinstance(x, torch.Tensor):
                return x
            return torch.from_numpy(np.ascontiguousarray(x).view(np.int32)).to(self.device)
        with torch.cuda.stream(self.stream):
            d_lo, d_hi, d_n = dev(lo), dev(hi), dev(nmask)
        self.stream.synchronize()
        co = np.ascontiguousarray(chrom_off, dtype=np.uint32)
        cl = np.ascontiguousarray(chrom_len, dtype=np.uint32)
        self.n_words = int(d_lo.numel())
        self._chk(self.lib.sd_set_reference(self.ctx, _t_ptr(d_lo), _t_ptr(d_hi), _t_ptr(d_n), self.n_words,
                                            _np_ptr(co), _np_ptr(cl), len(co)), "sd_set_reference")
        self.chrom_off, self.chrom_len = co, cl
        self._epoch += 1

    def set_snps(self, tc_mask, ag_mask):
        def dev(x):
            if x is None or isinstance(x, torch.Tensor):
                return x
            return torch.from_numpy(np.ascontiguousarray(x).view(np.int32)).to(self.device)
        with torch.cuda.stream(self.stream):
            d_tc, d_ag = dev(tc_mask), dev(ag_mask)
        self.stream.synchronize()
        self._chk(self.lib.sd_set_snps(self.ctx, _t_ptr(d_tc), _t_ptr(d_ag)), "sd_set_snps")

    def set_utrs(self, chrom_idx, start, stop, strand, in_bam=None):
        ci = np.ascontiguousarray(chrom_idx, dtype=np.uint32)
        st = np.ascontiguousarray(start, dtype=np.int64)
        sp = np.ascontiguousarray(stop, dtype=np.int64)
        sd = np.ascontiguousarray(strand, dtype=np.uint8)
        ib = np.ascontiguousarray(in_bam, dtype=np.uint8) if in_bam is not None else None
        n = len(ci)
        self._chk(self.lib.sd_set_utrs(self.ctx, _np_ptr(ci), _np_ptr(st), _np_ptr(sp), _np_ptr(sd),
                                       _np_ptr(ib) if ib is not None else None, n), "sd_set_utrs")
        self.n_utr = n
        self.n_utr_set = True
        self._epoch += 1
        npos = ctypes.c_uint64()
        self._chk(self.lib.sd_position_space(self.ctx, ctypes.byref(npos)), "sd_position_space")
        self.n_positions = int(npos.value)
        self.pos_off = np.zeros(n, dtype=np.uint64)
        self.pos_len = np.zeros(n, dtype=np.uint32)
        self._chk(self.lib.sd_utr_layout(self.ctx, _np_ptr(self.pos_off), _np_ptr(self.pos_len)), "sd_utr_layout")
        self.reset_outputs()

    def reset_outputs(self):
        with torch.cuda.stream(self.stream):
            self.counters = torch.empty(max(1, self.n_utr) * L.SD_NCOUNTER, dtype=torch.int64, device=self.device)
            self.pos_cov = torch.empty(max(1, self.n_positions), dtype=torch.int32, device=self.device)
            self.pos_tc = torch.empty(max(1, self.n_positions), dtype=torch.int32, device=self.device)
            self.stats = torch.empty(L.SD_NSTAT, dtype=torch.int64, device=self.device)
            self.tcontent = torch.zeros(max(1, self.n_utr), dtype=torch.int64, device=self.device)
        self._chk(self.lib.sd_bind_outputs(self.ctx, _t_ptr(self.counters), _t_ptr(self.pos_cov), _t_ptr(self.pos_tc),
                                           _t_ptr(self.stats)), "sd_bind_outputs")

    def new_output_set(self):
        """A second (third ...) set of output tensors for the current annotation: (counters, pos_cov, pos_tc, stats, tcontent).
        bind_output_set() makes one of them the target of the next counts (and zero-fills it)."""
        with torch.cuda.stream(self.stream):
            return (torch.empty(max(1, self.n_utr) * L.SD_NCOUNTER, dtype=torch.int64, device=self.device),
                    torch.empty(max(1, self.n_positions), dtype=torch.int32, device=self.device),
                    torch.empty(max(1, self.n_positions), dtype=torch.int32, device=self.device),
                    torch.empty(L.SD_NSTAT, dtype=torch.int64, device=self.device),
                    torch.zeros(max(1, self.n_utr), dtype=torch.int64, device=self.device))

    def bind_output_set(self, outs):
        self.counters, self.pos_cov, self.pos_tc, self.stats, self.tcontent = outs
        self._chk(self.lib.sd_bind_outputs(self.ctx, _t_ptr(self.counters), _t_ptr(self.pos_cov), _t_ptr(self.pos_tc),
                                           _t_ptr(self.stats)), "sd_bind_outputs")

    # -- the hot path ------------------------------------------------------------------
    def attach_spans(self, dbatch: DeviceBatch):
        """Per-tile span descriptors (sd_tile_spans) for the reference and annotation this engine holds: with them the count
        kernel stages each tile's reference window and annotation entries together with the tile.  Part of building a
        device batch; redo after set_reference / set_utrs (stale descriptors are ignored by the library, not misused)."""
        nt = int(dbatch.batch.n_tiles)
        with torch.cuda.stream(self.stream):
            spans = torch.empty(max(1, nt) * 64, dtype=torch.uint8, device=self.device)      # sizeof(sd_tile_span)
        self._chk(self.lib.sd_tile_spans(self.ctx, ctypes.byref(dbatch.batch), _t_ptr(spans)), "sd_tile_spans")
        dbatch.spans_tensor = spans      # owned by this DeviceBatch object (the tensors dict may be shared between views)
        dbatch.spans_key = (self.ctx.value, self._epoch)
        return dbatch

    def count(self, dbatch: DeviceBatch, params: L.Params, read_tc: Optional[torch.Tensor] = None, spans: bool = True):
        if spans and int(dbatch.batch.n_tiles) and self.n_utr_set and getattr(dbatch, "spans_key", None) != (self.ctx.value, self._epoch):
            self.attach_spans(dbatch)
        if not spans and dbatch.batch.spans:          # tests: the plain kernel on a batch that has descriptors
            dbatch.batch.spans = None
            dbatch.spans_key = None
        if read_tc is not None:
            self._chk(self.lib.sd_count_reads(self.ctx, ctypes.byref(dbatch.batch), ctypes.byref(params), _t_ptr(read_tc)),
                      "sd_count_reads")
        else:
            self._chk(self.lib.sd_count(self.ctx, ctypes.byref(dbatch.batch), ctypes.byref(params)), "sd_count")

    def count_host(self, hbatch: L.Batch, params: L.Params):
        self._chk(self.lib.sd_count_host(self.ctx, ctypes.byref(hbatch), ctypes.byref(params)), "sd_count_host")

    def count_wire(self, wire: L.Wire, params: L.Params):
        """sd_count_wire: a packed host batch (sd_wire) -> chunked copy, expansion on the device, count.  Synchronous."""
        self._chk(self.lib.sd_count_wire(self.ctx, ctypes.byref(wire), ctypes.byref(params)), "sd_count_wire")

    def finalize(self):
        self._chk(self.lib.sd_finalize(self.ctx), "sd_finalize")
        self._chk(self.lib.sd_tcontent(self.ctx, _t_ptr(self.tcontent)), "sd_tcontent")

    def last_kernel_ms(self) -> float:
        return float(self.lib.sd_last_kernel_ms(self.ctx))

    def last_kernel_split(self):
        """-> ([lean ms, staged ms, plain ms], [tiles lean -> staged, tiles staged -> plain]) of the most recent count"""
        ms = (ctypes.c_float * 3)()
        left = (ctypes.c_int64 * 2)()
        self._chk(self.lib.sd_last_kernel_split(self.ctx, ms, left), "sd_last_kernel_split")
        return [float(x) for x in ms], [int(x) for x in left]

    def leftover_tiles(self) -> int:
        return int(self.lib.sd_last_leftover_tiles(self.ctx))

    def launch_count(self) -> int:
        return int(self.lib.sd_launch_count(self.ctx))

    # -- read filter -------------------------------------------------------------------
    def filter(self, dbatch: DeviceBatch, params: L.Params, name_hash: Optional[torch.Tensor] = None, utrs=None):
        """filter.Filter decisions (sd_filter).  utrs: None for the default branch, or
        (chrom_idx, start, stop, name_id) numpy arrays in BED order for multimapper retention.
        -> (keep uint8 numpy [n_reads], stats int64 numpy [SD_NSTAT], state uint8 numpy [n_reads] for sd_rewrite_bam)"""
        n = int(dbatch.batch.n_reads)
        with torch.cuda.stream(self.stream):
            keep = torch.zeros(max(1, n), dtype=torch.uint8, device=self.device)
            stats = torch.zeros(L.SD_NSTAT, dtype=torch.int64, device=self.device)
            state = torch.zeros(max(1, n), dtype=torch.uint8, device=self.device)
        self.stream.synchronize()
        if utrs is None:
            args = (None, None, None, None, None, 0)
        else:
            ci = np.ascontiguousarray(utrs[0], dtype=np.uint32)
            st = np.ascontiguousarray(utrs[1], dtype=np.int64)
            sp = np.ascontiguousarray(utrs[2], dtype=np.int64)
            ni = np.ascontiguousarray(utrs[3], dtype=np.uint32)
            args = (_t_ptr(name_hash), _np_ptr(ci), _np_ptr(st), _np_ptr(sp), _np_ptr(ni), len(ci))
        self._chk(self.lib.sd_filter(self.ctx, ctypes.byref(dbatch.batch), ctypes.byref(params), args[0], args[1], args[2],
                                     args[3], args[4], args[5], _t_ptr(keep), _t_ptr(stats), _t_ptr(state)), "sd_filter")
        return keep.cpu().numpy()[:n], stats.cpu().numpy(), state.cpu().numpy()[:n]

    # -- genome-wide tracks ------------------------------------------------------------
    def track_runs(self, row: int, kind: int, coverage_cutoff: int, track_len: int):
        """Change points of one per-position track of annotation row `row` (sd_track_runs), after finalize().
        -> (pos uint32[n], val float64[n], isfloat uint8[n]) host arrays."""
        n = ctypes.c_uint64(0)
        self._chk(self.lib.sd_track_runs(self.ctx, row, kind, coverage_cutoff, track_len, None, None, None, 0, ctypes.byref(n)),
                  "sd_track_runs")
        cnt = int(n.value)
        if cnt == 0:
            return np.zeros(0, np.uint32), np.zeros(0, np.float64), np.zeros(0, np.uint8)
        with torch.cuda.stream(self.stream):
            pos = torch.empty(cnt, dtype=torch.int32, device=self.device)
            val = torch.empty(cnt, dtype=torch.float64, device=self.device)
            isf = torch.zeros(cnt, dtype=torch.uint8, device=self.device)
        self.stream.synchronize()
        self._chk(self.lib.sd_track_runs(self.ctx, row, kind, coverage_cutoff, track_len, _t_ptr(pos), _t_ptr(val), _t_ptr(isf),
                                         cnt, ctypes.byref(n)), "sd_track_runs")
        assert int(n.value) == cnt
        return pos.cpu().numpy().view(np.uint32), val.cpu().numpy(), isf.cpu().numpy()

    # -- outputs -----------------------------------------------------------------------
    def results(self):
        """-> dict of host numpy arrays: counters [n_utr, 5], tcontent [n_utr], stats [SD_NSTAT],
        pos_cov / pos_tc [n_positions]."""
        self.synchronize()
        n = self.n_utr
        return {
            "counters": self.counters.cpu().numpy().reshape(-1, L.SD_NCOUNTER)[:n],
            "tcontent": self.tcontent.cpu().numpy()[:n],
            "stats": self.stats.cpu().numpy(),
            "pos_cov": self.pos_cov.cpu().numpy()[:self.n_positions],
            "pos_tc": self.pos_tc.cpu().numpy()[:self.n_positions],
        }
