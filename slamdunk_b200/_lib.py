"""ctypes binding of libslamdunk_b200.so (include/slamdunk_b200.h).

There is no CPU fallback: if the library has not been built, loading raises.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SLAMDUNK_B200_LIB") or os.path.join(HERE, "libslamdunk_b200.so")   # env: kernel experiments

SD_OK = 0
SD_ABI_VERSION = 9   # 9 = tile-class counts in sd_batch; 8 = 64-byte sd_tile_span (lean-kernel tile description); 7 = SNP alleles in the generator (sd_synth_tables); 6 = sd_hbatch.n_mapped_no_xi/nm, wire form; include/slamdunk_b200.h: 2 = group-major plane-sliced SEQ column, sd_params.flags, sd_batch.max_span; 3 = coded qual; 4 = BAI writer; 5 = sd_read_stats
SD_F_REVERSE, SD_F_UNMAPPED, SD_F_SECONDARY, SD_F_SUPPLEMENTARY = 0x01, 0x02, 0x04, 0x08
SD_F_HAS_MP, SD_F_SKIP, SD_F_NEWNAME, SD_F_PAD = 0x10, 0x20, 0x40, 0x80
SD_REF_NONE = 0xFFFF
SD_MISMATCH_CIGAR, SD_MISMATCH_MP, SD_MISMATCH_VERIFY = 0, 1, 2
SD_PARAM_MP_ANY_BASE = 1
SD_TRACK_COVERAGE, SD_TRACK_COVERAGE_ON_BASE, SD_TRACK_CONVERSIONS, SD_TRACK_RATE = 0, 1, 2, 3
SD_NCOUNTER = 5
SD_DECODE_MP, SD_DECODE_GROUP, SD_DECODE_RAW_QUAL, SD_DECODE_MP_ALL, SD_DECODE_SEQ2 = 1, 2, 4, 8, 16
SD_NRATE = 25
SD_C_READS, SD_C_TCREADS, SD_C_MULTIMAP, SD_C_COV_ON_T, SD_C_CONV_ON_T = range(5)
SD_NSTAT = 10
(SD_S_MAPPED, SD_S_UNMAPPED, SD_S_FILTERED, SD_S_MQFILTERED, SD_S_IDFILTERED, SD_S_NMFILTERED,
 SD_S_COUNTED, SD_S_SLOWPATH, SD_S_MP_DISAGREE, SD_S_TC_TOTAL) = range(10)

c_u8p = ctypes.POINTER(ctypes.c_uint8)
c_u32p = ctypes.POINTER(ctypes.c_uint32)
c_u64p = ctypes.POINTER(ctypes.c_uint64)
c_i64p = ctypes.POINTER(ctypes.c_int64)


class ReadRec(ctypes.Structure):
    _fields_ = [("pos", ctypes.c_int32), ("ref_mapq_flag", ctypes.c_uint32), ("xi", ctypes.c_float),
                ("lseq_ncig", ctypes.c_uint32), ("nm_nmp", ctypes.c_uint32)]


class Tile(ctypes.Structure):
    _fields_ = [("seq_off", ctypes.c_uint64), ("qual_off", ctypes.c_uint64), ("cig_off", ctypes.c_uint64),
                ("mp_off", ctypes.c_uint64)]


class Batch(ctypes.Structure):
    _fields_ = [("n_reads", ctypes.c_uint64), ("n_tiles", ctypes.c_uint32), ("tile_reads", ctypes.c_uint32),
                ("rec", ctypes.c_void_p), ("seq", ctypes.c_void_p), ("qual", ctypes.c_void_p),
                ("cigar", ctypes.c_void_p), ("mp", ctypes.c_void_p), ("tiles", ctypes.c_void_p),
                ("max_lseq", ctypes.c_uint32), ("max_tile_seq", ctypes.c_uint32), ("max_tile_qual", ctypes.c_uint32),
                ("max_tile_cig", ctypes.c_uint32), ("max_tile_mp", ctypes.c_uint32), ("max_span", ctypes.c_uint32), ("qual_bits", ctypes.c_uint32),
                ("qual_dict", ctypes.c_uint8 * 16), ("seq_bits", ctypes.c_uint32), ("qual_form", ctypes.c_uint32),
                ("spans", ctypes.c_void_p), ("spans_epoch", ctypes.c_uint32), ("spans_counted", ctypes.c_uint32),
                ("spans_lean_tiles", ctypes.c_uint32), ("spans_staged_tiles", ctypes.c_uint32)]


class WireTile(ctypes.Structure):
    _fields_ = [("base_off", ctypes.c_uint64), ("var_off", ctypes.c_uint32), ("mp_off", ctypes.c_uint32), ("base_pos", ctypes.c_int32),
                ("ref", ctypes.c_uint16), ("flags", ctypes.c_uint16), ("lseq", ctypes.c_uint16), ("reserved0", ctypes.c_uint16),
                ("reserved1", ctypes.c_uint32)]


class Wire(ctypes.Structure):
    _fields_ = [("n_reads", ctypes.c_uint64), ("n_tiles", ctypes.c_uint32), ("tile_reads", ctypes.c_uint32), ("tiles", ctypes.c_void_p),
                ("pos16", ctypes.c_void_p), ("mqf", ctypes.c_void_p), ("xi16", ctypes.c_void_p), ("nm16", ctypes.c_void_p),
                ("nmp16", ctypes.c_void_p), ("xi32", ctypes.c_void_p), ("seq", ctypes.c_void_p), ("qual", ctypes.c_void_p),
                ("var", ctypes.c_void_p), ("mp", ctypes.c_void_p), ("xi_dict", ctypes.c_void_p), ("n_xi_dict", ctypes.c_uint32),
                ("qual_bits", ctypes.c_uint32), ("qual_dict", ctypes.c_uint8 * 16), ("max_lseq", ctypes.c_uint32),
                ("max_span", ctypes.c_uint32), ("max_tile_cig", ctypes.c_uint32), ("max_tile_mp", ctypes.c_uint32)]


SD_WT_POS32, SD_WT_REFS, SD_WT_SHAPE = 1, 2, 4
try:
    import numpy as _np
    WIRE_TILE_DT = _np.dtype([("base_off", "<u8"), ("var_off", "<u4"), ("mp_off", "<u4"), ("base_pos", "<i4"), ("ref", "<u2"),
                              ("flags", "<u2"), ("lseq", "<u2"), ("reserved0", "<u2"), ("reserved1", "<u4")])      # sd_wire_tile
except ImportError:      # numpy is only needed for the array views
    WIRE_TILE_DT = None
SD_DECODE_WIRE = 32


class Params(ctypes.Structure):
    _fields_ = [("min_qual", ctypes.c_int32), ("conversion_threshold", ctypes.c_int32),
                ("mismatch_source", ctypes.c_int32), ("filter_on", ctypes.c_int32), ("filter_mq", ctypes.c_int32),
                ("filter_max_nm", ctypes.c_int32), ("filter_min_identity", ctypes.c_double),
                ("force_slow_path", ctypes.c_int32), ("flags", ctypes.c_int32)]


class HBatch(ctypes.Structure):
    _fields_ = [("batch", Batch), ("name_hash", c_u64p), ("order", c_u32p), ("header_text", ctypes.c_char_p), ("n_ref", ctypes.c_uint32),
                ("ref_names", ctypes.POINTER(ctypes.c_char_p)), ("ref_lengths", c_u32p),
                ("seconds_inflate", ctypes.c_double), ("seconds_decode", ctypes.c_double),
                ("compressed_bytes", ctypes.c_uint64), ("inflated_bytes", ctypes.c_uint64),
                ("n_mapped_no_xi", ctypes.c_uint64), ("n_mapped_no_nm", ctypes.c_uint64), ("wire", Wire), ("opaque", ctypes.c_void_p)]


class HGenome(ctypes.Structure):
    _fields_ = [("n_chrom", ctypes.c_uint32), ("names", ctypes.POINTER(ctypes.c_char_p)), ("chrom_len", c_u32p),
                ("chrom_off", c_u32p), ("n_words", ctypes.c_uint64), ("lo", c_u32p), ("hi", c_u32p), ("nmask", c_u32p),
                ("seconds", ctypes.c_double), ("opaque", ctypes.c_void_p)]


class SynthParams(ctypes.Structure):
    _fields_ = [("seed", ctypes.c_uint64), ("first_read", ctypes.c_uint64), ("n_reads", ctypes.c_uint64),
                ("read_len", ctypes.c_uint32), ("tile_reads", ctypes.c_uint32), ("p_conv", ctypes.c_float),
                ("p_err", ctypes.c_float), ("frac_softclip", ctypes.c_float), ("frac_indel", ctypes.c_float),
                ("frac_multimap", ctypes.c_float), ("frac_antisense", ctypes.c_float), ("frac_labelled", ctypes.c_float),
                ("utr_first", ctypes.c_uint32), ("utr_count", ctypes.c_uint32), ("order", ctypes.c_void_p)]


class SynthTables(ctypes.Structure):
    _fields_ = [("lo", ctypes.c_void_p), ("hi", ctypes.c_void_p), ("nmask", ctypes.c_void_p),
                ("chrom_off", ctypes.c_void_p), ("chrom_len", ctypes.c_void_p), ("n_chrom", ctypes.c_uint32),
                ("n_utr", ctypes.c_uint32), ("utr_chrom", ctypes.c_void_p), ("utr_start", ctypes.c_void_p),
                ("utr_end", ctypes.c_void_p), ("utr_strand", ctypes.c_void_p), ("utr_cum_weight", ctypes.c_void_p),
                ("utr_label_frac", ctypes.c_void_p), ("snp_tc", ctypes.c_void_p), ("snp_ag", ctypes.c_void_p),
                ("snp_other", ctypes.c_void_p), ("p_snp_allele", ctypes.c_float)]


class StatsArgs(ctypes.Structure):
    _fields_ = [("min_qual", ctypes.c_int32), ("conversion_threshold", ctypes.c_int32), ("strict", ctypes.c_int32),
                ("max_read_length", ctypes.c_uint32), ("utr_norm", ctypes.c_uint32), ("d_order", ctypes.c_void_p),
                ("d_rates", ctypes.c_void_p), ("d_readpos", ctypes.c_void_p), ("d_utr_rates", ctypes.c_void_p),
                ("d_utr_snp", ctypes.c_void_p), ("d_utrpos", ctypes.c_void_p), ("d_context", ctypes.c_void_p), ("d_pairs", ctypes.c_void_p),
                ("pair_capacity", ctypes.c_uint64), ("d_n_pairs", ctypes.c_void_p)]


SD_SYNTH_MAX_TRIPLES = 40

# name -> (restype, argtypes); every symbol declared in include/slamdunk_b200.h
_VP = ctypes.c_void_p
PROTOTYPES = {
    "sd_abi_version": (ctypes.c_int, []),
    "sd_create": (ctypes.c_int, [ctypes.c_int, _VP, ctypes.POINTER(_VP)]),
    "sd_destroy": (None, [_VP]),
    "sd_last_error": (ctypes.c_char_p, [_VP]),
    "sd_device_synchronize": (ctypes.c_int, [_VP]),
    "sd_set_reference": (ctypes.c_int, [_VP, _VP, _VP, _VP, ctypes.c_uint64, _VP, _VP, ctypes.c_uint32]),
    "sd_set_snps": (ctypes.c_int, [_VP, _VP, _VP]),
    "sd_set_utrs": (ctypes.c_int, [_VP, _VP, _VP, _VP, _VP, _VP, ctypes.c_uint32]),
    "sd_position_space": (ctypes.c_int, [_VP, c_u64p]),
    "sd_utr_layout": (ctypes.c_int, [_VP, _VP, _VP]),
    "sd_bind_outputs": (ctypes.c_int, [_VP, _VP, _VP, _VP, _VP]),
    "sd_count": (ctypes.c_int, [_VP, ctypes.POINTER(Batch), ctypes.POINTER(Params)]),
    "sd_tile_spans": (ctypes.c_int, [_VP, ctypes.POINTER(Batch), _VP]),
    "sd_count_wire": (ctypes.c_int, [_VP, ctypes.POINTER(Wire), ctypes.POINTER(Params)]),
    "sd_wire_plan": (ctypes.c_int, [_VP, ctypes.POINTER(Batch), _VP, ctypes.POINTER(ctypes.c_uint64)]),
    "sd_wire_fill": (ctypes.c_int, [_VP, ctypes.POINTER(Batch), _VP, ctypes.POINTER(Wire)]),
    "sd_count_reads": (ctypes.c_int, [_VP, ctypes.POINTER(Batch), ctypes.POINTER(Params), _VP]),
    "sd_count_host": (ctypes.c_int, [_VP, ctypes.POINTER(Batch), ctypes.POINTER(Params)]),
    "sd_finalize": (ctypes.c_int, [_VP]),
    "sd_tcontent": (ctypes.c_int, [_VP, _VP]),
    "sd_launch_count": (ctypes.c_int64, [_VP]),
    "sd_last_leftover_tiles": (ctypes.c_int64, [_VP]),
    "sd_last_kernel_ms": (ctypes.c_float, [_VP]),
    "sd_last_kernel_split": (ctypes.c_int, [_VP, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int64)]),
    "sd_filter": (ctypes.c_int, [_VP, ctypes.POINTER(Batch), ctypes.POINTER(Params), _VP, _VP, _VP, _VP, _VP,
                                 ctypes.c_uint32, _VP, _VP, _VP]),
    "sd_rewrite_bam": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, _VP, _VP, ctypes.c_uint64, ctypes.c_int,
                                      ctypes.c_int, ctypes.c_int, c_u64p, ctypes.c_char_p, ctypes.c_int]),
    "sd_qual_unpack": (ctypes.c_int, [_VP, _VP, ctypes.c_uint64, ctypes.c_uint32, _VP, _VP]),
    "sd_qual_pack": (ctypes.c_int, [_VP, _VP, ctypes.c_uint64, _VP, ctypes.POINTER(ctypes.c_uint32), _VP]),
    "sd_qual_dict": (ctypes.c_int, [_VP, _VP, ctypes.c_uint64, ctypes.POINTER(ctypes.c_uint32), _VP]),
    "sd_qual_to_planes": (ctypes.c_int, [_VP, ctypes.POINTER(Batch), ctypes.c_uint32, _VP, _VP, _VP]),
    "sd_seq_pack": (ctypes.c_int, [_VP, ctypes.POINTER(Batch), _VP, _VP]),
    "sd_rewrite_bam_indexed": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, _VP, _VP, ctypes.c_uint64,
                                              ctypes.c_int, ctypes.c_int, ctypes.c_int, c_u64p, ctypes.c_char_p, ctypes.c_int]),
    "sd_decode_bam": (ctypes.c_int, [ctypes.c_char_p, ctypes.POINTER(ctypes.c_char_p), ctypes.c_uint32, ctypes.c_int,
                                     ctypes.c_uint32, ctypes.c_int, ctypes.POINTER(ctypes.POINTER(HBatch)),
                                     ctypes.c_char_p, ctypes.c_int]),
    "sd_hbatch_free": (None, [ctypes.POINTER(HBatch)]),
    "sd_load_fasta": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_int, ctypes.POINTER(ctypes.POINTER(HGenome)),
                                     ctypes.c_char_p, ctypes.c_int]),
    "sd_hgenome_free": (None, [ctypes.POINTER(HGenome)]),
    "sd_inflate_raw": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_uint64, _VP, ctypes.c_uint64]),
    "sd_crc32": (ctypes.c_uint32, [ctypes.c_char_p, ctypes.c_uint64]),
    "sd_deflate_raw": (ctypes.c_uint64, [ctypes.c_char_p, ctypes.c_uint64, _VP, ctypes.c_uint64, ctypes.c_int]),
    "sd_format_repr": (ctypes.c_int, [ctypes.c_double, ctypes.c_char_p]),
    "sd_format_tcount_rows": (ctypes.c_int, [ctypes.c_uint32, ctypes.POINTER(ctypes.c_char_p), _VP, _VP, ctypes.POINTER(ctypes.c_char_p), _VP,
                                             _VP, _VP, ctypes.c_int64, ctypes.POINTER(ctypes.c_void_p), c_u64p]),
    "sd_free_text": (None, [ctypes.c_void_p]),
    "sd_write_bedgraphs": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_uint32, ctypes.POINTER(ctypes.c_char_p), _VP, _VP,
                                          _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, ctypes.c_uint64, ctypes.c_int]),
    "sd_track_runs": (ctypes.c_int, [_VP, ctypes.c_uint32, ctypes.c_int, ctypes.c_int, ctypes.c_uint32, _VP, _VP, _VP, ctypes.c_uint64, c_u64p]),
    "sd_write_track": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_uint64, _VP, _VP, _VP]),
    "sd_read_stats": (ctypes.c_int, [_VP, ctypes.POINTER(Batch), ctypes.POINTER(StatsArgs)]),
    "sd_synth_positions": (ctypes.c_int, [_VP, ctypes.POINTER(SynthParams), ctypes.POINTER(SynthTables), _VP, _VP]),
    "sd_synth_plan": (ctypes.c_int, [_VP, ctypes.POINTER(SynthParams), ctypes.POINTER(SynthTables), _VP, _VP, c_u64p]),
    "sd_synth_fill": (ctypes.c_int, [_VP, ctypes.POINTER(SynthParams), ctypes.POINTER(SynthTables), _VP,
                                     ctypes.POINTER(Batch), _VP]),
    "sd_synth_genome": (ctypes.c_int, [_VP, ctypes.c_uint64, _VP, _VP, _VP, ctypes.c_uint64, _VP, _VP, ctypes.c_uint32,
                                       ctypes.c_float]),
}

_lib = None


def load():
    """Load the native library or raise -- the product path has no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libslamdunk_b200.so is not built (run `python -m slamdunk_b200.build`); "
                           "slamdunk_b200 has no CPU fallback")
    L = ctypes.CDLL(LIB_PATH)
    missing = []
    for name, (res, args) in PROTOTYPES.items():
        try:
            fn = getattr(L, name)
        except AttributeError:
            missing.append(name)
            continue
        fn.restype = res
        fn.argtypes = args
    if missing:
        raise RuntimeError("libslamdunk_b200.so is stale, it lacks %s; rebuild with `python -m slamdunk_b200.build`"
                           % ", ".join(missing))
    assert ctypes.sizeof(ReadRec) == 20 and ctypes.sizeof(Tile) == 32 and ctypes.sizeof(Batch) == 144
    if L.sd_abi_version() != SD_ABI_VERSION:
        raise RuntimeError("libslamdunk_b200.so has ABI version %d, these bindings expect %d; rebuild with `python -m slamdunk_b200.build`"
                           % (L.sd_abi_version(), SD_ABI_VERSION))
    _lib = L
    return L


class NativeError(RuntimeError):
    pass


def check(rc, ctx=None, what=""):
    if rc != SD_OK:
        msg = ""
        if ctx:
            msg = (load().sd_last_error(ctx) or b"").decode("utf-8", "replace")
        raise NativeError("%s failed with code %d: %s" % (what or "native call", rc, msg))
