/*
 * slamdunk_b200.h -- C ABI of libslamdunk_b200.so, the B200-native `slamdunk count` hot path.
 *
 * The reference (t-neumann/slamdunk v0.4.3) is pure Python and has NO FFI / plugin boundary
 * for this path; its boundary is the Python function
 *     tcounter.computeTconversions(ref, bed, snpsFile, bam, maxReadLength, minQual, outputCSV,
 *                                  outputBedgraphPlus, outputBedgraphMinus, conversionThreshold, log)
 *     (slamdunk/dunks/tcounter.py:125) called from slamdunk.runCount (slamdunk/slamdunk.py:202),
 * and filter.Filter(inputBAM, outputBAM, log, bed, MQ, minIdentity, NM, ...) (slamdunk/dunks/filter.py:213).
 * slamdunk_b200/dunks/tcounter.py and filter.py keep those signatures and call the entry points
 * below through ctypes (INTEGRATION.md shows the binding a reference maintainer would add).
 * Each entry point names the reference code whose work it takes over.
 *
 * Conventions
 *   - every function returns SD_OK (0) or a negative SD_ERR_* code; sd_last_error() gives text;
 *   - plain pointers and sizes only; no torch / C++ types cross the boundary;
 *   - "device pointer" arguments are owned by the caller (a torch tensor's data_ptr()) and must
 *     stay alive until the stream work that uses them has completed; the library owns only the
 *     tables it derives (UTR index, reference planes) and frees them in sd_destroy;
 *   - all device work is enqueued on the stream given to sd_create; functions documented as
 *     "synchronous" wait for it, the others do not;
 *   - a context is bound to one device and must be used from one host thread at a time;
 *   - only integers cross the boundary: the three float divisions of a tcount row
 *     (tcounter.py:252,297,303) are done by the Python host so the text is byte-identical.
 */
#ifndef SLAMDUNK_B200_H
#define SLAMDUNK_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SD_OK 0
#define SD_ERR_CUDA (-1)
#define SD_ERR_ARG (-2)
#define SD_ERR_STATE (-3)
#define SD_ERR_IO (-4)
#define SD_ERR_FORMAT (-5)
#define SD_ERR_LIMIT (-6)
#define SD_ERR_NOMEM (-7)

#define SD_ABI_VERSION 9

typedef struct sd_ctx sd_ctx;

/* ------------------------------------------------------------------ columnar read batch
 *
 * One record per BAM alignment, 20 bytes, array-of-structs so that one tile's records are a
 * single contiguous bulk copy.  Replaces the pysam AlignedSegment attributes the reference
 * reads (slamseq/SlamSeqFile.py:364-402, dunks/filter.py:233-256).
 */
#define SD_F_REVERSE 0x01u       /* BAM flag 0x10 */
#define SD_F_UNMAPPED 0x02u      /* BAM flag 0x4, or no CIGAR (pysam reference_end is None) */
#define SD_F_SECONDARY 0x04u     /* BAM flag 0x100 */
#define SD_F_SUPPLEMENTARY 0x08u /* BAM flag 0x800 */
#define SD_F_HAS_MP 0x10u        /* record carries NextGenMap's MP:Z tag (SlamSeqFile.py:318) */
#define SD_F_SKIP 0x20u          /* not to be counted: padding record, chromosome not in the FASTA */
#define SD_F_NEWNAME 0x40u       /* query name differs from the previous record in file order */
#define SD_F_PAD 0x80u           /* tile padding, not a record */

#define SD_REF_NONE 0xFFFFu      /* ref field: chromosome absent from the FASTA */

typedef struct sd_read_rec {
    int32_t pos;            /* 0-based leftmost reference position on its chromosome (BAM pos) */
    uint32_t ref_mapq_flag; /* bits 0-15 chromosome index (FASTA order), 16-23 MAPQ, 24-31 SD_F_* */
    float xi;               /* NextGenMap XI:f alignment identity (filter.py:246); NaN if absent */
    uint32_t lseq_ncig;     /* bits 0-15 l_seq, 16-31 number of CIGAR ops */
    uint32_t nm_nmp;        /* bits 0-15 NM tag (saturated), 16-31 number of MP entries stored */
} sd_read_rec;

/* Byte offsets of one tile's variable-length rows inside the seq / qual / cigar / mp columns.
 * tiles[t+1] - tiles[t] is the tile's size; every offset is a multiple of 16. */
typedef struct sd_tile {
    uint64_t seq_off;
    uint64_t qual_off;
    uint64_t cig_off;
    uint64_t mp_off;
} sd_tile;

/*
 * Column layouts (rows of one tile are consecutive, in record order, tiles padded to 16 bytes):
 *   seq   "plane-sliced 4-bit", group-major per tile: a group is 32 bases of one read as four 32-bit words, word p
 *         (p = 0..3) being bit-plane p of the BAM nibble code (bit0 A, bit1 C, bit2 G, bit3 T; N = all four,
 *         '=' = none), bit i of each word = base 32g+i, pad bits zero.  A tile's block holds, for g = 0 ..
 *         G-1 (G = ceil(longest l_seq of the tile / 32)), the g-th group of each of the tile_reads row slots:
 *         group g of row r is at byte 16 * (g * tile_reads + r) of the block (shorter reads and pad rows: zeros).
 *         One warp reads a group of all 32 rows as one conflict-free 512-byte access.
 *   qual  l_seq bytes per read, BAM phred values as stored.
 *   cigar n_cigar 32-bit words per read, BAM encoding (len << 4 | op).
 *   mp    per stored MP entry one 32-bit word: bits 0-15 readPos-1, bits 16-31 refPos-1
 *         (0-based, already decremented; entries with conv == 16 on forward reads and
 *         conv == 2 on reverse reads only -- the only ones isTCMismatch can accept,
 *         SlamSeqFile.py:137-141).  May be NULL when mismatches come from the CIGAR walk and
 *         no cross-check is wanted.
 */
typedef struct sd_batch {
    uint64_t n_reads;       /* real records (<= n_tiles * tile_reads) */
    uint32_t n_tiles;
    uint32_t tile_reads;    /* records per tile: 32 (one tile per warp); rec[] is padded with SD_F_PAD */
    const sd_read_rec *rec; /* [n_tiles * tile_reads] */
    const uint32_t *seq;
    const uint8_t *qual;
    const uint32_t *cigar;
    const uint32_t *mp;     /* may be NULL */
    const sd_tile *tiles;   /* [n_tiles + 1] */
    uint32_t max_lseq;      /* largest l_seq in the batch */
    uint32_t max_tile_seq;  /* largest per-tile byte sizes (size the shared-memory stages) */
    uint32_t max_tile_qual;
    uint32_t max_tile_cig;
    uint32_t max_tile_mp;
    uint32_t max_span;      /* largest reference span (M/D/N/=/X) among alignments with span <= l_seq + 64 (spliced
                               alignments are left out: they take the per-base path anyway); 0 = not known */
    uint32_t qual_bits;     /* bits per base quality in `qual`: 0 or 8 = one byte each (the only form the device entry
                               points take); 2 or 4 = dictionary codes, four or two per byte starting at the low bits, the
                               quality at column offset o being code (qual[o * bits / 8] >> (o * bits % 8)) & (2^bits - 1).
                               HOST batches only: sd_decode_bam codes a file with at most 4 / 16 distinct qualities this
                               way (Illumina's binned qualities), sd_count_host copies the codes and expands them on the
                               device -- a quarter of the column's PCIe bytes.  Tile qual offsets are always in bases. */
    uint8_t qual_dict[16];  /* code -> phred value */
    uint32_t seq_bits;      /* form of the `seq` column: 0 or 4 = the four bit-planes described above (16 bytes per group);
                               2 = two words per group, bit i of word 0 / word 1 = low / high bit of base 32g+i coded A=0 C=1
                               G=2 T=3, group g of row r at byte 8 * (g * tile_reads + r) of the tile block.  A base that is
                               not exactly one of A, C, G, T (N, IUPAC codes, '=') is coded A, and so are the pad bits: the
                               count path only ever asks "is this base C" / "is it G" (SlamSeqFile.py:137-141), for which the
                               2-bit form is lossless -- half the SEQ bytes in HBM and on PCIe.  Tile seq offsets are byte
                               offsets into the column as stored.  sd_count* / sd_count_host take both forms;
                               sd_read_stats (exact base counts) needs the 4-plane form. */
    uint32_t qual_form;     /* 0: `qual` holds one row of l_seq qualities per read (bytes, or codes on a host batch, see qual_bits);
                               1: DEVICE batches, group-major bit-planes of the dictionary code, like the 2-bit SEQ column:
                               qual_bits (2 or 4) words per 32-base group, word b = bit b of the codes of bases 32g .. 32g+31,
                               group g of row r at byte 4 * qual_bits * (g * tile_reads + r) of the tile block; pad bits zero;
                               qual_dict ascends (sd_decode_bam / sd_qual_dict), so `quality >= minQual` is a bit-sliced
                               `code >= c` on 32 bases at once and the block travels with the tile instead of being fetched
                               byte by byte.  Tile qual offsets are then byte offsets into the column as stored, max_tile_qual
                               the largest block.  Made by sd_qual_to_planes; needs seq_bits == 2; sd_count* only. */
    const void *spans;      /* optional, DEVICE batches: one sd_tile_span per tile, made by sd_tile_spans for the reference and
                               annotation the context holds; NULL = none.  With them sd_count stages each tile's reference
                               window and the annotation entries its reads can overlap together with the tile's columns
                               (two more bulk copies per tile) instead of one lookup chain per read. */
    uint32_t spans_epoch;   /* set by sd_tile_spans; descriptors made before a later sd_set_reference / sd_set_utrs are ignored */
    uint32_t spans_counted; /* 1: the two counts below are valid (sd_tile_spans sets them; 0 = not known, every launch is made) */
    uint32_t spans_lean_tiles;   /* tiles the lean kernel takes */
    uint32_t spans_staged_tiles; /* tiles whose reference window and annotation entries fit a stage (the lean ones included) */
} sd_batch;

/* What the staged count kernels copy for one tile besides its columns, as [begin, end) byte offsets into two
 * library-internal tables: the strand-merged reference planes (16 bytes per 32 linear positions) over the stretch of the
 * genome the tile's countable reads cover -- mapped, on a FASTA chromosome, position inside it -- and the run of
 * annotation entries (both strands, sorted by start) that can overlap that stretch.  A tile with reads all over the genome
 * (unsorted input) is described by an impossible size; sd_count then counts it with per-read lookups.
 * The second half describes the tile for the LEAN kernel, which only takes tiles whose countable reads are all alike:
 * one match block (M / = / X) of the same l_seq bases, on one chromosome, fully inside it, with at most 16 annotation
 * entries in reach.  `lean` then holds the common l_seq (0 = not such a tile), g0 the linear-genome offset of the tile's
 * chromosome (linear position = g0 + pos), ent = byte offset of the tile's first entry << 16 | bytes of its entries, win the
 * same for the reference window (informative). */
typedef struct sd_tile_span {
    uint64_t win_begin, win_end;
    uint64_t ent_begin, ent_end;
    uint64_t win, ent;
    uint32_t lean, g0;
    uint64_t reserved;
} sd_tile_span;

/* Where the per-read mismatch list comes from. */
#define SD_MISMATCH_CIGAR 0 /* walk the CIGAR, compare SEQ with the reference planes (north star) */
#define SD_MISMATCH_MP 1    /* take NextGenMap's MP entries as the reference code does */
#define SD_MISMATCH_VERIFY 2 /* CIGAR walk, and count reads whose MP entries disagree with it */

typedef struct sd_params {
    int32_t min_qual;             /* minQual (tcounter.py:125; `>=` test SlamSeqFile.py:330) */
    int32_t conversion_threshold; /* conversionThreshold (SlamSeqFile.py:397) */
    int32_t mismatch_source;      /* SD_MISMATCH_* */
    int32_t filter_on;            /* apply filter.py:241-251 predicates before counting */
    int32_t filter_mq;            /* MQ */
    int32_t filter_max_nm;        /* NM, -1 = off */
    double filter_min_identity;   /* minIdentity */
    int32_t force_slow_path;      /* testing: run every read through the per-base path */
    int32_t flags;                /* SD_PARAM_* */
} sd_params;
/* SD_MISMATCH_MP only: keep every accepted MP conversion, also where the FASTA base is not T / A.  For the genome-wide
 * tracks: genomewideConversionRates never looks at the FASTA for its conversion arrays (tcounter.py:384-389).  The
 * per-UTR conversionsOnTs counter of `count` does (tcounter.py:279) and must not be read from a run with this flag. */
#define SD_PARAM_MP_ANY_BASE 1

/* Per-UTR integer counters, row-major [n_utr][SD_NCOUNTER], BED order. */
#define SD_NCOUNTER 5
#define SD_C_READS 0     /* readCount        tcounter.py:216-227 */
#define SD_C_TCREADS 1   /* tcReadCount */
#define SD_C_MULTIMAP 2  /* multimapCount */
#define SD_C_COV_ON_T 3  /* coverageOnTs     tcounter.py:281 */
#define SD_C_CONV_ON_T 4 /* conversionsOnTs  tcounter.py:282 */

/* Run-level counters, int64[SD_NSTAT]. */
#define SD_NSTAT 10
#define SD_S_MAPPED 0       /* filter.py:235-239 (primary, mapped) */
#define SD_S_UNMAPPED 1
#define SD_S_FILTERED 2     /* filteredReads */
#define SD_S_MQFILTERED 3
#define SD_S_IDFILTERED 4
#define SD_S_NMFILTERED 5
#define SD_S_COUNTED 6      /* records that entered the T>C scan */
#define SD_S_SLOWPATH 7     /* records handled by the per-base path */
#define SD_S_MP_DISAGREE 8  /* SD_MISMATCH_VERIFY: records whose MP list != CIGAR walk */
#define SD_S_TC_TOTAL 9     /* sum of per-read tcCount over counted reads (diagnostic) */

/* ------------------------------------------------------------------ context */
int sd_abi_version(void);
/* device: CUDA ordinal.  stream: cudaStream_t (NULL = a private non-blocking stream). */
int sd_create(int device, void *stream, sd_ctx **out);
void sd_destroy(sd_ctx *ctx);
const char *sd_last_error(const sd_ctx *ctx);
int sd_device_synchronize(sd_ctx *ctx);

/* ------------------------------------------------------------------ reference genome
 * Replaces pysam.FastaFile.fetch(...).upper() (tcounter.py:181, SlamSeqFile.py:436).
 * The genome is one linear bit space: chromosome c occupies bits [chrom_off[c], chrom_off[c] +
 * chrom_len[c]); chrom_off[c] is a multiple of 32 and consecutive chromosomes are separated by
 * at least 32 pad positions.  Three bit-planes of n_words uint32 (bit j of word w = position
 * 32w+j): lo/hi = the two bits of the base code A=0 C=1 G=2 T=3, nmask = 1 where the upper-cased
 * FASTA character is not one of ACGT (and on every pad position).  Device pointers; the planes
 * are consumed (a derived table is built) before the call returns control to the stream.
 */
int sd_set_reference(sd_ctx *ctx, const uint32_t *d_lo, const uint32_t *d_hi, const uint32_t *d_nmask,
                     uint64_t n_words, const uint32_t *h_chrom_off, const uint32_t *h_chrom_len,
                     uint32_t n_chrom);
/* SNP masks in the same bit space (replaces SNPDictionary.isTCSnp / isAGSnp, utils/SNPtools.py:53-60;
 * the chrom+pos string-key collisions are resolved by the host when it builds the masks).
 * Device pointers, NULL = no SNPs.  Call after sd_set_reference. */
int sd_set_snps(sd_ctx *ctx, const uint32_t *d_tc_mask, const uint32_t *d_ag_mask);

/* ------------------------------------------------------------------ annotation
 * Replaces BedIterator + the per-UTR indexed fetch (tcounter.py:165, SlamSeqFile.py:441).
 * Host arrays in BED order.  chrom[i] = FASTA chromosome index or SD_REF_NONE; start/stop as in
 * the BED file (stop is clipped to the chromosome length internally, SlamSeqFile.py:441);
 * strand 0 '+', 1 '-' (2: no strand -- the row then takes reads of both strands as the reference's iterator does for
 * strand "."; sd_read_stats only, `count` refuses such rows).  in_bam[i] = 0 suppresses reads for rows whose chromosome is not in the
 * BAM header (SlamSeqFile.py:440-445).
 */
int sd_set_utrs(sd_ctx *ctx, const uint32_t *h_chrom, const int64_t *h_start, const int64_t *h_stop,
                const uint8_t *h_strand, const uint8_t *h_in_bam, uint32_t n_utr);
/* Size (in int32 elements) of each per-position array; 0 before sd_set_utrs. */
int sd_position_space(const sd_ctx *ctx, uint64_t *n_positions);
/* For every BED row: offset of its first base in the per-position arrays (UINT64_MAX if the row
 * has no device entry), and the number of in-chromosome bases it spans. */
int sd_utr_layout(const sd_ctx *ctx, uint64_t *h_pos_off, uint32_t *h_pos_len);

/* ------------------------------------------------------------------ outputs
 * Caller-owned device buffers (so they can be handed to NCCL as torch tensors):
 *   counters  int64 [n_utr * SD_NCOUNTER]
 *   pos_cov   int32 [n_positions]  coverage (difference form until sd_finalize)
 *   pos_tc    int32 [n_positions]  T>C conversions per position
 *   stats     int64 [SD_NSTAT]
 * sd_bind_outputs zero-fills them on the stream.
 */
int sd_bind_outputs(sd_ctx *ctx, int64_t *d_counters, int32_t *d_pos_cov, int32_t *d_pos_tc,
                    int64_t *d_stats);

/* ------------------------------------------------------------------ the hot path
 * sd_count: filter predicates (filter.py:241-251) + per-read conversion scan
 * (SlamSeqFile.py:313-402) + interval join and per-UTR reduction (tcounter.py:207-248)
 * + per-position scatter, for one device-resident batch.  Asynchronous.
 */
int sd_count(sd_ctx *ctx, const sd_batch *d_batch, const sd_params *params);
/* Fill d_spans (device, n_tiles * sizeof(sd_tile_span)) for a device-resident batch and attach it (d_batch->spans,
 * d_batch->spans_epoch, and how many tiles each kind of kernel will take: sd_count skips a launch that would take next to
 * none -- unsorted input goes straight to the per-read kernel).  Call after sd_set_reference and sd_set_utrs; part of building the batch, like its tile table.
 * A tile whose CIGAR block is empty (tiles[t + 1].cig_off == tiles[t].cig_off) is read as: every record with one CIGAR op
 * is a single M block of l_seq bases -- plain-match tiles need not carry the column.  Synchronous (the two counts are read back). */
int sd_tile_spans(sd_ctx *ctx, sd_batch *d_batch, void *d_spans);
/* sd_count that also writes every record's tcCount (after the conversionThreshold rule, saturated at
 * 255; 0 for records that were not counted) to d_read_tc[n_tiles * tile_reads] -- the per-read
 * quantity SlamSeqBamIterator yields (SlamSeqFile.py:386,397); tests and read separation use it. */
int sd_count_reads(sd_ctx *ctx, const sd_batch *d_batch, const sd_params *params, uint8_t *d_read_tc);
/* Same, from HOST buffers: copies the batch to the device in chunks (pinned staging, copy
 * stream overlapped with the kernel) and waits for completion.  Synchronous. */
int sd_count_host(sd_ctx *ctx, const sd_batch *h_batch, const sd_params *params);
/* Turn pos_cov from difference form into coverage (segment-local prefix sums).  Call once after
 * the last sd_count (and after any cross-GPU reduction of the outputs).  Asynchronous. */
int sd_finalize(sd_ctx *ctx);
/* Tcontent per BED row (tcounter.py:177-189): count of T ('+') or A ('-') in the upper-cased
 * FASTA slice.  d_tcontent: device int64[n_utr].  Asynchronous. */
int sd_tcontent(sd_ctx *ctx, int64_t *d_tcontent);
/* Kernel launches issued by this context since creation (bench.py's gpu_launches). */
int64_t sd_launch_count(const sd_ctx *ctx);
/* Tiles of the most recent sd_count that the staged kernel handed on to the plain one (diagnostic; waits for the stream). */
int64_t sd_last_leftover_tiles(sd_ctx *ctx);
/* The same split over the launches a batch with span descriptors takes -- ms[0] lean kernel (tiles whose reads are all one match
 * block of the same length), ms[1] staged kernel (the other tiles), ms[2] plain kernel (tiles spread too widely to stage) -- and the
 * number of tiles handed from the first to the second (left[0]) and from the second to the third launch (left[1]).  Diagnostic;
 * waits for the stream. */
int sd_last_kernel_split(sd_ctx *ctx, float ms[3], int64_t left[2]);
/* Duration of the most recent sd_count kernel on its stream (CUDA events), ms; <0 if none. */
float sd_last_kernel_ms(sd_ctx *ctx);

/* ------------------------------------------------------------------ packed host batch ("wire" form)
 * What travels over PCIe for `count`: the same records as an sd_batch, with every field in its own column and nothing
 * padded or repeated, so that sd_count_wire copies only what the call needs (no XI when the filter is off, no NM unless
 * the NM filter is on, no MP entries for a plain CIGAR walk) and expands the rest on the device, next to the kernel that
 * consumes it.  At 100 bp: 2-bit SEQ codes 25 B + 2-bit quality codes 25 B + pos / mapq+flags / XI code 6 B per read, plus
 * a 32-byte header per 32-read tile; plain-match tiles carry neither l_seq nor CIGAR (58 B per read against the 82 B of
 * the sd_batch host form).  Made by sd_decode_bam (SD_DECODE_WIRE) or, from a device batch, by sd_wire_plan + sd_wire_fill.
 *
 * Per-read columns [n_tiles * 32], slot order as in sd_batch.rec:
 *   pos16   pos - tile.base_pos (tiles without SD_WT_POS32)
 *   mqf     MAPQ | SD_F_* << 8
 *   xi16    index into xi_dict (files with at most 65536 distinct XI values), else xi32 holds the floats
 *   nm16    NM tag, saturated
 *   nmp16   number of MP entries of the record (only with mp)
 * Per-base columns, rows back to back at base granularity, a tile's first base at tiles[t].base_off (a multiple of 64):
 *   seq     2-bit base codes A=0 C=1 G=2 T=3 (anything else A, see sd_batch.seq_bits), base i of the column at bits
 *           2 (i % 4) of byte i / 4
 *   qual    qual_bits-bit dictionary codes (2 or 4, low bits first), or bytes (8)
 * var: per tile one block of 32-bit words at tiles[t].var_off (a multiple of 4 words), made of, in this order and each
 * only when the tile's flag is set: int32 pos[32] (SD_WT_POS32: positions that do not fit base_pos + 16 bits), uint16
 * ref[32] (SD_WT_REFS: the tile's records are not all on tiles[t].ref), uint32 l_seq | n_cigar << 16 [32] followed by
 * the CIGAR ops of the 32 rows back to back (SD_WT_SHAPE; a tile without it holds records that are one M block of
 * tiles[t].lseq bases each -- padding records, flagged SD_F_PAD in mqf, have no bases).
 * mp: as sd_batch.mp, tile blocks at tiles[t].mp_off (words, multiple of 4).
 */
#define SD_WT_POS32 1u
#define SD_WT_REFS 2u
#define SD_WT_SHAPE 4u
typedef struct sd_wire_tile {
    uint64_t base_off;      /* bases */
    uint32_t var_off;       /* 32-bit words */
    uint32_t mp_off;        /* 32-bit words */
    int32_t base_pos;
    uint16_t ref;
    uint16_t flags;         /* SD_WT_* */
    uint16_t lseq;
    uint16_t reserved0;
    uint32_t reserved1;
} sd_wire_tile;

typedef struct sd_wire {
    uint64_t n_reads;
    uint32_t n_tiles;
    uint32_t tile_reads;           /* 32 */
    const sd_wire_tile *tiles;     /* [n_tiles + 1]; the last entry holds the column totals */
    const uint16_t *pos16, *mqf, *xi16, *nm16, *nmp16;
    const float *xi32;             /* used when xi16 is NULL */
    const uint8_t *seq, *qual;
    const uint32_t *var, *mp;      /* mp may be NULL */
    const float *xi_dict;          /* [n_xi_dict] */
    uint32_t n_xi_dict;
    uint32_t qual_bits;            /* 2, 4 or 8 */
    uint8_t qual_dict[16];         /* code -> phred value, ascending */
    uint32_t max_lseq, max_span;   /* as in sd_batch */
    uint32_t max_tile_cig;         /* largest CIGAR block of a tile in bytes (rounded up to 16), 0 when no tile has SD_WT_SHAPE */
    uint32_t max_tile_mp;          /* largest MP block of a tile in bytes */
} sd_wire;

/* sd_count_host for the wire form: copies the columns this call needs to the device in chunks of whole tiles (copy
 * stream overlapped with the kernels), expands each chunk into an sd_batch with 2-bit SEQ, quality bit-planes and span
 * descriptors, and counts it.  Host pointers (page-locked for full speed).  Synchronous. */
int sd_count_wire(sd_ctx *ctx, const sd_wire *h_wire, const sd_params *params);
/* Wire form of a device-resident batch (2-bit SEQ column, byte qualities): pass 1 fills d_tiles [n_tiles + 1] and reports
 * h_totals = {bases, var words, mp words, largest CIGAR block in bytes}; pass 2 writes the columns into caller-allocated
 * DEVICE buffers of those sizes (out->seq / out->qual zero-filled by the caller; out->xi16 / xi_dict / nmp16 / mp may be
 * NULL; d_xi_code: optional uint16 [n_tiles * 32] dictionary code of every record's XI, copied to out->xi16; out->tiles =
 * d_tiles of pass 1; qual_bits and qual_dict from sd_qual_dict).  Both synchronous.  Copy the buffers to the host to get
 * what sd_decode_bam would have produced for the same records. */
int sd_wire_plan(sd_ctx *ctx, const sd_batch *d_batch, sd_wire_tile *d_tiles, uint64_t h_totals[4]);
int sd_wire_fill(sd_ctx *ctx, const sd_batch *d_batch, const uint16_t *d_xi_code, sd_wire *out);

/* ------------------------------------------------------------------ read filter
 * filter.Filter decision logic (dunks/filter.py:225-271) for one device-resident batch in FILE
 * order.  keep: device uint8[n_reads], 0 drop / 1 write / 2 write with secondary+supplementary
 * cleared (multimapper retained by dumpBufferToBam, filter.py:56-65).  With h_utr_* == NULL the
 * default branch (MQ / identity / NM) runs; otherwise the multimapper-retention branch, whose
 * UTR intervals are [start, stop + 1) (utils/BedReader.py:28) and compare by name id.
 * stats: device int64[SD_NSTAT].  d_state (optional): device uint8[n_reads], bits 0-1 survivor class (0 dropped by the
 * predicates, 1 unique survivor, 2 multimapper survivor), bit 2: the alignment overlaps a UTR interval.  Synchronous.
 */
int sd_filter(sd_ctx *ctx, const sd_batch *d_batch, const sd_params *params, const uint64_t *d_name_hash,
              const uint32_t *h_utr_chrom, const int64_t *h_utr_start, const int64_t *h_utr_stop,
              const uint32_t *h_utr_name_id, uint32_t n_utr, uint8_t *d_keep, int64_t *d_stats, uint8_t *d_state);

/* ------------------------------------------------------------------ host decode (C++ threads)
 * Replaces htslib/pysam for this path: BGZF inflate + BAM record decode into the columnar
 * layout above, FASTA into bit-planes.  Buffers are page-locked and owned by the returned
 * object; free with the matching *_free.
 */
#define SD_DECODE_MP 1
#define SD_DECODE_GROUP 2
#define SD_DECODE_RAW_QUAL 4   /* keep one byte per base quality even when the file has <= 16 distinct values */
#define SD_DECODE_SEQ2 16      /* write the SEQ column in the 2-bit form (sd_batch.seq_bits = 2) -- for sd_count* only */
#define SD_DECODE_WIRE 32      /* also build the wire form of the batch (sd_hbatch.wire): what `count` sends over PCIe */
#define SD_DECODE_MP_ALL 8     /* mp column holds EVERY MP entry, two words each: {readPos-1 | refPos-1 << 16, conv (0..24)} --
                                  the form sd_read_stats takes (sd_count cannot read it); implies SD_DECODE_MP */

typedef struct sd_hbatch {
    sd_batch batch;          /* host pointers */
    uint64_t *name_hash;     /* [n_reads] 64-bit hash of the query name, in FILE order (filter multimapper pass) */
    uint32_t *order;         /* [n_reads] batch slot -> index of the record in the file (identity unless grouped) */
    char *header_text;       /* SAM header text, NUL-terminated */
    uint32_t n_ref;          /* BAM reference sequences */
    char **ref_names;
    uint32_t *ref_lengths;
    double seconds_inflate;  /* decode cost, reported separately from the counting kernels */
    double seconds_decode;
    uint64_t compressed_bytes;
    uint64_t inflated_bytes;
    uint64_t n_mapped_no_xi; /* mapped records without an XI tag (their sd_read_rec.xi is NaN) / without an NM tag (nm = 0): */
    uint64_t n_mapped_no_nm; /* read.get_tag() raises KeyError for them in the reference (filter.py:246,249) */
    sd_wire wire;            /* SD_DECODE_WIRE: host pointers; n_tiles = 0 otherwise */
    void *opaque;
} sd_hbatch;

/* fasta_names / n_fasta: chromosome table of the FASTA (index = sd_read_rec ref field); n_fasta = 0:
 * the ref field is the BAM reference id itself (what sd_filter works in).
 * options: SD_DECODE_MP also decodes the MP tags into the mp column; SD_DECODE_GROUP orders the batch by
 * alignment shape and strand (plain single-match-block reads first, each group in file order) so that the
 * count kernel's warps are homogeneous -- use file order (no grouping) for sd_filter.  Base qualities are written as
 * 2- or 4-bit dictionary codes when the file has at most 4 / 16 distinct values (sd_batch.qual_bits) unless
 * SD_DECODE_RAW_QUAL is given.  tile_reads: 32 (0 = default).
 * threads <= 0: hardware concurrency. */
int sd_decode_bam(const char *path, const char *const *fasta_names, uint32_t n_fasta, int options,
                  uint32_t tile_reads, int threads, sd_hbatch **out, char *err, int errlen);
void sd_hbatch_free(sd_hbatch *hb);

typedef struct sd_hgenome {
    uint32_t n_chrom;
    char **names;
    uint32_t *chrom_len;
    uint32_t *chrom_off;   /* start of each chromosome in the linear bit space (multiple of 32) */
    uint64_t n_words;
    uint32_t *lo, *hi, *nmask;
    double seconds;
    void *opaque;
} sd_hgenome;

int sd_load_fasta(const char *path, int threads, sd_hgenome **out, char *err, int errlen);
void sd_hgenome_free(sd_hgenome *g);

/* ------------------------------------------------------------------ base-quality transfer coding
 * sd_qual_unpack: device codes -> device bytes (what sd_count_host does per chunk; also for uploading a host batch by
 * hand).  sd_qual_pack: the reverse for a device-resident column: *bits receives 2, 4 or 8 (8: more than 16 distinct
 * values, nothing written), dict the code table.  n_bases = length of the qual column in bases (a multiple of 16);
 * d_packed needs n_bases * bits / 8 bytes (n_bases / 2 always suffices for a packable column). */
int sd_qual_unpack(sd_ctx *ctx, const uint8_t *d_packed, uint64_t n_bases, uint32_t bits, const uint8_t *dict16, uint8_t *d_qual);
int sd_qual_pack(sd_ctx *ctx, const uint8_t *d_qual, uint64_t n_bases, uint8_t *d_packed, uint32_t *bits, uint8_t *dict16);
/* sd_seq_pack: the 2-bit form (sd_batch.seq_bits = 2) of a device-resident 4-plane SEQ column: d_seq2 receives half the
 * column's bytes, d_tiles2 [n_tiles + 1] the tile table with the seq offsets of the new column (other offsets copied).
 * What sd_decode_bam does with SD_DECODE_SEQ2, for batches that were made on the device.  Asynchronous. */
int sd_seq_pack(sd_ctx *ctx, const sd_batch *d_batch, uint32_t *d_seq2, sd_tile *d_tiles2);
/* sd_qual_dict: the distinct values of a device-resident byte qual column: *bits = 2 / 4 when there are at most 4 / 16 of
 * them (dict16 = the values in ascending order, code = index), else 8.  Synchronous.
 * sd_qual_to_planes: the plane form (sd_batch.qual_form = 1) of a device batch whose qual column holds bytes: d_planes
 * receives 4 * bits bytes per (group, row slot) of the batch's SEQ blocks, d_tiles_out [n_tiles + 1]
 * the tile table with the qual offsets of the new column (other offsets copied).  Values missing from dict16 get code 0.
 * Asynchronous. */
int sd_qual_dict(sd_ctx *ctx, const uint8_t *d_qual, uint64_t n_bases, uint32_t *bits, uint8_t *dict16);
int sd_qual_to_planes(sd_ctx *ctx, const sd_batch *d_batch, uint32_t bits, const uint8_t *dict16, uint32_t *d_planes,
                      sd_tile *d_tiles_out);

/* ------------------------------------------------------------------ filtered BAM output (host, C++ threads + zlib)
 * Replaces pysam's outfile.write / dumpBufferToBam and `samtools sort` after the decisions of sd_filter
 * (dunks/filter.py:56-65, 191, 273-310).  keep[i] as produced by sd_filter; state[i] (sd_filter's optional output:
 * bits 0-1 survivor class 0 none / 1 unique / 2 multimapper, bit 2 "overlaps a UTR") is needed when any keep[i] == 2, to
 * rebuild the reference's RD:Z text.  Kept records are copied byte for byte; sort != 0: coordinate sort like samtools
 * (tid, pos, strand; stable).  header_text: the new SAM header text.  level: zlib level 1-9. */
int sd_rewrite_bam(const char *in_path, const char *out_path, const char *header_text, const uint8_t *keep,
                   const uint8_t *state, uint64_t n_records, int sort, int level, int threads, uint64_t *n_written,
                   char *err, int errlen);
/* The same, and a BAI index of the output next to it (pysam.index, filter.py:313 / tcounter.py:526-527): binning index,
 * 16 kb linear index and htslib's metadata pseudo-bin, over the virtual offsets of the BGZF blocks just written.  The
 * output must come out coordinate sorted (sort != 0, or a sorted input); bai_path NULL = no index. */
int sd_rewrite_bam_indexed(const char *in_path, const char *out_path, const char *bai_path, const char *header_text,
                           const uint8_t *keep, const uint8_t *state, uint64_t n_records, int sort, int level, int threads,
                           uint64_t *n_written, char *err, int errlen);

/* ------------------------------------------------------------------ BGZF block inflate
 * sd_inflate_raw: one raw DEFLATE stream (RFC 1951; the payload of a BGZF block, htslib's bgzf.c inflate_block) decoded
 * into exactly out_len bytes.  Returns 1 on success and 0 when the stream is malformed, does not produce exactly
 * out_len bytes, or uses a form this decoder leaves to zlib (an incomplete Huffman code other than a single one-bit
 * code); it never reads or writes outside the two buffers.  sd_decode_bam tries it on every block, checks the block's
 * CRC32 and falls back to zlib's inflate on a 0 or a CRC mismatch.  Host pointers. */
int sd_inflate_raw(const uint8_t *in, uint64_t in_len, uint8_t *out, uint64_t out_len);
/* CRC-32 of a buffer as gzip / BGZF define it (zlib's crc32(0, p, n)); carry-less-multiply folding where the CPU has
 * PCLMULQDQ, zlib's table code otherwise and for the last n % 16 bytes. */
uint32_t sd_crc32(const uint8_t *p, uint64_t n);
/* sd_deflate_raw: the other direction, for the BAM writer -- at most 65 535 bytes in, one complete raw DEFLATE stream out
 * (a single dynamic-Huffman block, or a stored block when that is smaller); level 1..9 sets the depth of the match search
 * (1: two candidates per position, greedy ... 6: 12 candidates with one step of lazy evaluation ... 9: 96).  Returns the
 * number of bytes written, 0 when n > 65535 or the stream does not fit `cap` (the writer then uses zlib's deflate). */
uint64_t sd_deflate_raw(const uint8_t *in, uint64_t n, uint8_t *out, uint64_t cap, int level);

/* ------------------------------------------------------------------ host text output
 * sd_format_repr: x printed like Python's repr(float) (what the reference's print() of a rate produces); buf >= 32 bytes.
 * sd_write_bedgraphs: the two per-position conversion bedgraphs (tcounter.py:277-285, 315-326) from host copies of
 * the per-position arrays.  Per BED row: chromosome text, first in-chromosome position max(0, start), strand,
 * whether the row had reads, its offset / length in the position arrays (sd_utr_layout) and the linear genome
 * coordinate of its first position; ref_lo / ref_hi / ref_nmask: the host planes of sd_load_fasta ("base is T" = lo & hi & ~nmask
 * on '+' rows, "base is A" = ~lo & ~hi & ~nmask on '-' rows).  The rows are
 * formatted by `threads` host threads (0 = all cores) and written in BED order. */
int sd_format_repr(double x, char *buf);
/* The per-UTR rows of the tcount TSV (SlamSeqFile.py:99-100 as filled by tcounter.py:251-307) for BED rows in file order:
 * counters = int64 [n_utr][SD_NCOUNTER] and tcontent = int64 [n_utr] as read back from the device, read_number = the
 * header's FilteredReads (CPM denominator), row_strand 0 = '+', 1 = '-'.  *text is malloc'ed (release: sd_free_text). */
int sd_format_tcount_rows(uint32_t n_utr, const char *const *row_chrom, const int64_t *row_start, const int64_t *row_stop,
                          const char *const *row_name, const uint8_t *row_strand, const int64_t *counters,
                          const int64_t *tcontent, int64_t read_number, char **text, uint64_t *text_len);
void sd_free_text(char *text);
int sd_write_bedgraphs(const char *path_plus, const char *path_minus, uint32_t n_utr, const char *const *row_chrom,
                       const int64_t *row_first_pos, const uint8_t *row_strand, const uint8_t *row_has_reads,
                       const uint64_t *row_pos_off, const uint32_t *row_pos_len, const uint64_t *row_gstart,
                       const uint32_t *ref_lo, const uint32_t *ref_hi, const uint32_t *ref_nmask, const int32_t *pos_cov,
                       const int32_t *pos_tc, uint64_t n_positions, int threads);

/* ------------------------------------------------------------------ genome-wide per-position tracks
 * `alleyoop positional-tracks` (dunks/tcounter.py:332-487 genomewideConversionRates).  The per-position arrays come from
 * sd_count + sd_finalize run with one annotation row per chromosome strand starting at position 1 (the reference's
 * iterator has startPosition 1, SlamSeqFile.py:451: array index p = 0-based genomic position p + 1).
 * sd_track_runs compacts one track of one row into its change points on the device: d_pos[j] = index whose value
 * differs from the previous index's (index -1 counts as 0), d_val[j] = the value from there on, d_isfloat[j] (RATE
 * track, may be NULL otherwise) = 1 when the reference holds that value as a float (coverage >= cutoff) rather than
 * the int 0.  track_len = length of the reference's array (the chromosome length); indices past the row's clipped
 * length read as 0.  Call with d_pos == NULL (or capacity 0) to get the number of change points in *n_runs.
 * sd_write_track appends the reference's run-length lines (tcounter.py:417-476) for those change points to `path`:
 * one line per change point, "<chrom>\t<previous change + 1>\t<this change + 1>\t<value before>"; the run still
 * open at the end of the chromosome is never written, exactly as in the reference. */
#define SD_TRACK_COVERAGE 0         /* reads covering the position (coverage_plus / coverage_minus) */
#define SD_TRACK_COVERAGE_ON_BASE 1 /* the same where the FASTA base is T (plus) / A (minus), else 0 */
#define SD_TRACK_CONVERSIONS 2      /* T>C (plus) / A>G (minus) conversions at the position */
#define SD_TRACK_RATE 3             /* conversions / coverage where coverage >= max(1, cutoff), else int 0 */
int sd_track_runs(sd_ctx *ctx, uint32_t row, int kind, int coverage_cutoff, uint32_t track_len, uint32_t *d_pos,
                  double *d_val, uint8_t *d_isfloat, uint64_t capacity, uint64_t *n_runs);
int sd_write_track(const char *path, const char *chrom, int kind, uint64_t n_runs, const uint32_t *h_pos,
                   const double *h_val, const uint8_t *h_isfloat);

/* ------------------------------------------------------------------ conversion statistics over the same per-read scan
 * `alleyoop rates / tcperreadpos / utrrates / tcperutrpos / snpeval / tccontext` (dunks/stats.py:85-131, 591-657, 331-528,
 * 659-772, 774-849, 226-330) and the per-read lists of computeTconversions(mle=True) (dunks/tcounter.py:233-243) all iterate
 * SlamSeqBamIterator (slamseq/SlamSeqFile.py:364-402) and look at the COMPLETE mismatch list of every read -- all 25
 * conversion types of the MP tag that pass the base-quality test -- plus the read's 5 x 5 conversion matrix
 * (computeRatesForRead, SlamSeqFile.py:231-249).  sd_read_stats does that scan for one device-resident batch decoded
 * with SD_DECODE_MP_ALL: one thread per read, per-UTR quantities through the interval join of sd_set_utrs (strand
 * specific, SlamSeqFile.py:369), genome-wide ones over every placed record (readsInChromosome, strand ".").  SNP masking
 * follows sd_set_snps (none set = snps None).  Every output is optional (NULL = not wanted), caller-owned device memory,
 * ACCUMULATED into (zero it first); all integers.  Asynchronous.
 */
typedef struct sd_read_pair {  /* one (read, BED row) incidence for the mle per-read lists */
    uint32_t row;              /* BED row */
    uint32_t read;             /* index of the record in the file (d_order[slot]; the slot itself without d_order) */
    uint32_t t_count;          /* testN: T (A on reverse reads) in the read sequence + in-UTR mismatches on a T/A reference base, tcounter.py:233-238 */
    uint32_t tc_count;         /* testk: in-UTR T>C / A>G conversions of the read, tcounter.py:239-240 */
} sd_read_pair;

#define SD_NRATE 25            /* conv = 5 * refBase + readBase over A,C,G,T,N (SlamSeqFile.py:251-257) */
typedef struct sd_stats_args {
    int32_t min_qual;              /* minQual / minBaseQual: an MP entry counts when its base quality >= this (SlamSeqFile.py:330) */
    int32_t conversion_threshold;  /* isTcRead = tcCount >= this (SlamSeqFile.py:397); the stats use the iterator's default 1 */
    int32_t strict;                /* what happens to reads that are not isTcRead: utr_snp and pairs see an empty mismatch list
                                      (stats.py:806-810 with -m; tcounter.py:210-214 always); utr_rates skips them when
                                      tcCount > 0 (stats.py:376 with -m) */
    uint32_t max_read_length;      /* readpos: reads longer than this are left out (stats.py:621) */
    uint32_t utr_norm;             /* utrpos: utrNormFactor, 250 (stats.py:35) */
    const uint32_t *d_order;       /* optional device uint32[n_reads]: batch slot -> record index in the file */
    int64_t *d_rates;              /* [2][25] summed conversion matrices of forward / reverse reads (stats.py:112-117) */
    int64_t *d_readpos;            /* [6][max_read_length + 1]: per read position (reverse reads counted from their end,
                                      SlamSeqFile.py:335) other mismatches fwd, rev; T>C fwd, A>G rev; then reads covering
                                      the position fwd, rev in DIFFERENCE form (+1 at 0, -1 at the read length) */
    int64_t *d_utr_rates;          /* [n_utr][26]: summed matrices of the row's reads, then their number (stats.py:369-384) */
    int64_t *d_utr_snp;            /* [n_utr][3]: readCount, unmaskedTCCount, maskedTCCount (stats.py:799-835) */
    int64_t *d_utrpos;             /* [6][utr_norm + 1] over the last utr_norm positions of '+' rows / the first of '-' rows
                                      (stats.py:693-747): other mismatches fwd, rev; T>C fwd, A>G rev; coverage fwd, rev in
                                      DIFFERENCE form; the reverse arrays in the reference's own (unreversed) index */
    int64_t *d_context;            /* [2][2][5] `alleyoop tccontext` (stats.py:226-330): bases next to every T of a forward read / every A
                                      of a reverse read in the READ sequence -- [0] 5' neighbour, [1] 3' neighbour (of the T strand:
                                      reverse reads are complemented and their sides swapped), [.][0] forward / [.][1] reverse reads,
                                      [.][.][A, C, G, T, N].  Over every record placed on a FASTA chromosome, unmapped ones included
                                      (bamFile.fetch(region=chromosome), no iterator) */
    sd_read_pair *d_pairs;         /* [pair_capacity] */
    uint64_t pair_capacity;
    unsigned long long *d_n_pairs; /* number of pairs found (may exceed the capacity: then only the first are stored) */
} sd_stats_args;
int sd_read_stats(sd_ctx *ctx, const sd_batch *d_batch, const sd_stats_args *args);

/* ------------------------------------------------------------------ synthetic SLAM-seq reads
 * Device-side generator of the BASELINE.json configurations (per-T Bernoulli conversions as in
 * the reference's simulator, dunks/simulator.py:212-229).  Counter-based RNG keyed by
 * (seed, read index): any sharding of the index range yields the same reads.
 */
typedef struct sd_synth_params {
    uint64_t seed;
    uint64_t first_read;      /* global index of the first read of this batch */
    uint64_t n_reads;
    uint32_t read_len;
    uint32_t tile_reads;
    float p_conv;             /* per reference T (A on '-') conversion probability */
    float p_err;              /* other substitution errors per base */
    float frac_softclip;      /* reads with <= 5 soft-clipped bases */
    float frac_indel;         /* reads with one 1-3 bp insertion or deletion */
    float frac_multimap;      /* mapq 0 */
    float frac_antisense;     /* reads on the strand opposite to their UTR */
    float frac_labelled;      /* reads that carry conversions at all */
    uint32_t utr_first, utr_count; /* restrict placement to this range of the expression table */
    const uint32_t *order;    /* optional device uint32[n_reads]: batch slot j holds read first_read + order[j]
                                 (e.g. the coordinate-sort permutation, as `slamdunk filter` sorts its output) */
} sd_synth_params;

/* Device-resident tables the generator reads. */
typedef struct sd_synth_tables {
    const uint32_t *lo, *hi, *nmask; /* genome planes as for sd_set_reference */
    const uint32_t *chrom_off;       /* device copies of the chromosome table */
    const uint32_t *chrom_len;
    uint32_t n_chrom;
    uint32_t n_utr;                  /* expression table: one entry per UTR */
    const uint32_t *utr_chrom;       /* chromosome index */
    const uint32_t *utr_start;       /* chromosome-local start / end (end clipped to the chromosome) */
    const uint32_t *utr_end;
    const uint8_t *utr_strand;
    const double *utr_cum_weight;    /* inclusive cumulative expression weight */
    const float *utr_label_frac;     /* optional: per-UTR fraction of labelled reads, overriding sd_synth_params.frac_labelled --
                                        a 4sU time point, 1 - exp(-ln2 t / halflife(u)) (dunks/simulator.py:291-302) */
    const uint32_t *snp_tc, *snp_ag, *snp_other; /* optional bit planes over the linear genome (like nmask): true SNP sites of the
                                        sample -- reference T with alternative allele C, reference A with alternative G, any
                                        other substitution (the read then shows the next base code: A>C, C>G, G>T, T>A) */
    float p_snp_allele;              /* probability that a read covering a SNP site carries the alternative allele (BASELINE.json
                                        config 4: 0.9); reads of both strands, labelled or not */
} sd_synth_tables;

#define SD_SYNTH_MAX_TRIPLES 40

/* Linear genome position of every read's alignment start (placement only), device uint32[n_reads], and
 * optionally its grouping class (device uint8[n_reads]: bit 0 strand, bit 1 "not a plain match block", the
 * classes sd_decode_bam groups by); argsort them to build sd_synth_params.order.  Ignores p->order.  Synchronous. */
int sd_synth_positions(sd_ctx *ctx, const sd_synth_params *p, const sd_synth_tables *t, uint32_t *d_gpos, uint8_t *d_cls);
/* Pass 1: simulate every read once to learn its CIGAR and MP entry counts.
 * d_counts: device uint32[n_reads] scratch (n_cigar | n_mp << 16), d_tiles: device sd_tile[n_tiles + 1]
 * (filled), h_totals: host uint64[4] = total bytes of the seq, qual, cigar and mp columns.  Synchronous. */
int sd_synth_plan(sd_ctx *ctx, const sd_synth_params *p, const sd_synth_tables *t, uint32_t *d_counts,
                  sd_tile *d_tiles, uint64_t h_totals[4]);
/* Pass 2: fill caller-allocated device columns.  `out` is a host struct whose pointers are device
 * buffers of the sizes pass 1 reported (rec: n_tiles * tile_reads records; tiles = d_tiles of pass 1);
 * its scalar fields are filled in.  d_triples (optional): int32 [n_reads][1 + 3 * SD_SYNTH_MAX_TRIPLES],
 * per read the number of mismatching aligned columns followed by (conv, readPos, refPos) triples as
 * NextGenMap would list them in MP:Z -- the generator's own record of what it mutated, not a re-walk
 * of the encoded alignment.  Synchronous. */
int sd_synth_fill(sd_ctx *ctx, const sd_synth_params *p, const sd_synth_tables *t, const uint32_t *d_counts,
                  sd_batch *out, int32_t *d_triples);
/* Random genome planes: uniform ACGT, a fraction n_frac of 64-bp blocks all N; pad positions N.
 * Device pointers; chromosome table on the host.  Synchronous. */
int sd_synth_genome(sd_ctx *ctx, uint64_t seed, uint32_t *d_lo, uint32_t *d_hi, uint32_t *d_nmask,
                    uint64_t n_words, const uint32_t *h_chrom_off, const uint32_t *h_chrom_len,
                    uint32_t n_chrom, float n_frac);

#ifdef __cplusplus
}
#endif
#endif /* SLAMDUNK_B200_H */
