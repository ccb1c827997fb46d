// sd_count_kernel.cuh -- the count kernel (device code) and its launcher.  Included by sd_count.cu (host API, dispatch)
// and by the sd_count_inst_*.cu translation units, each of which instantiates the launcher for one column form so that
// the 24 kernel variants compile side by side instead of in one 80-second translation unit.
#pragma once
// The `slamdunk count` hot path for B200 (sm_100a).
//
// One persistent kernel streams the columnar read batch through shared memory with TMA bulk
// copies (cp.async.bulk + mbarrier, a ring of stages per CTA) and does, per read (one thread per
// read, one warp per 32-read group):
//   1. the read filter predicates                      (reference: dunks/filter.py:241-251)
//   2. the T>C / A>G conversion scan                   (slamseq/SlamSeqFile.py:313-348, 364-402)
//      as word-parallel bit operations: the read's "is C" ("is G" on the minus strand) mask is
//      one AND-reduction over the sliced 4-bit SEQ planes, aligned to reference coordinates by
//      the CIGAR, and ANDed with the reference's "is T / not a SNP" planes of the read's window;
//      base qualities are looked up only at the surviving candidate positions;
//   3. the read -> UTR interval join (binned start-sorted table) and the per-UTR counters
//      readCount / tcReadCount / multimapCount / coverageOnTs / conversionsOnTs
//      (dunks/tcounter.py:207-248, 277-283) with warp-aggregated int64 atomics;
//   4. the per-position coverage (difference form) and conversion scatter that the bedgraphs
//      need (tcounter.py:229-231, 246-248, 285).
// The path is integer streaming work bound by HBM bandwidth; there is no tensor-core use.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <numeric>

#include "sd_internal.h"

// -DSD_DEBUG_BOUNDS: device-side asserts on every scatter / gather index (compute-sanitizer stand-in for the test suite)
#ifdef SD_DEBUG_BOUNDS
#include <assert.h>
#define SD_CHECK(cond) assert(cond)
#else
#define SD_CHECK(cond) ((void)0)
#endif

// ------------------------------------------------------------------------------------------
// small PTX wrappers (mbarrier + 1-D bulk async copy; SASS: SYNCS.*, UBLKCP)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// Blocks in hardware (try_wait with a suspend-time hint) instead of spinning on issue slots.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok = 0;
    const uint32_t addr = smem_u32(bar);
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity), "r"(200000u)
            : "memory");
    }
}

// ------------------------------------------------------------------------------------------
// kernel parameters
// ------------------------------------------------------------------------------------------
#define SD_TILE_READS 32

struct CountArgs {
    // batch (device pointers)
    const sd_read_rec *rec;
    const uint32_t *seq;
    const uint8_t *qual;
    const uint32_t *cigar;
    const uint32_t *mp;
    const sd_tile *tiles;
    uint32_t n_tiles;
    uint32_t tile_reads;
    uint32_t seq2;            // the SEQ column holds 2-bit base codes (sd_batch.seq_bits == 2): two words per group
    uint32_t qual_planes;     // 0: qual column = bytes, read in place; 2 / 4: group-major bit-planes of the dictionary code
                              // (sd_batch.qual_form == 1), staged with the tile
    uint32_t q_cmin;          // plane form: quality >= min_qual  <=>  code >= q_cmin (codes ascend with quality); 2^planes: none
    uint32_t q_m[4];          // the same test as all-ones / all-zeros masks (constant-bank operands of the LOP3s):
                              // 2 planes: ok = (((p1 & (p0 | m0)) | (p0 & m1)) | m2) & m3;  4 planes: m[b] = bit b of q_cmin
    // shared-memory stage layout (bytes, all multiples of 128)
    uint32_t cap_rec, cap_seq, cap_qual, cap_cig, cap_mp;
    uint32_t cap_win, cap_ent;   // staged reference window / annotation entries (0: the batch has no span descriptors)
    uint32_t n_stages;
    uint32_t qual_prefetch;   // L2-prefetch each tile's qual bytes when the tile is issued
    // per-tile span descriptors (sd_batch.spans) and what they index: the strand-merged reference and annotation tables
    const void *spans;        // sd_tile_span per tile or null
    uint32_t hint_lean, hint_staged;   // tiles the lean kernel will take / that fit a stage (sd_batch.spans_*), 0xFFFFFFFF: not known
    const uint4 *refm;        // per 32 bp {isT, T>C SNPs, isA, A>G SNPs}
    const uint4 *ment;        // two uint4 per entry, see StageView
    // reference
    const uint2 *refx[2];
    const uint32_t *chrom_off;
    const uint32_t *chrom_len;
    uint32_t n_chrom;
    // annotation
    const uint32_t *ustart[2];
    const uint32_t *uend[2];
    const uint32_t *urow[2];
    const uint32_t *useg[2];
    const uint4 *utab[2];   // {start, end, row, seg}
    const uint4 *stab[2];   // {start, end, off_lo, off_hi}
    const uint32_t *bin_first[2];
    uint32_t n_u[2];
    uint32_t n_bins;
    const uint32_t *seg_start[2];
    const uint32_t *seg_end[2];
    const uint64_t *seg_off[2];
    // outputs
    unsigned long long *counters;
    int *pos_cov;
    int *pos_tc;
    unsigned long long *stats;
    uint8_t *read_tc;  // optional per-read tcCount (saturated), may be null
    unsigned int *tile_counter;   // work queue head: warps claim chunks of consecutive tiles
    // The batch's tiles are counted by two launches when it carries span descriptors: the STAGED kernel takes every tile whose
    // reference window and annotation entries fit a stage and appends the others (reads spread over the genome) to `leftover`;
    // the plain kernel then counts that list (tile_list / n_list) with per-read lookups.  Without descriptors the plain kernel
    // runs over all tiles (tile_list null).
    uint32_t *leftover;
    unsigned int *n_leftover;
    const uint32_t *tile_list;
    const unsigned int *n_list;
    // parameters
    int min_qual;
    int conv_threshold;
    int mismatch_source;
    int filter_on;
    int filter_mq;
    int filter_max_nm;
    float filter_min_identity;   // smallest float >= the double threshold: xi < this  <=>  (double)xi < threshold
    int force_slow;
    int mp_any_base;   // SD_PARAM_MP_ANY_BASE
    // bounds (checked in SD_DEBUG_BOUNDS builds only)
    unsigned long long n_positions, qual_bytes;
    uint32_t n_utr;
};

__device__ __forceinline__ bool cigar_consumes_ref(uint32_t op) { return op == 0 || op == 2 || op == 3 || op == 7 || op == 8; }
__device__ __forceinline__ bool cigar_is_match(uint32_t op) { return op == 0 || op == 7 || op == 8; }
__device__ __forceinline__ bool cigar_consumes_query(uint32_t op) { return op == 0 || op == 1 || op == 4 || op == 7 || op == 8; }

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

// @region helpers
// bits [lo, hi) of the window restricted to word k
__device__ __forceinline__ uint32_t range_word(int k, int lo, int hi) {
    int a = lo - 32 * k, b = hi - 32 * k;
    a = a < 0 ? 0 : a;
    b = b > 32 ? 32 : b;
    if (b <= a) return 0u;
    uint32_t m = (b == 32) ? 0xFFFFFFFFu : ((1u << b) - 1u);
    return m & ~((1u << a) - 1u);
}

// @region helpers
// bits [0, hi) of the window restricted to word k: the low max(0, min(32, hi - 32 k)) bits, as one add-and-clamp and one
// clamped funnel shift (the high word of 0x00000000FFFFFFFF << n)
__device__ __forceinline__ uint32_t prefix_word(int k, int hi) {
    int b = hi - 32 * k;
    b = b < 0 ? 0 : b;
    return __funnelshift_lc(0xFFFFFFFFu, 0u, (uint32_t)b);
}

// @region utr_accumulate
// Warp-aggregated add into one UTR's counter row: lanes that hit the same row elect a leader.
__device__ __forceinline__ void utr_accumulate(unsigned long long *counters, uint32_t row, uint32_t tcread,
                                               uint32_t multimap, uint32_t cov_t, uint32_t conv_t) {
    const unsigned active = __activemask();
    const unsigned grp = __match_any_sync(active, row);
    const uint32_t n = __popc(grp);
    const uint32_t s_tc = __reduce_add_sync(grp, tcread);
    const uint32_t s_mm = __reduce_add_sync(grp, multimap);
    const uint32_t s_cov = __reduce_add_sync(grp, cov_t);
    const uint32_t s_conv = __reduce_add_sync(grp, conv_t);
    if ((threadIdx.x & 31) == (__ffs(grp) - 1)) {
        unsigned long long *c = counters + (size_t)row * SD_NCOUNTER;
        atomicAdd(c + SD_C_READS, (unsigned long long)n);
        if (s_tc) atomicAdd(c + SD_C_TCREADS, (unsigned long long)s_tc);
        if (s_mm) atomicAdd(c + SD_C_MULTIMAP, (unsigned long long)s_mm);
        if (s_cov) atomicAdd(c + SD_C_COV_ON_T, (unsigned long long)s_cov);
        if (s_conv) atomicAdd(c + SD_C_CONV_ON_T, (unsigned long long)s_conv);
    }
}

// one base quality, fetched where a candidate survived (the tile's qualities were pulled into L2 when the tile was issued)
__device__ __forceinline__ int ld_qual(const uint8_t *p) { return (int)__ldg(p); }
#define SD_CHECK_QUAL(A, p) SD_CHECK((unsigned long long)((p) - (A).qual) < (A).qual_bytes)

// @region sliced_is_base
// Read mask of one 32-base group of the plane-sliced SEQ column (four words: bit-planes A, C, G, T of the BAM nibble):
// bit i set where base i is exactly C (plus strand) / exactly G (minus strand).  sm = all ones on the minus strand.
__device__ __forceinline__ uint32_t group_is_base(const uint4 p, uint32_t sm) {
    const uint32_t f = ((p.y & ~p.z) & ~sm) | ((p.z & ~p.y) & sm);   // one LOP3: the wanted plane without the other of C/G
    return f & ~p.x & ~p.w;                                          // one LOP3: neither A nor T
}

// The same for the 2-bit SEQ column (two words per group: low / high bit of the base code A=0 C=1 G=2 T=3): C is
// lo & ~hi, G is hi & ~lo -- one LOP3.  Bases that are not one of ACGT are coded A there, pad bits too: never C or G.
__device__ __forceinline__ uint32_t group_is_base2(const uint2 p, uint32_t sm) {
    return ((p.x & ~p.y) & ~sm) | ((p.y & ~p.x) & sm);
}
// group g of the row whose slot of group 0 is `row` (group-major tile block in shared memory)
template <int SB>
__device__ __forceinline__ uint32_t row_group_is_base(const uint32_t *row, uint32_t g, uint32_t sm) {
    if (SB == 2) return group_is_base2(reinterpret_cast<const uint2 *>(row)[g * SD_TILE_READS], sm);
    return group_is_base(reinterpret_cast<const uint4 *>(row)[g * SD_TILE_READS], sm);
}

// Bit-plane quality column (sd_batch.qual_form == 1): QB words per 32-base group, bit-plane b of the base's dictionary
// code.  sd_decode_bam / sd_qual_dict assign codes in ascending quality order, so the reference's test
// `readQlty >= minQual` (SlamSeqFile.py:330) is the bit-sliced comparison code >= q_cmin: 2 LOP3 per plane for 32 bases.
// CM: take the comparison constant as ready-made masks from the parameter block (A.q_m) instead of deriving them from
// A.q_cmin.  Same result; the count kernel uses the mask form (150 bp: 1.83 vs 2.31 ms), the per-base lookups the other.
template <int QB, bool CM, class AR>
__device__ __forceinline__ uint32_t planes_ge(const uint32_t (&p)[QB], const AR &A) {
    if (CM && QB == 2)   // code >= c over two planes: c = 3: p1 & p0; 2: p1; 1: p1 | p0; 0: all; 4: none
        return (((p[1] & (p[0] | A.q_m[0])) | (p[0] & A.q_m[1])) | A.q_m[2]) & A.q_m[3];
    const uint32_t cmin = A.q_cmin;
    uint32_t gt = 0u, eq = 0xFFFFFFFFu;
#pragma unroll
    for (int b = QB - 1; b >= 0; b--) {
        const uint32_t cb = CM ? A.q_m[b] : 0u - ((cmin >> b) & 1u);
        gt |= eq & p[b] & ~cb;
        eq &= ~(p[b] ^ cb);
    }
    return (cmin >> QB) ? 0u : (gt | eq);
}
template <int QB, bool CM = false, class AR = CountArgs>
__device__ __forceinline__ uint32_t row_group_qual_ok(const uint32_t *row, uint32_t g, const AR &A) {
    uint32_t p[QB];
    if (QB == 2) {
        const uint2 x = reinterpret_cast<const uint2 *>(row)[g * SD_TILE_READS];
        p[0] = x.x; p[1] = x.y;
    } else {
        const uint4 x = reinterpret_cast<const uint4 *>(row)[g * SD_TILE_READS];
        p[0] = x.x; p[1] = x.y; p[QB > 2 ? 2 : 0] = x.z; p[QB > 2 ? 3 : 1] = x.w;
    }
    return planes_ge<QB, CM>(p, A);
}
// "Quality" of read base qi (0 <= qi < l_seq) for a `>= A.min_qual` test: the byte itself, or on the plane form 255 / -1
// for pass / fail (min_qual <= 0 makes q_cmin 0: everything passes there too).
template <int QB>
__device__ __forceinline__ int qual_for_test(const CountArgs &A, const uint8_t *qual, int qi) {
    if (QB == 0) return (int)__ldg(qual + qi);
    const uint32_t ok = row_group_qual_ok<(QB ? QB : 2)>(reinterpret_cast<const uint32_t *>(qual), (uint32_t)qi >> 5, A);
    return ((ok >> (qi & 31)) & 1u) ? 255 : -1;
}
__device__ __forceinline__ int qual_for_test_rt(const CountArgs &A, const uint8_t *qual, int qi) {   // per-base path: form known at run time
    return A.qual_planes == 0 ? qual_for_test<0>(A, qual, qi) : (A.qual_planes == 2 ? qual_for_test<2>(A, qual, qi) : qual_for_test<4>(A, qual, qi));
}

// @region slow_and_generic
// Map a reference offset (relative to the alignment start) to the read index aligned to it.
__device__ __forceinline__ int ref_to_read(const uint32_t *cig, uint32_t ncig, int o) {
    int q = 0, r = 0;
    for (uint32_t i = 0; i < ncig; i++) {
        const uint32_t c = cig[i], op = c & 15u;
        const int len = (int)(c >> 4);
        if (cigar_is_match(op)) {
            if (o < r + len) return q + (o - r);
            q += len;
            r += len;
        } else if (op == 1 || op == 4) {
            q += len;
        } else if (op == 2 || op == 3) {
            r += len;
        }
    }
    return -1;
}

struct ReadView {
    const uint32_t *seq;   // this read's slot of group 0 in the tile's group-major plane block (shared memory)
    const uint8_t *qual;   // the read's quality bytes in the global qual column, or (plane form) its slot of group 0 in the tile's
                           // group-major plane block in shared memory
    const uint32_t *cig;   // ops (shared memory)
    const uint32_t *mp;    // entries (shared memory) or null
    uint32_t lseq, ncig, nmp;
    uint32_t g;            // window start in the linear genome
    int span;              // reference bases spanned (clipped to the chromosome)
    int strand;            // 0 '+', 1 '-'
    bool has_mp;
};

// ---------------------------------------------------------------------------------------
// Per-base path: reads whose window does not fit the register window (long deletions / N skips,
// reads longer than the instantiated window) and the testing cross-check (force_slow).
// Same semantics as the word-parallel path, written as plain loops.
// ---------------------------------------------------------------------------------------
template <class F>
__device__ __forceinline__ void slow_for_each_hit(const CountArgs &A, const ReadView &v, F f) {
    const uint2 *rx = A.refx[v.strand];
    if (A.mismatch_source == SD_MISMATCH_MP) {
        if (!v.has_mp || v.mp == nullptr) return;
        for (uint32_t i = 0; i < v.nmp; i++) {
            const uint32_t e = v.mp[i];
            const int rp = (int)(e & 0xFFFFu), o = (int)(e >> 16);
            const int q = (rp >= (int)v.lseq) ? 0 : qual_for_test_rt(A, v.qual, rp);
            if (q < A.min_qual) continue;
            const uint64_t gp = (uint64_t)v.g + (uint64_t)o;
            const uint2 x = __ldg(&rx[gp >> 5]);
            if ((x.y >> (gp & 31)) & 1u) continue;  // SNP
            f(o, (bool)((x.x >> (gp & 31)) & 1u));
        }
        return;
    }
    if (!v.has_mp) return;  // no MP tag: the reference sees no mismatches (SlamSeqFile.py:318)
    const uint32_t sm = v.strand ? 0xFFFFFFFFu : 0u;
    int q = 0, r = 0;
    for (uint32_t i = 0; i < v.ncig; i++) {
        const uint32_t c = v.cig[i], op = c & 15u;
        const int len = (int)(c >> 4);
        if (cigar_is_match(op)) {
            for (int j = 0; j < len; j++) {
                const int o = r + j, qi = q + j;
                if (o >= v.span || qi >= (int)v.lseq) break;
                const uint32_t isx = A.seq2 ? row_group_is_base<2>(v.seq, (uint32_t)qi >> 5, sm) : row_group_is_base<4>(v.seq, (uint32_t)qi >> 5, sm);
                if (!((isx >> (qi & 31)) & 1u)) continue;
                const uint64_t gp = (uint64_t)v.g + (uint64_t)o;
                const uint2 x = __ldg(&rx[gp >> 5]);
                const uint32_t bit = 1u << (gp & 31);
                if (!(x.x & bit) || (x.y & bit)) continue;
                if (qual_for_test_rt(A, v.qual, qi) < A.min_qual) continue;
                f(o, true);
            }
            q += len;
            r += len;
        } else if (op == 1 || op == 4) {
            q += len;
        } else if (op == 2 || op == 3) {
            r += len;
        }
    }
}

__device__ __forceinline__ uint32_t slow_count_t(const uint2 *rx, uint64_t ga, uint64_t gb) {
    // number of set isX bits in [ga, gb)
    uint32_t n = 0;
    if (gb <= ga) return 0;
    uint64_t w0 = ga >> 5, w1 = (gb - 1) >> 5;
    for (uint64_t w = w0; w <= w1; w++) {
        uint32_t x = __ldg(&rx[w]).x;
        if (w == w0) x &= ~((1u << (ga & 31)) - 1u);
        if (w == w1 && ((gb & 31) != 0)) x &= ((1u << (gb & 31)) - 1u);
        n += __popc(x);
    }
    return n;
}

static __device__ __noinline__ int slow_read(const CountArgs &A, const ReadView v, uint32_t mapq) {   // by value: the view stays in registers on the fast path
    int tc = 0;
    slow_for_each_hit(A, v, [&](int, bool) { tc++; });
    const bool is_tc = tc >= A.conv_threshold;
    if (!is_tc) tc = 0;
    const int s = v.strand;
    const uint32_t g = v.g, e = v.g + (uint32_t)v.span;
    const uint32_t nU = A.n_u[s];
    uint32_t i = A.bin_first[s][g >> SD_BIN_SHIFT];
    uint32_t last_seg = 0xFFFFFFFFu;
    while (i < nU) {
        const uint32_t st = A.ustart[s][i];
        if (st >= e) break;
        const uint32_t en = A.uend[s][i];
        if (en > g) {
            const uint32_t a = st > g ? st : g, b = en < e ? en : e;
            const uint32_t cov_t = slow_count_t(A.refx[s], a, b);
            uint32_t conv = 0;
            if (tc > 0)
                slow_for_each_hit(A, v, [&](int o, bool on_t) {
                    const uint32_t gp = g + (uint32_t)o;
                    if (on_t && gp >= a && gp < b) conv++;
                });
            utr_accumulate(A.counters, A.urow[s][i], tc > 0 ? 1u : 0u, mapq == 0 ? 1u : 0u, cov_t, conv);
            const uint32_t seg = A.useg[s][i];
            if (seg != last_seg) {
                last_seg = seg;
                const uint32_t ss = A.seg_start[s][seg], se = A.seg_end[s][seg];
                const uint64_t off = A.seg_off[s][seg];
                const uint32_t pa = ss > g ? ss : g, pb = se < e ? se : e;
                SD_CHECK(pb >= pa && off + (pb - ss) < A.n_positions);
                atomicAdd(&A.pos_cov[off + (pa - ss)], 1);
                atomicAdd(&A.pos_cov[off + (pb - ss)], -1);
                if (tc > 0)
                    slow_for_each_hit(A, v, [&](int o, bool on_t) {
                        const uint32_t gp = g + (uint32_t)o;
                        if ((on_t || (A.mp_any_base && A.mismatch_source == SD_MISMATCH_MP)) && gp >= pa && gp < pb)
                            {
                            SD_CHECK(off + (gp - ss) < A.n_positions);
                            atomicAdd(&A.pos_tc[off + (gp - ss)], 1);
                        }
                    });
            }
        }
        i++;
    }
    return tc;
}

// ---------------------------------------------------------------------------------------
// Word-parallel path.  W = window words (32 reference bases each).
//
// The alignment is reduced to at most three match blocks [r0, r1) of the window, each with the
// offset `sh` that turns a window position into a read index (one block: plain or soft-clipped
// read; two: one insertion or deletion; three: two indels).  With |sh| < 32 every block is a
// multi-word funnel shift of the read mask, all in registers.  Anything else (more blocks,
// larger shifts) goes through a shared-memory scratch copy of the read mask.
// ---------------------------------------------------------------------------------------
struct Blocks {
    int n;               // match blocks found (only the first three are described below)
    int r1a, r1b;        // window end of block 0 / 1 (block 2 runs to the end)
    int r0b, r0c;        // window start of block 1 / 2 (block 0 starts at 0 ... its own r0a below)
    int r0a, r1c;
    int sha, shb, shc;   // read index = window position + sh
    bool regs_ok;        // describable by <= 3 blocks with |sh| < 32
};

// @region parse_cigar
__device__ __forceinline__ int parse_cigar(const uint32_t *cig, uint32_t ncig, Blocks &B) {
    int q = 0, r = 0;
    B.n = 0;
    B.r0a = B.r1a = B.r0b = B.r1b = B.r0c = B.r1c = 0;
    B.sha = B.shb = B.shc = 0;
    for (uint32_t i = 0; i < ncig; i++) {
        const uint32_t c = cig[i], op = c & 15u;
        const int len = (int)(c >> 4);
        // op classes as bit sets over MIDNSHP=X: match {M,=,X}, query-only {I,S}, reference-only {D,N}
        if ((0x181u >> op) & 1u) {
            if (B.n == 0) { B.r0a = r; B.r1a = r + len; B.sha = q - r; }
            else if (B.n == 1) { B.r0b = r; B.r1b = r + len; B.shb = q - r; }
            else if (B.n == 2) { B.r0c = r; B.r1c = r + len; B.shc = q - r; }
            B.n++;
        }
        q += ((0x193u >> op) & 1u) ? len : 0;
        r += ((0x18Du >> op) & 1u) ? len : 0;
    }
    auto small = [](int s) { return s > -32 && s < 32; };
    B.regs_ok = B.n <= 3 && small(B.sha) && small(B.shb) && small(B.shc);
    return r;  // reference length
}

// @region block_index
// window word k of the read mask M shifted so that window position p reads M bit p + sh, |sh| < 32
template <int W>
__device__ __forceinline__ uint32_t shifted_word(const uint32_t (&M)[W], int k, int sh) {
    if (sh >= 0) return __funnelshift_r(M[k], (k + 1 < W) ? M[k + 1] : 0u, (uint32_t)sh);
    return __funnelshift_r(k > 0 ? M[k - 1] : 0u, M[k], (uint32_t)(32 + sh));
}

__device__ __forceinline__ int block_read_index(const Blocks &B, int o) {
    if (B.n <= 1 || o < B.r1a) return o + B.sha;
    if (B.n == 2 || o < B.r1b) return o + B.shb;
    return o + B.shc;
}

// SIMPLE: every read of the tile is one match block without clips (the decoder groups such reads into
// their own tiles), so the read mask needs no placement and a window position is a read index.
//
// The per-read scan is the same whether the tile's reference window and annotation entries were staged in shared
// memory with the tile (staged_read, the common case for coordinate-sorted input) or are fetched from global memory
// lane by lane (fast_read, the fallback): read_side() builds the candidate mask, reference_side() applies it.

// @region read_mask
// --- read-side candidate mask R in window coordinates
template <int W, bool SIMPLE, int SB, int QB>
__device__ __forceinline__ void read_side(const CountArgs &A, const ReadView &v, const Blocks &B, bool regs, uint32_t (&R)[W]) {
    const int s = v.strand;
    if (A.mismatch_source == SD_MISMATCH_MP) {
#pragma unroll
        for (int k = 0; k < W; k++) R[k] = 0u;
        if (v.has_mp && v.mp != nullptr) {
            for (uint32_t i = 0; i < v.nmp; i++) {
                const uint32_t e = v.mp[i];
                const int rp = (int)(e & 0xFFFFu), o = (int)(e >> 16);
                const int q = (rp >= (int)v.lseq) ? 0 : qual_for_test<QB>(A, v.qual, rp);
                if (q >= A.min_qual) {
#pragma unroll
                    for (int k = 0; k < W; k++)
                        if ((o >> 5) == k) R[k] |= 1u << (o & 31);
                }
            }
        }
        return;
    }
    uint32_t sm = 0u - (uint32_t)s;   // all ones on the minus strand
    asm volatile("" : "+r"(sm));      // keep it a mask operand of the LOP3s (not a predicate with selects)
    const uint32_t ng = (v.lseq + 31u) >> 5;
    uint32_t M[W];
#pragma unroll
    for (int k = 0; k < W; k++) {
        // unconditional load of an existing group (index clamped), masked afterwards: no branch per group
        const uint32_t kc = ng ? min((uint32_t)k, ng - 1u) : 0u;
        uint32_t m = row_group_is_base<SB>(v.seq, kc, sm);   // this row's slot of group kc; the tile block is group-major
        // plane-form qualities, plain CIGAR walk: the quality test is one more mask on the read side, for 32 bases at once
        if (QB > 0 && A.mismatch_source == SD_MISMATCH_CIGAR)
            m &= row_group_qual_ok<(QB ? QB : 2), true>(reinterpret_cast<const uint32_t *>(v.qual), kc, A);
        M[k] = ((uint32_t)k < ng) ? m : 0u;
    }
    if (SIMPLE) {
#pragma unroll
        for (int k = 0; k < W; k++) R[k] = M[k];
    } else if (regs) {
        if (B.n <= 1) {
            // one block: bits past its end are cleared by T (masked to the span) later on
            if (B.sha == 0) {
#pragma unroll
                for (int k = 0; k < W; k++) R[k] = M[k];
            } else {
#pragma unroll
                for (int k = 0; k < W; k++) R[k] = shifted_word<W>(M, k, B.sha) & range_word(k, B.r0a, B.r1a);
            }
        } else {
#pragma unroll
            for (int k = 0; k < W; k++) {
                uint32_t r = shifted_word<W>(M, k, B.sha) & range_word(k, B.r0a, B.r1a);
                r |= shifted_word<W>(M, k, B.shb) & range_word(k, B.r0b, B.r1b);
                if (B.n > 2) r |= shifted_word<W>(M, k, B.shc) & range_word(k, B.r0c, B.r1c);
                R[k] = r;
            }
        }
    } else {
        // generic placement through a scratch copy of the read mask (dynamically indexed, so it lives in
        // local memory; only alignments with > 3 match blocks or shifts >= 32 come here)
        uint32_t scratch[W + 1];
#pragma unroll
        for (int k = 0; k < W; k++) {
            R[k] = 0u;
            scratch[k] = M[k];
        }
        scratch[W] = 0u;
        int q = 0, r = 0;
        for (uint32_t i = 0; i < v.ncig; i++) {
            const uint32_t c = v.cig[i], op = c & 15u;
            const int len = (int)(c >> 4);
            if (cigar_is_match(op)) {
#pragma unroll
                for (int k = 0; k < W; k++) {
                    const uint32_t m = range_word(k, r, r + len);
                    if (m) {
                        const int b = q - r + 32 * k;  // read-mask bit aligned to window bit 32k
                        const int idx = b >> 5;        // arithmetic shift: floor
                        const uint32_t lo = (idx >= 0 && idx <= W) ? scratch[idx] : 0u;
                        const uint32_t hi = (idx + 1 >= 0 && idx + 1 <= W) ? scratch[idx + 1] : 0u;
                        R[k] |= __funnelshift_r(lo, hi, (uint32_t)(b & 31)) & m;
                    }
                }
                q += len;
                r += len;
            } else if (op == 1 || op == 4) {
                q += len;
            } else if (op == 2 || op == 3) {
                r += len;
            }
        }
    }
}

// --- R against the reference planes T (is T / A, masked to the span) and S (SNPs): MP cross-check, candidates H
// (quality-tested, threshold applied), the read's tcCount and its T coverage.
template <int W, bool SIMPLE, int SB, int QB>
__device__ __forceinline__ int reference_side(const CountArgs &A, const ReadView &v, const Blocks &B, bool regs, const uint32_t (&R)[W],
                                              const uint32_t (&T)[W], const uint32_t (&S)[W], uint32_t (&H)[W], uint32_t &pop_t,
                                              bool &mp_disagree) {
    // @region verify
    // --- cross-check against NextGenMap's MP list (before quality / SNP masking)
    if (A.mismatch_source == SD_MISMATCH_VERIFY && v.has_mp && v.mp != nullptr) {
        uint32_t D[W];
#pragma unroll
        for (int k = 0; k < W; k++) D[k] = R[k] & T[k];
        bool bad = false;
        for (uint32_t i = 0; i < v.nmp; i++) {
            const uint32_t e = v.mp[i];
            const int rp = (int)(e & 0xFFFFu), o = (int)(e >> 16);
            bool found = false;
#pragma unroll
            for (int k = 0; k < W; k++)
                if ((o >> 5) == k && ((D[k] >> (o & 31)) & 1u)) {
                    D[k] &= ~(1u << (o & 31));
                    found = true;
                }
            if (!found || (SIMPLE ? o : (regs ? block_read_index(B, o) : ref_to_read(v.cig, v.ncig, o))) != rp) bad = true;
        }
#pragma unroll
        for (int k = 0; k < W; k++) bad |= (D[k] != 0u);
        mp_disagree = bad;
    }
    // @region candidates_qual
    // --- candidates: reference is T (A), read is C (G), not a SNP
    int tc = 0;
    if (A.mismatch_source == SD_MISMATCH_MP) {
        // the reference counts tcCount from the MP entry alone (SlamSeqFile.py:343); the FASTA base is
        // only consulted when conversions are summed over T positions (tcounter.py:279)
#pragma unroll
        for (int k = 0; k < W; k++) {
            const uint32_t h = R[k] & ~S[k];
            tc += __popc(h);
            H[k] = A.mp_any_base ? h : (h & T[k]);
        }
    } else if (QB > 0 && A.mismatch_source == SD_MISMATCH_CIGAR) {
        // the read mask already carries the quality test (plane-form qual column)
#pragma unroll
        for (int k = 0; k < W; k++) {
            H[k] = v.has_mp ? (R[k] & T[k] & ~S[k]) : 0u;
            tc += __popc(H[k]);
        }
    } else {
#pragma unroll
        for (int k = 0; k < W; k++) H[k] = v.has_mp ? (R[k] & T[k] & ~S[k]) : 0u;
        // Base quality only where a candidate survived: one byte of the qual column per candidate.
        // The first candidate of every 32-position word is tested in straight-line code, so these quality bytes are in
        // flight together and the common case (at most one candidate per word) needs no loop at all.
        int qv[W];
#pragma unroll
        for (int k = 0; k < W; k++) {
            qv[k] = 255;
            if (H[k]) {
                const int o = 32 * k + (__ffs((int)H[k]) - 1);
                const int qi = SIMPLE ? o : (regs ? block_read_index(B, o) : ref_to_read(v.cig, v.ncig, o));
                if (QB == 0) SD_CHECK_QUAL(A, v.qual + (qi < 0 || qi >= (int)v.lseq ? 0 : qi));
                qv[k] = (qi < 0 || qi >= (int)v.lseq) ? -1 : qual_for_test<QB>(A, v.qual, qi);
            }
        }
        uint32_t C[W], more = 0u;   // C: candidates after the first of each word, still to be tested
#pragma unroll
        for (int k = 0; k < W; k++) {
            C[k] = H[k] & (H[k] - 1u);
            H[k] = (qv[k] >= A.min_qual) ? (H[k] ^ C[k]) : 0u;   // keep the lowest candidate iff its quality passes
            more |= C[k];
        }
        if (more) {
#pragma unroll
            for (int k2 = 0; k2 < W; k2 += 2) {
                unsigned long long rest = (unsigned long long)C[k2] | ((unsigned long long)C[k2 + 1] << 32), fail = 0ull;
                while (rest) {
                    const int bit = __ffsll((long long)rest) - 1;
                    rest &= rest - 1ull;
                    const int o = 32 * k2 + bit;
                    const int qi = SIMPLE ? o : (regs ? block_read_index(B, o) : ref_to_read(v.cig, v.ncig, o));
                    if (qi < 0 || qi >= (int)v.lseq || qual_for_test<QB>(A, v.qual, qi) < A.min_qual) fail |= 1ull << bit;
                }
                H[k2] |= C[k2] & ~(uint32_t)fail;
                H[k2 + 1] |= C[k2 + 1] & ~(uint32_t)(fail >> 32);
            }
        }
#pragma unroll
        for (int k = 0; k < W; k++) tc += __popc(H[k]);
    }
    // @region threshold_popt
    if (tc < A.conv_threshold) {  // not a T>C read: conversions are dropped (tcounter.py:210-214)
        tc = 0;
#pragma unroll
        for (int k = 0; k < W; k++) H[k] = 0u;
    }
    pop_t = 0;
#pragma unroll
    for (int k = 0; k < W; k++) pop_t += __popc(T[k]);
    return tc;
}

// @region scatter
// Per-position outputs of one (read, merged segment) incidence: the coverage difference (+1 at the first covered position,
// -1 after the last) and one add per surviving conversion.  [ss, se) is the segment, posbase = its offset in the
// per-position arrays minus ss, so that position p lives at index posbase + p.
template <int W>
__device__ __forceinline__ void scatter_segment(const CountArgs &A, const uint32_t (&H)[W], int tc, uint32_t g, uint32_t e,
                                                uint32_t ss, uint32_t se, unsigned long long posbase, bool dense_tile) {
    const uint32_t pa = ss > g ? ss : g, pb = se < e ? se : e;
    const unsigned long long ia = posbase + pa, ib = posbase + pb;
    SD_CHECK(pb >= pa && ib < A.n_positions);
    if (dense_tile) {   // the tile's reads start (and end) at a handful of positions: lanes with the same address elect one to add the count
        const unsigned act = __activemask();
        const unsigned ga = __match_any_sync(act, ia), gb = __match_any_sync(act, ib);   // lanes may sit in different segments: match the full index
        if (lane_id() == (uint32_t)__ffs((int)ga) - 1u) atomicAdd(&A.pos_cov[ia], (int)__popc(ga));
        if (lane_id() == (uint32_t)__ffs((int)gb) - 1u) atomicAdd(&A.pos_cov[ib], -(int)__popc(gb));
    } else {
        atomicAdd(&A.pos_cov[ia], 1);
        atomicAdd(&A.pos_cov[ib], -1);
    }
    if (tc > 0) {
        int *tcbase = A.pos_tc + (long long)(posbase + g);   // may point before the segment; only in-range bits are used
        const bool whole = (pa == g && pb == e);
        asm volatile("" : "+l"(tcbase));   // keep it a pointer: each scatter address is then one IMAD.WIDE
        if (whole) {
            uint32_t more = 0u;
#pragma unroll
            for (int k = 0; k < W; k++) {   // first conversion of every word without a loop
                const uint32_t h = H[k];
                if (h) {
                    SD_CHECK((unsigned long long)(tcbase + (32 * k + __ffs((int)h) - 1) - A.pos_tc) < A.n_positions);
                    atomicAdd(tcbase + (32 * k + __ffs((int)h) - 1), 1);
                }
                more |= h & (h - 1u);
            }
            if (more) {
#pragma unroll
                for (int k = 0; k < W; k++) {
                    uint32_t h = H[k] & (H[k] - 1u);
                    while (h) {
                        SD_CHECK((unsigned long long)(tcbase + (32 * k + __ffs((int)h) - 1) - A.pos_tc) < A.n_positions);
                        atomicAdd(tcbase + (32 * k + __ffs((int)h) - 1), 1);
                        h &= h - 1u;
                    }
                }
            }
        } else {   // the segment holds only part of the read (rare): masked, plain loops
            for (int k = 0; k < W; k++) {
                uint32_t h = H[k] & prefix_word(k, (int)(pb - g)) & ~prefix_word(k, (int)(pa - g));
                while (h) {
                    SD_CHECK((unsigned long long)(tcbase + (32 * k + __ffs((int)h) - 1) - A.pos_tc) < A.n_positions);
                    atomicAdd(tcbase + (32 * k + __ffs((int)h) - 1), 1);
                    h &= h - 1u;
                }
            }
        }
    }
}

// T coverage and conversions of the read inside one UTR [us, ue)
template <int W>
__device__ __forceinline__ void utr_part(const CountArgs &A, const uint32_t (&T)[W], const uint32_t (&H)[W], uint32_t g, uint32_t e, int span,
                                         uint32_t us, uint32_t ue, uint32_t pop_t, int tc, uint32_t &cov_t, uint32_t &conv) {
    cov_t = pop_t;
    conv = (uint32_t)tc;
    if (us > g || ue < e) {  // the UTR covers only part of the read
        const int lo = us > g ? (int)(us - g) : 0;
        const int hi = ue < e ? (int)(ue - g) : span;
        cov_t = 0;
        conv = 0;
#pragma unroll
        for (int k = 0; k < W; k++) {
            const uint32_t m = prefix_word(k, hi) & ~prefix_word(k, lo);
            cov_t += __popc(T[k] & m);
            conv += __popc(H[k] & m);
        }
    } else if (A.mismatch_source == SD_MISMATCH_MP) {
        conv = 0;
#pragma unroll
        for (int k = 0; k < W; k++) conv += __popc(H[k]);
    }
}

// Fallback: reference window and annotation entries come from global memory, lane by lane (tiles without a span
// descriptor, or whose reads are spread over more of the genome than a stage holds: unsorted input).
template <int W, bool SIMPLE, int SB, int QB>
__device__ __forceinline__ int fast_read(const CountArgs &A, const ReadView &v, const Blocks &B, uint32_t mapq,
                                         bool dense_tile, bool &mp_disagree) {
    const int s = v.strand;
    const uint32_t g = v.g, e = v.g + (uint32_t)v.span;
    const uint32_t nU = A.n_u[s];
    const uint4 *ut = A.utab[s];
    // the interval-join entry point is fetched up front so that its latency overlaps the bit work below
    uint32_t i = __ldg(&A.bin_first[s][g >> SD_BIN_SHIFT]);
    // @region window
    // --- reference window: isX (T or A) and SNP planes, shifted so that bit j = position g + j
    uint32_t T[W], S[W];
    {
        const uint2 *rx = A.refx[s] + (v.g >> 5);
        const uint32_t sh = v.g & 31u;
        uint2 x[W + 1];
#pragma unroll
        for (int k = 0; k <= W; k++) x[k] = __ldg(rx + k);
#pragma unroll
        for (int k = 0; k < W; k++) {
            T[k] = __funnelshift_r(x[k].x, x[k + 1].x, sh) & prefix_word(k, v.span);
            S[k] = __funnelshift_r(x[k].y, x[k + 1].y, sh);
        }
    }
    uint32_t R[W], H[W], pop_t;
    const bool regs = SIMPLE || B.regs_ok;
    read_side<W, SIMPLE, SB, QB>(A, v, B, regs, R);
    uint4 u = make_uint4(0xFFFFFFFFu, 0u, 0u, 0u);   // first table entry {start, end, row, seg}, in flight during the quality test
    if (i < nU) u = __ldg(&ut[i]);
    int tc = reference_side<W, SIMPLE, SB, QB>(A, v, B, regs, R, T, S, H, pop_t, mp_disagree);
    // @region utr_join
    // --- interval join + per-UTR reduction + per-position scatter
    uint32_t last_seg = 0xFFFFFFFFu;
    while (i < nU) {
        if (u.x >= e) break;
        if (u.y > g) {
            uint32_t cov_t, conv;
            utr_part<W>(A, T, H, g, e, v.span, u.x, u.y, pop_t, tc, cov_t, conv);
            SD_CHECK(u.z < A.n_utr);
            utr_accumulate(A.counters, u.z, tc > 0 ? 1u : 0u, mapq == 0 ? 1u : 0u, cov_t, conv);
            if (u.w != last_seg) {
                last_seg = u.w;
                const uint4 sg = __ldg(&A.stab[s][u.w]);   // {start, end, off_lo, off_hi}
                const unsigned long long off = (unsigned long long)sg.z | ((unsigned long long)sg.w << 32);
                scatter_segment<W>(A, H, tc, g, e, sg.x, sg.y, off - sg.x, dense_tile);
            }
        }
        i++;
        if (i < nU) u = __ldg(&ut[i]);
    }
    return tc;
}

// What a stage holds besides the tile's columns (see the kernel): the reference window of the tile's reads, both strands
// side by side, and the annotation entries that can overlap them.
struct StageView {
    const uint4 *win;     // staged words: {isT, T>C SNPs, isA, A>G SNPs} per 32 linear positions
    const uint4 *ent;     // entry j: ent[2j] = {start, end, row | strand << 31, segment}, ent[2j + 1] = {segment start, end, posbase lo, hi}
    uint32_t d, n_ent;    // first window word of this lane's read among the staged ones; entries staged
};

// Staged tile: EVERY lane of the warp comes here (live = this lane has a read to count), so the loop over the tile's
// annotation entries is warp-uniform and a UTR's counters are one set of warp votes / reductions and at most five atomics.
template <int W, bool SIMPLE, int SB, int QB>
__device__ __forceinline__ int staged_read(const CountArgs &A, const ReadView &v, const Blocks &B, uint32_t mapq, bool dense_tile,
                                           bool live, const StageView &sv, bool &mp_disagree) {
    const int s = v.strand;
    const uint32_t g = v.g, e = v.g + (uint32_t)v.span;
    // @region window
    uint32_t T[W], S[W];
    {
        const uint2 *wp = reinterpret_cast<const uint2 *>(sv.win + sv.d) + s;   // this strand's half of each 16-byte word
        const uint32_t sh = v.g & 31u;
        uint2 x[W + 1];
#pragma unroll
        for (int k = 0; k <= W; k++) x[k] = wp[2 * k];
#pragma unroll
        for (int k = 0; k < W; k++) {
            T[k] = __funnelshift_r(x[k].x, x[k + 1].x, sh) & prefix_word(k, v.span);
            S[k] = __funnelshift_r(x[k].y, x[k + 1].y, sh);
        }
    }
    uint32_t R[W], H[W], pop_t;
    const bool regs = SIMPLE || B.regs_ok;
    read_side<W, SIMPLE, SB, QB>(A, v, B, regs, R);
    int tc = reference_side<W, SIMPLE, SB, QB>(A, v, B, regs, R, T, S, H, pop_t, mp_disagree);
    // @region utr_join
    uint32_t last_seg = 0xFFFFFFFFu;
    for (uint32_t j = 0; j < sv.n_ent; j++) {   // warp-uniform trip count
        const uint4 u = sv.ent[2 * j];          // broadcast
        const bool hit = live && (u.z >> 31) == (uint32_t)s && u.x < e && u.y > g;
        const unsigned hm = __ballot_sync(0xFFFFFFFFu, hit);
        if (hm == 0u) continue;
        uint32_t cov_t = 0u, conv = 0u;
        if (hit) utr_part<W>(A, T, H, g, e, v.span, u.x, u.y, pop_t, tc, cov_t, conv);
        const unsigned b_tc = __ballot_sync(0xFFFFFFFFu, hit && tc > 0), b_mm = __ballot_sync(0xFFFFFFFFu, hit && mapq == 0u);
        const uint32_t s_cov = __reduce_add_sync(0xFFFFFFFFu, cov_t), s_conv = __reduce_add_sync(0xFFFFFFFFu, conv);
        if (lane_id() == 0u) {
            const uint32_t row = u.z & 0x7FFFFFFFu;
            SD_CHECK(row < A.n_utr);
            unsigned long long *c = A.counters + (size_t)row * SD_NCOUNTER;
            atomicAdd(c + SD_C_READS, (unsigned long long)__popc(hm));
            if (b_tc) atomicAdd(c + SD_C_TCREADS, (unsigned long long)__popc(b_tc));
            if (b_mm) atomicAdd(c + SD_C_MULTIMAP, (unsigned long long)__popc(b_mm));
            if (s_cov) atomicAdd(c + SD_C_COV_ON_T, (unsigned long long)s_cov);
            if (s_conv) atomicAdd(c + SD_C_CONV_ON_T, (unsigned long long)s_conv);
        }
        // per-position outputs, once per (read, merged segment).  Still the whole warp: the reads of a sorted tile start and
        // end at few distinct positions in lane order, so equal addresses sit in runs of neighbouring lanes -- the first lane of
        // a run adds the run's length (one shuffle and two votes per address instead of a match over the warp)
        const bool part = hit && u.w != last_seg;
        const unsigned pm = __ballot_sync(0xFFFFFFFFu, part);
        if (pm == 0u) continue;
        if (part) last_seg = u.w;
        const uint4 sg = sv.ent[2 * j + 1];   // {segment start, end, posbase lo, hi}: position p lives at index posbase + p
        const unsigned long long posbase = (unsigned long long)sg.z | ((unsigned long long)sg.w << 32);
        const uint32_t pa = sg.x > g ? sg.x : g, pb = sg.y < e ? sg.y : e;
        const uint32_t lane = lane_id();
#pragma unroll
        for (int side = 0; side < 2; side++) {
            const uint32_t p = side ? pb : pa;   // all participants share the segment: equal positions <=> equal addresses
            const uint32_t prev = __shfl_up_sync(0xFFFFFFFFu, p, 1);
            const bool lead = part && !(lane > 0u && ((pm >> (lane - 1u)) & 1u) && prev == p);
            const unsigned lm = __ballot_sync(0xFFFFFFFFu, lead);
            if (lead) {
                const unsigned stop = (~pm | lm) & ~((2u << lane) - 1u);   // first lane after this one that is not in its run
                const int run = (stop ? __ffs((int)stop) - 1 : 32) - (int)lane;
                SD_CHECK(pb >= pa && posbase + p < A.n_positions);
                atomicAdd(&A.pos_cov[posbase + p], side ? -run : run);
            }
        }
        if (part && tc > 0) {
            int *tcbase = A.pos_tc + (long long)(posbase + g);   // may point before the segment; only in-range bits are used
            asm volatile("" : "+l"(tcbase));   // keep it a pointer: each scatter address is then one IMAD.WIDE
            if (pa == g && pb == e) {
#pragma unroll
                for (int k = 0; k < W; k++) {
                    uint32_t h = H[k];
                    while (h) {
                        SD_CHECK((unsigned long long)(tcbase + (32 * k + __ffs((int)h) - 1) - A.pos_tc) < A.n_positions);
                        atomicAdd(tcbase + (32 * k + __ffs((int)h) - 1), 1);
                        h &= h - 1u;
                    }
                }
            } else {   // the segment holds only part of the read (rare)
                for (int k = 0; k < W; k++) {
                    uint32_t h = H[k] & prefix_word(k, (int)(pb - g)) & ~prefix_word(k, (int)(pa - g));
                    while (h) {
                        SD_CHECK((unsigned long long)(tcbase + (32 * k + __ffs((int)h) - 1) - A.pos_tc) < A.n_positions);
                        atomicAdd(tcbase + (32 * k + __ffs((int)h) - 1), 1);
                        h &= h - 1u;
                    }
                }
            }
        }
    }
    return tc;
}

// ---------------------------------------------------------------------------------------
// The persistent streaming kernel.
//
// A tile is 32 consecutive reads = one warp's worth; a warp is a self-contained stream processor.
// Each warp owns a small ring of shared-memory stages and its own mbarriers: its elected lane
// issues the TMA bulk copies (records, sliced SEQ, CIGAR ops and, when needed, MP entries of one
// tile per stage) two tiles ahead, all lanes wait on the stage's mbarrier, each lane handles one
// read, and the lane re-arms the stage.  No warp ever waits for another warp, there is no block
// barrier in the loop and no producer thread spinning.  Base qualities come in one of two forms.
// Plane form (QB = 2 / 4, files with at most 4 / 16 distinct qualities): bit-planes of the dictionary
// code in the SEQ column's geometry, staged with the tile (8 / 16 bytes per 32 bases); the test
// `quality >= minQual` is a bit-sliced compare ANDed into the read mask.  Byte form (QB = 0): the column
// is NOT staged -- a quality is needed only where the read shows a C over a reference T (about one base
// per read), so it is fetched from global memory right there, which keeps the shared-memory footprint
// small and leaves most of the column's DRAM sectors untouched.
// ---------------------------------------------------------------------------------------
// @region kernel_setup
#define SD_SMEM_STATS 8
#ifndef SD_DENSE_SPAN
#define SD_DENSE_SPAN 8   // <= 8 start positions for 32 reads: >= 4 reads per position
#endif
#ifndef SD_CLAIM
#define SD_CLAIM 8   // consecutive tiles claimed per atomic
#endif
#ifndef SD_CTA_THREADS
#define SD_CTA_THREADS 256
#endif
#define SD_WARP_TRAILER 512u   // per-warp bookkeeping behind the stages (at most two stages)
#ifndef SD_MIN_CTAS
#define SD_MIN_CTAS(W) ((W) <= 6 ? 4 : ((W) <= 10 ? 3 : 2))   // resident 256-thread CTAs per SM the register budget is tuned for
#endif

template <int W, int SB, int QB, bool STAGED>
__global__ void __launch_bounds__(SD_CTA_THREADS, SD_MIN_CTAS(W)) sd_count_kernel(const __grid_constant__ CountArgs A) {
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t stage_bytes = A.cap_rec + A.cap_seq + A.cap_cig + A.cap_mp + A.cap_qual + A.cap_win + A.cap_ent;
    const uint32_t warp_bytes = A.n_stages * stage_bytes + SD_WARP_TRAILER;   // stages | barriers + tile qual offsets | per-lane copy descriptors | span info
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t nwarp = blockDim.x >> 5;
    uint32_t *s_stats = reinterpret_cast<uint32_t *>(smem);                 // [SD_SMEM_STATS], first 128 bytes
    unsigned char *mine = smem + 128u + (size_t)warp * warp_bytes;
    unsigned char *stages = mine;
    uint64_t *full = reinterpret_cast<uint64_t *>(mine + (size_t)A.n_stages * stage_bytes);   // [n_stages]

    if (tid < SD_SMEM_STATS) s_stats[tid] = 0u;
    if (lane == 0) {
        for (uint32_t s = 0; s < A.n_stages; s++) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    // Tiles are handed out dynamically in chunks of SD_CLAIM consecutive tiles (one atomic per chunk): warps that hit
    // expensive tiles simply claim fewer, so the grid drains evenly, and a warp's consecutive tiles share reference
    // windows and table rows in L1.
    uint32_t *s_tile = reinterpret_cast<uint32_t *>(mine + (size_t)A.n_stages * stage_bytes + 16u);   // [n_stages] tile held by each stage
    uint32_t claim_next = 0, claim_end = 0;   // lane 0 only
    const uint32_t n_queue = A.tile_list ? *A.n_list : A.n_tiles;
    auto claim = [&]() -> uint32_t {          // next tile for this warp, or 0xFFFFFFFF when the queue is empty
        if (claim_next == claim_end) {
            const unsigned int chunk = (unsigned int)SD_CLAIM;
            claim_next = atomicAdd(A.tile_counter, chunk);
            claim_end = claim_next + chunk;
        }
        const uint32_t t = claim_next++;
        if (A.tile_list) return t < n_queue ? A.tile_list[t] : 0xFFFFFFFFu;
        return t < n_queue ? t : 0xFFFFFFFFu;
    };
    // @region tma_issue
    // A tile is one bulk copy per column, issued by lanes 0..6 side by side -- lane c owns column c: records, SEQ, CIGAR, MP,
    // quality planes and, in the staged kernel, the tile's reference window and its annotation entries.  Every lane runs the
    // same few instructions: two 64-bit loads from a per-tile table (the batch's tile table, or its span table) give the
    // column's [begin, end) in bytes.  The per-lane constants live in the warp's shared-memory region, and so does what the
    // lanes found out about the tile (s_meta), which the consumers read after the wait.
    uint4 *s_col = reinterpret_cast<uint4 *>(mine + (size_t)A.n_stages * stage_bytes + 64u);     // [7][2] {src lo, src hi, table lo, table hi}, {stage offset, end - begin words apart, extra bytes, most bytes}
    uint4 *s_meta = reinterpret_cast<uint4 *>(mine + (size_t)A.n_stages * stage_bytes + 288u);   // [n_stages][7] {begin lo, begin hi, bytes, copied}
    if (lane < 7u) {
        const unsigned char *src = reinterpret_cast<const unsigned char *>(A.rec);
        const unsigned long long *tbl = reinterpret_cast<const unsigned long long *>(A.tiles);   // begin of tile t at tbl[4 t], end `apart` words on
        uint32_t dst = 0u, apart = 4u, extra = 0u, most = 0xFFFFFFFFu;
        bool col = true;   // does this lane's column travel with the tile at all
        if (lane == 1) { src = reinterpret_cast<const unsigned char *>(A.seq); dst = A.cap_rec; }
        if (lane == 2) { src = reinterpret_cast<const unsigned char *>(A.cigar); dst = A.cap_rec + A.cap_seq; tbl += 2; }
        if (lane == 3) { src = reinterpret_cast<const unsigned char *>(A.mp); dst = A.cap_rec + A.cap_seq + A.cap_cig; tbl += 3; col = A.mp != nullptr; }
        if (lane == 4) { src = A.qual; dst = A.cap_rec + A.cap_seq + A.cap_cig + A.cap_mp; tbl += 1; col = QB > 0; }
        if (lane >= 5) {
            col = STAGED;
            tbl = reinterpret_cast<const unsigned long long *>(STAGED ? (const void *)A.spans : (const void *)A.tiles);
            apart = 1u;
        }
        if (lane == 5) { src = reinterpret_cast<const unsigned char *>(A.refm); dst = A.cap_rec + A.cap_seq + A.cap_cig + A.cap_mp + A.cap_qual; extra = 16u * W; most = A.cap_win; }
        if (lane == 6) { src = reinterpret_cast<const unsigned char *>(A.ment); dst = A.cap_rec + A.cap_seq + A.cap_cig + A.cap_mp + A.cap_qual + A.cap_win; tbl += 2; most = A.cap_ent; }
        if (!col) most = 0u;
        const unsigned long long a = reinterpret_cast<unsigned long long>(src), t = reinterpret_cast<unsigned long long>(tbl);
        s_col[2 * lane] = make_uint4((uint32_t)a, (uint32_t)(a >> 32), (uint32_t)t, (uint32_t)(t >> 32));
        s_col[2 * lane + 1] = make_uint4(dst, apart, extra, most);
    }
    __syncwarp();
    auto issue = [&](uint64_t tile, uint32_t stage) {   // lanes 0..6 together
        unsigned char *base = stages + (size_t)stage * stage_bytes;
        const uint4 c0 = s_col[2 * lane], c1 = s_col[2 * lane + 1];
        const unsigned char *my_src = reinterpret_cast<const unsigned char *>((unsigned long long)c0.x | ((unsigned long long)c0.y << 32));
        const unsigned long long *tp = reinterpret_cast<const unsigned long long *>((unsigned long long)c0.z | ((unsigned long long)c0.w << 32)) + ((STAGED && lane >= 5u) ? 8ull : 4ull) * tile;   // sd_tile is 4 words, sd_tile_span 8
        unsigned long long off0 = __ldg(tp), off1 = __ldg(tp + c1.y);
        if (lane == 0) {
            off0 = tile * (SD_TILE_READS * (unsigned long long)sizeof(sd_read_rec));
            off1 = off0 + SD_TILE_READS * (unsigned long long)sizeof(sd_read_rec);
        }
        // an empty column stays empty; the reference window gets the W words a read starting in its last word still loads
        unsigned long long want = off1 > off0 ? off1 - off0 + c1.z : 0ull;
        // the span table marks a tile whose reads are spread too widely with an impossible size; then neither of the two
        // span-driven copies is made and the consumers see `copied` = 0 for them
        const unsigned fit = __ballot_sync(0x7Fu, want <= (unsigned long long)c1.w);
        const bool copied = (fit >> lane) & 1u & ((lane < 5u) | ((fit >> 5) & (fit >> 6) & 1u));
        const uint32_t bytes = copied ? (uint32_t)want : 0u;
        s_meta[7u * stage + lane] = make_uint4((uint32_t)off0, (uint32_t)(off0 >> 32), bytes, copied ? 1u : 0u);
        if (QB == 0 && A.qual_prefetch && lane == 4) {   // byte form: pull the tile's base qualities into L2 ahead of use
            const uint32_t b_q = (uint32_t)(off1 - off0) & ~15u;
            if (b_q) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(A.qual + off0), "r"(b_q) : "memory");
        }
        const uint32_t total = __reduce_add_sync(0x7Fu, bytes);
        if (lane == 0) mbar_arrive_expect_tx(&full[stage], total);
        __syncwarp(0x7Fu);
        if (bytes) bulk_g2s(base + c1.x, my_src + off0, bytes, &full[stage]);
    };

    for (uint32_t s = 0; s < A.n_stages; s++) {
        uint32_t t = 0u;
        if (lane == 0) {
            t = claim();
            s_tile[s] = t;
        }
        t = __shfl_sync(0xFFFFFFFFu, t, 0);
        if (lane < 7u && t != 0xFFFFFFFFu) issue(t, s);
    }
    __syncwarp();

    // @region tile_wait_rec
    uint32_t my_tc_total = 0;   // per-lane run total (a lane sees far fewer than 2^32 conversions per launch)
    uint32_t cnt_a = 0, cnt_b = 0, cnt_tiles = 0;   // packed per-lane run counters, see filter_step
    auto flush_counts = [&]() {
        const uint32_t f[6] = {cnt_a & 1023u, (cnt_a >> 10) & 1023u, cnt_a >> 20, cnt_b & 1023u, (cnt_b >> 10) & 1023u, cnt_b >> 20};
        const int slot[6] = {SD_S_MAPPED, SD_S_FILTERED, SD_S_COUNTED, SD_S_UNMAPPED, SD_S_MQFILTERED, SD_S_IDFILTERED};
#pragma unroll
        for (int k = 0; k < 6; k++) {
            const uint32_t t = __reduce_add_sync(0xFFFFFFFFu, f[k]);
            if (lane == 0 && t) atomicAdd(&s_stats[slot[k]], t);
        }
        cnt_a = cnt_b = cnt_tiles = 0u;
    };
    uint32_t stage = 0, phase = 0;
    for (;;) {
        const uint64_t tile = s_tile[stage];
        if (tile == 0xFFFFFFFFull) break;   // the queue ran dry (claims are monotonic: later stages are empty too)
        mbar_wait(&full[stage], phase);
        const uint4 m_win = STAGED ? s_meta[7u * stage + 5u] : make_uint4(0u, 0u, 0u, 0u);   // {begin lo, begin hi, bytes, copied}
        if (STAGED && m_win.w == 0u) {
            // The tile's reads are spread over more of the genome than a stage holds (unsorted input, a tile that crosses
            // chromosomes): leave it to the plain kernel that follows this launch.
            if (lane == 0) A.leftover[atomicAdd(A.n_leftover, 1u)] = (uint32_t)tile;
        } else {
        const unsigned char *base = stages + (size_t)stage * stage_bytes;
        const uint32_t *rw = reinterpret_cast<const uint32_t *>(base) + 5u * lane;
        const int32_t pos = (int32_t)rw[0];
        const uint32_t rmf = rw[1];
        const float xi = __uint_as_float(rw[2]);
        const uint32_t lseq = rw[3] & 0xFFFFu, ncig = rw[3] >> 16;
        const uint32_t nm = rw[4] & 0xFFFFu, nmp = A.mp ? (rw[4] >> 16) : 0u;

#ifndef SD_NO_DESC_L1PF
        // lane 0 will read the next tile's descriptors when it refills the stage: start that (DRAM-latency) load now
        if (lane == 0 && !A.tile_list && claim_next != claim_end && claim_next < A.n_tiles) {
            asm volatile("prefetch.global.L1 [%0];" ::"l"(A.tiles + claim_next));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(A.tiles + claim_next + 1));
        }
#endif
        // @region row_scan
        // row offsets inside the tile: exclusive warp scan of the row sizes
        uint32_t xs = ncig, qs = lseq, ms = nmp;
        int uni_pred = 0;
        __match_all_sync(0xFFFFFFFFu, lseq | (ncig << 16), &uni_pred);
        if (uni_pred && !A.mp) {   // every row of the tile has the same shape: offsets are lane multiples
            xs = (lane + 1u) * ncig;
            qs = (lane + 1u) * lseq;
        } else
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t tx = __shfl_up_sync(0xFFFFFFFFu, xs, d);
            const uint32_t tq = __shfl_up_sync(0xFFFFFFFFu, qs, d);
            const uint32_t tm = A.mp ? __shfl_up_sync(0xFFFFFFFFu, ms, d) : 0u;
            if (lane >= (uint32_t)d) {
                xs += tx;
                qs += tq;
                ms += tm;
            }
        }
        const uint32_t cig_row = xs - ncig;
        qs -= lseq;
        ms -= nmp;

        // @region filter
        const uint32_t flags = rmf >> 24, mapq = (rmf >> 16) & 0xFFu, ref = rmf & 0xFFFFu;
        // filter.py:235-256 on a record that is not tile padding -> does it go on to be counted.  The run counters are kept per
        // lane as 10-bit fields of two registers (cnt_a: mapped | kept | counted, cnt_b: unmapped | MQ | identity) and added up
        // by flush_counts() before a field can overflow -- a handful of integer instructions per read instead of six warp
        // votes and as many shared-memory atomics per tile.
        auto filter_step = [&](bool go) -> bool {
            if (A.filter_on) {
                const bool primary = go && !(flags & (SD_F_SECONDARY | SD_F_SUPPLEMENTARY));
                const bool mapped = go && !(flags & SD_F_UNMAPPED);
                const bool f_mq = mapped && (int)mapq < A.filter_mq;
                const bool f_id = mapped && !f_mq && xi < A.filter_min_identity;
                const bool f_nm = mapped && !f_mq && !f_id && A.filter_max_nm > -1 && (int)nm > A.filter_max_nm;
                go = mapped && !f_mq && !f_id && !f_nm;
                cnt_a += (uint32_t)(primary && mapped) | ((uint32_t)(go && primary) << 10);
                cnt_b += (uint32_t)(primary && !mapped) | ((uint32_t)f_mq << 10) | ((uint32_t)f_id << 20);
                if (f_nm) atomicAdd(&s_stats[SD_S_NMFILTERED], 1u);   // off by default
            }
            return go;
        };
        // A tile without a CIGAR block: every record with one op is one M block of l_seq bases (sd_batch: plain-match tiles
        // need not carry the column at all)
        const bool nocig = s_meta[7u * stage + 2u].z == 0u;
        // could this record be counted at all, whatever the filter says
        const bool cand = !(flags & (SD_F_PAD | SD_F_UNMAPPED | SD_F_SKIP)) && ref < A.n_chrom && !(nocig && ncig > 1u);
        const uint32_t cig0 = nocig ? (lseq << 4) : (ncig ? reinterpret_cast<const uint32_t *>(base + A.cap_rec + A.cap_seq)[cig_row] : 0u);
        // Deep coverage: a coordinate-sorted tile whose 32 reads start within a few positions would send dozens of atomics to
        // the same addresses; such tiles (warp-uniform test on the tile's first and last record) aggregate them in the warp.
        const bool dense_tile = !STAGED && __shfl_sync(0xFFFFFFFFu, rmf & 0xFFFFu, 0) == __shfl_sync(0xFFFFFFFFu, rmf & 0xFFFFu, 31) &&
                                (uint32_t)(__shfl_sync(0xFFFFFFFFu, pos, 31) - __shfl_sync(0xFFFFFFFFu, pos, 0)) <= SD_DENSE_SPAN;
        int tc = 0;
        bool bad = false;
        if (STAGED) {
            // Is every read of this tile a single match block (no clips, no indels)?  Warp-uniform: the decoder and
            // the generator group such reads into their own tiles, so most warps take the lean path.
            const bool uni_simple = __all_sync(0xFFFFFFFFu, !cand || (ncig == 1u && ((0x181u >> (cig0 & 15u)) & 1u)));
            // @region view_setup
            // Every lane builds its view; a lane without a countable read gets an empty one (no bases, no span), so that the
            // whole warp runs through the same code and a UTR's counters are warp-wide votes.
            const uint32_t refc = cand ? ref : 0u;
            const uint32_t clen = __ldg(&A.chrom_len[refc]);
            const bool placed = cand && pos >= 0 && (uint32_t)pos < clen;
            ReadView v;
            v.seq = reinterpret_cast<const uint32_t *>(base + A.cap_rec) + (SB == 2 ? 2u : 4u) * lane;
            v.qual = QB > 0 ? base + A.cap_rec + A.cap_seq + A.cap_cig + A.cap_mp + 4u * QB * lane : A.qual + ((unsigned long long)s_meta[7u * stage + 4u].x | ((unsigned long long)s_meta[7u * stage + 4u].y << 32)) + qs;
            v.cig = reinterpret_cast<const uint32_t *>(base + A.cap_rec + A.cap_seq) + cig_row;
            v.mp = A.mp ? reinterpret_cast<const uint32_t *>(base + A.cap_rec + A.cap_seq + A.cap_cig) + ms : nullptr;
            v.ncig = placed ? ncig : 0u;
            v.g = __ldg(&A.chrom_off[refc]) + (placed ? (uint32_t)pos : 0u);
            v.strand = (flags & SD_F_REVERSE) ? 1 : 0;
            v.has_mp = (flags & SD_F_HAS_MP) != 0;
            Blocks B;
            uint32_t reflen;
            if (uni_simple) {   // one match block starting at window position 0
                B.n = 1;
                B.r0a = 0;
                B.r1a = (int)(cig0 >> 4);
                B.r0b = B.r1b = B.r0c = B.r1c = 0;
                B.sha = B.shb = B.shc = 0;
                B.regs_ok = true;
                reflen = cig0 >> 4;
            } else reflen = (uint32_t)parse_cigar(v.cig, v.ncig, B);
            if (reflen == 0) reflen = 1;  // htslib bam_endpos
            if (placed && (uint64_t)pos + reflen > clen) reflen = clen - (uint32_t)pos;
            bool fits = !A.force_slow && reflen <= 32u * W && lseq <= 32u * W;
            if (placed && fits && A.mismatch_source == SD_MISMATCH_MP && v.mp) {
                for (uint32_t i = 0; i < nmp; i++)
                    if ((v.mp[i] >> 16) >= 32u * W) fits = false;
            }
            // First window word of this read among the staged ones.  The span table is made for this batch (sd_tile_spans), so
            // the read lies inside; a table that lies gets a wrong count but no wild access (or the tile is left over).
            const uint32_t w0 = (uint32_t)(((unsigned long long)m_win.x | ((unsigned long long)m_win.y << 32)) >> 4);
            const uint32_t dword = (v.g >> 5) - w0;
            // A read that needs the per-base path (wider than the register window) sends the whole tile to the plain kernel --
            // before anything was counted.
            if (__any_sync(0xFFFFFFFFu, placed && !(fits && dword < (m_win.z >> 4) && dword + (uint32_t)W < (m_win.z >> 4)))) {
                if (lane == 0) A.leftover[atomicAdd(A.n_leftover, 1u)] = (uint32_t)tile;
            } else {
                const bool live = filter_step(!(flags & SD_F_PAD)) && placed;
                cnt_a += (uint32_t)live << 20;
                v.lseq = live ? lseq : 0u;
                v.nmp = live ? nmp : 0u;
                v.span = live ? (int)reflen : 0;
                StageView sv;
                sv.win = reinterpret_cast<const uint4 *>(base + A.cap_rec + A.cap_seq + A.cap_cig + A.cap_mp + A.cap_qual);
                sv.ent = reinterpret_cast<const uint4 *>(base + A.cap_rec + A.cap_seq + A.cap_cig + A.cap_mp + A.cap_qual + A.cap_win);
                sv.d = live ? dword : 0u;
                sv.n_ent = s_meta[7u * stage + 6u].z >> 5;
                tc = uni_simple ? staged_read<W, true, SB, QB>(A, v, B, mapq, dense_tile, live, sv, bad)
                                : staged_read<W, false, SB, QB>(A, v, B, mapq, dense_tile, live, sv, bad);
                if (!live) tc = 0;
            }
        } else {
            const bool go = filter_step(!(flags & SD_F_PAD)) && cand;
            const bool uni_simple = __all_sync(0xFFFFFFFFu, !go || (ncig == 1u && ((0x181u >> (cig0 & 15u)) & 1u)));
            // @region view_setup
            if (go) {
                const uint32_t clen = __ldg(&A.chrom_len[ref]);
                if (pos >= 0 && (uint32_t)pos < clen) {
                    ReadView v;
                    v.seq = reinterpret_cast<const uint32_t *>(base + A.cap_rec) + (SB == 2 ? 2u : 4u) * lane;
                    v.qual = QB > 0 ? base + A.cap_rec + A.cap_seq + A.cap_cig + A.cap_mp + 4u * QB * lane : A.qual + ((unsigned long long)s_meta[7u * stage + 4u].x | ((unsigned long long)s_meta[7u * stage + 4u].y << 32)) + qs;
                    v.cig = reinterpret_cast<const uint32_t *>(base + A.cap_rec + A.cap_seq) + cig_row;
                    v.mp = A.mp ? reinterpret_cast<const uint32_t *>(base + A.cap_rec + A.cap_seq + A.cap_cig) + ms : nullptr;
                    v.lseq = lseq;
                    v.ncig = ncig;
                    v.nmp = nmp;
                    v.g = __ldg(&A.chrom_off[ref]) + (uint32_t)pos;
                    v.strand = (flags & SD_F_REVERSE) ? 1 : 0;
                    v.has_mp = (flags & SD_F_HAS_MP) != 0;
                    Blocks B;
                    uint32_t reflen;
                    if (uni_simple) {   // one match block starting at window position 0
                        B.n = 1;
                        B.r0a = 0;
                        B.r1a = (int)(cig0 >> 4);
                        B.r0b = B.r1b = B.r0c = B.r1c = 0;
                        B.sha = B.shb = B.shc = 0;
                        B.regs_ok = true;
                        reflen = cig0 >> 4;
                    } else reflen = (uint32_t)parse_cigar(v.cig, ncig, B);
                    if (reflen == 0) reflen = 1;  // htslib bam_endpos
                    if ((uint64_t)pos + reflen > clen) reflen = clen - (uint32_t)pos;
                    v.span = (int)reflen;
                    bool fits = !A.force_slow && reflen <= 32u * W && lseq <= 32u * W;
                    if (fits && A.mismatch_source == SD_MISMATCH_MP && v.mp) {
                        for (uint32_t i = 0; i < nmp; i++)
                            if ((v.mp[i] >> 16) >= 32u * W) fits = false;
                    }
                    cnt_a += 1u << 20;
                    if (fits) {
                        tc = uni_simple ? fast_read<W, true, SB, QB>(A, v, B, mapq, dense_tile, bad) : fast_read<W, false, SB, QB>(A, v, B, mapq, dense_tile, bad);
                    } else {
                        uint32_t cigw = lseq << 4;   // a tile without a CIGAR block: the per-base path wants the op spelled out
                        if (nocig) v.cig = &cigw;
                        atomicAdd(&A.stats[SD_S_SLOWPATH], 1ull);
                        tc = slow_read(A, v, mapq);
                    }
                }
            }
        }
        if (bad) atomicAdd(&A.stats[SD_S_MP_DISAGREE], 1ull);   // rare: straight to the global counter
        my_tc_total += (uint32_t)tc;
        if (++cnt_tiles == 1000u) flush_counts();
        if (A.read_tc) A.read_tc[tile * SD_TILE_READS + lane] = (uint8_t)(tc > 255 ? 255 : tc);
        }

        // @region tile_end
        __syncwarp();   // every lane is done with this stage
        uint32_t next = 0u;
        if (lane == 0) {
            next = claim();
            s_tile[stage] = next;
        }
        next = __shfl_sync(0xFFFFFFFFu, next, 0);
        if (lane < 7u && next != 0xFFFFFFFFu) issue(next, stage);
        __syncwarp();
        if (++stage == A.n_stages) {
            stage = 0;
            phase ^= 1u;
        }
    }

    // @region kernel_tail
    // flush run-level counters
    flush_counts();
    unsigned long long tc_total = my_tc_total;
    for (int d = 16; d > 0; d >>= 1) tc_total += __shfl_xor_sync(0xFFFFFFFFu, tc_total, d);
    if (lane == 0 && tc_total) atomicAdd(&A.stats[SD_S_TC_TOTAL], tc_total);
    __syncthreads();
    if (tid < 7 && s_stats[tid]) atomicAdd(&A.stats[tid], (unsigned long long)s_stats[tid]);
}


// ---------------------------------------------------------------------------------------
// launcher
// ---------------------------------------------------------------------------------------
inline uint32_t align128(uint32_t x) { return (x + 127u) & ~127u; }
inline uint32_t align16(uint32_t x) { return (x + 15u) & ~15u; }

template <int W, int SB, int QB, bool STAGED>
static int launch_one(sd_ctx *ctx, CountArgs &A, cudaStream_t stream) {
    const uint32_t stage_bytes = A.cap_rec + A.cap_seq + A.cap_cig + A.cap_mp + A.cap_qual + A.cap_win + A.cap_ent;
    int max_smem = 0;
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device);
    static int forced_stages = -1, forced_warps = -1, forced_pf = -1;
    if (forced_stages < 0) {
        const char *e = getenv("SD_STAGES");
        forced_stages = e ? atoi(e) : 0;
        e = getenv("SD_QUAL_PF");
        forced_pf = e ? atoi(e) : -1;
        e = getenv("SD_WARPS");
        forced_warps = e ? atoi(e) : 0;
    }
    // Per-warp ring of n_stages tiles.  Two stages hide the copy latency behind the previous tile, but shared memory is taken
    // from the same 228 KB as the L1 that serves the reference window / annotation table lookups, and it can cost a resident
    // CTA: the second stage is used only when it neither lowers the CTAs per SM nor leaves L1 less than 56 KB (24 KB for the
    // staged kernel, whose reference window and annotation entries arrive with the tile instead of through L1).
    auto warp_bytes = [&](uint32_t ns) { return (size_t)ns * stage_bytes + SD_WARP_TRAILER; };
    if (128u + warp_bytes(1) > (size_t)max_smem)
        return sd_fail(ctx, SD_ERR_LIMIT, "a 32-read tile needs %u bytes of shared memory; reads are too long for this kernel", stage_bytes);
    auto plan = [&](uint32_t ns, uint32_t &warps, size_t &smem, int &per_sm) -> int {
        warps = (forced_warps >= 1 && forced_warps <= SD_CTA_THREADS / 32) ? (uint32_t)forced_warps : (uint32_t)(SD_CTA_THREADS / 32);
        while (warps > 1 && 128u + warps * warp_bytes(ns) > (size_t)max_smem / 2) warps--;   // keep >= 2 CTAs per SM possible
        while (warps > 1 && 128u + warps * warp_bytes(ns) > (size_t)max_smem) warps--;
        smem = 128u + warps * warp_bytes(ns);
        if (smem > (size_t)max_smem) { per_sm = 0; return SD_OK; }
        SD_CUDA(ctx, cudaFuncSetAttribute(sd_count_kernel<W, SB, QB, STAGED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SD_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sd_count_kernel<W, SB, QB, STAGED>, (int)warps * 32, smem));
        return SD_OK;
    };
    uint32_t n_stages = 1, warps = 0;
    size_t smem = 0;
    int per_sm = 0, rc = SD_OK;
    struct Planned { uint32_t stage_bytes; int device; uint32_t n_stages, warps; size_t smem; int per_sm; };
    static Planned last = {0, -1, 0, 0, 0, 0};   // the shape rarely changes between launches: skip the occupancy queries
    if (last.device == ctx->device && last.stage_bytes == stage_bytes) {
        n_stages = last.n_stages; warps = last.warps; smem = last.smem; per_sm = last.per_sm;
    } else if ((rc = plan(1, warps, smem, per_sm)) != SD_OK) {
        return rc;
    } else if (forced_stages >= 1 && forced_stages <= 2) {
        n_stages = (uint32_t)forced_stages;
        if ((rc = plan(n_stages, warps, smem, per_sm)) != SD_OK) return rc;
    } else {
        uint32_t w2 = 0;
        size_t s2 = 0;
        int p2 = 0;
        if ((rc = plan(2, w2, s2, p2)) != SD_OK) return rc;
        int sm_total = 0;
        cudaDeviceGetAttribute(&sm_total, cudaDevAttrMaxSharedMemoryPerMultiprocessor, ctx->device);
        const size_t l1_keep = (STAGED ? 24u : 56u) * 1024u;
        if (p2 >= 1 && (uint64_t)p2 * w2 >= (uint64_t)per_sm * warps && (size_t)p2 * s2 + l1_keep <= (size_t)sm_total) {
            n_stages = 2; warps = w2; smem = s2; per_sm = p2;
        }
    }
    if (per_sm < 1) return sd_fail(ctx, SD_ERR_LIMIT, "count kernel does not fit on an SM (smem %zu)", smem);
    if (last.device != ctx->device || last.stage_bytes != stage_bytes) {
        SD_CUDA(ctx, cudaFuncSetAttribute(sd_count_kernel<W, SB, QB, STAGED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        last = {stage_bytes, ctx->device, n_stages, warps, smem, per_sm};
    }
    A.n_stages = n_stages;
    A.qual_prefetch = forced_pf >= 0 ? (uint32_t)forced_pf : 0u;   // off: on-demand quality bytes keep DRAM traffic at the touched sectors
    const int threads = (int)warps * 32;
    const uint64_t want_ctas = ((uint64_t)A.n_tiles + warps - 1) / warps;
    uint32_t grid = (uint32_t)std::min<uint64_t>(want_ctas, (uint64_t)ctx->num_sms * per_sm);
    if (grid == 0) grid = 1;
    sd_count_kernel<W, SB, QB, STAGED><<<grid, threads, smem, stream>>>(A);
    ctx->launches++;
    SD_CUDA(ctx, cudaGetLastError());
    return SD_OK;
}

// Up to three launches count a batch that carries span descriptors: the LEAN kernel (sd_count_lean.cuh) takes the tiles whose reads
// are all alike and appends the others to a list; the STAGED kernel counts that list and appends the tiles whose reads are spread
// too widely for a stage to a second list; the plain kernel counts that one with per-read lookups.
// A.tile_counter points at eight zeroed words: the three queue heads (0..2) and the lengths of the two lists (3, 4).
template <int W, int QB, bool VERIFY>
int launch_lean(sd_ctx *ctx, const CountArgs &A, cudaStream_t stream, uint32_t *list_out, unsigned int *n_out);

template <int W, int SB, int QB>
int launch_count(sd_ctx *ctx, CountArgs &A, cudaStream_t stream, bool timed) {
    if (timed) SD_CUDA(ctx, cudaEventRecord(ctx->ev_k0, stream));
    unsigned int *words = A.tile_counter;
    int rc = SD_OK;
    int marks = 0;
    if (A.spans) {
        static int lean_env = -1;
        if (lean_env < 0) { const char *e = getenv("SD_LEAN"); lean_env = e ? atoi(e) : 1; }
        uint32_t *list1 = A.leftover, *list2 = A.leftover + A.n_tiles;
        // a launch that would take next to nothing (unsorted input: no tile fits a stage, none is lean) is not made; its tiles go on
        const uint32_t few = A.n_tiles / 64u;
        const bool skip_lean = A.hint_lean != 0xFFFFFFFFu && A.hint_lean <= few;
        const bool skip_staged = A.hint_staged != 0xFFFFFFFFu && A.hint_staged <= few && (skip_lean || A.hint_lean >= A.hint_staged);
        if constexpr (SB == 2 && QB > 0) if (lean_env && !skip_lean && !A.force_slow && (A.mismatch_source == SD_MISMATCH_CIGAR || A.mismatch_source == SD_MISMATCH_VERIFY)) {
            if (A.mismatch_source == SD_MISMATCH_CIGAR) rc = launch_lean<W, (QB > 0 ? QB : 2), false>(ctx, A, stream, list1, words + 3);
            else rc = launch_lean<W, (QB > 0 ? QB : 2), true>(ctx, A, stream, list1, words + 3);
            if (rc != SD_OK) return rc;
            A.tile_list = list1;
            A.n_list = words + 3;
            if (timed) { SD_CUDA(ctx, cudaEventRecord(ctx->ev_split[0], stream)); marks = 2; }
        }
        if (!skip_staged) {
            A.tile_counter = words + 1;
            A.leftover = list2;
            A.n_leftover = words + 4;
            if ((rc = launch_one<W, SB, QB, true>(ctx, A, stream)) != SD_OK) return rc;
            A.tile_list = list2;
            A.n_list = words + 4;
        }
        if (timed) { SD_CUDA(ctx, cudaEventRecord(ctx->ev_split[1], stream)); if (!marks) marks = 1; }
        A.tile_counter = words + 2;
        A.spans = nullptr;
        A.cap_win = A.cap_ent = 0;
    }
    if ((rc = launch_one<W, SB, QB, false>(ctx, A, stream)) != SD_OK) return rc;
    if (timed) {
        SD_CUDA(ctx, cudaEventRecord(ctx->ev_k1, stream));
        ctx->have_kernel_time = true;
        ctx->split_marks = marks;
    }
    return SD_OK;
}
