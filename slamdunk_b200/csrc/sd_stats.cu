// sd_stats.cu -- the conversion statistics that share the count path's per-read scan:
//   alleyoop rates         statsComputeOverallRates         dunks/stats.py:85-131
//   alleyoop tcperreadpos  tcPerReadPos                     dunks/stats.py:591-657
//   alleyoop utrrates      statsComputeOverallRatesPerUTR   dunks/stats.py:331-528
//   alleyoop tcperutrpos   tcPerUtr                         dunks/stats.py:659-772
//   alleyoop snpeval       computeSNPMaskedRates            dunks/stats.py:774-849
//   count --mle lists      computeTconversions(mle=True)    dunks/tcounter.py:233-243
// All of them iterate SlamSeqBamIterator (slamseq/SlamSeqFile.py:364-402) and use the read's complete mismatch list
// (fillMismatchesNGM, :313-348: every MP entry whose base quality passes) and its 5 x 5 conversion matrix
// (computeRatesForRead, :231-249).  One thread per read, one 32-read tile per warp; the per-UTR quantities go through
// the same bin table + sorted interval table as the count kernel.  Integer work only; HBM-bound on the columns it
// reads (records, SEQ planes, CIGAR, MP entries, the quality bytes of the mismatching bases).
#include "sd_internal.h"

namespace {

struct StatsArgs {
    const sd_read_rec *rec;
    const uint32_t *seq;
    const uint8_t *qual;
    const uint32_t *cigar;
    const uint32_t *mp;
    const sd_tile *tiles;
    uint32_t n_tiles;
    const uint32_t *order;
    const uint2 *refx[2];
    const uint32_t *chrom_off;
    uint64_t n_bits;            // linear genome positions covered by refx
    const uint4 *utab[2];
    const uint32_t *bin_first[2];
    uint32_t n_u[2];
    const uint32_t *row_blen;   // BED stop - start per row (utr.getLength(), not clipped to the chromosome)
    const uint8_t *row_strand;  // 0 '+', 1 '-', 2 none
    int min_qual, conv_threshold, strict;
    uint32_t max_read_length, utr_norm;
    uint32_t hist_smem;         // the read-position and UTR-position histograms are accumulated per CTA in shared memory
    unsigned long long *rates, *readpos, *utr_rates, *utr_snp, *utrpos, *context;
    sd_read_pair *pairs;
    unsigned long long pair_capacity;
    unsigned long long *n_pairs;
};

__device__ __forceinline__ uint32_t warp_excl_scan(uint32_t v, uint32_t lane) {
    uint32_t x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, d);
        if ((int)lane >= d) x += y;
    }
    return x - v;
}

// bits [lo, hi) of the 32-base group g
__device__ __forceinline__ uint32_t group_range(int g, int lo, int hi) {
    const int a = max(lo - 32 * g, 0), b = min(hi - 32 * g, 32);
    if (b <= a) return 0u;
    const uint32_t upto_b = b >= 32 ? 0xFFFFFFFFu : ((1u << b) - 1u);
    return upto_b & ~((1u << a) - 1u);   // a < 32 here
}

__device__ __forceinline__ void add64(unsigned long long *p, long long v) {
    if (v) atomicAdd(p, (unsigned long long)v);
}

struct ReadCtx {
    const uint8_t *qual;
    const uint32_t *mp;
    uint32_t n_mp, l_seq;
    uint64_t g;          // linear genome coordinate of reference_start
    int strand;
    const uint2 *rx;
    uint64_t n_bits;
    int min_qual;
};

// fillMismatchesNGM (SlamSeqFile.py:313-348): f(conv, readPos0, refPos0, isSnpPosition) for every kept entry
template <class F>
__device__ __forceinline__ void for_each_kept(const ReadCtx &r, F f) {
    for (uint32_t k = 0; k < r.n_mp; k++) {
        const uint32_t w = __ldg(&r.mp[2 * k]), conv = __ldg(&r.mp[2 * k + 1]);
        const uint32_t rp = w & 0xFFFFu, fp = w >> 16;
        const int q = rp < r.l_seq ? (int)__ldg(&r.qual[rp]) : 0;          // :326-329
        if (q < r.min_qual) continue;                                      // :330
        const uint64_t gp = r.g + fp;
        bool snp = false;
        if (gp < r.n_bits) snp = (__ldg(&r.rx[gp >> 5]).y >> (gp & 31)) & 1u;   // isAGSnp on reverse reads, isTCSnp else (:333-337)
        f(conv, rp, fp, snp);
    }
}

// Accumulation.  The genome-wide outputs are a few hundred counters that every read adds to, so nothing goes straight to
// global memory: the matrix diagonals (base counts) and the tccontext counts run in per-thread registers over all of a
// thread's reads and are warp-reduced once at the end; the off-diagonal cells (one or two per mismatch) and the read- /
// UTR-position histograms are shared-memory atomics on per-CTA copies, flushed with one 64-bit atomic per non-zero cell.
// RATES / CONTEXT: compile the register accumulators of those two outputs in only when they are wanted (30 registers).
template <bool RATES, bool CONTEXT>
__global__ void __launch_bounds__(256) sd_stats_kernel(const StatsArgs A) {
    extern __shared__ int s_hist[];   // [2 * 25] matrices fwd / rev | [6 * (max_read_length + 1)] | [6 * (utr_norm + 1)]
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    const bool want_utr = A.utr_rates || A.utr_snp || A.utrpos || A.pairs;
    const uint32_t n_rp = (A.readpos && A.hist_smem) ? 6u * (A.max_read_length + 1u) : 0u;
    const uint32_t n_up = (A.utrpos && A.hist_smem) ? 6u * (A.utr_norm + 1u) : 0u;
    int *s_rates = s_hist, *s_readpos = s_hist + 2 * SD_NRATE, *s_utrpos = s_readpos + n_rp;
    for (uint32_t i = threadIdx.x; i < 2u * SD_NRATE + n_rp + n_up; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    int dF[5] = {0, 0, 0, 0, 0}, dR[5] = {0, 0, 0, 0, 0};                   // diagonal of the fwd / rev matrix: exact base counts
    int c5F[5] = {0, 0, 0, 0, 0}, c5R[5] = {0, 0, 0, 0, 0}, c3F[5] = {0, 0, 0, 0, 0}, c3R[5] = {0, 0, 0, 0, 0};   // tccontext
    for (uint32_t t = warp; t < A.n_tiles; t += n_warps) {
        const sd_read_rec rec = A.rec[(size_t)t * 32 + lane];
        const uint32_t flags = rec.ref_mapq_flag >> 24, chrom = rec.ref_mapq_flag & 0xFFFFu;
        const uint32_t l_seq = rec.lseq_ncig & 0xFFFFu, n_cig = rec.lseq_ncig >> 16, n_mp = rec.nm_nmp >> 16;
        const sd_tile tile = A.tiles[t];
        const uint32_t qoff = warp_excl_scan(l_seq, lane), coff = warp_excl_scan(n_cig, lane), moff = warp_excl_scan(n_mp, lane);
        const bool valid = !(flags & (SD_F_PAD | SD_F_SKIP | SD_F_UNMAPPED)) && chrom != SD_REF_NONE;
        const int strand = (flags & SD_F_REVERSE) ? 1 : 0;

        int m[SD_NRATE];                       // the read's own matrix: only the per-UTR sums need it (dynamically indexed: local memory)
        const bool want_m = A.utr_rates != nullptr;
        if (want_m) {
#pragma unroll
            for (int c = 0; c < SD_NRATE; c++) m[c] = 0;
        }
        int tc_count = 0;
        uint32_t span = 0, t_in_read = 0;
        ReadCtx R;
        R.qual = A.qual + tile.qual_off + qoff;
        R.mp = A.mp + tile.mp_off / 4 + 2 * (size_t)moff;
        R.n_mp = valid ? n_mp : 0;
        R.l_seq = l_seq;
        R.strand = strand;
        R.rx = A.refx[strand];
        R.n_bits = A.n_bits;
        R.min_qual = A.min_qual;
        R.g = 0;
        if (valid) {
            R.g = (uint64_t)A.chrom_off[chrom] + (uint64_t)(uint32_t)rec.pos;
            // reference span and the soft clips that query_alignment_sequence leaves out
            const uint32_t *cig = A.cigar + tile.cig_off / 4 + coff;
            int qs = 0, qe = (int)l_seq;
            bool leading = true;
            uint32_t trail = 0;
            for (uint32_t k = 0; k < n_cig; k++) {
                const uint32_t c = __ldg(&cig[k]), op = c & 15u, len = c >> 4;
                if ((0x18Du >> op) & 1u) span += len;          // M D N = X
                if (op == 4) {
                    if (leading) qs += (int)len;
                    else trail += len;
                } else if (op != 5) {
                    leading = false;
                    trail = 0;
                }
            }
            qe -= (int)trail;
            if (qe < qs) qe = qs;
            // computeRatesForRead: diagonal = base counts of the aligned part of the read (SlamSeqFile.py:234-237)
            const uint4 *blk = reinterpret_cast<const uint4 *>(reinterpret_cast<const uint8_t *>(A.seq) + tile.seq_off);
            const int n_grp = (int)((l_seq + 31u) >> 5);
            for (int g = 0; g < n_grp; g++) {
                const uint4 p = __ldg(&blk[(size_t)g * 32 + lane]);
                const uint32_t in = group_range(g, qs, qe), all = group_range(g, 0, (int)l_seq);
                const uint32_t onlyA = p.x & ~p.y & ~p.z & ~p.w, onlyC = ~p.x & p.y & ~p.z & ~p.w;
                const uint32_t onlyG = ~p.x & ~p.y & p.z & ~p.w, onlyT = ~p.x & ~p.y & ~p.z & p.w;
                const uint32_t isN = p.x & p.y & p.z & p.w;
                const int nA = __popc(onlyA & in), nC = __popc(onlyC & in), nG = __popc(onlyG & in), nT = __popc(onlyT & in), nN = __popc(isN & in);
                if (want_m) { m[0] += nA; m[6] += nC; m[12] += nG; m[18] += nT; m[24] += nN; }
                if (RATES) {
                    if (!strand) { dF[0] += nA; dF[1] += nC; dF[2] += nG; dF[3] += nT; dF[4] += nN; }
                    else { dR[0] += nA; dR[1] += nC; dR[2] += nG; dR[3] += nT; dR[4] += nN; }
                }
                t_in_read += __popc((strand ? onlyA : onlyT) & all);   // SlamSeqRead.getTcount, SlamSeqFile.py:181-185
            }
            const uint32_t tc_conv = strand ? 2u : 16u;
            for_each_kept(R, [&](uint32_t conv, uint32_t, uint32_t, bool snp) {
                const uint32_t rb = conv / 5u, qb = conv % 5u;
                if (want_m) {
                    if (!snp) m[conv]++;                // SlamSeqFile.py:241-243
                    else m[6 * rb]++;                   // :245
                    m[6 * qb]--;
                }
                if (RATES) {
                    atomicAdd(&s_rates[strand * SD_NRATE + (snp ? 6 * rb : conv)], 1);
                    atomicAdd(&s_rates[strand * SD_NRATE + 6 * qb], -1);
                }
                if (conv == tc_conv && !snp) tc_count++;   // isTCMismatch, :137-141
            });
        }
        const bool is_tc_read = tc_count >= A.conv_threshold;               // SlamSeqFile.py:397-400

        // ---- alleyoop tccontext (stats.py:256-287): neighbours of every T (forward reads) / A (reverse reads) of the read sequence
        if (CONTEXT) {
            const bool in = !(flags & SD_F_PAD) && chrom != SD_REF_NONE;      // bamFile.fetch(region=chromosome): flag-unmapped records too
            if (in) {
                const uint4 *blk = reinterpret_cast<const uint4 *>(reinterpret_cast<const uint8_t *>(A.seq) + tile.seq_off);
                const int n_grp = (int)((l_seq + 31u) >> 5);
                uint32_t prev[5] = {0u, 0u, 0u, 0u, 0u};                         // exact-base masks of the previous group
                uint4 p = n_grp ? __ldg(&blk[lane]) : make_uint4(0u, 0u, 0u, 0u);
                for (int g = 0; g < n_grp; g++) {
                    const uint4 nx = g + 1 < n_grp ? __ldg(&blk[(size_t)(g + 1) * 32 + lane]) : make_uint4(0u, 0u, 0u, 0u);
                    const uint32_t cur[5] = {p.x & ~p.y & ~p.z & ~p.w, ~p.x & p.y & ~p.z & ~p.w, ~p.x & ~p.y & p.z & ~p.w,
                                             ~p.x & ~p.y & ~p.z & p.w, p.x & p.y & p.z & p.w};
                    const uint32_t nxt[5] = {nx.x & ~nx.y & ~nx.z & ~nx.w, ~nx.x & nx.y & ~nx.z & ~nx.w, ~nx.x & ~nx.y & nx.z & ~nx.w,
                                             ~nx.x & ~nx.y & ~nx.z & nx.w, nx.x & nx.y & nx.z & nx.w};
                    const uint32_t centre = strand ? cur[0] : cur[3];            // A on reverse reads, T on forward reads
#pragma unroll
                    for (int x = 0; x < 5; x++) {
                        const uint32_t left = (cur[x] << 1) | (prev[x] >> 31);   // bit i: base i-1 is x
                        const uint32_t right = (cur[x] >> 1) | (nxt[x] << 31);   // bit i: base i+1 is x
                        const int cx = x == 4 ? 4 : 3 - x;                       // complement (misc.py:276-280)
                        const int nl = __popc(centre & left), nr = __popc(centre & right);
                        if (!strand) {                                           // by the printed neighbour: A, C, G, T, N
                            c5F[x] += nl;
                            c3F[x] += nr;
                        } else {                                                 // front = next base, back = previous base, both complemented
                            c5R[cx] += nr;
                            c3R[cx] += nl;
                        }
                    }
#pragma unroll
                    for (int x = 0; x < 5; x++) prev[x] = cur[x];
                    p = nx;
                }
            }
        }

        // ---- genome-wide outputs (readsInChromosome: strand ".", every placed record)
        if (A.readpos) {
            const uint32_t W = A.max_read_length, ld = W + 1;
            const bool in = valid && l_seq <= W;                              // stats.py:621
            const uint32_t key = in ? ((uint32_t)strand << 16 | l_seq) : 0xFFFFFFFFu;
            const uint32_t peers = __match_any_sync(0xFFFFFFFFu, key);
            if (in && lane == (uint32_t)(__ffs(peers) - 1)) {
                const int n = __popc(peers);
                if (A.hist_smem) {
                    atomicAdd(&s_readpos[(4 + strand) * ld + 0], n);
                    atomicAdd(&s_readpos[(4 + strand) * ld + l_seq], -n);
                } else {
                    add64(&A.readpos[(4 + strand) * (size_t)ld + 0], n);
                    add64(&A.readpos[(4 + strand) * (size_t)ld + l_seq], -n);
                }
            }
            if (in) {
                const uint32_t tc_conv = strand ? 2u : 16u;
                for_each_kept(R, [&](uint32_t conv, uint32_t rp, uint32_t, bool snp) {
                    long long p = strand ? (long long)l_seq - (long long)rp - 1 : (long long)rp;   // SlamSeqFile.py:335
                    if (p < 0) p += W;                       // a Python list index below zero counts from the end
                    if (p < 0 || p >= (long long)W) return;  // (the reference raises IndexError here)
                    const bool is_tc = conv == tc_conv && !snp;
                    if (A.hist_smem) atomicAdd(&s_readpos[((is_tc ? 2 : 0) + strand) * ld + (uint32_t)p], 1);
                    else atomicAdd(&A.readpos[((is_tc ? 2 : 0) + strand) * (size_t)ld + (size_t)p], 1ull);
                });
            }
        }

        // ---- per-UTR outputs (readInRegion: strand-specific interval join)
        if (want_utr && valid) {
            const int s = strand;
            const uint64_t g = R.g, e = R.g + (span ? span : 1u);          // reference_end = pos + max(1, span)
            const uint32_t nU = A.n_u[s];
            uint32_t i = nU ? __ldg(&A.bin_first[s][g >> SD_BIN_SHIFT]) : 0;
            const uint32_t tc_conv = s ? 2u : 16u;
            const long long norm = A.utr_norm;
            for (; i < nU; i++) {
                const uint4 ent = __ldg(&A.utab[s][i]);
                if ((uint64_t)ent.x >= e) break;
                if ((uint64_t)ent.y <= g) continue;
                const uint32_t row = ent.z;
                const long long L = __ldg(&A.row_blen[row]);
                const bool plus_row = __ldg(&A.row_strand[row]) == 0;   // `if utr.strand == "+"` (a row without strand takes the else branch)
                const long long start_ref = (long long)g - (long long)ent.x, end_ref = (long long)e - (long long)ent.x;
                if (A.utr_rates) {
                    // lanes of a sorted tile mostly sit in the same UTR: the lanes that are here together and hit the same row
                    // sum their matrices in the warp and one of them adds (26 same-address atomics per lane otherwise)
                    const bool take = !(!is_tc_read && A.strict && tc_count > 0);       // stats.py:376
                    const unsigned grp = __match_any_sync(__activemask(), row);
                    const bool leader = lane == (uint32_t)(__ffs((int)grp) - 1);
                    unsigned long long *dst = A.utr_rates + (size_t)row * (SD_NRATE + 1);
#pragma unroll
                    for (int c = 0; c < SD_NRATE; c++) {
                        const int sum = __reduce_add_sync(grp, take ? m[c] : 0);
                        if (leader) add64(&dst[c], sum);
                    }
                    const int n_take = __reduce_add_sync(grp, take ? 1 : 0);
                    if (leader) add64(&dst[SD_NRATE], n_take);
                }
                const bool no_list = !is_tc_read && A.strict;                           // stats.py:806-810, tcounter.py:210-214
                if (A.utr_snp) {
                    bool is_tc = false, is_true_tc = false;
                    if (!no_list)
                        for_each_kept(R, [&](uint32_t conv, uint32_t, uint32_t fp, bool snp) {
                            const long long rel = start_ref + (long long)fp;           // SlamSeqFile.py:339
                            if (conv == tc_conv && rel >= 0 && rel < L) {
                                is_tc = true;                                          // stats.py:823-830
                                if (!snp) is_true_tc = true;                           // stats.py:820-821
                            }
                        });
                    unsigned long long *dst = A.utr_snp + (size_t)row * 3;
                    const unsigned grp = __match_any_sync(__activemask(), row);       // same-row lanes add once, as above
                    const int n_reads = __popc(grp), n_tc = __reduce_add_sync(grp, is_tc ? 1 : 0), n_true = __reduce_add_sync(grp, is_true_tc ? 1 : 0);
                    if (lane == (uint32_t)(__ffs((int)grp) - 1)) {
                        add64(&dst[0], n_reads);
                        add64(&dst[1], n_tc);
                        add64(&dst[2], n_true);
                    }
                }
                if (A.utrpos) {
                    const size_t ld = (size_t)norm + 1;
                    for_each_kept(R, [&](uint32_t conv, uint32_t, uint32_t fp, bool snp) {
                        const long long rel = start_ref + (long long)fp;
                        long long idx = -1;
                        if (plus_row) {
                            if (rel >= L - norm && rel < L) idx = norm - (L - rel);    // stats.py:703-706
                        } else if (rel >= 0 && rel < min(L, norm)) idx = rel;           // stats.py:714
                        if (idx < 0) return;
                        const bool is_tc = conv == tc_conv && !snp;
                        if (A.hist_smem) atomicAdd(&s_utrpos[((is_tc ? 2 : 0) + s) * ld + (size_t)idx], 1);
                        else atomicAdd(&A.utrpos[((is_tc ? 2 : 0) + s) * ld + (size_t)idx], 1ull);
                    });
                    long long a, b;
                    if (s) {                                                            // stats.py:725-726
                        const long long cap = min(L, norm);
                        a = max(0ll, min(cap, start_ref));
                        b = max(0ll, min(cap, end_ref));
                    } else {                                                            // stats.py:737-742
                        a = min(L, max(L - norm, start_ref)) + norm - L;
                        b = min(L, max(L - norm, end_ref)) + norm - L;
                    }
                    if (b > a) {
                        if (A.hist_smem) {
                            atomicAdd(&s_utrpos[(4 + s) * ld + (size_t)a], 1);
                            atomicAdd(&s_utrpos[(4 + s) * ld + (size_t)b], -1);
                        } else {
                            atomicAdd(&A.utrpos[(4 + s) * ld + (size_t)a], 1ull);
                            add64(&A.utrpos[(4 + s) * ld + (size_t)b], -1);
                        }
                    }
                }
                if (A.pairs) {
                    uint32_t n = t_in_read, k = 0;
                    if (!no_list)
                        for_each_kept(R, [&](uint32_t conv, uint32_t, uint32_t fp, bool snp) {
                            const long long rel = start_ref + (long long)fp;
                            if (rel < 0 || rel >= L) return;                            // tcounter.py:236
                            if (conv / 5u == (s ? 0u : 3u)) n++;                        // isT: reference base A / T (SlamSeqFile.py:143-147)
                            if (conv == tc_conv && !snp) k++;
                        });
                    const unsigned long long slot = atomicAdd(A.n_pairs, 1ull);
                    if (slot < A.pair_capacity) {
                        const size_t idx = (size_t)t * 32 + lane;
                        A.pairs[slot] = sd_read_pair{row, A.order ? __ldg(&A.order[idx]) : (uint32_t)idx, n, k};
                    }
                }
            }
        }
    }
    // ---- flush: register accumulators -> warp totals -> the CTA's shared copy (matrix diagonals) or global (tccontext)
    if (RATES) {
#pragma unroll
        for (int x = 0; x < 5; x++) {
            const int f = __reduce_add_sync(0xFFFFFFFFu, dF[x]), r = __reduce_add_sync(0xFFFFFFFFu, dR[x]);
            if (lane == 0) {
                if (f) atomicAdd(&s_rates[6 * x], f);
                if (r) atomicAdd(&s_rates[SD_NRATE + 6 * x], r);
            }
        }
    }
    if (CONTEXT) {
#pragma unroll
        for (int x = 0; x < 5; x++) {
            const int f5 = __reduce_add_sync(0xFFFFFFFFu, c5F[x]), r5 = __reduce_add_sync(0xFFFFFFFFu, c5R[x]);
            const int f3 = __reduce_add_sync(0xFFFFFFFFu, c3F[x]), r3 = __reduce_add_sync(0xFFFFFFFFu, c3R[x]);
            if (lane == 0) {
                add64(&A.context[0 + x], f5);
                add64(&A.context[5 + x], r5);
                add64(&A.context[10 + x], f3);
                add64(&A.context[15 + x], r3);
            }
        }
    }
    __syncthreads();
    if (RATES)
        for (uint32_t i = threadIdx.x; i < 2u * SD_NRATE; i += blockDim.x) add64(&A.rates[i], s_rates[i]);
    for (uint32_t i = threadIdx.x; i < n_rp; i += blockDim.x) add64(&A.readpos[i], s_readpos[i]);
    for (uint32_t i = threadIdx.x; i < n_up; i += blockDim.x) add64(&A.utrpos[i], s_utrpos[i]);
}

}  // namespace

extern "C" int sd_read_stats(sd_ctx *ctx, const sd_batch *b, const sd_stats_args *a) {
    if (!ctx || !b || !a) return sd_fail(ctx, SD_ERR_ARG, "sd_read_stats: null argument");
    if (!ctx->refx[0]) return sd_fail(ctx, SD_ERR_STATE, "sd_read_stats before sd_set_reference");
    if (b->tile_reads != 32) return sd_fail(ctx, SD_ERR_ARG, "sd_read_stats: tile_reads must be 32");
    if ((b->qual_bits != 0 && b->qual_bits != 8)) return sd_fail(ctx, SD_ERR_ARG, "sd_read_stats: expand the quality codes first (sd_qual_unpack)");
    const bool want_utr = a->d_utr_rates || a->d_utr_snp || a->d_utrpos || a->d_pairs;
    if (want_utr && !ctx->d_row_blen) return sd_fail(ctx, SD_ERR_STATE, "sd_read_stats: per-UTR outputs need sd_set_utrs");
    if (a->d_pairs && !a->d_n_pairs) return sd_fail(ctx, SD_ERR_ARG, "sd_read_stats: d_pairs without d_n_pairs");
    if (a->d_readpos && a->max_read_length == 0) return sd_fail(ctx, SD_ERR_ARG, "sd_read_stats: max_read_length 0");
    if (b->seq_bits == 2) return sd_fail(ctx, SD_ERR_ARG, "sd_read_stats needs the 4-plane SEQ column (exact base counts), not the 2-bit form");
    if (b->n_tiles == 0) return SD_OK;
    if (!b->mp && b->max_tile_mp) return sd_fail(ctx, SD_ERR_ARG, "sd_read_stats: batch without the MP column (decode with SD_DECODE_MP_ALL)");
    SD_CUDA(ctx, cudaSetDevice(ctx->device));
    StatsArgs A;
    A.rec = b->rec; A.seq = b->seq; A.qual = b->qual; A.cigar = b->cigar; A.mp = b->mp; A.tiles = b->tiles;
    A.n_tiles = b->n_tiles;
    A.order = a->d_order;
    A.chrom_off = ctx->d_chrom_off;
    A.n_bits = ctx->n_words * 32ull;
    for (int s = 0; s < 2; s++) {
        A.refx[s] = ctx->refx[s];
        A.utab[s] = ctx->d_utab[s];
        A.bin_first[s] = ctx->d_bin_first[s];
        A.n_u[s] = want_utr ? ctx->n_u[s] : 0;
    }
    A.row_blen = ctx->d_row_blen;
    A.row_strand = ctx->d_row_strand;
    A.min_qual = a->min_qual; A.conv_threshold = a->conversion_threshold; A.strict = a->strict;
    A.max_read_length = a->max_read_length; A.utr_norm = a->utr_norm;
    A.rates = (unsigned long long *)a->d_rates; A.readpos = (unsigned long long *)a->d_readpos;
    A.utr_rates = (unsigned long long *)a->d_utr_rates; A.utr_snp = (unsigned long long *)a->d_utr_snp;
    A.utrpos = (unsigned long long *)a->d_utrpos;
    A.context = (unsigned long long *)a->d_context;
    A.pairs = a->d_pairs; A.pair_capacity = a->pair_capacity; A.n_pairs = a->d_n_pairs;
    const uint32_t warps_per_block = 8;
    uint32_t blocks = (b->n_tiles + warps_per_block - 1) / warps_per_block;
    const uint32_t cap = (uint32_t)ctx->num_sms * 8u;    // grid-stride over tiles: a multiple of the SM count
    if (blocks > cap) blocks = cap;
    // per-CTA shared copies of the read-position / UTR-position histograms when they fit (else global atomics)
    size_t hist = (size_t)(a->d_readpos ? 6u * ((size_t)a->max_read_length + 1u) : 0u) + (size_t)(a->d_utrpos ? 6u * ((size_t)a->utr_norm + 1u) : 0u);
    A.hist_smem = hist * sizeof(int) <= 40u * 1024u ? 1u : 0u;
    const size_t smem = (2u * SD_NRATE + (A.hist_smem ? hist : 0u)) * sizeof(int);
    const unsigned threads = warps_per_block * 32;
    if (a->d_rates && a->d_context) sd_stats_kernel<true, true><<<blocks, threads, smem, ctx->stream>>>(A);
    else if (a->d_rates) sd_stats_kernel<true, false><<<blocks, threads, smem, ctx->stream>>>(A);
    else if (a->d_context) sd_stats_kernel<false, true><<<blocks, threads, smem, ctx->stream>>>(A);
    else sd_stats_kernel<false, false><<<blocks, threads, smem, ctx->stream>>>(A);
    ctx->launches++;
    SD_CUDA(ctx, cudaGetLastError());
    return SD_OK;
}
