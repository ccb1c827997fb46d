// sd_count_lean.cuh -- the count kernel for the common tile: 32 coordinate-sorted reads that are all one match block of
// the same length on one chromosome (sd_tile_span.lean, decided when the batch's span table is made).  Everything such a
// tile does not need is compiled out: no CIGAR, no row offsets, no per-read chromosome lookups, no placement of the read
// mask, mismatches from the CIGAR walk only (SD_MISMATCH_CIGAR, or SD_MISMATCH_VERIFY through VERIFY = true), 2-bit SEQ column
// and plane-form qualities only.  Tiles that are not of that kind go, untouched, to the general kernels that follow
// (sd_count_kernel<..., true> staged, then <..., false> plain).  Same outputs, bit for bit.
//
// The columns are read straight from global memory into registers: in the group-major layout the 32 rows of one 32-base
// group are 256 (SEQ, 2 quality planes) or 512 consecutive bytes, so "lane r loads group g of row r" is one fully coalesced
// warp access -- no shared-memory stage and no bulk copy.  (A first version staged every tile with five cp.async.bulk copies
// like the general kernel does.  It ran at exactly one bulk copy per 46 cycles per SM whatever else changed in the kernel:
// the copy engine's request rate, not instruction issue or DRAM, was the limit -- 7.1 M copies of 64 B .. 1 KB for 50 M reads.
// profiles/r2_lean_tma_*.)  The tile's reference window and annotation entries come through L1: without the stages almost
// all of the SM's 228 KB is cache, and consecutive tiles of a sorted batch look at the same few lines.
// One thread per read:
//   filter predicates (dunks/filter.py:235-256)  ->  read mask "is C / is G and quality >= minQual" (SlamSeqFile.py:137-141, 330)
//   AND reference planes "is T / is A, not a SNP" of the read's window (SlamSeqFile.py:333-344)  ->  tcCount, T coverage
//   ->  warp-uniform walk over the tile's annotation entries: per-UTR counters as warp votes / reductions, kept in registers
//   while consecutive tiles stay in the same UTR (tcounter.py:216-248, 277-283), per-position coverage differences and
//   conversions (tcounter.py:229-231, 246-248).
#pragma once
#include "sd_count_kernel.cuh"

struct LeanArgs {
    const sd_read_rec *rec;
    const unsigned char *seq, *qual, *mp;   // 2-bit SEQ column, plane-form qualities, MP entries (VERIFY) or null
    const sd_tile *tiles;
    const sd_tile_span *spans;
    const uint4 *refm, *ment;
    uint32_t n_tiles, n_chrom;
    uint32_t claim;               // consecutive tiles per queue access
    uint32_t q_cmin, q_m[4];
    int conv_threshold, filter_on, filter_mq, filter_max_nm;
    float filter_min_identity;
    unsigned long long *counters;
    int *pos_cov, *pos_tc;
    unsigned long long *stats;
    uint8_t *read_tc;
    unsigned int *tile_counter;   // queue head
    uint32_t *leftover;           // tiles left to the general kernels
    unsigned int *n_leftover;
    unsigned long long n_positions;
    uint32_t n_utr;
};

#ifndef SD_LEAN_DIAG
#define SD_LEAN_DIAG 0   // timing experiments only (wrong results): 1 no conversion scatter, 2 no coverage adds, 4 no annotation walk, 8 no window loads
#endif
#ifndef SD_LEAN_CLAIM
#define SD_LEAN_CLAIM 32   // most consecutive tiles per claim: a warp stays in one stretch of the genome (window and entry lines in L1, one
                           // UTR's sums in registers) -- 1.05 instead of 1.16 ms per 50 M reads against chunks of 8.  Small batches get
                           // smaller chunks, at least eight per warp, so that the grid drains evenly
#endif
#ifndef SD_LEAN_MIN_CTAS
#define SD_LEAN_MIN_CTAS(W) SD_MIN_CTAS(W)
#endif

template <int W, int QB, bool VERIFY>
__global__ void __launch_bounds__(SD_CTA_THREADS, SD_LEAN_MIN_CTAS(W)) sd_lean_kernel(const __grid_constant__ LeanArgs A) {
    __shared__ uint32_t s_stats[SD_SMEM_STATS];
    const uint32_t lane = threadIdx.x & 31u;
    if (threadIdx.x < SD_SMEM_STATS) s_stats[threadIdx.x] = 0u;
    __syncthreads();

    // tile queue: chunks of SD_LEAN_CLAIM consecutive tiles per atomic.  Every lane keeps the chunk's bounds (warp-uniform); the atomic for
    // the chunk AFTER the current one is already in flight (lane 0's `pend`), so taking a new chunk never waits for the queue head.
    uint32_t claim_next, claim_end, pend = 0u;
    {
        uint32_t c = 0u;
        if (lane == 0) c = atomicAdd(A.tile_counter, A.claim);
        claim_next = __shfl_sync(0xFFFFFFFFu, c, 0);
        claim_end = claim_next + A.claim;
        if (lane == 0) pend = atomicAdd(A.tile_counter, A.claim);
    }
    auto claim = [&]() -> uint32_t {
        if (claim_next == claim_end) {   // warp-uniform
            claim_next = __shfl_sync(0xFFFFFFFFu, pend, 0);
            claim_end = claim_next + A.claim;
            if (lane == 0 && claim_next < A.n_tiles) pend = atomicAdd(A.tile_counter, A.claim);
        }
        const uint32_t t = claim_next++;
        return t < A.n_tiles ? t : 0xFFFFFFFFu;
    };
    // A tile's description is four (five with VERIFY) 64-bit words in two tables; lane c fetches word c as soon as the tile is
    // claimed -- one tile ahead of its use -- and keeps it in a register:
    //   0 spans[t].{lean, g0}   1 tiles[t].seq_off   2 tiles[t].qual_off   3 spans[t].ent   4 tiles[t].mp_off
    auto fetch_desc = [&](uint32_t tile) -> unsigned long long {
        if (tile == 0xFFFFFFFFu || lane >= (VERIFY ? 5u : 4u)) return 0ull;
        const bool in_tiles = lane == 1u || lane == 2u || lane == 4u;
        const unsigned char *base = in_tiles ? reinterpret_cast<const unsigned char *>(A.tiles) : reinterpret_cast<const unsigned char *>(A.spans);
        //                      lane:   4    3    2    1    0
        const uint32_t field = (uint32_t)((0x1828080030ull >> (8u * lane)) & 0xFFull);   // byte offset of the word in its table entry
        return __ldg(reinterpret_cast<const unsigned long long *>(base + (size_t)tile * (in_tiles ? sizeof(sd_tile) : sizeof(sd_tile_span)) + field));
    };
    auto bcast64 = [&](unsigned long long x, int src) -> unsigned long long {
        const uint32_t lo = __shfl_sync(0xFFFFFFFFu, (uint32_t)x, src), hi = __shfl_sync(0xFFFFFFFFu, (uint32_t)(x >> 32), src);
        return (unsigned long long)lo | ((unsigned long long)hi << 32);
    };

    uint32_t cnt_a = 0, cnt_b = 0, cnt_tiles = 0, my_tc_total = 0;   // packed per-lane run counters (10-bit fields), see below
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
    // Counters of the annotation row the warp is in: consecutive tiles of a sorted batch hit the same UTR, so every lane keeps its own
    // part of the row's five sums in registers (reads | T>C reads << 10 | multimappers << 20 in one word: a lane adds at most one
    // of each per tile) and the warp adds them up when the row changes -- five reductions and atomics per row change, not per tile.
    uint32_t cur_row = 0xFFFFFFFFu, acc_pk = 0u, acc_cov = 0u, acc_conv = 0u;
    auto flush_row = [&]() {
        if (cur_row != 0xFFFFFFFFu) {
            const uint32_t n = __reduce_add_sync(0xFFFFFFFFu, acc_pk & 1023u), n_tc = __reduce_add_sync(0xFFFFFFFFu, (acc_pk >> 10) & 1023u),
                           n_mm = __reduce_add_sync(0xFFFFFFFFu, acc_pk >> 20);
            const uint32_t cov = __reduce_add_sync(0xFFFFFFFFu, acc_cov), conv = __reduce_add_sync(0xFFFFFFFFu, acc_conv);
            if (lane == 0u) {
                SD_CHECK(cur_row < A.n_utr);
                unsigned long long *c = A.counters + (size_t)cur_row * SD_NCOUNTER;
                if (n) atomicAdd(c + SD_C_READS, (unsigned long long)n);
                if (n_tc) atomicAdd(c + SD_C_TCREADS, (unsigned long long)n_tc);
                if (n_mm) atomicAdd(c + SD_C_MULTIMAP, (unsigned long long)n_mm);
                if (cov) atomicAdd(c + SD_C_COV_ON_T, (unsigned long long)cov);
                if (conv) atomicAdd(c + SD_C_CONV_ON_T, (unsigned long long)conv);
            }
        }
        acc_pk = acc_cov = acc_conv = 0u;
    };

    // the three record words the scan needs (pos, chromosome | MAPQ | flags, XI), loaded one tile ahead like the description
    const bool want_w4 = VERIFY || A.filter_max_nm > -1;   // NM | MP entries << 16: only the NM filter and the MP cross-check look at it
    auto fetch_rec = [&](uint32_t tile, uint32_t &w0, uint32_t &w1, uint32_t &w2, uint32_t &w4) {
        w0 = w1 = w2 = w4 = 0u;
        if (tile != 0xFFFFFFFFu) {
            const uint32_t *rw = reinterpret_cast<const uint32_t *>(A.rec + (size_t)tile * SD_TILE_READS + lane);
            w0 = __ldcs(rw);
            w1 = __ldcs(rw + 1);
            w2 = __ldcs(rw + 2);
            if (want_w4) w4 = __ldcs(rw + 4);
        }
    };
    uint32_t tile = claim();
    unsigned long long desc = fetch_desc(tile);
    uint32_t r_pos, r_rmf, r_xi, r_w4;
    fetch_rec(tile, r_pos, r_rmf, r_xi, r_w4);
    while (tile != 0xFFFFFFFFu) {   // claims are monotonic: nothing behind the first empty one
        // the next tile: claimed now, its description and records on the way into registers while this tile is counted
        const uint32_t next = claim();
        const unsigned long long next_desc = fetch_desc(next);
        uint32_t n_pos, n_rmf, n_xi, n_w4;
        fetch_rec(next, n_pos, n_rmf, n_xi, n_w4);
        const uint32_t lseq = __shfl_sync(0xFFFFFFFFu, (uint32_t)desc, 0);
        if (lseq == 0u) {
            if (lane == 0) A.leftover[atomicAdd(A.n_leftover, 1u)] = tile;
        } else {
            const uint32_t g0 = __shfl_sync(0xFFFFFFFFu, (uint32_t)(desc >> 32), 0);
            const unsigned long long seq_off = bcast64(desc, 1), qual_off = bcast64(desc, 2), entw = bcast64(desc, 3);
            const uint32_t ne = ((uint32_t)entw & 0xFFFFu) >> 5;
            const uint4 *ent = A.ment + (size_t)(entw >> 20);   // entw = byte offset << 16 | bytes; 16 bytes per uint4
            const uint32_t G = (lseq + 31u) >> 5;
            // --- the tile's column loads up front: the read's SEQ groups and quality planes (streamed: evict first).  (Pulling the blocks of
            // the tile two ahead into L2 with prefetch.global.L2 was measured: 50 more instructions per tile, no gain.)
            const int32_t pos = (int32_t)r_pos;
            const uint32_t rmf = r_rmf;
            const float xi = __uint_as_float(r_xi);
            const uint32_t nm_nmp = r_w4;
            // VERIFY: where this read's MP entries sit in the tile's block (exclusive scan of the entry counts), and its first two
            // entries on their way now, next to the SEQ / quality loads -- most reads list at most two conversions
            const uint32_t nmp = nm_nmp >> 16;
            const uint32_t *mp = nullptr;
            uint32_t en0 = 0u, en1 = 0u;
            if (VERIFY) {
                const unsigned long long mp_off = bcast64(desc, 4);
                uint32_t ms = nmp;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, ms, d);
                    if (lane >= (uint32_t)d) ms += t;
                }
                mp = reinterpret_cast<const uint32_t *>(A.mp + mp_off) + (ms - nmp);
                if (nmp > 0u) en0 = __ldcs(mp);
                if (nmp > 1u) en1 = __ldcs(mp + 1);
            }
            uint2 xs[W];
            uint32_t xq[W][QB];
            {
                const uint2 *sq = reinterpret_cast<const uint2 *>(A.seq + seq_off) + lane;
#pragma unroll
                for (int k = 0; k < W; k++) {
                    const uint32_t kc = min((uint32_t)k, G - 1u);   // an existing group; behind the read's last one T is zero anyway
                    xs[k] = __ldcs(sq + kc * SD_TILE_READS);
                    if (QB == 2) {
                        const uint2 q = __ldcs(reinterpret_cast<const uint2 *>(A.qual + qual_off) + lane + kc * SD_TILE_READS);
                        xq[k][0] = q.x; xq[k][1] = q.y;
                    } else {
                        const uint4 q = __ldcs(reinterpret_cast<const uint4 *>(A.qual + qual_off) + lane + kc * SD_TILE_READS);
                        xq[k][0] = q.x; xq[k][1] = q.y; xq[k][QB > 2 ? 2 : 0] = q.z; xq[k][QB > 2 ? 3 : 1] = q.w;
                    }
                }
            }
            const uint32_t flags = rmf >> 24, mapq = (rmf >> 16) & 0xFFu;
            const bool real = !(flags & SD_F_PAD);
            const bool cand = !(flags & (SD_F_PAD | SD_F_UNMAPPED | SD_F_SKIP)) && (rmf & 0xFFFFu) < A.n_chrom;
            // --- filter.py:235-256; run counters as 10-bit fields of two registers (cnt_a: mapped | kept | counted,
            // cnt_b: unmapped | MQ | identity), added up by flush_counts() before a field can overflow
            bool go = real;
            if (A.filter_on) {
                const bool primary = real && !(flags & (SD_F_SECONDARY | SD_F_SUPPLEMENTARY));
                const bool mapped = real && !(flags & SD_F_UNMAPPED);
                const bool f_mq = mapped && (int)mapq < A.filter_mq;
                const bool f_id = mapped && !f_mq && xi < A.filter_min_identity;
                bool f_nm = false;
                if (A.filter_max_nm > -1) {
                    f_nm = mapped && !f_mq && !f_id && (int)(nm_nmp & 0xFFFFu) > A.filter_max_nm;
                    if (f_nm) atomicAdd(&s_stats[SD_S_NMFILTERED], 1u);
                }
                go = mapped && !f_mq && !f_id && !f_nm;
                cnt_a += (uint32_t)(primary && mapped) | ((uint32_t)(go && primary) << 10);
                cnt_b += (uint32_t)(primary && !mapped) | ((uint32_t)f_mq << 10) | ((uint32_t)f_id << 20);
            }
            const bool live = go && cand;
            cnt_a += (uint32_t)live << 20;
            const uint32_t s = flags & SD_F_REVERSE;   // 0 '+', 1 '-'
            const uint32_t lm = live ? 0xFFFFFFFFu : 0u;
            // --- geometry: one match block of lseq bases at linear position g
            const uint32_t g = g0 + (live ? (uint32_t)pos : 0u), e = g + lseq;
            const uint32_t sh = g & 31u;
            // --- reference window of this strand, shifted to the read's start: T = is T / A inside the read, S = SNPs
            uint32_t T[W], S[W];
            {
                const uint2 *wp = reinterpret_cast<const uint2 *>(A.refm + (g >> 5)) + s;
                uint2 x[W + 1];
#pragma unroll
                for (int k = 0; k <= W; k++) x[k] = (SD_LEAN_DIAG & 8) ? make_uint2(0x11111111u * (uint32_t)(k + 1) + g, 0u) : __ldg(wp + 2 * k);
#pragma unroll
                for (int k = 0; k < W; k++) {
                    T[k] = __funnelshift_r(x[k].x, x[k + 1].x, sh) & prefix_word(k, (int)lseq) & lm;
                    S[k] = __funnelshift_r(x[k].y, x[k + 1].y, sh);
                }
            }
            // --- read mask: base is C ('+') / G ('-') and its quality passes; candidates H = M & T & ~S
            uint32_t H[W], D[VERIFY ? W : 1];
            {
                const uint32_t sm = 0u - s;
                const uint32_t hm = (flags & SD_F_HAS_MP) ? 0xFFFFFFFFu : 0u;   // no MP tag: the reference sees no mismatches (SlamSeqFile.py:318)
#pragma unroll
                for (int k = 0; k < W; k++) {
                    const uint32_t m = group_is_base2(xs[k], sm);
                    const uint32_t q = planes_ge<QB, true>(xq[k], A);
                    if (VERIFY) D[k] = m & T[k] & hm;   // before quality / SNP masking: what NextGenMap lists
                    H[k] = m & q & T[k] & ~S[k] & hm;
                }
            }
            if (VERIFY) {
                // the read's T>C / A>G candidates against its MP entries (reference offset and read index; plain match: equal)
                bool bad = false;
                if (live && (flags & SD_F_HAS_MP)) {
                    for (uint32_t i = 0; i < nmp; i++) {
                        const uint32_t en = i == 0u ? en0 : (i == 1u ? en1 : __ldcs(mp + i));
                        const uint32_t rp = en & 0xFFFFu, o = en >> 16;
                        bool found = false;
#pragma unroll
                        for (int k = 0; k < W; k++)
                            if ((o >> 5) == (uint32_t)k && ((D[k] >> (o & 31u)) & 1u)) {
                                D[k] &= ~(1u << (o & 31u));
                                found = true;
                            }
                        if (!found || o != rp) bad = true;
                    }
#pragma unroll
                    for (int k = 0; k < W; k++) bad |= (D[k] != 0u);
                }
                if (bad) atomicAdd(&A.stats[SD_S_MP_DISAGREE], 1ull);
            }
            int tc = 0;
            uint32_t pop_t = 0;
#pragma unroll
            for (int k = 0; k < W; k++) {
                tc += __popc(H[k]);
                pop_t += __popc(T[k]);
            }
            if (tc < A.conv_threshold) {   // not a T>C read: its conversions are dropped (tcounter.py:210-214)
                tc = 0;
#pragma unroll
                for (int k = 0; k < W; k++) H[k] = 0u;
            }
            my_tc_total += (uint32_t)tc;
            if (A.read_tc) A.read_tc[(size_t)tile * SD_TILE_READS + lane] = (uint8_t)(tc > 255 ? 255 : tc);

            // --- interval join over the tile's annotation entries (warp-uniform loop)
            uint32_t last_seg = 0xFFFFFFFFu;
            for (uint32_t j = 0; j < ((SD_LEAN_DIAG & 4) ? 0u : ne); j++) {
                const uint4 u = __ldg(ent + 2 * j);   // {start, end, row | strand << 31, segment}: one address for the warp
                const bool hit = live && (u.z >> 31) == s && u.x < e && u.y > g;
                const unsigned hmask = __ballot_sync(0xFFFFFFFFu, hit);
                if (hmask == 0u) continue;
                uint32_t cov_t = hit ? pop_t : 0u, conv = hit ? (uint32_t)tc : 0u;
                const bool partial = hit && (u.x > g || u.y < e);
                if (__any_sync(0xFFFFFFFFu, partial)) {
                    if (partial) {   // the UTR covers only part of the read
                        const int lo = u.x > g ? (int)(u.x - g) : 0;
                        const int hi = u.y < e ? (int)(u.y - g) : (int)lseq;
                        cov_t = 0u;
                        conv = 0u;
#pragma unroll
                        for (int k = 0; k < W; k++) {
                            const uint32_t m = prefix_word(k, hi) & ~prefix_word(k, lo);
                            cov_t += __popc(T[k] & m);
                            conv += __popc(H[k] & m);
                        }
                    }
                }
                const uint32_t row = u.z & 0x7FFFFFFFu;
                if (row != cur_row) {   // warp-uniform
                    flush_row();
                    cur_row = row;
                }
                acc_pk += (uint32_t)hit | ((uint32_t)(hit && tc > 0) << 10) | ((uint32_t)(hit && mapq == 0u) << 20);
                acc_cov += cov_t;
                acc_conv += conv;
                // per-position outputs, once per (read, merged segment).  The reads of a sorted tile start and end at few
                // distinct positions in lane order, so equal addresses sit in runs of neighbouring lanes: the first lane of a
                // run adds the run's length
                const bool part = hit && u.w != last_seg;
                const unsigned pm = __ballot_sync(0xFFFFFFFFu, part);
                if (pm == 0u) continue;
                if (part) last_seg = u.w;
                const uint4 sg = __ldg(ent + 2 * j + 1);   // {segment start, end, posbase lo, hi}: position p lives at index posbase + p
                const unsigned long long posbase = (unsigned long long)sg.z | ((unsigned long long)sg.w << 32);
                const uint32_t pa = sg.x > g ? sg.x : g, pb = sg.y < e ? sg.y : e;
                const bool sticks_out = part && !(pa == g && pb == e);
                if (SD_LEAN_DIAG & 2) {
                } else if (!__any_sync(0xFFFFFFFFu, sticks_out)) {
                    // every read lies inside the segment: ends are starts + l_seq, so runs of equal starts are runs of equal ends
                    const uint32_t prev = __shfl_up_sync(0xFFFFFFFFu, pa, 1);
                    const bool lead = part && !(lane > 0u && ((pm >> (lane - 1u)) & 1u) && prev == pa);
                    const unsigned lmk = __ballot_sync(0xFFFFFFFFu, lead);
                    if (lead) {
                        const unsigned stop = (~pm | lmk) & ~((2u << lane) - 1u);   // first lane after this one that is not in its run
                        const int run = (stop ? __ffs((int)stop) - 1 : 32) - (int)lane;
                        SD_CHECK(posbase + pb < A.n_positions);
                        int *cp = A.pos_cov + (long long)(posbase + pa);
                        atomicAdd(cp, run);
                        atomicAdd(cp + lseq, -run);
                    }
                } else {
#pragma unroll
                    for (int side = 0; side < 2; side++) {
                        const uint32_t p = side ? pb : pa;
                        const uint32_t prev = __shfl_up_sync(0xFFFFFFFFu, p, 1);
                        const bool lead = part && !(lane > 0u && ((pm >> (lane - 1u)) & 1u) && prev == p);
                        const unsigned lmk = __ballot_sync(0xFFFFFFFFu, lead);
                        if (lead) {
                            const unsigned stop = (~pm | lmk) & ~((2u << lane) - 1u);
                            const int run = (stop ? __ffs((int)stop) - 1 : 32) - (int)lane;
                            SD_CHECK(pb >= pa && posbase + p < A.n_positions);
                            atomicAdd(&A.pos_cov[posbase + p], side ? -run : run);
                        }
                    }
                }
                const bool conv_here = part && tc > 0 && !(SD_LEAN_DIAG & 1);
                int *tcbase = A.pos_tc + (long long)(posbase + g);   // may point before the segment; only in-range bits are used
                asm volatile("" : "+l"(tcbase));
                if (__any_sync(0xFFFFFFFFu, conv_here && sticks_out)) {   // some read sticks out of the segment (rare)
                    if (conv_here) {
#pragma unroll
                        for (int k = 0; k < W; k++) {
                            uint32_t h = H[k] & prefix_word(k, (int)(pb - g)) & ~prefix_word(k, (int)(pa - g));
                            while (h) {
                                SD_CHECK((unsigned long long)(tcbase + (32 * k + __ffs((int)h) - 1) - A.pos_tc) < A.n_positions);
                                atomicAdd(tcbase + (32 * k + __ffs((int)h) - 1), 1);
                                h &= h - 1u;
                            }
                        }
                    }
                } else if (conv_here) {
#pragma unroll
                    for (int k = 0; k < W; k++) {
                        uint32_t h = H[k];
                        while (h) {
                            SD_CHECK((unsigned long long)(tcbase + (32 * k + __ffs((int)h) - 1) - A.pos_tc) < A.n_positions);
                            atomicAdd(tcbase + (32 * k + __ffs((int)h) - 1), 1);
                            h &= h - 1u;
                        }
                    }
                }
            }
            if (++cnt_tiles == 1000u) {   // before a 10-bit field of the run counters (or a 32-bit row sum) can overflow
                flush_counts();
                flush_row();
                cur_row = 0xFFFFFFFFu;
            }
        }
        tile = next;
        desc = next_desc;
        r_pos = n_pos; r_rmf = n_rmf; r_xi = n_xi; r_w4 = n_w4;
    }

    flush_counts();
    flush_row();
    unsigned long long tc_total = my_tc_total;
    for (int d = 16; d > 0; d >>= 1) tc_total += __shfl_xor_sync(0xFFFFFFFFu, tc_total, d);
    if (lane == 0 && tc_total) atomicAdd(&A.stats[SD_S_TC_TOTAL], tc_total);
    __syncthreads();
    if (threadIdx.x < 7 && s_stats[threadIdx.x]) atomicAdd(&A.stats[threadIdx.x], (unsigned long long)s_stats[threadIdx.x]);
}

// ---------------------------------------------------------------------------------------
// launcher
// ---------------------------------------------------------------------------------------
template <int W, int QB, bool VERIFY>
int launch_lean(sd_ctx *ctx, const CountArgs &C, cudaStream_t stream, uint32_t *list_out, unsigned int *n_out) {
    LeanArgs A;
    memset(&A, 0, sizeof A);
    A.rec = C.rec;
    A.seq = reinterpret_cast<const unsigned char *>(C.seq);
    A.qual = C.qual;
    A.mp = VERIFY ? reinterpret_cast<const unsigned char *>(C.mp) : nullptr;
    A.tiles = C.tiles;
    A.spans = reinterpret_cast<const sd_tile_span *>(C.spans);
    A.refm = C.refm;
    A.ment = C.ment;
    A.n_tiles = C.n_tiles;
    A.n_chrom = C.n_chrom;
    A.q_cmin = C.q_cmin;
    for (int k = 0; k < 4; k++) A.q_m[k] = C.q_m[k];
    A.conv_threshold = C.conv_threshold;
    A.filter_on = C.filter_on; A.filter_mq = C.filter_mq; A.filter_max_nm = C.filter_max_nm;
    A.filter_min_identity = C.filter_min_identity;
    A.counters = C.counters; A.pos_cov = C.pos_cov; A.pos_tc = C.pos_tc; A.stats = C.stats; A.read_tc = C.read_tc;
    A.tile_counter = C.tile_counter;
    A.leftover = list_out;
    A.n_leftover = n_out;
    A.n_positions = C.n_positions;
    A.n_utr = C.n_utr;
    static int per_sm = 0, device = -1;
    if (device != ctx->device) {
        SD_CUDA(ctx, cudaFuncSetAttribute(sd_lean_kernel<W, QB, VERIFY>, cudaFuncAttributePreferredSharedMemoryCarveout, 0));   // all of it L1
        SD_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sd_lean_kernel<W, QB, VERIFY>, SD_CTA_THREADS, 0));
        if (per_sm < 1) return sd_fail(ctx, SD_ERR_LIMIT, "lean count kernel does not fit on an SM");
        device = ctx->device;
    }
    const uint32_t warps = SD_CTA_THREADS / 32;
    const uint64_t want_ctas = ((uint64_t)A.n_tiles + warps - 1) / warps;
    uint32_t grid = (uint32_t)std::min<uint64_t>(want_ctas, (uint64_t)ctx->num_sms * per_sm);
    if (grid == 0) grid = 1;
    A.claim = (uint32_t)std::min<uint64_t>(SD_LEAN_CLAIM, std::max<uint64_t>(1, (uint64_t)A.n_tiles / ((uint64_t)grid * warps * 8u)));
    sd_lean_kernel<W, QB, VERIFY><<<grid, SD_CTA_THREADS, 0, stream>>>(A);
    ctx->launches++;
    SD_CUDA(ctx, cudaGetLastError());
    return SD_OK;
}
