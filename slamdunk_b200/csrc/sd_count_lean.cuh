// sd_count_lean.cuh -- the count kernel for the common tile: 32 coordinate-sorted reads that are all one match block of
// the same length on one chromosome (sd_tile_span.lean, decided when the batch's span table is made).  Everything such a
// tile does not need is compiled out: no CIGAR, no row offsets, no per-read chromosome lookups, no placement of the read
// mask, mismatches from the CIGAR walk only (SD_MISMATCH_CIGAR, or SD_MISMATCH_VERIFY through VERIFY = true), 2-bit SEQ column
// and plane-form qualities only.  Tiles that are not of that kind go, untouched, to the general kernels that follow
// (sd_count_kernel<..., true> staged, then <..., false> plain).  Same outputs, bit for bit.
//
// Per warp: a ring of two shared-memory stages filled by TMA bulk copies (records, SEQ codes, quality planes, the tile's
// reference window and annotation entries -- one copy per lane 0..4), one thread per read:
//   filter predicates (dunks/filter.py:235-256)  ->  read mask "is C / is G and quality >= minQual" (SlamSeqFile.py:137-141, 330)
//   AND reference planes "is T / is A, not a SNP" of the read's window (SlamSeqFile.py:333-344)  ->  tcCount, T coverage
//   ->  warp-uniform walk over the tile's annotation entries: per-UTR counters as warp votes / reductions and at most five
//   atomics (tcounter.py:216-248, 277-283), per-position coverage differences and conversions (tcounter.py:229-231, 246-248).
#pragma once
#include "sd_count_kernel.cuh"

struct LeanArgs {
    const sd_read_rec *rec;
    const unsigned char *seq, *qual, *mp;   // 2-bit SEQ column, plane-form qualities, MP entries (VERIFY) or null
    const sd_tile *tiles;
    const sd_tile_span *spans;
    const unsigned char *refm, *ment;
    uint32_t n_tiles, n_chrom;
    uint32_t cap_seq, cap_qual, cap_win, cap_ent, cap_mp;   // stage layout: rec | seq | qual | win | ent | mp, multiples of 128
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

#define SD_LEAN_STAGES 2
#ifndef SD_LEAN_MIN_CTAS
#define SD_LEAN_MIN_CTAS(W) SD_MIN_CTAS(W)
#endif
#define SD_LEAN_REC_BYTES (SD_TILE_READS * 20u)

template <int W, int QB, bool VERIFY>
__global__ void __launch_bounds__(SD_CTA_THREADS, SD_LEAN_MIN_CTAS(W)) sd_lean_kernel(const __grid_constant__ LeanArgs A) {
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t off_seq = SD_LEAN_REC_BYTES, off_qual = off_seq + A.cap_seq, off_win = off_qual + A.cap_qual, off_ent = off_win + A.cap_win,
                   off_mp = off_ent + A.cap_ent;
    const uint32_t stage_bytes = off_mp + A.cap_mp;
    uint32_t *s_stats = reinterpret_cast<uint32_t *>(smem);   // [SD_SMEM_STATS]
    unsigned char *mine = smem + 128u + (size_t)warp * (SD_LEAN_STAGES * stage_bytes + 128u);
    uint64_t *full = reinterpret_cast<uint64_t *>(mine + SD_LEAN_STAGES * stage_bytes);   // [2]
    uint4 *s_info = reinterpret_cast<uint4 *>(mine + SD_LEAN_STAGES * stage_bytes + 32u);   // [2][2] {lean, g0, win lo, win hi}, {ent lo, ent hi, tile, 0} of the tile in each stage
    if (threadIdx.x < SD_SMEM_STATS) s_stats[threadIdx.x] = 0u;
    if (lane == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        mbar_fence_init();
    }
    __syncthreads();

    // --- per-lane copy constants: lane c < 5 (6 with VERIFY) moves column c of a tile into the stage
    const unsigned char *my_src = reinterpret_cast<const unsigned char *>(A.rec);
    uint32_t my_dst = 0u;
    if (lane == 1) { my_src = A.seq; my_dst = off_seq; }
    if (lane == 2) { my_src = A.qual; my_dst = off_qual; }
    if (lane == 3) { my_src = A.refm; my_dst = off_win; }
    if (lane == 4) { my_src = A.ment; my_dst = off_ent; }
    if (lane == 5) { my_src = A.mp; my_dst = off_mp; }
    constexpr uint32_t NCOPY = VERIFY ? 6u : 5u;
    constexpr unsigned COPY_MASK = (1u << NCOPY) - 1u;

    // tile queue: chunks of SD_CLAIM consecutive tiles per atomic.  Every lane keeps the chunk's bounds (warp-uniform); the atomic for
    // the chunk AFTER the current one is already in flight (lane 0's `pend`), so taking a new chunk never waits for the queue head.
    uint32_t claim_next, claim_end, pend = 0u;
    {
        uint32_t c = 0u;
        if (lane == 0) c = atomicAdd(A.tile_counter, (unsigned int)SD_CLAIM);
        claim_next = __shfl_sync(0xFFFFFFFFu, c, 0);
        claim_end = claim_next + SD_CLAIM;
        if (lane == 0) pend = atomicAdd(A.tile_counter, (unsigned int)SD_CLAIM);
    }
    auto claim = [&]() -> uint32_t {
        if (claim_next == claim_end) {   // warp-uniform
            claim_next = __shfl_sync(0xFFFFFFFFu, pend, 0);
            claim_end = claim_next + SD_CLAIM;
            if (lane == 0 && claim_next < A.n_tiles) pend = atomicAdd(A.tile_counter, (unsigned int)SD_CLAIM);
        }
        const uint32_t t = claim_next++;
        return t < A.n_tiles ? t : 0xFFFFFFFFu;
    };
    // A tile's description is five (six / seven with VERIFY) 64-bit words in two tables; lane c fetches word c as soon as the tile is
    // claimed -- one tile ahead of the copies that need it -- and keeps it in a register:
    //   0 spans[t].{lean, g0}   1 tiles[t].seq_off   2 tiles[t].qual_off   3 spans[t].win   4 spans[t].ent   5, 6 tiles[t], tiles[t + 1].mp_off
    auto fetch_desc = [&](uint32_t tile) -> unsigned long long {
        if (tile == 0xFFFFFFFFu || lane >= NCOPY + (VERIFY ? 1u : 0u)) return 0ull;
        const bool in_tiles = lane == 1u || lane == 2u || lane >= 5u;
        const unsigned char *base = in_tiles ? reinterpret_cast<const unsigned char *>(A.tiles) : reinterpret_cast<const unsigned char *>(A.spans);
        //                      lane:     6    5    4    3    2    1    0
        const uint32_t field = (uint32_t)((0x38182820080030ull >> (8u * lane)) & 0xFFull);   // byte offset of the word in its table entry
        return __ldg(reinterpret_cast<const unsigned long long *>(base + (size_t)tile * (in_tiles ? sizeof(sd_tile) : sizeof(sd_tile_span)) + field));
    };
    // Issue the copies of `tile` (description words `x`, one per lane, from fetch_desc) into `stage`.  Returns lq = l_seq of a lean
    // tile, 0 for any other -- no copies are made and the consumer hands the tile on without waiting.  The description -- {lean, g0},
    // win, ent, the tile index -- waits in s_info[stage] for the consumer.
    auto issue = [&](uint32_t tile, uint32_t stage, unsigned long long x) -> uint32_t {
        const uint32_t lseq = __shfl_sync(0xFFFFFFFFu, (uint32_t)x, 0);
        unsigned long long *info = reinterpret_cast<unsigned long long *>(&s_info[2u * stage]);
        if (lane == 0u) info[3] = tile;
        if (lseq == 0u) { __syncwarp(); return 0u; }   // warp-uniform
        if (lane == 0u || lane == 3u || lane == 4u) info[lane == 0u ? 0 : lane - 2u] = x;   // {lean, g0} | win | ent
        // every lane goes through the same instructions (lanes without a column copy nothing): the warp stays converged for the reduction
        const uint32_t G = (lseq + 31u) >> 5;
        unsigned long long off = x;
        uint32_t bytes = 0u;
        if (lane == 0) { off = (unsigned long long)tile * SD_LEAN_REC_BYTES; bytes = SD_LEAN_REC_BYTES; }
        if (lane == 1) bytes = G * (8u * SD_TILE_READS);
        if (lane == 2) bytes = G * (4u * QB * SD_TILE_READS);
        if (lane == 3 || lane == 4) { off = x >> 16; bytes = (uint32_t)x & 0xFFFFu; }
        if (VERIFY) {
            const unsigned long long mp_end = __shfl_sync(0xFFFFFFFFu, x, 6);
            if (lane == 5) bytes = (uint32_t)(mp_end - x);
        }
        const uint32_t total = __reduce_add_sync(0xFFFFFFFFu, bytes);
        if (lane == 0) mbar_arrive_expect_tx(&full[stage], total);
        __syncwarp();
        if (bytes) bulk_g2s(mine + stage * stage_bytes + my_dst, my_src + off, bytes, &full[stage]);
        return lseq;
    };

    // two tiles in flight; lq_[st] = l_seq of the (lean) tile in stage st, 0 = not a lean tile, 0xFFFFFFFF = the queue ran dry
    uint32_t lq_[2];
#pragma unroll
    for (int st = 0; st < 2; st++) {
        const uint32_t t = claim();
        lq_[st] = t != 0xFFFFFFFFu ? issue(t, (uint32_t)st, fetch_desc(t)) : 0xFFFFFFFFu;
    }

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

    // Counters of the annotation row the warp is in: consecutive tiles of a sorted batch hit the same UTR, so its five sums are kept
    // in (warp-uniform) registers and go to global memory when the row changes -- five atomics per row change instead of per tile.
    uint32_t cur_row = 0xFFFFFFFFu, acc_n = 0u, acc_tc = 0u, acc_mm = 0u, acc_cov = 0u, acc_conv = 0u;
    auto flush_row = [&]() {
        if (cur_row != 0xFFFFFFFFu && lane == 0u) {
            SD_CHECK(cur_row < A.n_utr);
            unsigned long long *c = A.counters + (size_t)cur_row * SD_NCOUNTER;
            atomicAdd(c + SD_C_READS, (unsigned long long)acc_n);
            if (acc_tc) atomicAdd(c + SD_C_TCREADS, (unsigned long long)acc_tc);
            if (acc_mm) atomicAdd(c + SD_C_MULTIMAP, (unsigned long long)acc_mm);
            if (acc_cov) atomicAdd(c + SD_C_COV_ON_T, (unsigned long long)acc_cov);
            if (acc_conv) atomicAdd(c + SD_C_CONV_ON_T, (unsigned long long)acc_conv);
        }
        acc_n = acc_tc = acc_mm = acc_cov = acc_conv = 0u;
    };

    uint32_t phase = 0u;   // bit st: parity the next wait on stage st looks for (a stage's barrier advances only when its tile was copied)
    for (;;) {
        // ------------------------------------------------------------------ stage 0, then stage 1 (unrolled by hand via the lambda)
        bool done = false;
#pragma unroll
        for (int st = 0; st < SD_LEAN_STAGES; st++) {
            if (lq_[st] == 0xFFFFFFFFu) { done = true; break; }   // claims are monotonic: nothing behind it either
            const uint32_t lseq = lq_[st];
            // the tile that will take this stage over: claimed now, its description on the way into registers while this tile is counted
            const uint32_t next = claim();
            const unsigned long long next_desc = fetch_desc(next);
            if (lseq == 0u) {
                if (lane == 0) A.leftover[atomicAdd(A.n_leftover, 1u)] = s_info[2 * st + 1].z;
            } else {
                const unsigned char *base = mine + st * stage_bytes;
                mbar_wait(&full[st], (phase >> st) & 1u);
                phase ^= 1u << st;
                const uint4 info = s_info[2 * st], info2 = s_info[2 * st + 1];   // {lean, g0, win lo, win hi}, {ent lo, ent hi, tile, .}
                const uint32_t tile = info2.z, ne = (info2.x & 0xFFFFu) >> 5;
                const uint32_t w0 = (info.z >> 20) | (info.w << 12);   // win = first window byte << 16 | bytes; 16 bytes per word
                // --- the record
                const uint32_t *rw = reinterpret_cast<const uint32_t *>(base) + 5u * lane;
                const int32_t pos = (int32_t)rw[0];
                const uint32_t rmf = rw[1];
                const float xi = __uint_as_float(rw[2]);
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
                        f_nm = mapped && !f_mq && !f_id && (int)(rw[4] & 0xFFFFu) > A.filter_max_nm;
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
                const uint32_t g = info.y + (uint32_t)pos, e = g + lseq;
                const uint32_t dword = live ? (g >> 5) - w0 : 0u;   // window word of the read among the staged ones
                const uint32_t sh = g & 31u;
                // --- reference window of this strand, shifted to the read's start: T = is T / A inside the read, S = SNPs
                uint32_t T[W], S[W];
                {
                    const uint2 *wp = reinterpret_cast<const uint2 *>(base + off_win) + 2u * dword + s;
                    uint2 x[W + 1];
#pragma unroll
                    for (int k = 0; k <= W; k++) x[k] = wp[2 * k];
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
                    const uint2 *sq = reinterpret_cast<const uint2 *>(base + off_seq) + lane;
                    const uint32_t *ql = reinterpret_cast<const uint32_t *>(base + off_qual) + QB * lane;
                    const uint32_t hm = (flags & SD_F_HAS_MP) ? 0xFFFFFFFFu : 0u;   // no MP tag: the reference sees no mismatches (SlamSeqFile.py:318)
#pragma unroll
                    for (int k = 0; k < W; k++) {
                        // groups behind the read's last one hold whatever an earlier tile left in the stage: T is zero there
                        const uint32_t m = group_is_base2(sq[k * SD_TILE_READS], sm);
                        const uint32_t q = row_group_qual_ok<QB, true>(ql, (uint32_t)k, A);
                        if (VERIFY) D[k] = m & T[k] & hm;   // before quality / SNP masking: what NextGenMap lists
                        H[k] = m & q & T[k] & ~S[k] & hm;
                    }
                }
                bool bad = false;
                if (VERIFY) {
                    // the read's T>C / A>G candidates against its MP entries (reference offset and read index; plain match: equal)
                    const uint32_t nmp_all = live ? (rw[4] >> 16) : 0u;
                    uint32_t ms = rw[4] >> 16;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, ms, d);
                        if (lane >= (uint32_t)d) ms += t;
                    }
                    const uint32_t *mp = reinterpret_cast<const uint32_t *>(base + off_mp) + (ms - (rw[4] >> 16));
                    if (flags & SD_F_HAS_MP) {
                        for (uint32_t i = 0; i < nmp_all; i++) {
                            const uint32_t en = mp[i];
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
                const uint4 *ent = reinterpret_cast<const uint4 *>(base + off_ent);
                uint32_t last_seg = 0xFFFFFFFFu;
                for (uint32_t j = 0; j < ne; j++) {
                    const uint4 u = ent[2 * j];   // {start, end, row | strand << 31, segment}: broadcast
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
                    const unsigned b_tc = __ballot_sync(0xFFFFFFFFu, hit && tc > 0), b_mm = __ballot_sync(0xFFFFFFFFu, hit && mapq == 0u);
                    const uint32_t s_cov = __reduce_add_sync(0xFFFFFFFFu, cov_t), s_conv = __reduce_add_sync(0xFFFFFFFFu, conv);
                    const uint32_t row = u.z & 0x7FFFFFFFu;
                    if (row != cur_row) {   // warp-uniform
                        flush_row();
                        cur_row = row;
                    }
                    acc_n += __popc(hmask);
                    acc_tc += __popc(b_tc);
                    acc_mm += __popc(b_mm);
                    acc_cov += s_cov;
                    acc_conv += s_conv;
                    // per-position outputs, once per (read, merged segment).  The reads of a sorted tile start and end at few
                    // distinct positions in lane order, so equal addresses sit in runs of neighbouring lanes: the first lane of a
                    // run adds the run's length
                    const bool part = hit && u.w != last_seg;
                    const unsigned pm = __ballot_sync(0xFFFFFFFFu, part);
                    if (pm == 0u) continue;
                    if (part) last_seg = u.w;
                    const uint4 sg = ent[2 * j + 1];   // {segment start, end, posbase lo, hi}: position p lives at index posbase + p
                    const unsigned long long posbase = (unsigned long long)sg.z | ((unsigned long long)sg.w << 32);
                    const uint32_t pa = sg.x > g ? sg.x : g, pb = sg.y < e ? sg.y : e;
                    const bool sticks_out = part && !(pa == g && pb == e);
                    if (!__any_sync(0xFFFFFFFFu, sticks_out)) {
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
                    const bool conv_here = part && tc > 0;
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
            // --- refill this stage
            __syncwarp();
            lq_[st] = next != 0xFFFFFFFFu ? issue(next, (uint32_t)st, next_desc) : 0xFFFFFFFFu;
        }
        if (done) break;
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
    A.refm = reinterpret_cast<const unsigned char *>(C.refm);
    A.ment = reinterpret_cast<const unsigned char *>(C.ment);
    A.n_tiles = C.n_tiles;
    A.n_chrom = C.n_chrom;
    A.cap_seq = align128(8u * SD_TILE_READS * W);
    A.cap_qual = align128(4u * QB * SD_TILE_READS * W);
    A.cap_win = C.cap_win;
    A.cap_ent = C.cap_ent;
    A.cap_mp = VERIFY ? align128(C.cap_mp) : 0u;
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
    const uint32_t stage_bytes = SD_LEAN_REC_BYTES + A.cap_seq + A.cap_qual + A.cap_win + A.cap_ent + A.cap_mp;
    int max_smem = 0;
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device);
    uint32_t warps = SD_CTA_THREADS / 32;
    auto need = [&](uint32_t w) { return 128u + (size_t)w * (SD_LEAN_STAGES * stage_bytes + 128u); };
    while (warps > 1 && need(warps) > (size_t)max_smem / 2) warps--;   // at least two CTAs per SM
    if (need(warps) > (size_t)max_smem) return sd_fail(ctx, SD_ERR_LIMIT, "a lean tile needs %u bytes of shared memory per stage", stage_bytes);
    const size_t smem = need(warps);
    struct Planned { uint32_t stage_bytes; int device; int per_sm; };
    static Planned last = {0, -1, 0};
    if (last.device != ctx->device || last.stage_bytes != stage_bytes) {
        int per_sm = 0;
        SD_CUDA(ctx, cudaFuncSetAttribute(sd_lean_kernel<W, QB, VERIFY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SD_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sd_lean_kernel<W, QB, VERIFY>, (int)warps * 32, smem));
        if (per_sm < 1) return sd_fail(ctx, SD_ERR_LIMIT, "lean count kernel does not fit on an SM (smem %zu)", smem);
        last = {stage_bytes, ctx->device, per_sm};
    }
    const uint64_t want_ctas = ((uint64_t)A.n_tiles + warps - 1) / warps;
    uint32_t grid = (uint32_t)std::min<uint64_t>(want_ctas, (uint64_t)ctx->num_sms * last.per_sm);
    if (grid == 0) grid = 1;
    sd_lean_kernel<W, QB, VERIFY><<<grid, (int)warps * 32, smem, stream>>>(A);
    ctx->launches++;
    SD_CUDA(ctx, cudaGetLastError());
    return SD_OK;
}
