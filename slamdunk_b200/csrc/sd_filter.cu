// sd_filter.cu -- decision logic of `slamdunk filter` on the device (reference: dunks/filter.py).
//
// Default branch (filter.py:233-256): independent per-read predicates.
// Multimapper-retention branch (filter.py:72-210): a sequential state machine in the reference; here
//   pass 1 (one warp per 32-read tile) classifies every record -- dropped by the identity / NM
//          predicates, unique survivor, or multimapper survivor (mapq 0) -- and materialises its
//          reference interval [pos, end) from the CIGAR;
//   pass 2 (one thread per multimapper group head) replays the group's alignments in file order:
//          a group is a maximal run of consecutive multimapper survivors with the same query name
//          (runs are cut by any unique survivor or a name change, filter.py:115,173-187); the first
//          alignment that hits UTRs seeds the set of UTR names, a later alignment hitting a name
//          outside the set makes the group ambiguous (dropped), otherwise ONE alignment is kept:
//          the last one that hit the most recently inserted UTR name (dumpBufferToBam, filter.py:56-65).
// UTR intervals are [start, stop + 1) as in the reference's interval tree (utils/BedReader.py:28) and
// are visited in BED order (the reference iterates a set; see tests/shims/intervaltree).
#include <algorithm>
#include <vector>

#include "sd_internal.h"

namespace {

struct FilterTables {
    const uint32_t *chrom_begin;   // [n_chrom + 1] range of each chromosome in the sorted table
    const int64_t *start;          // sorted by (chrom, start)
    const int64_t *stop1;          // stop + 1
    const int64_t *maxstop1;       // running maximum of stop1 inside the chromosome
    const uint32_t *name_id;
    const uint32_t *bed_index;     // position of the row in the BED file
    uint32_t n_chrom;
};

struct ReadInfo {      // per record, written by pass 1
    int32_t pos, end;
    uint32_t chrom;
    uint32_t state;    // 0 not a survivor, 1 unique survivor, 2 multimapper survivor
};

#define SD_FILTER_MAX_KEYS 32
#define SD_FILTER_MAX_HITS 32

__global__ void filter_pass1(const sd_read_rec *rec, const uint32_t *cigar, const sd_tile *tiles, uint32_t n_tiles,
                             int multimap_mode, int mq, double min_identity, int max_nm, uint8_t *keep, ReadInfo *info,
                             uint8_t *state_out, uint64_t n_reads, unsigned long long *stats) {
    // grid-stride over tiles; the six run counters stay in per-lane registers over all of a warp's tiles and reach global
    // memory as one atomic per counter and CTA (one per counter and WARP serialised ~2 M same-address atomics at 50 M reads:
    // 3.2 of the kernel's 3.6 ms)
    __shared__ unsigned int s_cnt[6];
    if (threadIdx.x < 6) s_cnt[threadIdx.x] = 0u;
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t warp0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    uint32_t a_mapped = 0, a_unmapped = 0, a_filtered = 0, a_mq = 0, a_id = 0, a_nm = 0;
    for (uint64_t tile = warp0; tile < n_tiles; tile += n_warps) {
    const uint64_t i = tile * 32 + lane;
    const sd_read_rec r = rec[i];
    const uint32_t flags = r.ref_mapq_flag >> 24, mapq = (r.ref_mapq_flag >> 16) & 0xFFu;
    const uint32_t ncig = r.lseq_ncig >> 16, nm = r.nm_nmp & 0xFFFFu;
    uint32_t cs = ncig;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, cs, d);
        if (lane >= (uint32_t)d) cs += t;
    }
    uint32_t reflen = 0;
    if (multimap_mode) {            // the alignment's reference interval is only needed by the UTR lookups of pass 2
        const uint32_t *cg = cigar + tiles[tile].cig_off / 4 + (cs - ncig);
        for (uint32_t k = 0; k < ncig; k++) {
            const uint32_t c = cg[k];
            if ((0x18Du >> (c & 15u)) & 1u) reflen += c >> 4;
        }
    }
    if (reflen == 0) reflen = 1;
    const bool pad = (flags & SD_F_PAD) != 0;
    const bool primary = !(flags & (SD_F_SECONDARY | SD_F_SUPPLEMENTARY));
    const bool unmapped = (flags & SD_F_UNMAPPED) != 0;
    uint32_t state = 0;
    uint32_t c_mapped = 0, c_unmapped = 0, c_filtered = 0, c_mq = 0, c_id = 0, c_nm = 0;
    if (!pad) {
        if (primary) { if (unmapped) c_unmapped = 1; else c_mapped = 1; }
        if (!unmapped) {
            if (!multimap_mode && (int)mapq < mq) c_mq = 1;
            else if ((double)r.xi < min_identity) c_id = 1;
            else if (max_nm > -1 && (int)nm > max_nm) c_nm = 1;
            else if (!multimap_mode) { state = 1; if (primary) c_filtered = 1; }
            else if (mapq != 0) { state = 1; c_filtered = 1; }       // filter.py:189-192 counts every written unique read
            else state = 2;
        }
        keep[i] = state == 1 ? 1 : 0;
    }
    if (multimap_mode) {            // pass 2 replays the groups from these
        ReadInfo o;
        o.pos = r.pos;
        o.end = r.pos + (int32_t)reflen;
        o.chrom = r.ref_mapq_flag & 0xFFFFu;
        o.state = pad ? 3u : state; // 3: padding, ends every scan
        info[i] = o;
    } else if (state_out && i < n_reads) state_out[i] = (uint8_t)state;   // default branch: the survivor class is all there is
    a_mapped += c_mapped; a_unmapped += c_unmapped; a_filtered += c_filtered; a_mq += c_mq; a_id += c_id; a_nm += c_nm;
    }
    a_mapped = __reduce_add_sync(0xFFFFFFFFu, a_mapped);
    a_unmapped = __reduce_add_sync(0xFFFFFFFFu, a_unmapped);
    a_filtered = __reduce_add_sync(0xFFFFFFFFu, a_filtered);
    a_mq = __reduce_add_sync(0xFFFFFFFFu, a_mq);
    a_id = __reduce_add_sync(0xFFFFFFFFu, a_id);
    a_nm = __reduce_add_sync(0xFFFFFFFFu, a_nm);
    if (lane == 0) {
        if (a_mapped) atomicAdd(&s_cnt[SD_S_MAPPED], a_mapped);
        if (a_unmapped) atomicAdd(&s_cnt[SD_S_UNMAPPED], a_unmapped);
        if (a_filtered) atomicAdd(&s_cnt[SD_S_FILTERED], a_filtered);
        if (a_mq) atomicAdd(&s_cnt[SD_S_MQFILTERED], a_mq);
        if (a_id) atomicAdd(&s_cnt[SD_S_IDFILTERED], a_id);
        if (a_nm) atomicAdd(&s_cnt[SD_S_NMFILTERED], a_nm);
    }
    __syncthreads();
    if (threadIdx.x < 6 && s_cnt[threadIdx.x]) atomicAdd(&stats[threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]);
}

// UTR rows hit by [pos, end) on `chrom`, as (bed_index, name_id) sorted by bed_index.  Returns the count
// (capped at SD_FILTER_MAX_HITS; *overflow set when more exist).
__device__ int utr_hits(const FilterTables &T, uint32_t chrom, int64_t pos, int64_t end, uint32_t *bed, uint32_t *name, bool *overflow) {
    if (chrom >= T.n_chrom) return 0;
    const uint32_t b = T.chrom_begin[chrom], e = T.chrom_begin[chrom + 1];
    // first entry with start >= end
    uint32_t lo = b, hi = e;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (T.start[mid] < end) lo = mid + 1; else hi = mid;
    }
    int n = 0;
    for (uint32_t j = lo; j > b; j--) {
        const uint32_t k = j - 1;
        if (T.maxstop1[k] <= pos) break;      // nothing at or before k can reach pos
        if (T.stop1[k] > pos) {
            if (n == SD_FILTER_MAX_HITS) { *overflow = true; break; }
            // insertion sort by BED index
            int p = n++;
            const uint32_t bi = T.bed_index[k], ni = T.name_id[k];
            while (p > 0 && bed[p - 1] > bi) { bed[p] = bed[p - 1]; name[p] = name[p - 1]; p--; }
            bed[p] = bi;
            name[p] = ni;
        }
    }
    return n;
}

__global__ void filter_pass2(const ReadInfo *info, const uint64_t *name_hash, uint64_t n, FilterTables T, uint8_t *keep,
                             uint8_t *utr_hit, unsigned long long *stats) {
    // every thread stays to the end: the retained groups are counted per CTA (one global atomic per CTA, not per group)
    __shared__ unsigned int s_kept;
    if (threadIdx.x == 0) s_kept = 0u;
    __syncthreads();
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool head = i < n && info[i].state == 2;
    // group head?  the previous survivor must be unique or carry another name (filter.py:115,173)
    if (head) {
        uint64_t j = i;
        while (j > 0) {
            j--;
            const uint32_t st = info[j].state;
            if (st == 0) continue;
            head = (st != 2) || (name_hash[j] != name_hash[i]);
            break;
        }
    }
    bool retained = false;
    if (head) {
    uint32_t keys[SD_FILTER_MAX_KEYS];
    uint64_t last[SD_FILTER_MAX_KEYS];
    int n_keys = 0;
    bool dump = true, overflow = false;
    const uint64_t h0 = name_hash[i];
    for (uint64_t k = i; k < n; k++) {
        const ReadInfo r = info[k];
        if (r.state == 0) continue;
        if (r.state != 2 || name_hash[k] != h0) break;
        uint32_t bed[SD_FILTER_MAX_HITS], name[SD_FILTER_MAX_HITS];
        const int nh = utr_hits(T, r.chrom, r.pos, r.end, bed, name, &overflow);
        if (nh > 0 && utr_hit) utr_hit[k] = 4u;   // remembered for the RD tag text (every multimapper belongs to one group)
        const bool first_group = (n_keys == 0);                       // filter.py:147
        for (int h = 0; h < nh; h++) {
            int found = -1;
            for (int q = 0; q < n_keys; q++)
                if (keys[q] == name[h]) { found = q; break; }
            if (found < 0) {
                if (n_keys == SD_FILTER_MAX_KEYS) { overflow = true; continue; }
                keys[n_keys] = name[h];
                last[n_keys] = k;
                n_keys++;
                if (!first_group) dump = false;                       // filter.py:154-158
            } else last[found] = k;
        }
    }
    if (overflow) atomicAdd(&stats[SD_S_SLOWPATH], 1ull);             // host treats this as an error
    if (dump && n_keys > 0) {                                         // filter.py:118-120,178-180,197-199
        keep[last[n_keys - 1]] = 2;
        retained = true;
    }
    }
    const unsigned int kept = __popc(__ballot_sync(0xFFFFFFFFu, retained));
    if ((threadIdx.x & 31u) == 0 && kept) atomicAdd(&s_kept, kept);
    __syncthreads();
    if (threadIdx.x == 0 && s_kept) atomicAdd(&stats[SD_S_FILTERED], (unsigned long long)s_kept);
}

__global__ void export_state(const ReadInfo *info, uint64_t n, uint8_t *state) {   // state[] arrives holding the UTR-hit bits
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) state[i] = (uint8_t)((info[i].state == 3u ? 0u : info[i].state) | state[i]);
}

template <class T>
int to_device(sd_ctx *ctx, T **dst, const std::vector<T> &v) {
    SD_CUDA(ctx, cudaMalloc(dst, std::max<size_t>(1, v.size()) * sizeof(T)));
    if (v.size()) SD_CUDA(ctx, cudaMemcpyAsync(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    return SD_OK;
}

}  // namespace

extern "C" int sd_filter(sd_ctx *ctx, const sd_batch *b, const sd_params *p, const uint64_t *d_name_hash,
                         const uint32_t *h_utr_chrom, const int64_t *h_utr_start, const int64_t *h_utr_stop,
                         const uint32_t *h_utr_name_id, uint32_t n_utr, uint8_t *d_keep, int64_t *d_stats, uint8_t *d_state) {
    if (!ctx || !b || !p || !d_keep || !d_stats) return sd_fail(ctx, SD_ERR_ARG, "sd_filter: null argument");
    if (b->tile_reads != 32) return sd_fail(ctx, SD_ERR_ARG, "tile_reads must be 32");
    const bool multimap = h_utr_chrom != nullptr;
    if (multimap && (!d_name_hash || !h_utr_start || !h_utr_stop || !h_utr_name_id))
        return sd_fail(ctx, SD_ERR_ARG, "sd_filter: multimapper mode needs name hashes and the UTR table");
    SD_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint64_t n_pad = (uint64_t)b->n_tiles * 32;
    SD_CUDA(ctx, cudaMemsetAsync(d_stats, 0, SD_NSTAT * sizeof(int64_t), ctx->stream));
    SD_CUDA(ctx, cudaMemsetAsync(d_keep, 0, std::max<uint64_t>(1, b->n_reads), ctx->stream));
    if (d_state) SD_CUDA(ctx, cudaMemsetAsync(d_state, 0, std::max<uint64_t>(1, b->n_reads), ctx->stream));
    if (b->n_tiles == 0) { SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); return SD_OK; }
    // per-record scratch (16-byte info + padded keep flags) lives in a grow-only buffer of the context: `slamdunk filter`
    // runs one BAM after the other, and allocating ~17 bytes per read on every call cost more than the kernels
    const size_t need = (size_t)n_pad * sizeof(ReadInfo) + (size_t)n_pad;
    if (need > ctx->filter_scratch_cap) {
        cudaFree(ctx->d_filter_scratch);
        ctx->d_filter_scratch = nullptr;
        ctx->filter_scratch_cap = 0;
        SD_CUDA(ctx, cudaMalloc(&ctx->d_filter_scratch, need));
        ctx->filter_scratch_cap = need;
    }
    ReadInfo *info = reinterpret_cast<ReadInfo *>(ctx->d_filter_scratch);
    uint8_t *keep_pad = reinterpret_cast<uint8_t *>(ctx->d_filter_scratch) + (size_t)n_pad * sizeof(ReadInfo);
    SD_CUDA(ctx, cudaMemsetAsync(keep_pad, 0, n_pad, ctx->stream));
    unsigned long long *st = reinterpret_cast<unsigned long long *>(d_stats);
    const unsigned p1_blocks = (unsigned)std::min<uint64_t>((n_pad + 255) / 256, (uint64_t)ctx->num_sms * 16u);
    filter_pass1<<<p1_blocks, 256, 0, ctx->stream>>>(b->rec, b->cigar, b->tiles, b->n_tiles, multimap ? 1 : 0,
                                                                          p->filter_mq, p->filter_min_identity, p->filter_max_nm,
                                                                          keep_pad, info, d_state, b->n_reads, st);
    ctx->launches++;
    int rc = SD_OK;
    uint32_t *d_cb = nullptr, *d_nm = nullptr, *d_bi = nullptr;
    int64_t *d_s = nullptr, *d_e = nullptr, *d_m = nullptr;
    if (multimap) {
        uint32_t n_chrom = 0;
        for (uint32_t i = 0; i < n_utr; i++)
            if (h_utr_chrom[i] != SD_REF_NONE) n_chrom = std::max(n_chrom, h_utr_chrom[i] + 1);
        std::vector<uint32_t> order;
        for (uint32_t i = 0; i < n_utr; i++)
            if (h_utr_chrom[i] != SD_REF_NONE && h_utr_stop[i] + 1 > h_utr_start[i]) order.push_back(i);
        std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) {
            return h_utr_chrom[x] < h_utr_chrom[y] || (h_utr_chrom[x] == h_utr_chrom[y] && h_utr_start[x] < h_utr_start[y]);
        });
        std::vector<uint32_t> cb(n_chrom + 1, 0), nm(order.size()), bi(order.size());
        std::vector<int64_t> s(order.size()), e(order.size()), m(order.size());
        for (size_t k = 0; k < order.size(); k++) {
            const uint32_t i = order[k];
            s[k] = h_utr_start[i];
            e[k] = h_utr_stop[i] + 1;
            nm[k] = h_utr_name_id[i];
            bi[k] = i;
            cb[h_utr_chrom[i] + 1]++;
        }
        for (uint32_t c = 0; c < n_chrom; c++) cb[c + 1] += cb[c];
        for (uint32_t c = 0; c < n_chrom; c++)
            for (uint32_t k = cb[c]; k < cb[c + 1]; k++) m[k] = (k == cb[c]) ? e[k] : std::max(m[k - 1], e[k]);
        if ((rc = to_device(ctx, &d_cb, cb)) || (rc = to_device(ctx, &d_s, s)) || (rc = to_device(ctx, &d_e, e)) ||
            (rc = to_device(ctx, &d_m, m)) || (rc = to_device(ctx, &d_nm, nm)) || (rc = to_device(ctx, &d_bi, bi))) {
            return rc;
        }
        FilterTables T{d_cb, d_s, d_e, d_m, d_nm, d_bi, n_chrom};
        filter_pass2<<<(unsigned)((n_pad + 127) / 128), 128, 0, ctx->stream>>>(info, d_name_hash, b->n_reads, T, keep_pad, d_state, st);
        ctx->launches++;
    }
    SD_CUDA(ctx, cudaMemcpyAsync(d_keep, keep_pad, b->n_reads, cudaMemcpyDeviceToDevice, ctx->stream));
    if (d_state && multimap) {
        export_state<<<(unsigned)((b->n_reads + 255) / 256), 256, 0, ctx->stream>>>(info, b->n_reads, d_state);
        ctx->launches++;
    }
    cudaError_t e1 = cudaGetLastError();
    cudaError_t e2 = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_cb); cudaFree(d_s); cudaFree(d_e); cudaFree(d_m); cudaFree(d_nm); cudaFree(d_bi);
    if (e1 != cudaSuccess || e2 != cudaSuccess)
        return sd_fail(ctx, SD_ERR_CUDA, "sd_filter: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
    return SD_OK;
}
