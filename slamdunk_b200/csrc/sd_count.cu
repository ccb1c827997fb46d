// sd_count.cu -- the `slamdunk count` hot path for B200 (sm_100a).
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
#include "sd_count_kernel.cuh"

#include <atomic>

// the launcher is instantiated in sd_count_inst_*.cu
#define SD_EXTERN_W(SB, QB)                                                                  \
    extern template int launch_count<4, SB, QB>(sd_ctx *, CountArgs &, cudaStream_t, bool);  \
    extern template int launch_count<6, SB, QB>(sd_ctx *, CountArgs &, cudaStream_t, bool);  \
    extern template int launch_count<8, SB, QB>(sd_ctx *, CountArgs &, cudaStream_t, bool);  \
    extern template int launch_count<10, SB, QB>(sd_ctx *, CountArgs &, cudaStream_t, bool); \
    extern template int launch_count<12, SB, QB>(sd_ctx *, CountArgs &, cudaStream_t, bool); \
    extern template int launch_count<16, SB, QB>(sd_ctx *, CountArgs &, cudaStream_t, bool);
SD_EXTERN_W(4, 0)
SD_EXTERN_W(2, 0)
SD_EXTERN_W(2, 2)
SD_EXTERN_W(2, 4)
#undef SD_EXTERN_W

// ---------------------------------------------------------------------------------------
// small kernels: reference planes, finalize scan, Tcontent
// ---------------------------------------------------------------------------------------
__global__ void sd_refx_kernel(const uint32_t *lo, const uint32_t *hi, const uint32_t *nm, uint64_t n_words,
                               uint2 *rx_plus, uint2 *rx_minus, uint4 *rm) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_words + SD_REF_PAD_WORDS) return;
    uint32_t t = 0, a = 0;
    if (i < n_words) {
        const uint32_t l = lo[i], h = hi[i], n = nm[i];
        t = l & h & ~n;
        a = ~l & ~h & ~n;
    }
    rx_plus[i] = make_uint2(t, 0u);
    rx_minus[i] = make_uint2(a, 0u);
    rm[i] = make_uint4(t, 0u, a, 0u);
}

__global__ void sd_snp_kernel(const uint32_t *tc, const uint32_t *ag, uint64_t n_words, uint2 *rx_plus, uint2 *rx_minus, uint4 *rm) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_words) return;
    const uint32_t t = tc ? tc[i] : 0u, a = ag ? ag[i] : 0u;
    rx_plus[i].y = t;
    rx_minus[i].y = a;
    rm[i].y = t;
    rm[i].w = a;
}

// segment-local inclusive prefix sum of the coverage differences; one warp per segment
__global__ void sd_finalize_kernel(const uint64_t *seg_off0, const uint32_t *seg_start0, const uint32_t *seg_end0, uint32_t n0,
                                   const uint64_t *seg_off1, const uint32_t *seg_start1, const uint32_t *seg_end1, uint32_t n1,
                                   int *pos_cov) {
    const uint32_t wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (wid >= n0 + n1) return;
    uint64_t off;
    uint32_t len;
    if (wid < n0) {
        off = seg_off0[wid];
        len = seg_end0[wid] - seg_start0[wid] + 1u;
    } else {
        off = seg_off1[wid - n0];
        len = seg_end1[wid - n0] - seg_start1[wid - n0] + 1u;
    }
    int carry = 0;
    for (uint32_t b = 0; b < len; b += 32) {
        const uint32_t i = b + lane;
        int v = i < len ? pos_cov[off + i] : 0;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xFFFFFFFFu, v, d);
            if (lane >= (uint32_t)d) v += t;
        }
        v += carry;
        if (i < len) pos_cov[off + i] = v;
        carry = __shfl_sync(0xFFFFFFFFu, v, 31);
    }
}

__global__ void sd_tcontent_kernel(const uint2 *rx_plus, const uint2 *rx_minus, const uint32_t *gstart, const uint32_t *gend,
                                   const uint8_t *strand, uint32_t n, long long *out) {
    const uint32_t wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (wid >= n) return;
    const uint32_t a = gstart[wid], b = gend[wid];
    const uint2 *rx = strand[wid] ? rx_minus : rx_plus;
    uint32_t cnt = 0;
    if (b > a) {
        const uint32_t w0 = a >> 5, w1 = (b - 1) >> 5;
        for (uint32_t w = w0 + lane; w <= w1; w += 32) {
            uint32_t x = rx[w].x;
            if (w == w0) x &= ~((1u << (a & 31)) - 1u);
            if (w == w1 && (b & 31)) x &= (1u << (b & 31)) - 1u;
            cnt += __popc(x);
        }
    }
    cnt = __reduce_add_sync(0xFFFFFFFFu, cnt);
    if (lane == 0) out[wid] = (long long)cnt;
}

// Span descriptor of every tile (sd_batch.spans): the stretch of the linear genome its countable reads cover and the
// run of strand-merged annotation entries that can overlap it.  One warp per tile, one lane per read.
struct SpanArgs {
    const sd_read_rec *rec;
    const uint32_t *cigar;
    const sd_tile *tiles;
    uint32_t n_tiles;
    const uint32_t *chrom_off, *chrom_len;
    uint32_t n_chrom;
    const uint4 *ment;
    const uint32_t *bin_first_m;
    uint32_t n_ment;
    sd_tile_span *out;
    uint64_t n_refm;
    uint32_t W, cap_win, cap_ent;   // the count kernels' window words and stage room for this batch (span_caps)
    unsigned int *counts;           // optional: [0] tiles the lean kernel takes, [1] tiles that fit a stage
};

__global__ void sd_tile_spans_kernel(const SpanArgs A) {
    const uint32_t tile = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (tile >= A.n_tiles) return;
    const sd_read_rec r = A.rec[(size_t)tile * 32u + lane];
    const uint32_t flags = r.ref_mapq_flag >> 24, ref = r.ref_mapq_flag & 0xFFFFu;
    const uint32_t lseq = r.lseq_ncig & 0xFFFFu, ncig = r.lseq_ncig >> 16;
    const unsigned long long c0 = A.tiles[tile].cig_off, c1 = A.tiles[tile + 1].cig_off;
    const bool nocig = c1 == c0;
    uint32_t xs = ncig;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, xs, d);
        if (lane >= (uint32_t)d) xs += t;
    }
    const bool cand = !(flags & (SD_F_PAD | SD_F_UNMAPPED | SD_F_SKIP)) && ref < A.n_chrom;   // what the lean kernel takes for a countable record
    bool countable = cand && !(nocig && ncig > 1u);
    uint32_t clen = 0, g = 0xFFFFFFFFu, e = 0u;
    bool alike = true;   // lean: one match block of l_seq bases, inside the chromosome
    if (cand) {
        clen = A.chrom_len[ref];
        alike = ncig == 1u && lseq >= 1u && r.pos >= 0 && (uint64_t)r.pos + lseq <= clen;
        if (alike && !nocig) {
            const uint32_t c = A.cigar[c0 / 4u + (xs - ncig)];
            alike = ((0x181u >> (c & 15u)) & 1u) && (c >> 4) == lseq;
        }
    }
    if (countable) countable = r.pos >= 0 && (uint32_t)r.pos < clen;
    if (countable) {
        uint32_t reflen = 0;
        if (nocig) reflen = ncig ? lseq : 0u;
        else {
            const uint32_t *cg = A.cigar + c0 / 4u + (xs - ncig);
            for (uint32_t i = 0; i < ncig; i++) {
                const uint32_t c = cg[i];
                if ((0x18Du >> (c & 15u)) & 1u) reflen += c >> 4;
            }
        }
        if (reflen == 0) reflen = 1;
        if ((uint64_t)r.pos + reflen > clen) reflen = clen - (uint32_t)r.pos;
        g = A.chrom_off[ref] + (uint32_t)r.pos;
        e = g + reflen;
    }
    const uint32_t gmin = __reduce_min_sync(0xFFFFFFFFu, g), gend = __reduce_max_sync(0xFFFFFFFFu, e);
    // {window begin, end, entries begin, end} as byte offsets into the strand-merged reference and annotation tables
    sd_tile_span out;
    memset(&out, 0, sizeof out);   // no countable read: nothing to stage
    const unsigned cm = __ballot_sync(0xFFFFFFFFu, cand);
    const int first = cm ? __ffs((int)cm) - 1 : 0;
    const uint32_t ref0 = __shfl_sync(0xFFFFFFFFu, ref, first), lseq0 = __shfl_sync(0xFFFFFFFFu, lseq, first);
    bool lean = cm != 0u && __all_sync(0xFFFFFFFFu, !cand || (alike && ref == ref0 && lseq == lseq0)) && lseq0 <= 32u * A.W;
    if (gend != 0u) {
        const uint32_t i0 = A.bin_first_m[gmin >> SD_BIN_SHIFT];
        const uint32_t i = i0 + lane;
        const bool can = i < A.n_ment && __ldg(&A.ment[2 * (size_t)i]).x < gend;   // sorted by start: a prefix of the lanes
        const unsigned b = __ballot_sync(0xFFFFFFFFu, can);
        const unsigned long long w0 = gmin >> 5, w1 = ((unsigned long long)(gend - 1u) >> 5) + 1ull;
        out.win_begin = w0 * 16ull;
        out.win_end = w1 * 16ull;
        out.ent_begin = (unsigned long long)i0 * 32ull;
        out.ent_end = out.ent_begin + (unsigned long long)__popc(b) * 32ull;
        const unsigned long long win_bytes = (w1 - w0 + A.W) * 16ull, ent_bytes = (unsigned long long)__popc(b) * 32ull;
        if (b == 0xFFFFFFFFu || w1 + 32ull > A.n_refm) {   // too many candidate entries: never staged
            out.win_end = out.win_begin + (1ull << 40);
            lean = false;
        }
        if (lean && ent_bytes <= 16u * 32u) {   // the lean kernel walks the entries in a warp-uniform loop: at most 16
            out.win = (out.win_begin << 16) | (win_bytes & 0xFFFFull);
            out.ent = (out.ent_begin << 16) | ent_bytes;
            out.lean = lseq0;
            out.g0 = A.chrom_off[ref0];
        }
    }
    if (lane == 0) {
        A.out[tile] = out;
        if (A.counts) {
            if (out.lean) atomicAdd(&A.counts[0], 1u);
            if (gend != 0u && out.win_end - out.win_begin + 16ull * A.W <= A.cap_win && out.ent_end - out.ent_begin <= A.cap_ent) atomicAdd(&A.counts[1], 1u);
        }
    }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
int sd_fail(sd_ctx *ctx, int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    return code;
}

extern "C" int sd_abi_version(void) { return SD_ABI_VERSION; }

// identifies one (reference layout, annotation) state across all contexts of the process: span descriptors are valid for one
static uint32_t next_epoch() {
    static std::atomic<uint32_t> epoch(0);
    return ++epoch;
}

extern "C" const char *sd_last_error(const sd_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

extern "C" int sd_create(int device, void *stream, sd_ctx **out) {
    if (!out) return SD_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return SD_ERR_CUDA;
    sd_ctx *ctx = new sd_ctx();
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return SD_ERR_CUDA; }
    if (stream) ctx->stream = (cudaStream_t)stream;
    else {
        if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return SD_ERR_CUDA; }
        ctx->own_stream = true;
    }
    cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
    cudaEventCreate(&ctx->ev_k0);
    cudaEventCreate(&ctx->ev_k1);
    cudaEventCreate(&ctx->ev_split[0]);
    cudaEventCreate(&ctx->ev_split[1]);
    for (int i = 0; i < 2; i++) {
        cudaEventCreateWithFlags(&ctx->ev_copy[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ctx->ev_done[i], cudaEventDisableTiming);
    }
    cudaDeviceGetAttribute(&ctx->num_sms, cudaDevAttrMultiProcessorCount, device);
    *out = ctx;
    return SD_OK;
}

static void free_utr_tables(sd_ctx *ctx) {
    for (int s = 0; s < 2; s++) {
        cudaFree(ctx->d_ustart[s]); cudaFree(ctx->d_uend[s]); cudaFree(ctx->d_urow[s]); cudaFree(ctx->d_useg[s]);
        cudaFree(ctx->d_utab[s]); cudaFree(ctx->d_stab[s]); ctx->d_utab[s] = ctx->d_stab[s] = nullptr;
        cudaFree(ctx->d_bin_first[s]); cudaFree(ctx->d_seg_start[s]); cudaFree(ctx->d_seg_end[s]); cudaFree(ctx->d_seg_off[s]);
        ctx->d_ustart[s] = ctx->d_uend[s] = ctx->d_urow[s] = ctx->d_useg[s] = ctx->d_bin_first[s] = nullptr;
        ctx->d_seg_start[s] = ctx->d_seg_end[s] = nullptr;
        ctx->d_seg_off[s] = nullptr;
    }
    cudaFree(ctx->d_ment); cudaFree(ctx->d_bin_first_m);
    ctx->d_ment = nullptr;
    ctx->d_bin_first_m = nullptr;
    ctx->n_ment = 0;
    cudaFree(ctx->d_row_gstart); cudaFree(ctx->d_row_gend); cudaFree(ctx->d_row_strand); cudaFree(ctx->d_row_blen);
    ctx->d_row_gstart = ctx->d_row_gend = ctx->d_row_blen = nullptr;
    ctx->d_row_strand = nullptr;
}

extern "C" void sd_destroy(sd_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(ctx->d_track_scratch);
    cudaFree(ctx->d_filter_scratch);
    free_utr_tables(ctx);
    cudaFree(ctx->refx[0]); cudaFree(ctx->refx[1]); cudaFree(ctx->refm);
    cudaFree(ctx->d_chrom_off); cudaFree(ctx->d_chrom_len);
    cudaFree(ctx->d_stage[0]); cudaFree(ctx->d_stage[1]);
    cudaFree(ctx->d_tile_counter);
    cudaFree(ctx->d_leftover);
    cudaFree(ctx->d_xi_dict);
    cudaEventDestroy(ctx->ev_k0); cudaEventDestroy(ctx->ev_k1);
    cudaEventDestroy(ctx->ev_split[0]); cudaEventDestroy(ctx->ev_split[1]);
    for (int i = 0; i < 2; i++) { cudaEventDestroy(ctx->ev_copy[i]); cudaEventDestroy(ctx->ev_done[i]); }
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" int sd_device_synchronize(sd_ctx *ctx) {
    if (!ctx) return SD_ERR_ARG;
    SD_CUDA(ctx, cudaSetDevice(ctx->device));
    SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SD_OK;
}

extern "C" int sd_set_reference(sd_ctx *ctx, const uint32_t *d_lo, const uint32_t *d_hi, const uint32_t *d_nmask,
                                uint64_t n_words, const uint32_t *h_chrom_off, const uint32_t *h_chrom_len, uint32_t n_chrom) {
    if (!ctx || !d_lo || !d_hi || !d_nmask || !h_chrom_off || !h_chrom_len || n_chrom == 0) return sd_fail(ctx, SD_ERR_ARG, "sd_set_reference: null argument");
    if (n_chrom >= SD_REF_NONE) return sd_fail(ctx, SD_ERR_LIMIT, "more than %u chromosomes", SD_REF_NONE - 1);
    if (n_words * 32ull + 4096ull > 0xFFFFFFFFull) return sd_fail(ctx, SD_ERR_LIMIT, "genome exceeds the 32-bit linear coordinate space");
    for (uint32_t c = 0; c < n_chrom; c++) {
        if (h_chrom_off[c] & 31u) return sd_fail(ctx, SD_ERR_ARG, "chromosome offset %u not a multiple of 32", c);
        if ((uint64_t)h_chrom_off[c] + h_chrom_len[c] > n_words * 32ull) return sd_fail(ctx, SD_ERR_ARG, "chromosome %u exceeds the planes", c);
        if (c + 1 < n_chrom && (uint64_t)h_chrom_off[c] + h_chrom_len[c] + 32ull > h_chrom_off[c + 1])
            return sd_fail(ctx, SD_ERR_ARG, "chromosomes %u and %u need >= 32 pad positions between them", c, c + 1);
    }
    SD_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaFree(ctx->refx[0]); cudaFree(ctx->refx[1]); cudaFree(ctx->refm); cudaFree(ctx->d_chrom_off); cudaFree(ctx->d_chrom_len);
    ctx->refx[0] = ctx->refx[1] = nullptr;
    ctx->refm = nullptr;
    ctx->ann_epoch = next_epoch();
    ctx->n_words = n_words;
    ctx->n_chrom = n_chrom;
    ctx->chrom_off.assign(h_chrom_off, h_chrom_off + n_chrom);
    ctx->chrom_len.assign(h_chrom_len, h_chrom_len + n_chrom);
    const uint64_t tot = n_words + SD_REF_PAD_WORDS;
    SD_CUDA(ctx, cudaMalloc(&ctx->refx[0], tot * sizeof(uint2)));
    SD_CUDA(ctx, cudaMalloc(&ctx->refx[1], tot * sizeof(uint2)));
    SD_CUDA(ctx, cudaMalloc(&ctx->refm, tot * sizeof(uint4)));
    SD_CUDA(ctx, cudaMalloc(&ctx->d_chrom_off, n_chrom * sizeof(uint32_t)));
    SD_CUDA(ctx, cudaMalloc(&ctx->d_chrom_len, n_chrom * sizeof(uint32_t)));
    SD_CUDA(ctx, cudaMemcpyAsync(ctx->d_chrom_off, h_chrom_off, n_chrom * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    SD_CUDA(ctx, cudaMemcpyAsync(ctx->d_chrom_len, h_chrom_len, n_chrom * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    const unsigned blocks = (unsigned)((tot + 255) / 256);
    sd_refx_kernel<<<blocks, 256, 0, ctx->stream>>>(d_lo, d_hi, d_nmask, n_words, ctx->refx[0], ctx->refx[1], ctx->refm);
    ctx->launches++;
    SD_CUDA(ctx, cudaGetLastError());
    SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SD_OK;
}

extern "C" int sd_set_snps(sd_ctx *ctx, const uint32_t *d_tc_mask, const uint32_t *d_ag_mask) {
    if (!ctx) return SD_ERR_ARG;
    if (!ctx->refx[0]) return sd_fail(ctx, SD_ERR_STATE, "sd_set_snps before sd_set_reference");
    SD_CUDA(ctx, cudaSetDevice(ctx->device));
    const unsigned blocks = (unsigned)((ctx->n_words + 255) / 256);
    sd_snp_kernel<<<blocks, 256, 0, ctx->stream>>>(d_tc_mask, d_ag_mask, ctx->n_words, ctx->refx[0], ctx->refx[1], ctx->refm);
    ctx->launches++;
    SD_CUDA(ctx, cudaGetLastError());
    SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SD_OK;
}

template <class T>
static int upload(sd_ctx *ctx, T **dst, const std::vector<T> &v) {
    const size_t n = v.size() ? v.size() : 1;
    SD_CUDA(ctx, cudaMalloc(dst, n * sizeof(T)));
    if (v.size()) SD_CUDA(ctx, cudaMemcpyAsync(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    return SD_OK;
}

extern "C" int sd_set_utrs(sd_ctx *ctx, const uint32_t *h_chrom, const int64_t *h_start, const int64_t *h_stop,
                           const uint8_t *h_strand, const uint8_t *h_in_bam, uint32_t n_utr) {
    if (!ctx || (n_utr && (!h_chrom || !h_start || !h_stop || !h_strand))) return sd_fail(ctx, SD_ERR_ARG, "sd_set_utrs: null argument");
    if (!ctx->refx[0]) return sd_fail(ctx, SD_ERR_STATE, "sd_set_utrs before sd_set_reference");
    SD_CUDA(ctx, cudaSetDevice(ctx->device));
    SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    free_utr_tables(ctx);
    ctx->n_utr = n_utr;
    ctx->row_pos_off.assign(n_utr, UINT64_MAX);
    ctx->row_pos_len.assign(n_utr, 0);
    std::vector<uint32_t> row_gs(n_utr, 0), row_ge(n_utr, 0);
    std::vector<uint8_t> row_st(n_utr, 0);
    std::vector<uint32_t> row_bl(n_utr, 0);
    struct Ent { uint32_t gs, ge, row; };
    std::vector<Ent> ent[2];
    for (uint32_t i = 0; i < n_utr; i++) {
        const int s = h_strand[i] == 1 ? 1 : 0;
        row_st[i] = h_strand[i] > 1 ? (uint8_t)2 : (uint8_t)s;   // 2: unstranded row, joins reads of both strands (sd_read_stats only)
        row_bl[i] = h_stop[i] > h_start[i] ? (uint32_t)std::min<int64_t>(h_stop[i] - h_start[i], 0xFFFFFFFFll) : 0u;
        if (h_chrom[i] >= ctx->n_chrom) continue;
        const int64_t clen = ctx->chrom_len[h_chrom[i]];
        int64_t a = h_start[i] < 0 ? 0 : h_start[i];
        int64_t b = h_stop[i] > clen ? clen : h_stop[i];
        if (a >= b) continue;  // empty after clipping: no bases, no reads
        const uint32_t off = ctx->chrom_off[h_chrom[i]];
        row_gs[i] = off + (uint32_t)a;
        row_ge[i] = off + (uint32_t)b;
        ctx->row_pos_len[i] = (uint32_t)(b - a);
        if (h_in_bam && !h_in_bam[i]) continue;
        ent[s].push_back({row_gs[i], row_ge[i], i});
        if (h_strand[i] > 1) ent[1].push_back({row_gs[i], row_ge[i], i});
    }
    const uint64_t total_bits = ctx->n_words * 32ull;
    ctx->n_bins = (uint32_t)((total_bits >> SD_BIN_SHIFT) + 2);
    uint64_t pos_total = 0;
    struct MEnt { uint32_t gs, ge, rowst, seg, ss, se; uint64_t posbase; };
    std::vector<MEnt> merged;
    for (int s = 0; s < 2; s++) {
        auto &E = ent[s];
        std::stable_sort(E.begin(), E.end(), [](const Ent &x, const Ent &y) { return x.gs < y.gs || (x.gs == y.gs && x.ge < y.ge); });
        const uint32_t n = (uint32_t)E.size();
        std::vector<uint32_t> us(n), ue(n), ur(n), useg(n), maxend(n), sstart, send;
        std::vector<uint64_t> soff;
        uint32_t cur_end = 0;
        for (uint32_t i = 0; i < n; i++) {
            us[i] = E[i].gs; ue[i] = E[i].ge; ur[i] = E[i].row;
            maxend[i] = i ? std::max(maxend[i - 1], ue[i]) : ue[i];
            if (sstart.empty() || us[i] >= cur_end) {  // new merged segment
                sstart.push_back(us[i]);
                send.push_back(ue[i]);
                cur_end = ue[i];
            } else if (ue[i] > cur_end) {
                cur_end = ue[i];
                send.back() = cur_end;
            }
            useg[i] = (uint32_t)sstart.size() - 1;
        }
        soff.resize(sstart.size());
        for (size_t m = 0; m < sstart.size(); m++) {
            soff[m] = pos_total;
            pos_total += (uint64_t)(send[m] - sstart[m]) + 1ull;
        }
        for (uint32_t i = 0; i < n; i++) ctx->row_pos_off[ur[i]] = soff[useg[i]] + (us[i] - sstart[useg[i]]);
        for (uint32_t i = 0; i < n; i++)
            merged.push_back({us[i], ue[i], ur[i] | ((uint32_t)s << 31), useg[i], sstart[useg[i]], send[useg[i]], soff[useg[i]] - sstart[useg[i]]});
        // bin -> first table index that can still overlap a window starting in the bin
        std::vector<uint32_t> bins(ctx->n_bins);
        uint32_t j = 0;
        for (uint32_t b = 0; b < ctx->n_bins; b++) {
            const uint64_t bstart = (uint64_t)b << SD_BIN_SHIFT;
            while (j < n && (uint64_t)maxend[j] <= bstart) j++;
            bins[b] = j;
        }
        ctx->n_u[s] = n;
        ctx->n_seg[s] = (uint32_t)sstart.size();
        int rc;
        if ((rc = upload(ctx, &ctx->d_ustart[s], us))) return rc;
        if ((rc = upload(ctx, &ctx->d_uend[s], ue))) return rc;
        if ((rc = upload(ctx, &ctx->d_urow[s], ur))) return rc;
        if ((rc = upload(ctx, &ctx->d_useg[s], useg))) return rc;
        if ((rc = upload(ctx, &ctx->d_bin_first[s], bins))) return rc;
        if ((rc = upload(ctx, &ctx->d_seg_start[s], sstart))) return rc;
        if ((rc = upload(ctx, &ctx->d_seg_end[s], send))) return rc;
        if ((rc = upload(ctx, &ctx->d_seg_off[s], soff))) return rc;
        std::vector<uint4> utab(n), stab(sstart.size());
        for (uint32_t i = 0; i < n; i++) utab[i] = make_uint4(us[i], ue[i], ur[i], useg[i]);
        for (size_t m = 0; m < sstart.size(); m++) stab[m] = make_uint4(sstart[m], send[m], (uint32_t)(soff[m] & 0xFFFFFFFFull), (uint32_t)(soff[m] >> 32));
        if ((rc = upload(ctx, &ctx->d_utab[s], utab))) return rc;
        if ((rc = upload(ctx, &ctx->d_stab[s], stab))) return rc;
        SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // vectors go out of scope
    }
    ctx->n_positions = pos_total;
    int rc;
    {   // strand-merged table for the staged join (same order rule, so each strand's entries keep their relative order)
        if (n_utr >= 0x80000000u) return sd_fail(ctx, SD_ERR_LIMIT, "too many annotation rows");
        std::stable_sort(merged.begin(), merged.end(), [](const MEnt &x, const MEnt &y) { return x.gs < y.gs || (x.gs == y.gs && x.ge < y.ge); });
        const uint32_t n = (uint32_t)merged.size();
        std::vector<uint4> tab(2 * (size_t)n);
        std::vector<uint32_t> bins(ctx->n_bins);
        uint32_t maxend = 0, j = 0;
        std::vector<uint32_t> pm(n);
        for (uint32_t i = 0; i < n; i++) {
            const MEnt &m = merged[i];
            tab[2 * (size_t)i] = make_uint4(m.gs, m.ge, m.rowst, m.seg);
            tab[2 * (size_t)i + 1] = make_uint4(m.ss, m.se, (uint32_t)(m.posbase & 0xFFFFFFFFull), (uint32_t)(m.posbase >> 32));
            maxend = std::max(maxend, m.ge);
            pm[i] = maxend;
        }
        for (uint32_t b = 0; b < ctx->n_bins; b++) {
            const uint64_t bstart = (uint64_t)b << SD_BIN_SHIFT;
            while (j < n && (uint64_t)pm[j] <= bstart) j++;
            bins[b] = j;
        }
        ctx->n_ment = n;
        if ((rc = upload(ctx, &ctx->d_ment, tab))) return rc;
        if ((rc = upload(ctx, &ctx->d_bin_first_m, bins))) return rc;
        SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    ctx->ann_epoch = next_epoch();
    if ((rc = upload(ctx, &ctx->d_row_gstart, row_gs))) return rc;
    if ((rc = upload(ctx, &ctx->d_row_gend, row_ge))) return rc;
    if ((rc = upload(ctx, &ctx->d_row_strand, row_st))) return rc;
    if ((rc = upload(ctx, &ctx->d_row_blen, row_bl))) return rc;
    ctx->row_gstart = row_gs;
    ctx->row_strand = row_st;
    SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->counters = nullptr;
    ctx->pos_cov = ctx->pos_tc = nullptr;
    return SD_OK;
}

extern "C" int sd_position_space(const sd_ctx *ctx, uint64_t *n_positions) {
    if (!ctx || !n_positions) return SD_ERR_ARG;
    *n_positions = ctx->n_positions;
    return SD_OK;
}

extern "C" int sd_utr_layout(const sd_ctx *ctx, uint64_t *h_pos_off, uint32_t *h_pos_len) {
    if (!ctx) return SD_ERR_ARG;
    for (uint32_t i = 0; i < ctx->n_utr; i++) {
        if (h_pos_off) h_pos_off[i] = ctx->row_pos_off[i];
        if (h_pos_len) h_pos_len[i] = ctx->row_pos_len[i];
    }
    return SD_OK;
}

extern "C" int sd_bind_outputs(sd_ctx *ctx, int64_t *d_counters, int32_t *d_pos_cov, int32_t *d_pos_tc, int64_t *d_stats) {
    if (!ctx || !d_counters || !d_stats) return sd_fail(ctx, SD_ERR_ARG, "sd_bind_outputs: null argument");
    if (ctx->n_positions && (!d_pos_cov || !d_pos_tc)) return sd_fail(ctx, SD_ERR_ARG, "sd_bind_outputs: position arrays required");
    SD_CUDA(ctx, cudaSetDevice(ctx->device));
    ctx->counters = d_counters;
    ctx->pos_cov = d_pos_cov;
    ctx->pos_tc = d_pos_tc;
    ctx->stats = d_stats;
    SD_CUDA(ctx, cudaMemsetAsync(d_counters, 0, (size_t)ctx->n_utr * SD_NCOUNTER * sizeof(int64_t), ctx->stream));
    if (ctx->n_positions) {
        SD_CUDA(ctx, cudaMemsetAsync(d_pos_cov, 0, ctx->n_positions * sizeof(int32_t), ctx->stream));
        SD_CUDA(ctx, cudaMemsetAsync(d_pos_tc, 0, ctx->n_positions * sizeof(int32_t), ctx->stream));
    }
    SD_CUDA(ctx, cudaMemsetAsync(d_stats, 0, SD_NSTAT * sizeof(int64_t), ctx->stream));
    return SD_OK;
}

// Window words of the count kernels for a batch, and the room a stage gives a tile's reference window and annotation entries.
// The window must hold the alignment's reference span, i.e. the read plus its deletions -- a read that does not fit takes the
// per-base path, two orders of magnitude slower, and stalls its whole warp (190 bp reads with 2 % short indels in a 192-base
// window: 9.6 instead of 15 G reads/s).  The batch says how wide its alignments are (spliced ones aside); without that, 16
// bases of slack cover the usual gaps.
static void span_caps(const sd_batch *b, uint32_t &need, uint32_t &Wn, uint32_t &cap_win, uint32_t &cap_ent) {
    need = b->max_span ? std::max(b->max_lseq, b->max_span) : b->max_lseq + 16u;
    Wn = need <= 128 ? 4 : need <= 192 ? 6 : need <= 256 ? 8 : need <= 320 ? 10 : need <= 384 ? 12 : 16;
    // the reads' words, a few more for the spread of their starts, W behind the last: 864 bp of tile at W = 4 -- the thinly
    // populated classes of a grouped batch (clipped / indel reads) fit too; 8 entries
    cap_win = align128(16u * (6u * Wn + 8u));
    cap_ent = 256u;
}

int count_on_stream(sd_ctx *ctx, const sd_batch *b, const sd_params *p, cudaStream_t stream, uint8_t *read_tc, bool timed) {
    if (!ctx->refx[0]) return sd_fail(ctx, SD_ERR_STATE, "sd_count before sd_set_reference");
    if (!ctx->counters) return sd_fail(ctx, SD_ERR_STATE, "sd_count before sd_bind_outputs");
    if (b->n_tiles == 0) return SD_OK;
    if (b->tile_reads != 32) return sd_fail(ctx, SD_ERR_ARG, "tile_reads must be 32 (one tile per warp)");
    if (p->mismatch_source != SD_MISMATCH_CIGAR && !b->mp) return sd_fail(ctx, SD_ERR_ARG, "mismatch source needs the mp column");
    const bool qplanes = b->qual_form == 1;
    if (!qplanes && b->qual_bits != 0 && b->qual_bits != 8)
        return sd_fail(ctx, SD_ERR_ARG, "the qual column holds %u-bit codes: expand it with sd_qual_unpack (sd_count_host does)", b->qual_bits);
    if (qplanes && ((b->qual_bits != 2 && b->qual_bits != 4) || b->seq_bits != 2))
        return sd_fail(ctx, SD_ERR_ARG, "a plane-form qual column needs 2 or 4 planes and the 2-bit SEQ column");
    if (b->qual_form > 1) return sd_fail(ctx, SD_ERR_ARG, "unknown qual_form %u", b->qual_form);
    CountArgs A;
    memset(&A, 0, sizeof A);
    A.rec = b->rec; A.seq = b->seq; A.qual = b->qual; A.cigar = b->cigar;
    A.mp = (p->mismatch_source == SD_MISMATCH_CIGAR) ? nullptr : b->mp;
    A.tiles = b->tiles; A.n_tiles = b->n_tiles; A.tile_reads = b->tile_reads;
    A.cap_rec = align16(32u * (uint32_t)sizeof(sd_read_rec));
    if (b->seq_bits != 0 && b->seq_bits != 4 && b->seq_bits != 2) return sd_fail(ctx, SD_ERR_ARG, "seq_bits must be 4 (or 0) or 2");
    A.seq2 = b->seq_bits == 2 ? 1u : 0u;
    A.cap_seq = std::max(align16(b->max_tile_seq), (A.seq2 ? 8u : 16u) * SD_TILE_READS);   // at least one group per row: the kernel loads group 0 unconditionally
    A.cap_qual = 0;  // byte form: the qual column is read in place (see the kernel's header comment)
    A.cap_cig = align16(b->max_tile_cig);
    A.cap_mp = A.mp ? align16(b->max_tile_mp) : 0;
    {   // keep every warp's region 128-byte aligned
        const uint32_t sum = A.cap_rec + A.cap_seq + A.cap_cig + A.cap_mp;
        A.cap_cig += align128(sum) - sum;
    }
    if (qplanes) {   // plane form: the tile's block is staged behind the other columns
        A.qual_planes = b->qual_bits;
        A.cap_qual = align128(std::max(b->max_tile_qual, 4u * b->qual_bits * SD_TILE_READS));
        // codes ascend with quality (sd_decode_bam, sd_qual_dict): quality >= min_qual  <=>  code >= q_cmin
        uint32_t used = 1;
        for (uint32_t c = 1; c < (1u << b->qual_bits); c++)
            if (b->qual_dict[c] > b->qual_dict[c - 1]) used = c + 1;
        uint32_t cmin = 0;
        while (cmin < used && (int)b->qual_dict[cmin] < p->min_qual) cmin++;
        A.q_cmin = cmin < used ? cmin : (1u << b->qual_bits);
        if (b->qual_bits == 2) {
            const uint32_t c = A.q_cmin, ones = 0xFFFFFFFFu;
            A.q_m[0] = (c == 1 || c == 2) ? ones : 0u;   // p1 alone suffices
            A.q_m[1] = c == 1 ? ones : 0u;               // p0 alone suffices
            A.q_m[2] = c == 0 ? ones : 0u;               // everything passes
            A.q_m[3] = c >= 4 ? 0u : ones;               // nothing passes
        } else
            for (int k = 0; k < 4; k++) A.q_m[k] = 0u - ((A.q_cmin >> k) & 1u);
    }
    // span descriptors made for this reference layout and annotation: the kernel stages each tile's reference window and
    // annotation entries with the tile (otherwise every lane fetches its own from global memory)
    static int spans_env = -1;
    if (spans_env < 0) { const char *e = getenv("SD_SPANS"); spans_env = e ? atoi(e) : 1; }
    A.spans = (b->spans && b->spans_epoch == ctx->ann_epoch && spans_env) ? b->spans : nullptr;
    A.hint_lean = (A.spans && b->spans_counted) ? b->spans_lean_tiles : 0xFFFFFFFFu;       // 0xFFFFFFFF: not known
    A.hint_staged = (A.spans && b->spans_counted) ? b->spans_staged_tiles : 0xFFFFFFFFu;
    A.refm = ctx->refm;
    A.ment = ctx->d_ment;
    for (int s = 0; s < 2; s++) {
        A.refx[s] = ctx->refx[s];
        A.ustart[s] = ctx->d_ustart[s]; A.uend[s] = ctx->d_uend[s]; A.urow[s] = ctx->d_urow[s]; A.useg[s] = ctx->d_useg[s];
        A.utab[s] = ctx->d_utab[s]; A.stab[s] = ctx->d_stab[s];
        A.bin_first[s] = ctx->d_bin_first[s]; A.n_u[s] = ctx->n_u[s];
        A.seg_start[s] = ctx->d_seg_start[s]; A.seg_end[s] = ctx->d_seg_end[s]; A.seg_off[s] = ctx->d_seg_off[s];
    }
    A.chrom_off = ctx->d_chrom_off; A.chrom_len = ctx->d_chrom_len; A.n_chrom = ctx->n_chrom;
    A.n_bins = ctx->n_bins;
    A.counters = reinterpret_cast<unsigned long long *>(ctx->counters);
    A.pos_cov = ctx->pos_cov; A.pos_tc = ctx->pos_tc;
    A.stats = reinterpret_cast<unsigned long long *>(ctx->stats);
    A.read_tc = read_tc;
    if (!ctx->d_tile_counter) SD_CUDA(ctx, cudaMalloc(&ctx->d_tile_counter, 8 * sizeof(unsigned int)));
    SD_CUDA(ctx, cudaMemsetAsync(ctx->d_tile_counter, 0, 8 * sizeof(unsigned int), stream));
    A.tile_counter = ctx->d_tile_counter;
    if (A.spans) {   // two lists: tiles the lean kernel hands on to the staged one, and the staged one to the plain one
        if (ctx->leftover_cap < b->n_tiles) {
            SD_CUDA(ctx, cudaStreamSynchronize(stream));
            cudaFree(ctx->d_leftover);
            ctx->d_leftover = nullptr;
            ctx->leftover_cap = 0;
            SD_CUDA(ctx, cudaMalloc(&ctx->d_leftover, 2 * (size_t)b->n_tiles * sizeof(uint32_t)));
            ctx->leftover_cap = b->n_tiles;
        }
        A.leftover = ctx->d_leftover;
    }
    A.min_qual = p->min_qual; A.conv_threshold = p->conversion_threshold; A.mismatch_source = p->mismatch_source;
    A.filter_on = p->filter_on; A.filter_mq = p->filter_mq; A.filter_max_nm = p->filter_max_nm;
    {   // filter.py:246 compares the float32 XI tag, promoted to double, with the double threshold; the same predicate in float:
        float f = (float)p->filter_min_identity;
        if ((double)f < p->filter_min_identity) f = nextafterf(f, INFINITY);
        A.filter_min_identity = f;
    }
    A.force_slow = p->force_slow_path;
    A.mp_any_base = (p->flags & SD_PARAM_MP_ANY_BASE) ? 1 : 0;
    A.n_positions = ctx->n_positions;
    A.n_utr = ctx->n_utr;
    A.qual_bytes = 0;   // set below in debug builds
#ifdef SD_DEBUG_BOUNDS
    {
        sd_tile last;
        SD_CUDA(ctx, cudaDeviceSynchronize());   // the tile table may still be in flight on the copy stream (sd_count_host)
        SD_CUDA(ctx, cudaMemcpy(&last, b->tiles + b->n_tiles, sizeof last, cudaMemcpyDefault));
        A.qual_bytes = last.qual_off;
    }
#endif
    uint32_t need, Wn, cap_win, cap_ent;
    span_caps(b, need, Wn, cap_win, cap_ent);
    if (A.spans) {
        A.cap_win = cap_win;
        A.cap_ent = cap_ent;
    }
#define SD_DISPATCH_W(SB, QB)                                                        \
    do {                                                                             \
        if (need <= 128) return launch_count<4, SB, QB>(ctx, A, stream, timed);      \
        if (need <= 192) return launch_count<6, SB, QB>(ctx, A, stream, timed);      \
        if (need <= 256) return launch_count<8, SB, QB>(ctx, A, stream, timed);      \
        if (need <= 320) return launch_count<10, SB, QB>(ctx, A, stream, timed);     \
        if (need <= 384) return launch_count<12, SB, QB>(ctx, A, stream, timed);     \
        return launch_count<16, SB, QB>(ctx, A, stream, timed);                      \
    } while (0)
    if (A.qual_planes == 2) SD_DISPATCH_W(2, 2);
    if (A.qual_planes == 4) SD_DISPATCH_W(2, 4);
    if (A.seq2) SD_DISPATCH_W(2, 0);
    SD_DISPATCH_W(4, 0);
#undef SD_DISPATCH_W
}

int tile_spans_on(sd_ctx *ctx, cudaStream_t stream, const sd_batch *b, void *d_spans, unsigned int *d_counts) {
    if (b->n_tiles == 0) return SD_OK;
    SpanArgs S;
    S.counts = d_counts;
    S.rec = b->rec; S.cigar = b->cigar; S.tiles = b->tiles; S.n_tiles = b->n_tiles;
    S.chrom_off = ctx->d_chrom_off; S.chrom_len = ctx->d_chrom_len; S.n_chrom = ctx->n_chrom;
    S.ment = ctx->d_ment; S.bin_first_m = ctx->d_bin_first_m; S.n_ment = ctx->n_ment;
    S.out = reinterpret_cast<sd_tile_span *>(d_spans);
    S.n_refm = ctx->n_words + SD_REF_PAD_WORDS;
    uint32_t need;
    span_caps(b, need, S.W, S.cap_win, S.cap_ent);
    const unsigned blocks = (unsigned)(((uint64_t)b->n_tiles * 32u + 255u) / 256u);
    sd_tile_spans_kernel<<<blocks, 256, 0, stream>>>(S);
    ctx->launches++;
    SD_CUDA(ctx, cudaGetLastError());
    return SD_OK;
}

extern "C" int sd_tile_spans(sd_ctx *ctx, sd_batch *d_batch, void *d_spans) {
    if (!ctx || !d_batch || (d_batch->n_tiles && !d_spans)) return sd_fail(ctx, SD_ERR_ARG, "sd_tile_spans: null argument");
    if (!ctx->refx[0] || !ctx->d_ment) return sd_fail(ctx, SD_ERR_STATE, "sd_tile_spans before sd_set_reference / sd_set_utrs");
    if (d_batch->tile_reads != 32) return sd_fail(ctx, SD_ERR_ARG, "tile_reads must be 32");
    SD_CUDA(ctx, cudaSetDevice(ctx->device));
    // how many tiles each kind of kernel will take: two counters behind the queue heads of the count launches
    if (!ctx->d_tile_counter) SD_CUDA(ctx, cudaMalloc(&ctx->d_tile_counter, 8 * sizeof(unsigned int)));
    unsigned int *d_counts = ctx->d_tile_counter + 6;
    SD_CUDA(ctx, cudaMemsetAsync(d_counts, 0, 2 * sizeof(unsigned int), ctx->stream));
    const int rc = tile_spans_on(ctx, ctx->stream, d_batch, d_spans, d_counts);
    if (rc) return rc;
    unsigned int h[2] = {0, 0};
    SD_CUDA(ctx, cudaMemcpyAsync(h, d_counts, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
    SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    d_batch->spans = d_spans;
    d_batch->spans_epoch = ctx->ann_epoch;
    d_batch->spans_counted = 1;
    d_batch->spans_lean_tiles = h[0];
    d_batch->spans_staged_tiles = h[1];
    return SD_OK;
}

extern "C" int sd_count(sd_ctx *ctx, const sd_batch *d_batch, const sd_params *params) {
    if (!ctx || !d_batch || !params) return sd_fail(ctx, SD_ERR_ARG, "sd_count: null argument");
    SD_CUDA(ctx, cudaSetDevice(ctx->device));
    return count_on_stream(ctx, d_batch, params, ctx->stream, nullptr, true);
}

// per-read tcCount output (tests, diagnostics; `alleyoop read-separator` would consume it)
extern "C" int sd_count_reads(sd_ctx *ctx, const sd_batch *d_batch, const sd_params *params, uint8_t *d_read_tc) {
    if (!ctx || !d_batch || !params) return sd_fail(ctx, SD_ERR_ARG, "sd_count_reads: null argument");
    SD_CUDA(ctx, cudaSetDevice(ctx->device));
    return count_on_stream(ctx, d_batch, params, ctx->stream, d_read_tc, true);
}

extern "C" float sd_last_kernel_ms(sd_ctx *ctx) {
    if (!ctx || !ctx->have_kernel_time) return -1.f;
    cudaSetDevice(ctx->device);
    if (cudaEventSynchronize(ctx->ev_k1) != cudaSuccess) return -1.f;
    float ms = -1.f;
    if (cudaEventElapsedTime(&ms, ctx->ev_k0, ctx->ev_k1) != cudaSuccess) return -1.f;
    return ms;
}

extern "C" int sd_last_kernel_split(sd_ctx *ctx, float ms[3], int64_t left[2]) {
    if (!ctx || !ms || !left) return SD_ERR_ARG;
    ms[0] = ms[1] = ms[2] = 0.f;
    left[0] = left[1] = 0;
    if (!ctx->have_kernel_time || !ctx->d_tile_counter) return sd_fail(ctx, SD_ERR_STATE, "no timed sd_count yet");
    SD_CUDA(ctx, cudaSetDevice(ctx->device));
    SD_CUDA(ctx, cudaEventSynchronize(ctx->ev_k1));
    cudaEvent_t marks[4] = {ctx->ev_k0, ctx->ev_split[0], ctx->ev_split[1], ctx->ev_k1};
    if (ctx->split_marks == 2) {
        for (int i = 0; i < 3; i++) SD_CUDA(ctx, cudaEventElapsedTime(&ms[i], marks[i], marks[i + 1]));
    } else if (ctx->split_marks == 1) {   // no lean launch
        SD_CUDA(ctx, cudaEventElapsedTime(&ms[1], ctx->ev_k0, ctx->ev_split[1]));
        SD_CUDA(ctx, cudaEventElapsedTime(&ms[2], ctx->ev_split[1], ctx->ev_k1));
    } else SD_CUDA(ctx, cudaEventElapsedTime(&ms[2], ctx->ev_k0, ctx->ev_k1));
    unsigned int n[2] = {0, 0};
    SD_CUDA(ctx, cudaMemcpy(n, ctx->d_tile_counter + 3, sizeof n, cudaMemcpyDeviceToHost));
    left[0] = n[0];
    left[1] = n[1];
    return SD_OK;
}

extern "C" int64_t sd_launch_count(const sd_ctx *ctx) { return ctx ? ctx->launches : 0; }

extern "C" int64_t sd_last_leftover_tiles(sd_ctx *ctx) {
    if (!ctx || !ctx->d_tile_counter) return -1;
    cudaSetDevice(ctx->device);
    unsigned int n = 0;
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return -1;
    if (cudaMemcpy(&n, ctx->d_tile_counter + 4, sizeof n, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return (int64_t)n;
}

// Host batch -> device in chunks of whole tiles, double buffered: the copy of chunk i+1 overlaps the
// kernel of chunk i.  One staging allocation per buffer holds rec | seq | qual | cigar | mp | tiles.
extern "C" int sd_count_host(sd_ctx *ctx, const sd_batch *hb, const sd_params *params) {
    if (!ctx || !hb || !params) return sd_fail(ctx, SD_ERR_ARG, "sd_count_host: null argument");
    SD_CUDA(ctx, cudaSetDevice(ctx->device));
    if (hb->n_tiles == 0) return SD_OK;
    const bool use_mp = params->mismatch_source != SD_MISMATCH_CIGAR;
    if (use_mp && !hb->mp) return sd_fail(ctx, SD_ERR_ARG, "mismatch source needs the mp column");
    const uint64_t target = 256ull << 20;  // bytes per chunk
    const uint32_t qbits = hb->qual_bits ? hb->qual_bits : 8u;
    const bool coded = qbits != 8;
    if (coded && qbits != 2 && qbits != 4) return sd_fail(ctx, SD_ERR_ARG, "quality coding of %u bits is not supported", qbits);
    // coded qualities next to the 2-bit SEQ column: the count kernel takes them as bit-planes (sd_batch.qual_form = 1)
    static int planes_env = -1;
    if (planes_env < 0) { const char *e = getenv("SD_QUAL_PLANES"); planes_env = e ? atoi(e) : 1; }
    const bool to_planes = coded && hb->seq_bits == 2 && planes_env != 0;
    const sd_tile *T = hb->tiles;
    const uint64_t rec_tile = (uint64_t)hb->tile_reads * sizeof(sd_read_rec);
    auto chunk_bytes = [&](uint32_t t0, uint32_t t1) {
        uint64_t b = (uint64_t)(t1 - t0) * rec_tile + (T[t1].seq_off - T[t0].seq_off) + (T[t1].qual_off - T[t0].qual_off) +
                     (T[t1].cig_off - T[t0].cig_off) + (uint64_t)(t1 - t0 + 1) * sizeof(sd_tile) + 6 * 256;
        if (coded) b += (T[t1].qual_off - T[t0].qual_off) * qbits / 8;   // the codes as copied, next to the expanded column
        if (to_planes) b += (T[t1].seq_off - T[t0].seq_off) * qbits / 2 + (uint64_t)(t1 - t0 + 1) * sizeof(sd_tile) + 4 * 256;
        if (use_mp) b += T[t1].mp_off - T[t0].mp_off;
        b += (uint64_t)(t1 - t0) * sizeof(sd_tile_span) + 256;
        return b;
    };
    // chunk boundaries
    std::vector<uint32_t> cuts{0};
    {
        uint32_t t0 = 0;
        while (t0 < hb->n_tiles) {
            uint32_t lo = t0 + 1, hi = hb->n_tiles;
            if (chunk_bytes(t0, hi) > target) {
                while (lo < hi) {
                    const uint32_t mid = lo + (hi - lo + 1) / 2;
                    if (chunk_bytes(t0, mid) <= target) lo = mid; else hi = mid - 1;
                }
                hi = lo;
            }
            cuts.push_back(hi);
            t0 = hi;
        }
    }
    uint64_t need = 0;
    for (size_t c = 0; c + 1 < cuts.size(); c++) need = std::max(need, chunk_bytes(cuts[c], cuts[c + 1]));
    if (need > ctx->stage_cap) {
        for (int i = 0; i < 2; i++) { cudaFree(ctx->d_stage[i]); ctx->d_stage[i] = nullptr; }
        ctx->stage_cap = 0;
        for (int i = 0; i < 2; i++) SD_CUDA(ctx, cudaMalloc(&ctx->d_stage[i], need));
        ctx->stage_cap = need;
    }
    std::vector<std::vector<sd_tile>> rel_keep(2);
    auto up256 = [](uint64_t x) { return (x + 255ull) & ~255ull; };
    for (size_t c = 0; c + 1 < cuts.size(); c++) {
        const int buf = (int)(c & 1);
        const uint32_t t0 = cuts[c], t1 = cuts[c + 1], nt = t1 - t0;
        // the previous kernel that used this buffer must be done before it is overwritten
        if (c >= 2) SD_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_done[buf], 0));
        unsigned char *d = (unsigned char *)ctx->d_stage[buf];
        uint64_t o = 0;
        sd_batch db = *hb;
        db.n_tiles = nt;
        db.n_reads = (uint64_t)nt * hb->tile_reads;
        const uint64_t b_rec = nt * rec_tile, b_seq = T[t1].seq_off - T[t0].seq_off, b_qual = T[t1].qual_off - T[t0].qual_off,
                       b_cig = T[t1].cig_off - T[t0].cig_off, b_mp = use_mp ? T[t1].mp_off - T[t0].mp_off : 0;
        db.rec = (const sd_read_rec *)(d + o);
        SD_CUDA(ctx, cudaMemcpyAsync(d + o, (const unsigned char *)hb->rec + t0 * rec_tile, b_rec, cudaMemcpyHostToDevice, ctx->copy_stream));
        o = up256(o + b_rec);
        db.seq = (const uint32_t *)(d + o);
        if (b_seq) SD_CUDA(ctx, cudaMemcpyAsync(d + o, (const unsigned char *)hb->seq + T[t0].seq_off, b_seq, cudaMemcpyHostToDevice, ctx->copy_stream));
        o = up256(o + b_seq);
        db.qual = (const uint8_t *)(d + o);
        db.qual_bits = 8;
        unsigned char *d_codes = nullptr;
        if (!coded) {
            if (b_qual) SD_CUDA(ctx, cudaMemcpyAsync(d + o, hb->qual + T[t0].qual_off, b_qual, cudaMemcpyHostToDevice, ctx->copy_stream));
            o = up256(o + b_qual);
        } else {   // copy the dictionary codes; they are expanded into db.qual on the compute stream once the copy has landed
            o = up256(o + b_qual);
            d_codes = d + o;
            if (b_qual) SD_CUDA(ctx, cudaMemcpyAsync(d_codes, hb->qual + T[t0].qual_off * qbits / 8, b_qual * qbits / 8, cudaMemcpyHostToDevice, ctx->copy_stream));
            o = up256(o + b_qual * qbits / 8);
        }
        db.cigar = (const uint32_t *)(d + o);
        if (b_cig) SD_CUDA(ctx, cudaMemcpyAsync(d + o, (const unsigned char *)hb->cigar + T[t0].cig_off, b_cig, cudaMemcpyHostToDevice, ctx->copy_stream));
        o = up256(o + b_cig);
        db.mp = nullptr;
        if (use_mp) {
            db.mp = (const uint32_t *)(d + o);
            if (b_mp) SD_CUDA(ctx, cudaMemcpyAsync(d + o, (const unsigned char *)hb->mp + T[t0].mp_off, b_mp, cudaMemcpyHostToDevice, ctx->copy_stream));
            o = up256(o + b_mp);
        }
        auto &rl = rel_keep[buf];
        rl.resize(nt + 1);
        for (uint32_t t = 0; t <= nt; t++) {
            rl[t].seq_off = T[t0 + t].seq_off - T[t0].seq_off;
            rl[t].qual_off = T[t0 + t].qual_off - T[t0].qual_off;
            rl[t].cig_off = T[t0 + t].cig_off - T[t0].cig_off;
            rl[t].mp_off = T[t0 + t].mp_off - T[t0].mp_off;
        }
        db.tiles = (const sd_tile *)(d + o);
        SD_CUDA(ctx, cudaMemcpyAsync(d + o, rl.data(), (size_t)(nt + 1) * sizeof(sd_tile), cudaMemcpyHostToDevice, ctx->copy_stream));
        SD_CUDA(ctx, cudaEventRecord(ctx->ev_copy[buf], ctx->copy_stream));
        SD_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_copy[buf], 0));
        o = up256(o + (size_t)(nt + 1) * sizeof(sd_tile));
        int rc = SD_OK;
        if (coded && b_qual && (rc = sd_qual_unpack_on(ctx, ctx->stream, d_codes, b_qual, qbits, hb->qual_dict, const_cast<uint8_t *>(db.qual))))
            return rc;
        if (to_planes) {   // bytes -> bit-planes in the SEQ geometry; the kernel then stages them with the tile
            uint32_t *d_planes = (uint32_t *)(d + o);
            o = up256(o + b_seq * qbits / 2);
            sd_tile *d_tiles2 = (sd_tile *)(d + o);
            o = up256(o + (size_t)(nt + 1) * sizeof(sd_tile));
            if ((rc = sd_qual_to_planes_on(ctx, ctx->stream, &db, qbits, hb->qual_dict, d_planes, d_tiles2))) return rc;
            db.qual = (const uint8_t *)d_planes;
            db.tiles = d_tiles2;
            db.qual_bits = qbits;
            db.qual_form = 1;
            db.max_tile_qual = hb->max_tile_seq / 8u * 4u * qbits;
        }
        db.spans = nullptr;
        if (ctx->d_ment) {   // span descriptors of the chunk's tiles (the staged count kernel)
            void *d_spans = d + o;
            o = up256(o + (size_t)nt * sizeof(sd_tile_span));
            if ((rc = tile_spans_on(ctx, ctx->stream, &db, d_spans, nullptr))) return rc;
            db.spans = d_spans;
            db.spans_epoch = ctx->ann_epoch;
        }
        rc = count_on_stream(ctx, &db, params, ctx->stream, nullptr, false);
        if (rc) return rc;
        SD_CUDA(ctx, cudaEventRecord(ctx->ev_done[buf], ctx->stream));
    }
    SD_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
    SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SD_OK;
}

extern "C" int sd_finalize(sd_ctx *ctx) {
    if (!ctx) return SD_ERR_ARG;
    if (!ctx->pos_cov || ctx->n_seg[0] + ctx->n_seg[1] == 0) return SD_OK;
    SD_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint32_t nseg = ctx->n_seg[0] + ctx->n_seg[1];
    const unsigned blocks = (nseg * 32u + 255u) / 256u;
    sd_finalize_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->d_seg_off[0], ctx->d_seg_start[0], ctx->d_seg_end[0], ctx->n_seg[0],
                                                        ctx->d_seg_off[1], ctx->d_seg_start[1], ctx->d_seg_end[1], ctx->n_seg[1],
                                                        ctx->pos_cov);
    ctx->launches++;
    SD_CUDA(ctx, cudaGetLastError());
    return SD_OK;
}

extern "C" int sd_tcontent(sd_ctx *ctx, int64_t *d_tcontent) {
    if (!ctx || !d_tcontent) return SD_ERR_ARG;
    if (!ctx->refx[0] || ctx->n_utr == 0) return SD_OK;
    SD_CUDA(ctx, cudaSetDevice(ctx->device));
    const unsigned blocks = (ctx->n_utr * 32u + 255u) / 256u;
    sd_tcontent_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->refx[0], ctx->refx[1], ctx->d_row_gstart, ctx->d_row_gend,
                                                        ctx->d_row_strand, ctx->n_utr, reinterpret_cast<long long *>(d_tcontent));
    ctx->launches++;
    SD_CUDA(ctx, cudaGetLastError());
    return SD_OK;
}

