// sd_synth.cu -- device-side synthetic SLAM-seq generator for the BASELINE.json configurations.
//
// Model (after the reference's simulator, slamdunk/dunks/simulator.py:212-229 and splash.py:125):
// reads are placed on UTRs drawn from an expression profile, carry a per-reference-T Bernoulli
// T>C conversion (A>G for minus-strand reads) when the read is "labelled", plus uniform
// substitution errors; a share of reads is soft-clipped or has one short insertion / deletion.
// Everything derives from a counter-based RNG keyed by (seed, global read index), so any split of
// the index range over batches or GPUs produces the same reads.
//
// Output is written directly in the columnar batch layout (include/slamdunk_b200.h) together
// with the NextGenMap-style tags the pipeline consumes (XI, NM, MP entries) -- the MP entries and
// the optional triples come from the simulation's own record of which aligned columns it mutated.
#include <cub/cub.cuh>

#include "sd_internal.h"

namespace {

struct Rng {
    uint64_t s;
    __device__ explicit Rng(uint64_t seed, uint64_t idx) : s(seed * 0x9E3779B97F4A7C15ULL + idx * 0xD1B54A32D192ED03ULL + 0x632BE59BD9B4E019ULL) { next(); }
    __device__ uint64_t next() {  // splitmix64
        uint64_t z = (s += 0x9E3779B97F4A7C15ULL);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        return z ^ (z >> 31);
    }
    __device__ float uniform() { return (float)(next() >> 40) * (1.0f / 16777216.0f); }
    __device__ double uniform_d() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    __device__ uint32_t below(uint32_t n) { return (uint32_t)(((next() >> 32) * (uint64_t)n) >> 32); }
};

struct Plan {
    uint32_t chrom, pos, strand;  // chromosome index, local position, 0 '+' / 1 '-'
    uint32_t ncig;
    uint32_t cig[3];
    uint32_t mapq;
    bool labelled;
};

// The part of a read that does not depend on the genome: placement and alignment shape.
__device__ Plan make_plan(Rng &rng, const sd_synth_params &P, const sd_synth_tables &T) {
    Plan pl;
    const uint32_t L = P.read_len;
    // UTR by expression weight
    const uint32_t u0 = P.utr_first, u1 = P.utr_first + (P.utr_count ? P.utr_count : T.n_utr - P.utr_first);
    const double w0 = u0 ? T.utr_cum_weight[u0 - 1] : 0.0, w1 = T.utr_cum_weight[u1 - 1];
    const double r = w0 + rng.uniform_d() * (w1 - w0);
    uint32_t lo = u0, hi = u1 - 1;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (T.utr_cum_weight[mid] > r) hi = mid; else lo = mid + 1;
    }
    const uint32_t u = lo;
    pl.chrom = T.utr_chrom[u];
    const uint32_t clen = T.chrom_len[pl.chrom];
    // alignment shape
    const float rs = rng.uniform();
    uint32_t ref_len = L;
    if (rs < P.frac_softclip && L > 12) {
        const uint32_t k = 1 + rng.below(5);
        pl.ncig = 2;
        if (rng.below(2)) { pl.cig[0] = (k << 4) | 4u; pl.cig[1] = ((L - k) << 4) | 0u; }
        else { pl.cig[0] = ((L - k) << 4) | 0u; pl.cig[1] = (k << 4) | 4u; }
        ref_len = L - k;
    } else if (rs < P.frac_softclip + P.frac_indel && L > 30) {
        const uint32_t k = 1 + rng.below(3);
        const uint32_t a = 10 + rng.below(L - 20 - k);
        pl.ncig = 3;
        if (rng.below(2)) {  // insertion
            pl.cig[0] = (a << 4) | 0u; pl.cig[1] = (k << 4) | 1u; pl.cig[2] = ((L - a - k) << 4) | 0u;
            ref_len = L - k;
        } else {             // deletion
            pl.cig[0] = (a << 4) | 0u; pl.cig[1] = (k << 4) | 2u; pl.cig[2] = ((L - a) << 4) | 0u;
            ref_len = L + k;
        }
    } else {
        pl.ncig = 1;
        pl.cig[0] = (L << 4) | 0u;
    }
    // start so that the alignment overlaps the UTR by at least one base, inside the chromosome
    const uint32_t us = T.utr_start[u], ue = T.utr_end[u];
    const int64_t first = (int64_t)us - (int64_t)ref_len + 1, last = (int64_t)ue - 1;
    int64_t pos = first + (int64_t)(rng.uniform_d() * (double)(last - first + 1));
    if (pos > last) pos = last;
    if (pos + ref_len > clen) pos = (int64_t)clen - ref_len;
    if (pos < 0) pos = 0;
    pl.pos = (uint32_t)pos;
    pl.strand = T.utr_strand[u];
    if (rng.uniform() < P.frac_antisense) pl.strand ^= 1u;
    pl.mapq = (rng.uniform() < P.frac_multimap) ? 0u : 60u;
    pl.labelled = rng.uniform() < (T.utr_label_frac ? T.utr_label_frac[u] : P.frac_labelled);
    return pl;
}

__device__ __forceinline__ uint32_t ref_code(const sd_synth_tables &T, uint64_t g) {
    const uint32_t w = (uint32_t)(g >> 5), b = (uint32_t)(g & 31);
    if ((T.nmask[w] >> b) & 1u) return 4u;
    return ((T.lo[w] >> b) & 1u) | (((T.hi[w] >> b) & 1u) << 1);
}

// Simulate one read.  Em receives base(i, code 0..3), qual(i, q), mismatch(conv, readPos1, refPos1).
template <class Em>
__device__ void simulate(Rng &rng, const sd_synth_params &P, const sd_synth_tables &T, const Plan &pl, Em &em) {
    const uint64_t g0 = (uint64_t)T.chrom_off[pl.chrom] + pl.pos;
    const uint32_t conv_from = pl.strand ? 0u : 3u, conv_to = pl.strand ? 2u : 1u;  // A>G on '-', T>C on '+'
    uint32_t q = 0, r = 0;
    for (uint32_t i = 0; i < pl.ncig; i++) {
        const uint32_t op = pl.cig[i] & 15u, len = pl.cig[i] >> 4;
        if (op == 0) {
            for (uint32_t j = 0; j < len; j++) {
                const uint32_t rc = ref_code(T, g0 + r);
                uint32_t b = rc < 4 ? rc : rng.below(4);
                if (T.snp_tc && rc < 4) {   // the sample's own SNPs: the read carries the alternative allele
                    const uint64_t g = g0 + r;
                    const uint32_t w = (uint32_t)(g >> 5), bit = 1u << (g & 31);
                    const uint32_t alt = (T.snp_tc[w] & bit) ? 1u : ((T.snp_ag[w] & bit) ? 2u : ((T.snp_other && (T.snp_other[w] & bit)) ? ((rc + 1u) & 3u) : rc));
                    if (alt != rc && rng.uniform() < T.p_snp_allele) b = alt;
                }
                if (rc == conv_from && pl.labelled && rng.uniform() < P.p_conv) b = conv_to;
                if (rng.uniform() < P.p_err) b = (b + 1 + rng.below(3)) & 3u;
                em.base(q, b);
                if (b != rc) em.mismatch(5u * rc + b, q + 1, r + 1);
                q++;
                r++;
            }
        } else if (op == 1 || op == 4) {
            for (uint32_t j = 0; j < len; j++) em.base(q++, rng.below(4));
        } else if (op == 2) {
            r += len;
        }
    }
    for (uint32_t i = 0; i < P.read_len; i++) {
        const float x = rng.uniform();
        em.qual(i, x < 0.1f ? 20u : (x < 0.3f ? 30u : 37u));
    }
}

struct CountEm {
    uint32_t want, nmp, nmis;
    __device__ void base(uint32_t, uint32_t) {}
    __device__ void qual(uint32_t, uint32_t) {}
    __device__ void mismatch(uint32_t conv, uint32_t, uint32_t) {
        nmis++;
        if (conv == want) nmp++;
    }
};

__global__ void synth_plan_kernel(sd_synth_params P, sd_synth_tables T, uint32_t *counts) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n_reads) return;
    Rng rng(P.seed, P.first_read + (P.order ? (uint64_t)P.order[i] : i));
    const Plan pl = make_plan(rng, P, T);
    CountEm em{pl.strand ? 2u : 16u, 0u, 0u};
    simulate(rng, P, T, pl, em);
    counts[i] = pl.ncig | (em.nmp << 16);
}

__global__ void synth_positions_kernel(sd_synth_params P, sd_synth_tables T, uint32_t *gpos, uint8_t *cls) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n_reads) return;
    Rng rng(P.seed, P.first_read + i);
    const Plan pl = make_plan(rng, P, T);
    gpos[i] = T.chrom_off[pl.chrom] + pl.pos;
    if (cls) cls[i] = (uint8_t)((pl.ncig == 1 ? 0u : 2u) | pl.strand);   // same classes as sd_decode_bam's grouping
}

// per tile: padded byte sizes of the four columns
__global__ void synth_tile_sizes_kernel(sd_synth_params P, const uint32_t *counts, uint32_t n_tiles, unsigned long long *sizes /*[4][n_tiles+1]*/) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > n_tiles) return;
    unsigned long long bs = 0, bq = 0, bc = 0, bm = 0;
    if (t < n_tiles) {
        const uint64_t b = (uint64_t)t * P.tile_reads, e = min((uint64_t)P.n_reads, b + P.tile_reads);
        bs = 16ull * P.tile_reads * ((P.read_len + 31u) >> 5);   // group-major block with all row slots
        for (uint64_t i = b; i < e; i++) {
            bq += P.read_len;
            bc += 4ull * (counts[i] & 0xFFFFu);
            bm += 4ull * (counts[i] >> 16);
        }
    }
    const size_t stride = (size_t)n_tiles + 1;
    sizes[0 * stride + t] = (bs + 15ull) & ~15ull;
    sizes[1 * stride + t] = (bq + 15ull) & ~15ull;
    sizes[2 * stride + t] = (bc + 15ull) & ~15ull;
    sizes[3 * stride + t] = (bm + 15ull) & ~15ull;
}

__global__ void synth_pack_tiles_kernel(const unsigned long long *offs, uint32_t n_tiles, sd_tile *tiles) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > n_tiles) return;
    const size_t stride = (size_t)n_tiles + 1;
    sd_tile x;
    x.seq_off = offs[0 * stride + t];
    x.qual_off = offs[1 * stride + t];
    x.cig_off = offs[2 * stride + t];
    x.mp_off = offs[3 * stride + t];
    tiles[t] = x;
}

struct FillEm {
    uint32_t *seq;      // this read's slot of group 0 in the tile's group-major block
    uint8_t *qv;
    uint32_t *mp;       // mp entries of this read (may be null)
    int32_t *triples;   // [1 + 3 * SD_SYNTH_MAX_TRIPLES] (may be null)
    uint32_t want, nmp, nmis, ntri;
    __device__ void base(uint32_t i, uint32_t code) {
        // BAM nibble of A,C,G,T = 1,2,4,8 -> plane `code` gets the bit
        seq[(i >> 5) * (4u * 32u) + code] |= 1u << (i & 31u);   // group-major tile block: 32 row slots of 4 words per group
    }
    __device__ void qual(uint32_t i, uint32_t q) { qv[i] = (uint8_t)q; }
    __device__ void mismatch(uint32_t conv, uint32_t rp1, uint32_t fp1) {
        // columns over a reference N are left out of the triples (they can never be a T>C / A>G entry)
        if (triples && conv < 20u) {
            if (ntri < SD_SYNTH_MAX_TRIPLES) {
                triples[1 + 3 * ntri] = (int32_t)conv;
                triples[2 + 3 * ntri] = (int32_t)rp1;
                triples[3 + 3 * ntri] = (int32_t)fp1;
            }
            ntri++;
        }
        nmis++;
        if (conv == want) {
            if (mp) mp[nmp] = (rp1 - 1u) | ((fp1 - 1u) << 16);
            nmp++;
        }
    }
};

// one warp per tile (32 reads), one thread per read
__global__ void synth_fill_kernel(sd_synth_params P, sd_synth_tables T, const uint32_t *counts, const sd_tile *tiles,
                                  sd_read_rec *rec, uint32_t *seq, uint8_t *qual, uint32_t *cig, uint32_t *mp, int32_t *triples) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t tile = i >> 5;
    const uint64_t n_pad = ((P.n_reads + 31ull) >> 5) << 5;
    if (i >= n_pad) return;   // whole warps only: n_pad is a multiple of 32
    const bool live = i < P.n_reads;
    const uint32_t c = live ? counts[i] : 0u;
    uint32_t cs = c & 0xFFFFu, ms = c >> 16;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t tc_ = __shfl_up_sync(0xFFFFFFFFu, cs, d), tm_ = __shfl_up_sync(0xFFFFFFFFu, ms, d);
        if (lane >= (uint32_t)d) { cs += tc_; ms += tm_; }
    }
    const uint32_t cig_before = cs - (c & 0xFFFFu), mp_before = ms - (c >> 16);
    sd_read_rec &o = rec[i];
    const uint32_t groups = (P.read_len + 31u) >> 5;
    uint32_t *my_seq = seq + tiles[tile].seq_off / 4 + (size_t)lane * 4u;   // slot (group 0, row lane); groups are 128 words apart
    for (uint32_t g = 0; g < groups; g++)
        *reinterpret_cast<uint4 *>(my_seq + (size_t)g * 128u) = make_uint4(0u, 0u, 0u, 0u);   // pad rows too: deterministic columns
    if (!live) {
        o.pos = 0;
        o.ref_mapq_flag = (uint32_t)SD_REF_NONE | ((uint32_t)(SD_F_PAD | SD_F_SKIP | SD_F_UNMAPPED) << 24);
        o.xi = 0.f;
        o.lseq_ncig = 0;
        o.nm_nmp = 0;
        return;
    }
    const uint32_t k = lane;
    const sd_tile t = tiles[tile];
    FillEm em;
    em.seq = my_seq;
    em.qv = qual + t.qual_off + (size_t)k * P.read_len;
    em.mp = mp ? mp + t.mp_off / 4 + mp_before : nullptr;
    em.triples = triples ? triples + i * (1 + 3 * SD_SYNTH_MAX_TRIPLES) : nullptr;
    Rng rng(P.seed, P.first_read + (P.order ? (uint64_t)P.order[i] : i));
    const Plan pl = make_plan(rng, P, T);
    em.want = pl.strand ? 2u : 16u;
    em.nmp = 0;
    em.nmis = 0;
    em.ntri = 0;
    simulate(rng, P, T, pl, em);
    if (em.triples) em.triples[0] = (int32_t)em.ntri;
    uint32_t *pc = cig + t.cig_off / 4 + cig_before;
    uint32_t aligned = 0, indel = 0;
    for (uint32_t j = 0; j < pl.ncig; j++) {
        pc[j] = pl.cig[j];
        const uint32_t op = pl.cig[j] & 15u, len = pl.cig[j] >> 4;
        if (op == 0) aligned += len;
        if (op == 1 || op == 2) indel += len;
    }
    const uint32_t nm = em.nmis + indel;
    uint32_t f = pl.strand ? SD_F_REVERSE : 0u;
    if (em.nmis) f |= SD_F_HAS_MP;
    f |= SD_F_NEWNAME;
    o.pos = (int32_t)pl.pos;
    o.ref_mapq_flag = pl.chrom | (pl.mapq << 16) | (f << 24);
    o.xi = 1.0f - (float)nm / (float)(aligned ? aligned : 1u);
    o.lseq_ncig = P.read_len | (pl.ncig << 16);
    o.nm_nmp = (nm > 65535u ? 65535u : nm) | (em.nmp << 16);
}

__global__ void synth_genome_kernel(uint64_t seed, uint32_t *lo, uint32_t *hi, uint32_t *nm, uint64_t n_words, float n_frac) {
    const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_words) return;
    Rng rng(seed, w);
    lo[w] = (uint32_t)rng.next();
    hi[w] = (uint32_t)rng.next();
    Rng blk(seed ^ 0xABCDEF12345ULL, w >> 1);  // 64-bp blocks
    nm[w] = blk.uniform() < n_frac ? 0xFFFFFFFFu : 0u;
}

// pad positions (outside every chromosome) are N
__global__ void synth_genome_pad_kernel(uint32_t *nm, uint64_t n_words, const uint32_t *chrom_off, const uint32_t *chrom_len, uint32_t n_chrom) {
    const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_words) return;
    const uint64_t b0 = w * 32ull;
    // find the chromosome whose span could cover this word (linear scan, n_chrom is small)
    uint32_t keep = 0u;
    for (uint32_t c = 0; c < n_chrom; c++) {
        const uint64_t s = chrom_off[c], e = s + chrom_len[c];
        if (e <= b0 || s >= b0 + 32ull) continue;
        const uint32_t a = s > b0 ? (uint32_t)(s - b0) : 0u, b = e < b0 + 32ull ? (uint32_t)(e - b0) : 32u;
        keep |= (b == 32u ? 0xFFFFFFFFu : ((1u << b) - 1u)) & ~((1u << a) - 1u);
    }
    nm[w] |= ~keep;
}

}  // namespace

extern "C" int sd_synth_positions(sd_ctx *ctx, const sd_synth_params *p, const sd_synth_tables *t, uint32_t *d_gpos, uint8_t *d_cls) {
    if (!ctx || !p || !t || !d_gpos) return sd_fail(ctx, SD_ERR_ARG, "sd_synth_positions: null argument");
    if (t->n_utr == 0 || p->utr_first >= t->n_utr || p->utr_first + p->utr_count > t->n_utr) return sd_fail(ctx, SD_ERR_ARG, "bad UTR range");
    SD_CUDA(ctx, cudaSetDevice(ctx->device));
    synth_positions_kernel<<<(unsigned)((p->n_reads + 255) / 256), 256, 0, ctx->stream>>>(*p, *t, d_gpos, d_cls);
    ctx->launches++;
    SD_CUDA(ctx, cudaGetLastError());
    SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SD_OK;
}

extern "C" int sd_synth_plan(sd_ctx *ctx, const sd_synth_params *p, const sd_synth_tables *t, uint32_t *d_counts,
                             sd_tile *d_tiles, uint64_t h_totals[4]) {
    if (!ctx || !p || !t || !d_counts || !d_tiles || !h_totals) return sd_fail(ctx, SD_ERR_ARG, "sd_synth_plan: null argument");
    if (p->tile_reads != 32) return sd_fail(ctx, SD_ERR_ARG, "tile_reads must be 32");
    if (p->read_len == 0 || p->read_len > 65535 || p->n_reads == 0) return sd_fail(ctx, SD_ERR_ARG, "bad read_len / n_reads");
    if (t->n_utr == 0 || p->utr_first >= t->n_utr || p->utr_first + p->utr_count > t->n_utr) return sd_fail(ctx, SD_ERR_ARG, "bad UTR range");
    SD_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint64_t n_tiles64 = (p->n_reads + p->tile_reads - 1) / p->tile_reads;
    if (n_tiles64 > 0xFFFFFFF0ull) return sd_fail(ctx, SD_ERR_LIMIT, "too many reads for one batch");
    const uint32_t n_tiles = (uint32_t)n_tiles64;
    synth_plan_kernel<<<(unsigned)((p->n_reads + 255) / 256), 256, 0, ctx->stream>>>(*p, *t, d_counts);
    ctx->launches++;
    SD_CUDA(ctx, cudaGetLastError());
    const size_t stride = (size_t)n_tiles + 1;
    unsigned long long *sizes = nullptr, *offs = nullptr;
    SD_CUDA(ctx, cudaMalloc(&sizes, 4 * stride * sizeof(unsigned long long)));
    SD_CUDA(ctx, cudaMalloc(&offs, 4 * stride * sizeof(unsigned long long)));
    synth_tile_sizes_kernel<<<(unsigned)((stride + 255) / 256), 256, 0, ctx->stream>>>(*p, d_counts, n_tiles, sizes);
    ctx->launches++;
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, sizes, offs, (int)stride, ctx->stream);
    void *tmp = nullptr;
    SD_CUDA(ctx, cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 16));
    for (int c = 0; c < 4; c++) {
        cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, sizes + c * stride, offs + c * stride, (int)stride, ctx->stream);
        ctx->launches++;
    }
    synth_pack_tiles_kernel<<<(unsigned)((stride + 255) / 256), 256, 0, ctx->stream>>>(offs, n_tiles, d_tiles);
    ctx->launches++;
    sd_tile last;
    SD_CUDA(ctx, cudaMemcpyAsync(&last, d_tiles + n_tiles, sizeof last, cudaMemcpyDeviceToHost, ctx->stream));
    SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(sizes); cudaFree(offs); cudaFree(tmp);
    SD_CUDA(ctx, cudaGetLastError());
    h_totals[0] = last.seq_off; h_totals[1] = last.qual_off; h_totals[2] = last.cig_off; h_totals[3] = last.mp_off;
    return SD_OK;
}

extern "C" int sd_synth_fill(sd_ctx *ctx, const sd_synth_params *p, const sd_synth_tables *t, const uint32_t *d_counts,
                             sd_batch *out, int32_t *d_triples) {
    if (!ctx || !p || !t || !d_counts || !out || !out->rec || !out->seq || !out->qual || !out->cigar || !out->tiles)
        return sd_fail(ctx, SD_ERR_ARG, "sd_synth_fill: null argument");
    SD_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint32_t n_tiles = (uint32_t)((p->n_reads + p->tile_reads - 1) / p->tile_reads);
    synth_fill_kernel<<<(unsigned)(((uint64_t)n_tiles * 32 + 255) / 256), 256, 0, ctx->stream>>>(*p, *t, d_counts, out->tiles, const_cast<sd_read_rec *>(out->rec),
                                                        const_cast<uint32_t *>(out->seq), const_cast<uint8_t *>(out->qual),
                                                        const_cast<uint32_t *>(out->cigar), const_cast<uint32_t *>(out->mp), d_triples);
    ctx->launches++;
    SD_CUDA(ctx, cudaGetLastError());
    // largest tile sizes (cigar / mp vary per tile): read the tile table back
    std::vector<sd_tile> ht((size_t)n_tiles + 1);
    SD_CUDA(ctx, cudaMemcpyAsync(ht.data(), out->tiles, ht.size() * sizeof(sd_tile), cudaMemcpyDeviceToHost, ctx->stream));
    SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    uint32_t ms = 0, mq = 0, mc = 0, mm = 0;
    for (uint32_t i = 0; i < n_tiles; i++) {
        ms = std::max<uint32_t>(ms, (uint32_t)(ht[i + 1].seq_off - ht[i].seq_off));
        mq = std::max<uint32_t>(mq, (uint32_t)(ht[i + 1].qual_off - ht[i].qual_off));
        mc = std::max<uint32_t>(mc, (uint32_t)(ht[i + 1].cig_off - ht[i].cig_off));
        mm = std::max<uint32_t>(mm, (uint32_t)(ht[i + 1].mp_off - ht[i].mp_off));
    }
    out->n_reads = p->n_reads;
    out->n_tiles = n_tiles;
    out->tile_reads = p->tile_reads;
    out->max_lseq = p->read_len;
    out->qual_bits = 8;
    out->seq_bits = 4;
    out->max_span = p->read_len + 3;   // at most one deletion of up to three bases per read (make_plan)
    out->max_tile_seq = ms; out->max_tile_qual = mq; out->max_tile_cig = mc; out->max_tile_mp = mm;
    return SD_OK;
}

extern "C" int sd_synth_genome(sd_ctx *ctx, uint64_t seed, uint32_t *d_lo, uint32_t *d_hi, uint32_t *d_nmask, uint64_t n_words,
                               const uint32_t *h_chrom_off, const uint32_t *h_chrom_len, uint32_t n_chrom, float n_frac) {
    if (!ctx || !d_lo || !d_hi || !d_nmask || !h_chrom_off || !h_chrom_len) return sd_fail(ctx, SD_ERR_ARG, "sd_synth_genome: null argument");
    SD_CUDA(ctx, cudaSetDevice(ctx->device));
    uint32_t *d_off = nullptr, *d_len = nullptr;
    SD_CUDA(ctx, cudaMalloc(&d_off, n_chrom * sizeof(uint32_t)));
    SD_CUDA(ctx, cudaMalloc(&d_len, n_chrom * sizeof(uint32_t)));
    SD_CUDA(ctx, cudaMemcpyAsync(d_off, h_chrom_off, n_chrom * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    SD_CUDA(ctx, cudaMemcpyAsync(d_len, h_chrom_len, n_chrom * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    const unsigned blocks = (unsigned)((n_words + 255) / 256);
    synth_genome_kernel<<<blocks, 256, 0, ctx->stream>>>(seed, d_lo, d_hi, d_nmask, n_words, n_frac);
    synth_genome_pad_kernel<<<blocks, 256, 0, ctx->stream>>>(d_nmask, n_words, d_off, d_len, n_chrom);
    ctx->launches += 2;
    SD_CUDA(ctx, cudaGetLastError());
    SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(d_off); cudaFree(d_len);
    return SD_OK;
}
