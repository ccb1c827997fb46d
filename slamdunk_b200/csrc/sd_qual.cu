// sd_qual.cu -- transfer coding of the base-quality column.
//
// The count kernel reads one byte per base quality (and only at candidate positions).  Over PCIe the column is the largest
// part of a batch, and sequencers emit few distinct values (Illumina bins qualities into 4-8 levels), so a HOST batch may
// hold 2- or 4-bit dictionary codes instead (sd_batch.qual_bits / qual_dict): sd_count_host copies the codes and expands
// them on the device, next to the kernel that consumes them.
#include <algorithm>

#include "sd_internal.h"

namespace {

// one thread per 16 output bytes (4 or 8 bytes of codes)
__global__ void qual_unpack_kernel(const uint8_t *__restrict__ packed, uint64_t n_bases, uint32_t bits, uint4 dict_words, uint8_t *__restrict__ out) {
    const uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16ull;
    if (i >= n_bases) return;
    const uint32_t dw[4] = {dict_words.x, dict_words.y, dict_words.z, dict_words.w};
    auto dict = [&](uint32_t code) { return (dw[code >> 2] >> (8u * (code & 3u))) & 0xFFu; };
    uint32_t o[4] = {0, 0, 0, 0};
    if (bits == 2) {
        const uint32_t w = *reinterpret_cast<const uint32_t *>(packed + i / 4);   // 16 codes
#pragma unroll
        for (int k = 0; k < 16; k++) o[k >> 2] |= dict((w >> (2 * k)) & 3u) << (8 * (k & 3));
    } else {
        const uint2 w = *reinterpret_cast<const uint2 *>(packed + i / 2);
#pragma unroll
        for (int k = 0; k < 8; k++) o[k >> 2] |= dict((w.x >> (4 * k)) & 15u) << (8 * (k & 3));
#pragma unroll
        for (int k = 0; k < 8; k++) o[2 + (k >> 2)] |= dict((w.y >> (4 * k)) & 15u) << (8 * (k & 3));
    }
    *reinterpret_cast<uint4 *>(out + i) = make_uint4(o[0], o[1], o[2], o[3]);
}

__global__ void qual_seen_kernel(const uint8_t *__restrict__ q, uint64_t n, unsigned int *seen /*[8] bit per value*/) {
    __shared__ unsigned int s[8];
    if (threadIdx.x < 8) s[threadIdx.x] = 0u;
    __syncthreads();
    unsigned int mine[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16ull; i < n; i += (uint64_t)gridDim.x * blockDim.x * 16ull) {
        const uint4 w = *reinterpret_cast<const uint4 *>(q + i);
        const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const uint32_t v = (ww[k >> 2] >> (8 * (k & 3))) & 0xFFu;
            mine[v >> 5] |= 1u << (v & 31u);
        }
    }
#pragma unroll
    for (int k = 0; k < 8; k++)
        if (mine[k]) atomicOr(&s[k], mine[k]);
    __syncthreads();
    if (threadIdx.x < 8 && s[threadIdx.x]) atomicOr(&seen[threadIdx.x], s[threadIdx.x]);
}

struct CodeTable { uint8_t code[256]; };

__global__ void qual_pack_kernel(const uint8_t *__restrict__ q, uint64_t n_bases, uint32_t bits, CodeTable t, uint8_t *__restrict__ packed) {
    const uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16ull;
    if (i >= n_bases) return;
    const uint4 w = *reinterpret_cast<const uint4 *>(q + i);
    const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
    if (bits == 2) {
        uint32_t o = 0;
#pragma unroll
        for (int k = 0; k < 16; k++) o |= (uint32_t)t.code[(ww[k >> 2] >> (8 * (k & 3))) & 0xFFu] << (2 * k);
        *reinterpret_cast<uint32_t *>(packed + i / 4) = o;
    } else {
        uint32_t o0 = 0, o1 = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) o0 |= (uint32_t)t.code[(ww[k >> 2] >> (8 * (k & 3))) & 0xFFu] << (4 * k);
#pragma unroll
        for (int k = 0; k < 8; k++) o1 |= (uint32_t)t.code[(ww[2 + (k >> 2)] >> (8 * (k & 3))) & 0xFFu] << (4 * k);
        *reinterpret_cast<uint2 *>(packed + i / 2) = make_uint2(o0, o1);
    }
}

}  // namespace

// internal: used by sd_count_host on its own stream
int sd_qual_unpack_on(sd_ctx *ctx, cudaStream_t stream, const uint8_t *d_packed, uint64_t n_bases, uint32_t bits, const uint8_t *dict16,
                      uint8_t *d_qual) {
    if (bits != 2 && bits != 4) return sd_fail(ctx, SD_ERR_ARG, "quality coding of %u bits is not supported", bits);
    if (n_bases % 16) return sd_fail(ctx, SD_ERR_ARG, "quality column length must be a multiple of 16");
    if (n_bases == 0) return SD_OK;
    uint4 dw;
    memcpy(&dw, dict16, 16);
    const uint64_t threads = n_bases / 16;
    qual_unpack_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(d_packed, n_bases, bits, dw, d_qual);
    ctx->launches++;
    SD_CUDA(ctx, cudaGetLastError());
    return SD_OK;
}

extern "C" int sd_qual_unpack(sd_ctx *ctx, const uint8_t *d_packed, uint64_t n_bases, uint32_t bits, const uint8_t *dict16, uint8_t *d_qual) {
    if (!ctx || !d_packed || !dict16 || !d_qual) return sd_fail(ctx, SD_ERR_ARG, "sd_qual_unpack: null argument");
    SD_CUDA(ctx, cudaSetDevice(ctx->device));
    const int rc = sd_qual_unpack_on(ctx, ctx->stream, d_packed, n_bases, bits, dict16, d_qual);
    if (rc != SD_OK) return rc;
    SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SD_OK;
}

extern "C" int sd_qual_pack(sd_ctx *ctx, const uint8_t *d_qual, uint64_t n_bases, uint8_t *d_packed, uint32_t *bits, uint8_t *dict16) {
    if (!ctx || !d_qual || !bits || !dict16) return sd_fail(ctx, SD_ERR_ARG, "sd_qual_pack: null argument");
    if (n_bases % 16) return sd_fail(ctx, SD_ERR_ARG, "quality column length must be a multiple of 16");
    SD_CUDA(ctx, cudaSetDevice(ctx->device));
    *bits = 8;
    memset(dict16, 0, 16);
    if (n_bases == 0) return SD_OK;
    unsigned int *d_seen = nullptr, seen[8];
    SD_CUDA(ctx, cudaMalloc(&d_seen, sizeof seen));
    cudaMemsetAsync(d_seen, 0, sizeof seen, ctx->stream);
    qual_seen_kernel<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(d_qual, n_bases, d_seen);
    ctx->launches++;
    cudaError_t e = cudaMemcpyAsync(seen, d_seen, sizeof seen, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_seen);
    if (e != cudaSuccess) return sd_fail(ctx, SD_ERR_CUDA, "sd_qual_pack: %s", cudaGetErrorString(e));
    CodeTable t;
    memset(&t, 0, sizeof t);
    int distinct = 0;
    for (int v = 0; v < 256; v++)
        if ((seen[v >> 5] >> (v & 31)) & 1u) {
            if (distinct < 16) { dict16[distinct] = (uint8_t)v; t.code[v] = (uint8_t)distinct; }
            distinct++;
        }
    if (distinct > 16) return SD_OK;   // not packable: *bits stays 8
    *bits = distinct <= 4 ? 2u : 4u;
    if (!d_packed) return sd_fail(ctx, SD_ERR_ARG, "sd_qual_pack: d_packed is null");
    const uint64_t threads = n_bases / 16;
    qual_pack_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, ctx->stream>>>(d_qual, n_bases, *bits, t, d_packed);
    ctx->launches++;
    SD_CUDA(ctx, cudaGetLastError());
    SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SD_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// 2-bit form of the SEQ column (sd_batch.seq_bits = 2): the count path only asks "is this base C / G", so A=0 C=1 G=2 T=3
// with everything else coded A is lossless for it and halves the column.
namespace {
__global__ void seq_pack_kernel(const uint4 *seq4, uint64_t n_groups, uint2 *seq2) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_groups) return;
    const uint4 p = seq4[i];
    const uint32_t onlyC = ~p.x & p.y & ~p.z & ~p.w, onlyG = ~p.x & ~p.y & p.z & ~p.w, onlyT = ~p.x & ~p.y & ~p.z & p.w;
    seq2[i] = make_uint2(onlyC | onlyT, onlyG | onlyT);
}
__global__ void tiles_seq2_kernel(const sd_tile *in, uint32_t n, sd_tile *out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    sd_tile t = in[i];
    t.seq_off >>= 1;
    out[i] = t;
}
}  // namespace

extern "C" int sd_seq_pack(sd_ctx *ctx, const sd_batch *b, uint32_t *d_seq2, sd_tile *d_tiles2) {
    if (!ctx || !b || !d_seq2 || !d_tiles2) return sd_fail(ctx, SD_ERR_ARG, "sd_seq_pack: null argument");
    if (b->seq_bits != 0 && b->seq_bits != 4) return sd_fail(ctx, SD_ERR_ARG, "sd_seq_pack: the column is not in the 4-plane form");
    SD_CUDA(ctx, cudaSetDevice(ctx->device));
    sd_tile last;
    SD_CUDA(ctx, cudaMemcpyAsync(&last, b->tiles + b->n_tiles, sizeof last, cudaMemcpyDeviceToHost, ctx->stream));
    SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const uint64_t n_groups = last.seq_off / 16;
    if (n_groups) {
        seq_pack_kernel<<<(unsigned)((n_groups + 255) / 256), 256, 0, ctx->stream>>>(reinterpret_cast<const uint4 *>(b->seq), n_groups,
                                                                                   reinterpret_cast<uint2 *>(d_seq2));
        ctx->launches++;
    }
    tiles_seq2_kernel<<<(b->n_tiles + 1 + 255) / 256, 256, 0, ctx->stream>>>(b->tiles, b->n_tiles + 1, d_tiles2);
    ctx->launches++;
    SD_CUDA(ctx, cudaGetLastError());
    return SD_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Plane form of the quality column (sd_batch.qual_form = 1): bit-planes of the dictionary code per 32-base group, in the
// group-major tile geometry of the SEQ column, so that the count kernel stages a tile's qualities with the tile and tests
// 32 bases per LOP3 pair instead of fetching one byte per candidate.
extern "C" int sd_qual_dict(sd_ctx *ctx, const uint8_t *d_qual, uint64_t n_bases, uint32_t *bits, uint8_t *dict16) {
    if (!ctx || !d_qual || !bits || !dict16) return sd_fail(ctx, SD_ERR_ARG, "sd_qual_dict: null argument");
    if (n_bases % 16) return sd_fail(ctx, SD_ERR_ARG, "quality column length must be a multiple of 16");
    SD_CUDA(ctx, cudaSetDevice(ctx->device));
    *bits = 8;
    memset(dict16, 0, 16);
    if (n_bases == 0) { *bits = 2; return SD_OK; }
    unsigned int *d_seen = nullptr, seen[8];
    SD_CUDA(ctx, cudaMalloc(&d_seen, sizeof seen));
    cudaMemsetAsync(d_seen, 0, sizeof seen, ctx->stream);
    qual_seen_kernel<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(d_qual, n_bases, d_seen);
    ctx->launches++;
    cudaError_t e = cudaMemcpyAsync(seen, d_seen, sizeof seen, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_seen);
    if (e != cudaSuccess) return sd_fail(ctx, SD_ERR_CUDA, "sd_qual_dict: %s", cudaGetErrorString(e));
    int distinct = 0;
    for (int v = 0; v < 256; v++)
        if ((seen[v >> 5] >> (v & 31)) & 1u) {
            if (distinct < 16) dict16[distinct] = (uint8_t)v;
            distinct++;
        }
    if (distinct <= 16) *bits = distinct <= 4 ? 2u : 4u;
    return SD_OK;
}

namespace {
// one warp per tile, lane = row slot; every (group, row) slot of the tile's block is written (zeros past the read's end)
template <int QB>
__global__ void qual_planes_kernel(const sd_read_rec *rec, const uint8_t *qual, const sd_tile *tiles, uint32_t n_tiles,
                                   uint32_t seq_group_bytes, CodeTable t, uint32_t *planes) {
    __shared__ uint8_t lut[256];
    for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = t.code[i];
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t tile = warp; tile < n_tiles; tile += n_warps) {
        const uint32_t lseq = rec[(size_t)tile * 32 + lane].lseq_ncig & 0xFFFFu;
        uint32_t off = lseq;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, off, d);
            if ((int)lane >= d) off += y;
        }
        off -= lseq;
        const sd_tile t0 = tiles[tile], t1 = tiles[tile + 1];
        const uint64_t group0 = t0.seq_off / seq_group_bytes;                                  // (group, row) slots before this tile
        const uint32_t n_grp = (uint32_t)((t1.seq_off - t0.seq_off) / seq_group_bytes) / 32u;  // groups per row slot
        const uint8_t *q = qual + t0.qual_off + off;
        for (uint32_t g = 0; g < n_grp; g++) {
            uint32_t p[QB];
#pragma unroll
            for (int b = 0; b < QB; b++) p[b] = 0u;
            const uint32_t lo = 32u * g, hi = min(lseq, lo + 32u);
            for (uint32_t j = lo; j < hi; j++) {
                const uint32_t c = lut[q[j]];
#pragma unroll
                for (int b = 0; b < QB; b++) p[b] |= ((c >> b) & 1u) << (j - lo);
            }
            uint32_t *dst = planes + (group0 + (uint64_t)g * 32u + lane) * QB;
#pragma unroll
            for (int b = 0; b < QB; b++) dst[b] = p[b];
        }
    }
}
__global__ void tiles_qual_planes_kernel(const sd_tile *in, uint32_t n, uint32_t seq_group_bytes, uint32_t plane_group_bytes, sd_tile *out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    sd_tile t = in[i];
    t.qual_off = t.seq_off / seq_group_bytes * plane_group_bytes;
    out[i] = t;
}
}  // namespace

int sd_qual_to_planes_on(sd_ctx *ctx, cudaStream_t stream, const sd_batch *b, uint32_t bits, const uint8_t *dict16, uint32_t *d_planes,
                         sd_tile *d_tiles_out) {
    if (bits != 2 && bits != 4) return sd_fail(ctx, SD_ERR_ARG, "sd_qual_to_planes: 2 or 4 planes");
    if (b->tile_reads != 32) return sd_fail(ctx, SD_ERR_ARG, "sd_qual_to_planes: tile_reads must be 32");
    if (b->qual_form != 0 || (b->qual_bits != 0 && b->qual_bits != 8)) return sd_fail(ctx, SD_ERR_ARG, "sd_qual_to_planes: the qual column must hold bytes");
    CodeTable t;
    memset(&t, 0, sizeof t);
    for (uint32_t c = 0; c < (1u << bits); c++)
        if (c == 0 || dict16[c] > dict16[c - 1]) t.code[dict16[c]] = (uint8_t)c;   // entries past the used ones are zero
    const uint32_t sgb = b->seq_bits == 2 ? 8u : 16u;
    if (b->n_tiles) {
        const uint32_t blocks = std::min<uint32_t>((b->n_tiles + 7u) / 8u, (uint32_t)ctx->num_sms * 8u);
        if (bits == 2) qual_planes_kernel<2><<<blocks, 256, 0, stream>>>(b->rec, b->qual, b->tiles, b->n_tiles, sgb, t, d_planes);
        else qual_planes_kernel<4><<<blocks, 256, 0, stream>>>(b->rec, b->qual, b->tiles, b->n_tiles, sgb, t, d_planes);
        ctx->launches++;
    }
    tiles_qual_planes_kernel<<<(b->n_tiles + 1 + 255) / 256, 256, 0, stream>>>(b->tiles, b->n_tiles + 1, sgb, 4u * bits, d_tiles_out);
    ctx->launches++;
    SD_CUDA(ctx, cudaGetLastError());
    return SD_OK;
}

extern "C" int sd_qual_to_planes(sd_ctx *ctx, const sd_batch *b, uint32_t bits, const uint8_t *dict16, uint32_t *d_planes,
                                 sd_tile *d_tiles_out) {
    if (!ctx || !b || !dict16 || !d_planes || !d_tiles_out) return sd_fail(ctx, SD_ERR_ARG, "sd_qual_to_planes: null argument");
    SD_CUDA(ctx, cudaSetDevice(ctx->device));
    return sd_qual_to_planes_on(ctx, ctx->stream, b, bits, dict16, d_planes, d_tiles_out);
}
