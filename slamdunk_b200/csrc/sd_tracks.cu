// sd_tracks.cu -- run-length view of the per-position arrays: the eight genome-wide bedGraph tracks of
// `alleyoop positional-tracks` (reference: dunks/tcounter.py:332-487, genomewideConversionRates).
//
// The arrays themselves come from the count path (sd_count + sd_finalize) run with one annotation row per chromosome
// strand starting at position 1 -- the reference's iterator is built with startPosition 1 (SlamSeqFile.py:451), so its
// array index p stands for the 0-based genomic position p + 1.  Here a track is compacted on the device into its change
// points (the positions whose value differs from the previous position's); only those travel to the host, where
// sd_write_track prints the reference's run-length lines.
#include <cub/cub.cuh>

#include "sd_internal.h"

namespace {

struct TrackArgs {
    const int *cov, *tc;      // the row's slices of pos_cov (integrated) / pos_tc
    const uint2 *refx;        // {isX, snpX} planes of the row's strand
    uint32_t gstart;          // linear genome coordinate of the row's first position
    uint32_t len;             // positions held by the arrays; positions [len, track_len) read as zero
    uint32_t track_len;
    int kind, cutoff;
};

struct TrackValue { double v; bool is_float; };

// value of the track at position p in the reference's terms (tcounter.py:417-476)
__device__ __forceinline__ TrackValue track_value(const TrackArgs &a, long long p) {
    TrackValue r{0.0, false};
    if (p < 0 || p >= (long long)a.len) return r;
    const int cov = a.cov[p];
    switch (a.kind) {
        case SD_TRACK_COVERAGE: r.v = (double)cov; break;
        case SD_TRACK_COVERAGE_ON_BASE: {   // coverage where the FASTA base is T (plus) / A (minus), :427-439
            const uint64_t g = (uint64_t)a.gstart + (uint64_t)p;
            const bool isx = (__ldg(&a.refx[g >> 5]).x >> (g & 31)) & 1u;
            r.v = (cov > 0 && isx) ? (double)cov : 0.0;
            break;
        }
        case SD_TRACK_CONVERSIONS: r.v = (double)a.tc[p]; break;
        default:                            // conversions / coverage where coverage reaches the cutoff, else int 0 (:460-466)
            if (cov > 0 && cov >= a.cutoff) {
                r.v = (double)a.tc[p] / (double)cov;
                r.is_float = true;
            }
            break;
    }
    return r;
}

__global__ void track_flags_kernel(TrackArgs a, uint8_t *flags) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (long long)a.track_len) return;
    flags[p] = track_value(a, p).v != track_value(a, p - 1).v ? 1 : 0;   // position -1 counts as 0 (prev* start at 0)
}

__global__ void track_gather_kernel(TrackArgs a, const uint32_t *pos, uint64_t n, double *val, uint8_t *is_float) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const TrackValue t = track_value(a, (long long)pos[j]);
    val[j] = t.v;
    if (is_float) is_float[j] = t.is_float ? 1 : 0;
}

}  // namespace

extern "C" int sd_track_runs(sd_ctx *ctx, uint32_t row, int kind, int coverage_cutoff, uint32_t track_len, uint32_t *d_pos,
                             double *d_val, uint8_t *d_isfloat, uint64_t capacity, uint64_t *n_runs) {
    if (!ctx || !n_runs) return sd_fail(ctx, SD_ERR_ARG, "sd_track_runs: null argument");
    if (row >= ctx->n_utr) return sd_fail(ctx, SD_ERR_ARG, "sd_track_runs: row %u out of range", row);
    if (kind < SD_TRACK_COVERAGE || kind > SD_TRACK_RATE) return sd_fail(ctx, SD_ERR_ARG, "sd_track_runs: unknown track kind %d", kind);
    if (!ctx->pos_cov || !ctx->pos_tc) return sd_fail(ctx, SD_ERR_STATE, "sd_track_runs before sd_bind_outputs");
    if (track_len > 0x7FFFFFFFu) return sd_fail(ctx, SD_ERR_LIMIT, "sd_track_runs: track longer than 2^31 - 1 positions");
    SD_CUDA(ctx, cudaSetDevice(ctx->device));
    *n_runs = 0;
    if (track_len == 0) return SD_OK;
    TrackArgs a;
    const uint64_t off = ctx->row_pos_off[row];
    a.len = off == UINT64_MAX ? 0u : ctx->row_pos_len[row];
    a.cov = ctx->pos_cov + (a.len ? off : 0);
    a.tc = ctx->pos_tc + (a.len ? off : 0);
    a.refx = ctx->refx[ctx->row_strand[row] ? 1 : 0];
    a.gstart = ctx->row_gstart[row];
    a.track_len = track_len;
    a.kind = kind;
    a.cutoff = coverage_cutoff;
    const bool sizing = d_pos == nullptr || capacity == 0;
    if (!sizing && !d_val) return sd_fail(ctx, SD_ERR_ARG, "sd_track_runs: d_val is null");

    // scratch (kept in the context, grow only): flags | count | selected positions | cub temp
    cub::CountingInputIterator<uint32_t> positions(0);
    cub::TransformInputIterator<uint64_t, cub::CastOp<uint64_t>, const uint8_t *> wide((const uint8_t *)nullptr, cub::CastOp<uint64_t>());
    size_t tmp_select = 0, tmp_sum = 0;
    cub::DeviceSelect::Flagged(nullptr, tmp_select, positions, (const uint8_t *)nullptr, (uint32_t *)nullptr, (uint64_t *)nullptr, (int)track_len, ctx->stream);
    cub::DeviceReduce::Sum(nullptr, tmp_sum, wide, (uint64_t *)nullptr, (int)track_len, ctx->stream);
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t o_flags = 0, o_count = up(track_len), o_sel = o_count + 256, o_tmp = o_sel + up((size_t)track_len * sizeof(uint32_t));
    const size_t need = o_tmp + up(std::max(tmp_select, tmp_sum)) + 256;
    if (need > ctx->track_scratch_cap) {
        SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->d_track_scratch);
        ctx->d_track_scratch = nullptr;
        ctx->track_scratch_cap = 0;
        if (cudaMalloc(&ctx->d_track_scratch, need) != cudaSuccess) return sd_fail(ctx, SD_ERR_NOMEM, "sd_track_runs: out of device memory (%zu bytes)", need);
        ctx->track_scratch_cap = need;
    }
    unsigned char *base = static_cast<unsigned char *>(ctx->d_track_scratch);
    uint8_t *flags = base + o_flags;
    uint64_t *d_n = reinterpret_cast<uint64_t *>(base + o_count);
    uint32_t *sel = reinterpret_cast<uint32_t *>(base + o_sel);
    void *tmp = base + o_tmp;

    track_flags_kernel<<<(track_len + 255u) / 256u, 256, 0, ctx->stream>>>(a, flags);
    ctx->launches++;
    if (sizing) {
        wide = cub::TransformInputIterator<uint64_t, cub::CastOp<uint64_t>, const uint8_t *>(flags, cub::CastOp<uint64_t>());
        cub::DeviceReduce::Sum(tmp, tmp_sum, wide, d_n, (int)track_len, ctx->stream);
    } else {
        cub::DeviceSelect::Flagged(tmp, tmp_select, positions, flags, sel, d_n, (int)track_len, ctx->stream);
    }
    ctx->launches++;
    uint64_t n = 0;
    {
        cudaError_t e = cudaMemcpyAsync(&n, d_n, sizeof n, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) return sd_fail(ctx, SD_ERR_CUDA, "sd_track_runs: %s", cudaGetErrorString(e));
    }
    *n_runs = n;
    if (sizing || n == 0) return SD_OK;
    if (n > capacity) return sd_fail(ctx, SD_ERR_LIMIT, "sd_track_runs: %llu change points, room for %llu", (unsigned long long)n, (unsigned long long)capacity);
    SD_CUDA(ctx, cudaMemcpyAsync(d_pos, sel, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
    track_gather_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(a, d_pos, n, d_val, d_isfloat);
    ctx->launches++;
    SD_CUDA(ctx, cudaGetLastError());
    SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SD_OK;
}
