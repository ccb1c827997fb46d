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
    uint8_t *flags = nullptr;
    uint32_t *sel = nullptr;
    uint64_t *d_n = nullptr;
    void *tmp = nullptr;
    size_t tmp_bytes = 0;
    auto done = [&](int code) {
        cudaFree(flags); cudaFree(sel); cudaFree(d_n); cudaFree(tmp);
        return code;
    };
    SD_CUDA(ctx, cudaMalloc(&flags, track_len));
    if (cudaMalloc(&d_n, sizeof(uint64_t)) != cudaSuccess) return done(sd_fail(ctx, SD_ERR_NOMEM, "sd_track_runs: out of device memory"));
    track_flags_kernel<<<(track_len + 255u) / 256u, 256, 0, ctx->stream>>>(a, flags);
    ctx->launches++;
    cub::CountingInputIterator<uint32_t> positions(0);
    if (d_pos == nullptr || capacity == 0) {   // sizing call: count the change points
        // flags are bytes: sum them as 64-bit through a transform iterator
        cub::TransformInputIterator<uint64_t, cub::CastOp<uint64_t>, const uint8_t *> wide(flags, cub::CastOp<uint64_t>());
        cub::DeviceReduce::Sum(nullptr, tmp_bytes, wide, d_n, (int)track_len, ctx->stream);
        if (cudaMalloc(&tmp, std::max<size_t>(tmp_bytes, 16)) != cudaSuccess) return done(sd_fail(ctx, SD_ERR_NOMEM, "sd_track_runs: out of device memory"));
        cub::DeviceReduce::Sum(tmp, tmp_bytes, wide, d_n, (int)track_len, ctx->stream);
        ctx->launches++;
        uint64_t n = 0;
        cudaError_t e = cudaMemcpyAsync(&n, d_n, sizeof n, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) return done(sd_fail(ctx, SD_ERR_CUDA, "sd_track_runs: %s", cudaGetErrorString(e)));
        *n_runs = n;
        return done(SD_OK);
    }
    if (!d_val) return done(sd_fail(ctx, SD_ERR_ARG, "sd_track_runs: d_val is null"));
    if (cudaMalloc(&sel, (size_t)track_len * sizeof(uint32_t)) != cudaSuccess) return done(sd_fail(ctx, SD_ERR_NOMEM, "sd_track_runs: out of device memory"));
    cub::DeviceSelect::Flagged(nullptr, tmp_bytes, positions, flags, sel, d_n, (int)track_len, ctx->stream);
    if (cudaMalloc(&tmp, std::max<size_t>(tmp_bytes, 16)) != cudaSuccess) return done(sd_fail(ctx, SD_ERR_NOMEM, "sd_track_runs: out of device memory"));
    cub::DeviceSelect::Flagged(tmp, tmp_bytes, positions, flags, sel, d_n, (int)track_len, ctx->stream);
    ctx->launches++;
    uint64_t n = 0;
    {
        cudaError_t e = cudaMemcpyAsync(&n, d_n, sizeof n, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) return done(sd_fail(ctx, SD_ERR_CUDA, "sd_track_runs: %s", cudaGetErrorString(e)));
    }
    *n_runs = n;
    if (n > capacity) return done(sd_fail(ctx, SD_ERR_LIMIT, "sd_track_runs: %llu change points, room for %llu", (unsigned long long)n, (unsigned long long)capacity));
    if (n) {
        cudaError_t e = cudaMemcpyAsync(d_pos, sel, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream);
        if (e != cudaSuccess) return done(sd_fail(ctx, SD_ERR_CUDA, "sd_track_runs: %s", cudaGetErrorString(e)));
        track_gather_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(a, d_pos, n, d_val, d_isfloat);
        ctx->launches++;
    }
    {
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) return done(sd_fail(ctx, SD_ERR_CUDA, "sd_track_runs: %s", cudaGetErrorString(e)));
    }
    return done(SD_OK);
}
