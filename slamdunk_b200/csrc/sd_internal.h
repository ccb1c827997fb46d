// Internal declarations shared by the translation units of libslamdunk_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "slamdunk_b200.h"

#define SD_BIN_SHIFT 10   // interval-join bins of 1024 bp over the linear genome
#define SD_REF_PAD_WORDS 64

struct sd_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_k0 = nullptr, ev_k1 = nullptr;
    cudaEvent_t ev_split[2] = {nullptr, nullptr};   // between the lean, staged and plain launches of a timed sd_count
    int split_marks = 0;
    cudaEvent_t ev_copy[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
    bool have_kernel_time = false;
    std::string err;
    int num_sms = 0;
    int64_t launches = 0;

    // reference: per strand {isX, snpX} per 32 bp; strand 0: X = T / T>C SNPs, strand 1: X = A / A>G SNPs
    uint2 *refx[2] = {nullptr, nullptr};
    uint4 *refm = nullptr;   // both strands side by side {isT, T>C SNPs, isA, A>G SNPs}: what a tile's staged reference window is copied from
    uint64_t n_words = 0;
    std::vector<uint32_t> chrom_off, chrom_len;
    uint32_t *d_chrom_off = nullptr, *d_chrom_len = nullptr;
    uint32_t n_chrom = 0;

    // annotation tables, per strand, sorted by global start
    uint32_t n_utr = 0;
    uint32_t n_u[2] = {0, 0};
    uint32_t *d_ustart[2] = {nullptr, nullptr}, *d_uend[2] = {nullptr, nullptr};
    uint32_t *d_urow[2] = {nullptr, nullptr}, *d_useg[2] = {nullptr, nullptr};
    uint4 *d_utab[2] = {nullptr, nullptr};   // {start, end, row, seg} per table entry (one 128-bit load)
    uint4 *d_stab[2] = {nullptr, nullptr};   // {start, end, off_lo, off_hi} per merged segment
    uint32_t *d_bin_first[2] = {nullptr, nullptr};
    uint32_t n_bins = 0;
    uint32_t n_seg[2] = {0, 0};
    uint32_t *d_seg_start[2] = {nullptr, nullptr}, *d_seg_end[2] = {nullptr, nullptr};
    uint64_t *d_seg_off[2] = {nullptr, nullptr};
    uint64_t n_positions = 0;
    // strand-merged annotation for the staged join: entries of both strands sorted by start, two uint4 each
    // ({start, end, row | strand << 31, segment}, {segment start, segment end, posbase lo, hi}; posbase = segment offset - start),
    // and per bin the first entry that can still overlap a window starting in it
    uint4 *d_ment = nullptr;
    uint32_t *d_bin_first_m = nullptr;
    uint32_t n_ment = 0;
    uint32_t ann_epoch = 0;   // bumped by sd_set_reference / sd_set_utrs: span descriptors made before are ignored
    // all BED rows (for Tcontent and host layout)
    uint32_t *d_row_gstart = nullptr, *d_row_gend = nullptr;
    uint8_t *d_row_strand = nullptr;
    uint32_t *d_row_blen = nullptr;     // BED stop - start per row, not clipped (utr.getLength(); sd_read_stats)
    std::vector<uint64_t> row_pos_off;
    std::vector<uint32_t> row_pos_len;
    std::vector<uint32_t> row_gstart;   // host copies of d_row_gstart / d_row_strand
    std::vector<uint8_t> row_strand;

    // caller-owned outputs
    int64_t *counters = nullptr;
    int32_t *pos_cov = nullptr, *pos_tc = nullptr;
    int64_t *stats = nullptr;

    unsigned int *d_tile_counter = nullptr;   // work-queue heads of the count kernel's launches + number of left-over tiles
    uint32_t *d_leftover = nullptr;           // tiles the staged count kernel leaves to the plain one
    uint32_t leftover_cap = 0;

    // grow-only scratch of sd_track_runs (flags | selected positions | cub temp | count)
    void *d_track_scratch = nullptr;
    size_t track_scratch_cap = 0;

    // grow-only scratch of sd_filter (per-record info | padded keep flags)
    void *d_filter_scratch = nullptr;
    size_t filter_scratch_cap = 0;

    // XI dictionary of the wire batch being counted (sd_count_wire)
    float *d_xi_dict = nullptr;
    uint32_t xi_dict_cap = 0;

    // staging for sd_count_host / sd_count_wire
    void *d_stage[2] = {nullptr, nullptr};
    size_t stage_cap = 0;
};

int sd_fail(sd_ctx *ctx, int code, const char *fmt, ...);
int sd_qual_unpack_on(sd_ctx *ctx, cudaStream_t stream, const uint8_t *d_packed, uint64_t n_bases, uint32_t bits,
                      const uint8_t *dict16, uint8_t *d_qual);   // sd_qual.cu
int sd_qual_to_planes_on(sd_ctx *ctx, cudaStream_t stream, const sd_batch *b, uint32_t bits, const uint8_t *dict16, uint32_t *d_planes,
                         sd_tile *d_tiles_out);   // sd_qual.cu
#define SD_CUDA(ctx, call)                                                                        \
    do {                                                                                          \
        cudaError_t e__ = (call);                                                                 \
        if (e__ != cudaSuccess)                                                                   \
            return sd_fail((ctx), SD_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                           __FILE__, __LINE__);                                                   \
    } while (0)
