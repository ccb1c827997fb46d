// sd_wire.cu -- the packed host form of a read batch (sd_wire, include/slamdunk_b200.h) and its expansion on the device.
//
// `slamdunk count` from host memory is bound by PCIe, so what crosses the bus is stripped to what the call needs: one
// column per record field (copied only when used), SEQ and the base qualities as back-to-back rows of 2-bit codes, no
// l_seq / CIGAR for tiles of plain-match reads.  sd_count_wire copies chunks of whole tiles on the copy stream and, on the
// compute stream, turns each chunk into the device batch the count kernel streams: 20-byte records, group-major 2-bit SEQ
// planes, quality bit-planes, CIGAR / MP blocks, tile table and span descriptors.  The expansion reads ~60 and writes ~90
// bytes per read at HBM speed; it hides under the copy of the next chunk.
#include <algorithm>
#include <vector>

#include "sd_internal.h"

int count_on_stream(sd_ctx *ctx, const sd_batch *b, const sd_params *p, cudaStream_t stream, uint8_t *read_tc, bool timed);   // sd_count.cu
int tile_spans_on(sd_ctx *ctx, cudaStream_t stream, const sd_batch *b, void *d_spans, unsigned int *d_counts);                                         // sd_count.cu

namespace {

__device__ __forceinline__ uint32_t even16(uint32_t x) {   // bits 0, 2, 4 .. 30 -> bits 0 .. 15
    x &= 0x55555555u;
    x = (x | (x >> 1)) & 0x33333333u;
    x = (x | (x >> 2)) & 0x0F0F0F0Fu;
    x = (x | (x >> 4)) & 0x00FF00FFu;
    x = (x | (x >> 8)) & 0x0000FFFFu;
    return x;
}
__device__ __forceinline__ uint32_t spread16(uint32_t x) {   // bits 0 .. 15 -> bits 0, 2, 4 .. 30
    x &= 0x0000FFFFu;
    x = (x | (x << 8)) & 0x00FF00FFu;
    x = (x | (x << 4)) & 0x0F0F0F0Fu;
    x = (x | (x << 2)) & 0x33333333u;
    x = (x | (x << 1)) & 0x55555555u;
    return x;
}
__device__ __forceinline__ uint32_t nibble_bit8(uint32_t x, int b) {   // bit b of each of the 8 nibbles -> bits 0 .. 7
    x = (x >> b) & 0x11111111u;
    x = (x | (x >> 3)) & 0x03030303u;
    x = (x | (x >> 6)) & 0x000F000Fu;
    x = (x | (x >> 12)) & 0xFFu;
    return x;
}

// `nbits` bits of a bit column starting at bit `bit0` (word loads; the column is padded by 16 bytes)
__device__ __forceinline__ void load_bits64(const uint32_t *col, unsigned long long bit0, uint32_t &lo, uint32_t &hi) {
    const uint32_t *w = col + (bit0 >> 5);
    const uint32_t sh = (uint32_t)bit0 & 31u;
    const uint32_t a = __ldg(w), b = __ldg(w + 1), c = __ldg(w + 2);
    lo = __funnelshift_r(a, b, sh);
    hi = __funnelshift_r(b, c, sh);
}

struct ExpandArgs {
    const sd_wire_tile *wt;   // [nt + 1], offsets as in the whole wire batch; the column pointers below point at the chunk's first element
    const uint16_t *pos16, *mqf, *xi16, *nm16, *nmp16;
    const uint32_t *xi32, *xi_dict;
    const uint32_t *seq, *qual;
    const uint32_t *var, *mp;
    uint32_t nt, qual_bits;
    uint4 *sizes;             // [nt] device-batch bytes of each tile {seq, qual, cigar, mp}
    sd_read_rec *rec;
    uint32_t *dseq;
    uint8_t *dqual;
    uint32_t *dcig, *dmp;
    sd_tile *tiles;           // [nt + 1]
};

// Per-lane view of one wire tile: the record fields and where the variable-length parts sit.
struct WireRow {
    int32_t pos;
    uint32_t ref, mapq, flags, lseq, ncig;
    uint32_t ops;   // word offset of the tile's CIGAR ops in var (SD_WT_SHAPE), relative to the chunk
};

__device__ __forceinline__ WireRow wire_row(const ExpandArgs &A, uint32_t t, uint32_t lane) {
    const sd_wire_tile h = A.wt[t];
    WireRow r;
    uint32_t v = h.var_off - A.wt[0].var_off;
    const uint32_t m = A.mqf[(size_t)t * 32u + lane];
    r.mapq = m & 0xFFu;
    r.flags = m >> 8;
    const bool pad = (r.flags & SD_F_PAD) != 0;
    if (h.flags & SD_WT_POS32) { r.pos = (int32_t)A.var[v + lane]; v += 32u; }
    else r.pos = h.base_pos + (int32_t)A.pos16[(size_t)t * 32u + lane];
    if (h.flags & SD_WT_REFS) { r.ref = reinterpret_cast<const uint16_t *>(A.var + v)[lane]; v += 16u; }
    else r.ref = h.ref;
    if (h.flags & SD_WT_SHAPE) {
        const uint32_t sh = A.var[v + lane];
        r.lseq = sh & 0xFFFFu;
        r.ncig = sh >> 16;
        v += 32u;
    } else {
        r.lseq = pad ? 0u : h.lseq;
        r.ncig = pad ? 0u : 1u;
    }
    r.ops = v;
    if (pad) { r.pos = 0; r.ref = SD_REF_NONE; r.lseq = 0u; r.ncig = 0u; }
    return r;
}

__global__ void wire_sizes_kernel(const ExpandArgs A) {
    const uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (t >= A.nt) return;
    const WireRow r = wire_row(A, t, lane);
    const uint32_t maxl = __reduce_max_sync(0xFFFFFFFFu, r.lseq), sum_l = __reduce_add_sync(0xFFFFFFFFu, r.lseq);
    const uint32_t sum_c = (A.wt[t].flags & SD_WT_SHAPE) ? __reduce_add_sync(0xFFFFFFFFu, r.ncig) : 0u;
    if (lane == 0) {
        const uint32_t G = (maxl + 31u) >> 5;
        uint4 s;
        s.x = 8u * 32u * G;
        s.y = A.qual_bits == 8 ? ((sum_l + 15u) & ~15u) : 4u * A.qual_bits * 32u * G;
        s.z = (4u * sum_c + 15u) & ~15u;
        s.w = 4u * (A.wt[t + 1].mp_off - A.wt[t].mp_off);
        A.sizes[t] = s;
    }
}

// exclusive scan of the tile sizes into the device tile table; one block
__global__ void wire_scan_kernel(const uint4 *sizes, uint32_t nt, sd_tile *tiles) {
    __shared__ unsigned long long part[1024][4];
    const uint32_t tid = threadIdx.x, per = (nt + 1023u) / 1024u;
    const uint32_t b = min(nt, tid * per), e = min(nt, b + per);
    unsigned long long s[4] = {0, 0, 0, 0};
    for (uint32_t i = b; i < e; i++) {
        const uint4 z = sizes[i];
        s[0] += z.x; s[1] += z.y; s[2] += z.z; s[3] += z.w;
    }
    for (int k = 0; k < 4; k++) part[tid][k] = s[k];
    __syncthreads();
    if (tid == 0) {
        unsigned long long run[4] = {0, 0, 0, 0};
        for (int i = 0; i < 1024; i++)
            for (int k = 0; k < 4; k++) {
                const unsigned long long v = part[i][k];
                part[i][k] = run[k];
                run[k] += v;
            }
        tiles[nt].seq_off = run[0]; tiles[nt].qual_off = run[1]; tiles[nt].cig_off = run[2]; tiles[nt].mp_off = run[3];
    }
    __syncthreads();
    for (int k = 0; k < 4; k++) s[k] = part[tid][k];
    for (uint32_t i = b; i < e; i++) {
        const uint4 z = sizes[i];
        tiles[i].seq_off = s[0]; tiles[i].qual_off = s[1]; tiles[i].cig_off = s[2]; tiles[i].mp_off = s[3];
        s[0] += z.x; s[1] += z.y; s[2] += z.z; s[3] += z.w;
    }
}

__global__ void wire_expand_kernel(const ExpandArgs A) {
    const uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (t >= A.nt) return;
    const WireRow r = wire_row(A, t, lane);
    const sd_wire_tile h = A.wt[t];
    const sd_tile dt = A.tiles[t], dn = A.tiles[t + 1];
    const size_t slot = (size_t)t * 32u + lane;
    const bool pad = (r.flags & SD_F_PAD) != 0;
    // record
    uint32_t xi = 0u, nm = 0u, nmp = 0u;
    if (!pad) {
        if (A.xi16) xi = __ldg(&A.xi_dict[A.xi16[slot]]);
        else if (A.xi32) xi = A.xi32[slot];
        if (A.nm16) nm = A.nm16[slot];
        if (A.nmp16) nmp = A.nmp16[slot];
    }
    sd_read_rec o;
    o.pos = r.pos;
    o.ref_mapq_flag = (r.ref & 0xFFFFu) | (r.mapq << 16) | (r.flags << 24);
    o.xi = __uint_as_float(xi);
    o.lseq_ncig = r.lseq | (r.ncig << 16);
    o.nm_nmp = nm | (nmp << 16);
    A.rec[slot] = o;
    // row offsets (bases) inside the tile
    uint32_t xs = r.lseq, cs = r.ncig;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t a = __shfl_up_sync(0xFFFFFFFFu, xs, d), c = __shfl_up_sync(0xFFFFFFFFu, cs, d);
        if (lane >= (uint32_t)d) { xs += a; cs += c; }
    }
    const unsigned long long row0 = (h.base_off - A.wt[0].base_off) + (xs - r.lseq);   // first base of this row in the chunk's columns
    const uint32_t G = (uint32_t)((dn.seq_off - dt.seq_off) / (8u * 32u));
    uint2 *ds = reinterpret_cast<uint2 *>(A.dseq + dt.seq_off / 4u);
    for (uint32_t g = 0; g < G; g++) {
        uint2 pl = make_uint2(0u, 0u);
        if (32u * g < r.lseq) {
            uint32_t lo, hi;
            load_bits64(A.seq, 2ull * (row0 + 32u * g), lo, hi);
            const uint32_t nb = min(32u, r.lseq - 32u * g), m = nb == 32u ? 0xFFFFFFFFu : ((1u << nb) - 1u);
            pl.x = (even16(lo) | (even16(hi) << 16)) & m;
            pl.y = (even16(lo >> 1) | (even16(hi >> 1) << 16)) & m;
        }
        ds[g * 32u + lane] = pl;
    }
    if (A.qual_bits == 2) {
        uint2 *dq = reinterpret_cast<uint2 *>(A.dqual + dt.qual_off);
        for (uint32_t g = 0; g < G; g++) {
            uint2 pl = make_uint2(0u, 0u);
            if (32u * g < r.lseq) {
                uint32_t lo, hi;
                load_bits64(A.qual, 2ull * (row0 + 32u * g), lo, hi);
                const uint32_t nb = min(32u, r.lseq - 32u * g), m = nb == 32u ? 0xFFFFFFFFu : ((1u << nb) - 1u);
                pl.x = (even16(lo) | (even16(hi) << 16)) & m;
                pl.y = (even16(lo >> 1) | (even16(hi >> 1) << 16)) & m;
            }
            dq[g * 32u + lane] = pl;
        }
    } else if (A.qual_bits == 4) {
        uint4 *dq = reinterpret_cast<uint4 *>(A.dqual + dt.qual_off);
        for (uint32_t g = 0; g < G; g++) {
            uint32_t p[4] = {0u, 0u, 0u, 0u};
            if (32u * g < r.lseq) {
                uint32_t w[4];
                load_bits64(A.qual, 4ull * (row0 + 32u * g), w[0], w[1]);
                load_bits64(A.qual, 4ull * (row0 + 32u * g) + 64ull, w[2], w[3]);
                const uint32_t nb = min(32u, r.lseq - 32u * g), m = nb == 32u ? 0xFFFFFFFFu : ((1u << nb) - 1u);
#pragma unroll
                for (int b = 0; b < 4; b++)
                    p[b] = (nibble_bit8(w[0], b) | (nibble_bit8(w[1], b) << 8) | (nibble_bit8(w[2], b) << 16) | (nibble_bit8(w[3], b) << 24)) & m;
            }
            dq[g * 32u + lane] = make_uint4(p[0], p[1], p[2], p[3]);
        }
    } else {   // bytes: rows back to back, the tile padded to 16 bytes
        const uint8_t *sq = reinterpret_cast<const uint8_t *>(A.qual) + row0;
        uint8_t *dq = A.dqual + dt.qual_off + (xs - r.lseq);
        for (uint32_t i = 0; i < r.lseq; i++) dq[i] = sq[i];
        const uint32_t tot = __shfl_sync(0xFFFFFFFFu, xs, 31);
        for (uint32_t i = tot + lane; i < (uint32_t)(dn.qual_off - dt.qual_off); i += 32u) A.dqual[dt.qual_off + i] = 0;
    }
    // CIGAR ops and MP entries: whole blocks
    if (h.flags & SD_WT_SHAPE) {
        const uint32_t tot = __shfl_sync(0xFFFFFFFFu, cs, 31), words = (uint32_t)((dn.cig_off - dt.cig_off) / 4u);
        uint32_t *dc = A.dcig + dt.cig_off / 4u;
        for (uint32_t i = lane; i < words; i += 32u) dc[i] = i < tot ? A.var[r.ops + i] : 0u;
    }
    if (A.dmp) {
        const uint32_t words = h.mp_off == A.wt[t + 1].mp_off ? 0u : A.wt[t + 1].mp_off - h.mp_off;
        const uint32_t *sm = A.mp + (h.mp_off - A.wt[0].mp_off);
        uint32_t *dm = A.dmp + dt.mp_off / 4u;
        for (uint32_t i = lane; i < words; i += 32u) dm[i] = sm[i];
    }
}

// ------------------------------------------------------------------------------------------ device batch -> wire
struct PackArgs {
    const sd_read_rec *rec;
    const uint32_t *seq;
    const uint8_t *qual;
    const uint32_t *cigar, *mp;
    const sd_tile *tiles;
    uint32_t nt, qual_bits;
    uint8_t code[256];
    sd_wire_tile *wt;
    uint16_t *pos16, *mqf, *xi16, *nm16, *nmp16;
    uint32_t *xi32;
    const uint16_t *xi_code;
    uint32_t *oseq, *oqual, *var, *omp;
    unsigned int *max_cig;
};

struct TileShape {
    uint32_t flags, lseq_u, ref_u, var_words, bases, cig_words;
    int32_t base_pos;
};

__device__ __forceinline__ TileShape tile_shape(const PackArgs &A, uint32_t t, uint32_t lane, const sd_read_rec &r, uint32_t &cig_row) {
    const uint32_t fl = r.ref_mapq_flag >> 24, ref = r.ref_mapq_flag & 0xFFFFu;
    const uint32_t lseq = r.lseq_ncig & 0xFFFFu, ncig = r.lseq_ncig >> 16;
    const bool pad = (fl & SD_F_PAD) != 0;
    uint32_t cs = ncig;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t c = __shfl_up_sync(0xFFFFFFFFu, cs, d);
        if (lane >= (uint32_t)d) cs += c;
    }
    cig_row = cs - ncig;
    const uint32_t sum_c = __shfl_sync(0xFFFFFFFFu, cs, 31);
    const unsigned real = __ballot_sync(0xFFFFFFFFu, !pad);
    const int first = real ? __ffs((int)real) - 1 : 0;
    TileShape s;
    s.lseq_u = real ? __shfl_sync(0xFFFFFFFFu, lseq, first) : 0u;
    s.ref_u = real ? __shfl_sync(0xFFFFFFFFu, ref, first) : (uint32_t)SD_REF_NONE;
    uint32_t c0 = 0u;
    const bool has_block = A.tiles[t + 1].cig_off != A.tiles[t].cig_off;
    if (!pad && ncig == 1u && has_block) c0 = A.cigar[A.tiles[t].cig_off / 4u + cig_row];
    const bool plain = pad || (lseq == s.lseq_u && ncig == 1u && (!has_block || c0 == (lseq << 4)));
    const int32_t pmin = (int32_t)__reduce_min_sync(0xFFFFFFFFu, pad ? 0x7FFFFFFF : r.pos);
    const int32_t pmax = (int32_t)__reduce_max_sync(0xFFFFFFFFu, pad ? (int)0x80000000 : r.pos);
    s.base_pos = real ? pmin : 0;
    s.flags = 0u;
    if (!__all_sync(0xFFFFFFFFu, plain)) s.flags |= SD_WT_SHAPE;
    if (!__all_sync(0xFFFFFFFFu, pad || ref == s.ref_u)) s.flags |= SD_WT_REFS;
    if (real && (long long)pmax - (long long)pmin > 65535ll) s.flags |= SD_WT_POS32;
    s.cig_words = (s.flags & SD_WT_SHAPE) ? sum_c : 0u;
    s.var_words = ((s.flags & SD_WT_POS32) ? 32u : 0u) + ((s.flags & SD_WT_REFS) ? 16u : 0u) + ((s.flags & SD_WT_SHAPE) ? 32u + sum_c : 0u);
    s.var_words = (s.var_words + 3u) & ~3u;
    s.bases = (__reduce_add_sync(0xFFFFFFFFu, pad ? 0u : lseq) + 63u) & ~63u;
    return s;
}

__global__ void wire_plan_kernel(const PackArgs A) {
    const uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (t >= A.nt) return;
    const sd_read_rec r = A.rec[(size_t)t * 32u + lane];
    uint32_t cig_row;
    const TileShape s = tile_shape(A, t, lane, r, cig_row);
    if (lane == 0) {
        sd_wire_tile w;
        w.base_off = s.bases;          // sizes for now; wire_plan_scan turns them into offsets
        w.var_off = s.var_words;
        w.mp_off = (uint32_t)((A.tiles[t + 1].mp_off - A.tiles[t].mp_off) / 4u);
        w.base_pos = s.base_pos;
        w.ref = (uint16_t)s.ref_u;
        w.flags = (uint16_t)s.flags;
        w.lseq = (uint16_t)s.lseq_u;
        w.reserved0 = 0;
        w.reserved1 = 0;
        A.wt[t] = w;
        if (s.cig_words) atomicMax(A.max_cig, (4u * s.cig_words + 15u) & ~15u);
    }
}

__global__ void wire_plan_scan_kernel(sd_wire_tile *wt, uint32_t nt, int with_mp) {
    __shared__ unsigned long long part[1024][3];
    const uint32_t tid = threadIdx.x, per = (nt + 1023u) / 1024u;
    const uint32_t b = min(nt, tid * per), e = min(nt, b + per);
    unsigned long long s[3] = {0, 0, 0};
    for (uint32_t i = b; i < e; i++) { s[0] += wt[i].base_off; s[1] += wt[i].var_off; s[2] += with_mp ? wt[i].mp_off : 0u; }
    for (int k = 0; k < 3; k++) part[tid][k] = s[k];
    __syncthreads();
    if (tid == 0) {
        unsigned long long run[3] = {0, 0, 0};
        for (int i = 0; i < 1024; i++)
            for (int k = 0; k < 3; k++) {
                const unsigned long long v = part[i][k];
                part[i][k] = run[k];
                run[k] += v;
            }
        sd_wire_tile last;
        memset(&last, 0, sizeof last);
        last.base_off = run[0]; last.var_off = (uint32_t)run[1]; last.mp_off = (uint32_t)run[2];
        last.reserved1 = (run[1] >> 32 || run[2] >> 32) ? 1u : 0u;   // overflow of the 32-bit word offsets
        wt[nt] = last;
    }
    __syncthreads();
    for (int k = 0; k < 3; k++) s[k] = part[tid][k];
    for (uint32_t i = b; i < e; i++) {
        const unsigned long long nb = wt[i].base_off, nv = wt[i].var_off, nm = with_mp ? wt[i].mp_off : 0u;
        wt[i].base_off = s[0]; wt[i].var_off = (uint32_t)s[1]; wt[i].mp_off = (uint32_t)s[2];
        s[0] += nb; s[1] += nv; s[2] += nm;
    }
}

__device__ __forceinline__ void or_bits(uint32_t *col, unsigned long long bit0, uint32_t lo, uint32_t hi) {   // 64 bits into a zeroed bit column
    uint32_t *w = col + (bit0 >> 5);
    const uint32_t sh = (uint32_t)bit0 & 31u;
    const uint32_t a = lo << sh, b = __funnelshift_l(lo, hi, sh), c = sh ? (hi >> (32u - sh)) : 0u;
    if (a) atomicOr(w, a);
    if (b) atomicOr(w + 1, b);
    if (c) atomicOr(w + 2, c);
}

__global__ void wire_fill_kernel(const PackArgs A) {
    const uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (t >= A.nt) return;
    const size_t slot = (size_t)t * 32u + lane;
    const sd_read_rec r = A.rec[slot];
    const sd_wire_tile h = A.wt[t];
    const uint32_t fl = r.ref_mapq_flag >> 24, ref = r.ref_mapq_flag & 0xFFFFu, mapq = (r.ref_mapq_flag >> 16) & 0xFFu;
    const uint32_t lseq = r.lseq_ncig & 0xFFFFu, ncig = r.lseq_ncig >> 16;
    const bool pad = (fl & SD_F_PAD) != 0;
    A.pos16[slot] = (pad || (h.flags & SD_WT_POS32)) ? (uint16_t)0 : (uint16_t)(r.pos - h.base_pos);
    A.mqf[slot] = (uint16_t)(mapq | (fl << 8));
    if (A.xi16) A.xi16[slot] = A.xi_code ? A.xi_code[slot] : (uint16_t)0;
    if (A.xi32) A.xi32[slot] = __float_as_uint(r.xi);
    A.nm16[slot] = (uint16_t)(r.nm_nmp & 0xFFFFu);
    if (A.nmp16) A.nmp16[slot] = (uint16_t)(r.nm_nmp >> 16);
    uint32_t ls = pad ? 0u : lseq, cs = ncig;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t a = __shfl_up_sync(0xFFFFFFFFu, ls, d), c = __shfl_up_sync(0xFFFFFFFFu, cs, d);
        if (lane >= (uint32_t)d) { ls += a; cs += c; }
    }
    const uint32_t cig_row = cs - ncig;
    // var block
    uint32_t v = h.var_off;
    if (h.flags & SD_WT_POS32) { A.var[v + lane] = (uint32_t)r.pos; v += 32u; }
    if (h.flags & SD_WT_REFS) { reinterpret_cast<uint16_t *>(A.var + v)[lane] = (uint16_t)ref; v += 16u; }
    if (h.flags & SD_WT_SHAPE) {
        A.var[v + lane] = pad ? 0u : r.lseq_ncig;
        v += 32u;
        const bool has_block = A.tiles[t + 1].cig_off != A.tiles[t].cig_off;
        for (uint32_t i = 0; i < ncig; i++)
            A.var[v + cig_row + i] = has_block ? A.cigar[A.tiles[t].cig_off / 4u + cig_row + i] : (lseq << 4);
    }
    // per-base columns
    const unsigned long long row0 = h.base_off + (ls - (pad ? 0u : lseq));
    const uint2 *sp = reinterpret_cast<const uint2 *>(A.seq + A.tiles[t].seq_off / 4u);
    uint32_t qs = lseq;   // offset of this row's quality bytes inside the tile (pads have no bases)
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t a = __shfl_up_sync(0xFFFFFFFFu, qs, d);
        if (lane >= (uint32_t)d) qs += a;
    }
    const uint8_t *q = A.qual + A.tiles[t].qual_off + (qs - lseq);
    for (uint32_t g = 0; 32u * g < (pad ? 0u : lseq); g++) {
        const uint32_t nb = min(32u, lseq - 32u * g), m = nb == 32u ? 0xFFFFFFFFu : ((1u << nb) - 1u);
        const uint2 pl = sp[g * 32u + lane];
        const uint32_t lo = pl.x & m, hi = pl.y & m;
        or_bits(A.oseq, 2ull * (row0 + 32u * g), spread16(lo) | (spread16(hi) << 1), spread16(lo >> 16) | (spread16(hi >> 16) << 1));
        if (A.qual_bits == 2) {
            uint32_t c0 = 0u, c1 = 0u;
            for (uint32_t i = 0; i < nb; i++) {
                const uint32_t c = A.code[q[32u * g + i]];
                if (i < 16u) c0 |= c << (2u * i); else c1 |= c << (2u * (i - 16u));
            }
            or_bits(A.oqual, 2ull * (row0 + 32u * g), c0, c1);
        } else if (A.qual_bits == 4) {
            uint32_t c[4] = {0u, 0u, 0u, 0u};
            for (uint32_t i = 0; i < nb; i++) c[i >> 3] |= (uint32_t)A.code[q[32u * g + i]] << (4u * (i & 7u));
            or_bits(A.oqual, 4ull * (row0 + 32u * g), c[0], c[1]);
            or_bits(A.oqual, 4ull * (row0 + 32u * g) + 64ull, c[2], c[3]);
        } else {
            uint8_t *o = reinterpret_cast<uint8_t *>(A.oqual) + row0 + 32u * g;
            for (uint32_t i = 0; i < nb; i++) o[i] = q[32u * g + i];
        }
    }
    if (A.omp) {
        const uint32_t words = (uint32_t)((A.tiles[t + 1].mp_off - A.tiles[t].mp_off) / 4u);
        for (uint32_t i = lane; i < words; i += 32u) A.omp[h.mp_off + i] = A.mp[A.tiles[t].mp_off / 4u + i];
    }
}

int pack_args(sd_ctx *ctx, const sd_batch *b, PackArgs &P) {
    if (b->tile_reads != 32) return sd_fail(ctx, SD_ERR_ARG, "tile_reads must be 32");
    if (b->seq_bits != 2) return sd_fail(ctx, SD_ERR_ARG, "the wire form is made from the 2-bit SEQ column (sd_seq_pack)");
    if (b->qual_form != 0 || (b->qual_bits != 0 && b->qual_bits != 8)) return sd_fail(ctx, SD_ERR_ARG, "the wire form is made from byte qualities");
    memset(&P, 0, sizeof P);
    P.rec = b->rec; P.seq = b->seq; P.qual = b->qual; P.cigar = b->cigar; P.mp = b->mp; P.tiles = b->tiles; P.nt = b->n_tiles;
    return SD_OK;
}

}  // namespace

extern "C" int sd_wire_plan(sd_ctx *ctx, const sd_batch *d_batch, sd_wire_tile *d_tiles, uint64_t h_totals[4]) {
    if (!ctx || !d_batch || !d_tiles || !h_totals) return sd_fail(ctx, SD_ERR_ARG, "sd_wire_plan: null argument");
    SD_CUDA(ctx, cudaSetDevice(ctx->device));
    PackArgs P;
    int rc = pack_args(ctx, d_batch, P);
    if (rc) return rc;
    P.wt = d_tiles;
    unsigned int *d_max = nullptr;
    SD_CUDA(ctx, cudaMalloc(&d_max, sizeof(unsigned int)));
    cudaMemsetAsync(d_max, 0, sizeof(unsigned int), ctx->stream);
    P.max_cig = d_max;
    if (P.nt) {
        wire_plan_kernel<<<(unsigned)(((uint64_t)P.nt * 32u + 255u) / 256u), 256, 0, ctx->stream>>>(P);
        ctx->launches++;
    }
    wire_plan_scan_kernel<<<1, 1024, 0, ctx->stream>>>(d_tiles, P.nt, d_batch->mp != nullptr);
    ctx->launches++;
    sd_wire_tile last;
    unsigned int mc = 0;
    cudaError_t e = cudaMemcpyAsync(&last, d_tiles + P.nt, sizeof last, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&mc, d_max, sizeof mc, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_max);
    if (e != cudaSuccess) return sd_fail(ctx, SD_ERR_CUDA, "sd_wire_plan: %s", cudaGetErrorString(e));
    if (last.reserved1) return sd_fail(ctx, SD_ERR_LIMIT, "batch too large for the wire form's 32-bit word offsets");
    h_totals[0] = last.base_off; h_totals[1] = last.var_off; h_totals[2] = last.mp_off; h_totals[3] = mc;
    return SD_OK;
}

extern "C" int sd_wire_fill(sd_ctx *ctx, const sd_batch *d_batch, const uint16_t *d_xi_code, sd_wire *out) {
    if (!ctx || !d_batch || !out || !out->tiles || !out->pos16 || !out->mqf || !out->nm16 || !out->seq || !out->qual || !out->var)
        return sd_fail(ctx, SD_ERR_ARG, "sd_wire_fill: null argument");
    if (!out->xi16 && !out->xi32) return sd_fail(ctx, SD_ERR_ARG, "sd_wire_fill: one of xi16 / xi32 is needed");
    SD_CUDA(ctx, cudaSetDevice(ctx->device));
    PackArgs P;
    int rc = pack_args(ctx, d_batch, P);
    if (rc) return rc;
    if (out->qual_bits != 2 && out->qual_bits != 4 && out->qual_bits != 8) return sd_fail(ctx, SD_ERR_ARG, "qual_bits must be 2, 4 or 8");
    P.qual_bits = out->qual_bits;
    for (uint32_t c = 0; c < (1u << (P.qual_bits == 8 ? 0 : P.qual_bits)); c++) P.code[out->qual_dict[c]] = (uint8_t)c;
    P.wt = const_cast<sd_wire_tile *>(out->tiles);
    P.pos16 = const_cast<uint16_t *>(out->pos16); P.mqf = const_cast<uint16_t *>(out->mqf);
    P.xi16 = const_cast<uint16_t *>(out->xi16); P.nm16 = const_cast<uint16_t *>(out->nm16);
    P.nmp16 = d_batch->mp ? const_cast<uint16_t *>(out->nmp16) : nullptr;
    P.xi32 = out->xi16 ? nullptr : reinterpret_cast<uint32_t *>(const_cast<float *>(out->xi32));
    P.xi_code = d_xi_code;
    P.oseq = reinterpret_cast<uint32_t *>(const_cast<uint8_t *>(out->seq));
    P.oqual = reinterpret_cast<uint32_t *>(const_cast<uint8_t *>(out->qual));
    P.var = const_cast<uint32_t *>(out->var);
    P.omp = d_batch->mp ? const_cast<uint32_t *>(out->mp) : nullptr;
    if (d_batch->mp && (!out->nmp16 || !out->mp)) return sd_fail(ctx, SD_ERR_ARG, "sd_wire_fill: the batch has MP entries, out->nmp16 / out->mp are needed");
    if (P.nt) {
        wire_fill_kernel<<<(unsigned)(((uint64_t)P.nt * 32u + 255u) / 256u), 256, 0, ctx->stream>>>(P);
        ctx->launches++;
    }
    SD_CUDA(ctx, cudaGetLastError());
    SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    out->n_reads = d_batch->n_reads;
    out->n_tiles = d_batch->n_tiles;
    out->tile_reads = 32;
    out->max_lseq = d_batch->max_lseq;
    out->max_span = d_batch->max_span;
    out->max_tile_mp = d_batch->mp ? d_batch->max_tile_mp : 0;
    return SD_OK;
}

// ------------------------------------------------------------------------------------------ wire -> count
extern "C" int sd_count_wire(sd_ctx *ctx, const sd_wire *w, const sd_params *params) {
    if (!ctx || !w || !params) return sd_fail(ctx, SD_ERR_ARG, "sd_count_wire: null argument");
    SD_CUDA(ctx, cudaSetDevice(ctx->device));
    if (w->n_tiles == 0) return SD_OK;
    if (w->tile_reads != 32) return sd_fail(ctx, SD_ERR_ARG, "tile_reads must be 32");
    if (w->qual_bits != 2 && w->qual_bits != 4 && w->qual_bits != 8) return sd_fail(ctx, SD_ERR_ARG, "qual_bits must be 2, 4 or 8");
    const bool use_mp = params->mismatch_source != SD_MISMATCH_CIGAR;
    if (use_mp && (!w->mp || !w->nmp16)) return sd_fail(ctx, SD_ERR_ARG, "mismatch source needs the mp column");
    const bool use_xi = params->filter_on != 0, use_nm = params->filter_on != 0 && params->filter_max_nm > -1;
    if (use_xi && !w->xi16 && !w->xi32) return sd_fail(ctx, SD_ERR_ARG, "the filter needs the XI column");
    const uint32_t qb = w->qual_bits;
    const sd_wire_tile *T = w->tiles;
    const uint32_t Gmax = (w->max_lseq + 31u) / 32u;
    auto up256 = [](uint64_t x) { return (x + 255ull) & ~255ull; };
    struct Lay { uint64_t wt, pos16, mqf, xi, nm, nmp, seq, qual, var, mp, sizes, rec, dseq, dqual, dcig, dmp, tiles, spans, total; };
    auto layout = [&](uint32_t t0, uint32_t t1) {
        Lay L;
        const uint64_t nt = t1 - t0, nb = T[t1].base_off - T[t0].base_off;
        uint64_t o = 0;
        auto take = [&](uint64_t bytes) { const uint64_t at = o; o = up256(o + bytes); return at; };
        L.wt = take((nt + 1) * sizeof(sd_wire_tile));
        L.pos16 = take(nt * 64); L.mqf = take(nt * 64);
        L.xi = take(use_xi ? nt * (w->xi16 ? 64 : 128) : 0);
        L.nm = take(use_nm ? nt * 64 : 0);
        L.nmp = take(use_mp ? nt * 64 : 0);
        L.seq = take(nb / 4 + 16);
        L.qual = take(nb * qb / 8 + 16);
        L.var = take((uint64_t)(T[t1].var_off - T[t0].var_off) * 4 + 16);
        L.mp = take(use_mp ? (uint64_t)(T[t1].mp_off - T[t0].mp_off) * 4 + 16 : 0);
        L.sizes = take(nt * 16);
        L.rec = take(nt * 32 * sizeof(sd_read_rec));
        L.dseq = take(nt * 8 * 32 * Gmax);
        L.dqual = take(qb == 8 ? nb + nt * 16 : nt * 4 * qb * 32 * Gmax);
        L.dcig = take((uint64_t)(T[t1].var_off - T[t0].var_off) * 4 + nt * 16 + 16);
        L.dmp = take(use_mp ? (uint64_t)(T[t1].mp_off - T[t0].mp_off) * 4 + 16 : 0);
        L.tiles = take((nt + 1) * sizeof(sd_tile));
        L.spans = take(nt * sizeof(sd_tile_span));
        L.total = o;
        return L;
    };
    // chunk boundaries: about 192 MB of wire bytes per chunk
    const double wire_per_base = 0.25 + qb / 8.0;
    const uint64_t target = 192ull << 20;
    std::vector<uint32_t> cuts{0};
    {
        uint32_t t0 = 0;
        while (t0 < w->n_tiles) {
            uint32_t lo = t0 + 1, hi = w->n_tiles;
            auto wire_bytes = [&](uint32_t a, uint32_t b) {
                return (uint64_t)((T[b].base_off - T[a].base_off) * wire_per_base) + (uint64_t)(b - a) * (32 + 64 * 5) +
                       (uint64_t)(T[b].var_off - T[a].var_off) * 4 + (use_mp ? (uint64_t)(T[b].mp_off - T[a].mp_off) * 4 : 0);
            };
            if (wire_bytes(t0, hi) > target) {
                while (lo < hi) {
                    const uint32_t mid = lo + (hi - lo + 1) / 2;
                    if (wire_bytes(t0, mid) <= target) lo = mid; else hi = mid - 1;
                }
                hi = lo;
            }
            cuts.push_back(hi);
            t0 = hi;
        }
    }
    uint64_t need = 0;
    for (size_t c = 0; c + 1 < cuts.size(); c++) need = std::max(need, layout(cuts[c], cuts[c + 1]).total);
    if (need > ctx->stage_cap) {
        SD_CUDA(ctx, cudaDeviceSynchronize());
        for (int i = 0; i < 2; i++) { cudaFree(ctx->d_stage[i]); ctx->d_stage[i] = nullptr; }
        ctx->stage_cap = 0;
        for (int i = 0; i < 2; i++) SD_CUDA(ctx, cudaMalloc(&ctx->d_stage[i], need));
        ctx->stage_cap = need;
    }
    // XI dictionary: once per call
    if (use_xi && w->xi16) {
        if (!w->xi_dict || w->n_xi_dict == 0) return sd_fail(ctx, SD_ERR_ARG, "xi16 without a dictionary");
        if (ctx->xi_dict_cap < w->n_xi_dict) {
            SD_CUDA(ctx, cudaDeviceSynchronize());
            cudaFree(ctx->d_xi_dict);
            ctx->d_xi_dict = nullptr;
            ctx->xi_dict_cap = 0;
            SD_CUDA(ctx, cudaMalloc(&ctx->d_xi_dict, (size_t)std::max<uint32_t>(w->n_xi_dict, 65536u) * sizeof(float)));
            ctx->xi_dict_cap = std::max<uint32_t>(w->n_xi_dict, 65536u);
        }
        SD_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_done[0], 0));
        SD_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_done[1], 0));
        SD_CUDA(ctx, cudaMemcpyAsync(ctx->d_xi_dict, w->xi_dict, (size_t)w->n_xi_dict * sizeof(float), cudaMemcpyHostToDevice, ctx->copy_stream));
    }
    for (size_t c = 0; c + 1 < cuts.size(); c++) {
        const int buf = (int)(c & 1);
        const uint32_t t0 = cuts[c], t1 = cuts[c + 1], nt = t1 - t0;
        const Lay L = layout(t0, t1);
        if (c >= 2) SD_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_done[buf], 0));
        unsigned char *d = (unsigned char *)ctx->d_stage[buf];
        auto h2d = [&](uint64_t at, const void *src, uint64_t bytes) -> cudaError_t {
            return bytes ? cudaMemcpyAsync(d + at, src, bytes, cudaMemcpyHostToDevice, ctx->copy_stream) : cudaSuccess;
        };
        const uint64_t b0 = T[t0].base_off, nb = T[t1].base_off - b0;
        SD_CUDA(ctx, h2d(L.wt, T + t0, (uint64_t)(nt + 1) * sizeof(sd_wire_tile)));
        SD_CUDA(ctx, h2d(L.pos16, w->pos16 + (size_t)t0 * 32, (uint64_t)nt * 64));
        SD_CUDA(ctx, h2d(L.mqf, w->mqf + (size_t)t0 * 32, (uint64_t)nt * 64));
        if (use_xi && w->xi16) SD_CUDA(ctx, h2d(L.xi, w->xi16 + (size_t)t0 * 32, (uint64_t)nt * 64));
        if (use_xi && !w->xi16) SD_CUDA(ctx, h2d(L.xi, w->xi32 + (size_t)t0 * 32, (uint64_t)nt * 128));
        if (use_nm) SD_CUDA(ctx, h2d(L.nm, w->nm16 + (size_t)t0 * 32, (uint64_t)nt * 64));
        if (use_mp) SD_CUDA(ctx, h2d(L.nmp, w->nmp16 + (size_t)t0 * 32, (uint64_t)nt * 64));
        SD_CUDA(ctx, h2d(L.seq, w->seq + b0 / 4, nb / 4));
        SD_CUDA(ctx, h2d(L.qual, w->qual + b0 * qb / 8, nb * qb / 8));
        SD_CUDA(ctx, h2d(L.var, w->var + T[t0].var_off, (uint64_t)(T[t1].var_off - T[t0].var_off) * 4));
        if (use_mp) SD_CUDA(ctx, h2d(L.mp, w->mp + T[t0].mp_off, (uint64_t)(T[t1].mp_off - T[t0].mp_off) * 4));
        SD_CUDA(ctx, cudaEventRecord(ctx->ev_copy[buf], ctx->copy_stream));
        SD_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_copy[buf], 0));
        ExpandArgs E;
        memset(&E, 0, sizeof E);
        E.wt = (const sd_wire_tile *)(d + L.wt);
        E.pos16 = (const uint16_t *)(d + L.pos16); E.mqf = (const uint16_t *)(d + L.mqf);
        E.xi16 = (use_xi && w->xi16) ? (const uint16_t *)(d + L.xi) : nullptr;
        E.xi32 = (use_xi && !w->xi16) ? (const uint32_t *)(d + L.xi) : nullptr;
        E.xi_dict = (const uint32_t *)ctx->d_xi_dict;
        E.nm16 = use_nm ? (const uint16_t *)(d + L.nm) : nullptr;
        E.nmp16 = use_mp ? (const uint16_t *)(d + L.nmp) : nullptr;
        E.seq = (const uint32_t *)(d + L.seq); E.qual = (const uint32_t *)(d + L.qual);
        E.var = (const uint32_t *)(d + L.var); E.mp = use_mp ? (const uint32_t *)(d + L.mp) : nullptr;
        E.nt = nt; E.qual_bits = qb;
        E.sizes = (uint4 *)(d + L.sizes);
        E.rec = (sd_read_rec *)(d + L.rec); E.dseq = (uint32_t *)(d + L.dseq); E.dqual = d + L.dqual;
        E.dcig = (uint32_t *)(d + L.dcig); E.dmp = use_mp ? (uint32_t *)(d + L.dmp) : nullptr;
        E.tiles = (sd_tile *)(d + L.tiles);
        const unsigned blocks = (unsigned)(((uint64_t)nt * 32u + 255u) / 256u);
        wire_sizes_kernel<<<blocks, 256, 0, ctx->stream>>>(E);
        wire_scan_kernel<<<1, 1024, 0, ctx->stream>>>(E.sizes, nt, E.tiles);
        wire_expand_kernel<<<blocks, 256, 0, ctx->stream>>>(E);
        ctx->launches += 3;
        SD_CUDA(ctx, cudaGetLastError());
        sd_batch db;
        memset(&db, 0, sizeof db);
        db.n_tiles = nt;
        db.tile_reads = 32;
        db.n_reads = std::min<uint64_t>(w->n_reads, (uint64_t)t1 * 32) - (uint64_t)t0 * 32;
        db.rec = E.rec; db.seq = E.dseq; db.qual = E.dqual; db.cigar = E.dcig; db.mp = E.dmp; db.tiles = E.tiles;
        db.max_lseq = w->max_lseq;
        db.max_span = w->max_span;
        db.max_tile_seq = 8u * 32u * Gmax;
        db.max_tile_qual = qb == 8 ? ((32u * w->max_lseq + 15u) & ~15u) : 4u * qb * 32u * Gmax;
        db.max_tile_cig = w->max_tile_cig;
        db.max_tile_mp = use_mp ? w->max_tile_mp : 0u;
        db.seq_bits = 2;
        db.qual_bits = qb;
        db.qual_form = qb == 8 ? 0u : 1u;
        memcpy(db.qual_dict, w->qual_dict, 16);
        int rc;
        if (ctx->d_ment) {
            if ((rc = tile_spans_on(ctx, ctx->stream, &db, d + L.spans, nullptr))) return rc;
            db.spans = d + L.spans;
            db.spans_epoch = ctx->ann_epoch;
        }
        if ((rc = count_on_stream(ctx, &db, params, ctx->stream, nullptr, false))) return rc;
        SD_CUDA(ctx, cudaEventRecord(ctx->ev_done[buf], ctx->stream));
    }
    SD_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
    SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SD_OK;
}
