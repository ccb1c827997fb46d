// sd_bamwrite.cpp -- host-side writer of the filtered BAM of `slamdunk filter`.
//
// Takes over what pysam + `samtools sort` do for the reference after the filter decisions are made
// (dunks/filter.py:56-65 dumpBufferToBam, :191 outfile.write, :273-310 header rewrite + bamSort): the kept records are
// copied byte for byte from the (inflated) input, multimapper-retained records get secondary/supplementary cleared and the
// RD:Z tag, everything is coordinate sorted (tid, pos, strand; stable) and written as BGZF with deflate on all threads.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <zlib.h>

#include <algorithm>
#include <map>
#include <atomic>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "sd_bamio.h"
#include "slamdunk_b200.h"

namespace {

int wfail(char *err, int errlen, int code, const char *fmt, ...) {
    if (err && errlen > 0) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(err, (size_t)errlen, fmt, ap);
        va_end(ap);
    }
    return code;
}

inline uint16_t g16(const uint8_t *p) { return (uint16_t)(p[0] | (p[1] << 8)); }
inline uint32_t g32(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
inline void p16(uint8_t *p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); }
inline void p32(uint8_t *p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24); }

template <class F>
void pfor(size_t n, int threads, F f) {
    if (threads <= 1 || n < 2) {
        for (size_t i = 0; i < n; i++) f(i);
        return;
    }
    std::atomic<size_t> next(0);
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++)
        pool.emplace_back([&]() {
            for (;;) {
                const size_t i = next.fetch_add(1);
                if (i >= n) break;
                f(i);
            }
        });
    for (auto &th : pool) th.join();
}

// size of one BAM optional field starting at p (tag[2] type value), 0 on malformed data
size_t aux_field_size(const uint8_t *p, const uint8_t *end) {
    if (p + 3 > end) return 0;
    const uint8_t ty = p[2];
    size_t n = 3;
    switch (ty) {
        case 'A': case 'c': case 'C': n += 1; break;
        case 's': case 'S': n += 2; break;
        case 'i': case 'I': case 'f': n += 4; break;
        case 'Z': case 'H': {
            const uint8_t *z = (const uint8_t *)memchr(p + 3, 0, (size_t)(end - p - 3));
            if (!z) return 0;
            n = (size_t)(z - p) + 1;
            break;
        }
        case 'B': {
            if (p + 8 > end) return 0;
            const uint8_t sub = p[3];
            const size_t cnt = g32(p + 4), sz = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : 4;
            n = 8 + cnt * sz;
            break;
        }
        default: return 0;
    }
    return p + n <= end ? n : 0;
}

struct Span { const uint8_t *p; size_t n; };

}  // namespace

namespace {

// ---- BAI (SAM spec 5.2): binning index + 16 kb linear index over virtual offsets (compressed block start << 16 | offset)
inline uint32_t reg2bin(int64_t beg, int64_t end) {
    --end;
    if (beg >> 14 == end >> 14) return (uint32_t)(((1 << 15) - 1) / 7 + (beg >> 14));
    if (beg >> 17 == end >> 17) return (uint32_t)(((1 << 12) - 1) / 7 + (beg >> 17));
    if (beg >> 20 == end >> 20) return (uint32_t)(((1 << 9) - 1) / 7 + (beg >> 20));
    if (beg >> 23 == end >> 23) return (uint32_t)(((1 << 6) - 1) / 7 + (beg >> 23));
    if (beg >> 26 == end >> 26) return (uint32_t)(((1 << 3) - 1) / 7 + (beg >> 26));
    return 0;
}

struct BaiRef {
    std::map<uint32_t, std::vector<std::pair<uint64_t, uint64_t>>> bins;   // bin -> chunks (begin, end)
    uint32_t last_bin = 0;
    bool has_last = false;
    std::vector<std::pair<uint64_t, uint64_t>> *last_chunks = nullptr;
    std::vector<uint64_t> linear;
    uint64_t off_beg = 0, off_end = 0, n_mapped = 0, n_unmapped = 0;
    bool used = false;
};

void put32(std::vector<uint8_t> &o, uint32_t v) { for (int k = 0; k < 4; k++) o.push_back((uint8_t)(v >> (8 * k))); }
void put64(std::vector<uint8_t> &o, uint64_t v) { for (int k = 0; k < 8; k++) o.push_back((uint8_t)(v >> (8 * k))); }

}  // namespace

extern "C" int sd_rewrite_bam_indexed(const char *in_path, const char *out_path, const char *bai_path, const char *header_text,
                                      const uint8_t *keep, const uint8_t *state, uint64_t n_records, int sort, int level, int threads,
                                      uint64_t *n_written, char *err, int errlen) {
    if (!in_path || !out_path || !header_text || !keep) return wfail(err, errlen, SD_ERR_ARG, "sd_rewrite_bam: null argument");
    if (threads <= 0) threads = (int)std::max(1u, std::thread::hardware_concurrency());
    if (level < 1 || level > 9) level = 6;
    RawBuf raw;
    int rc = sd_bam_inflate_file(in_path, threads, raw, nullptr, err, errlen);
    if (rc != SD_OK) return rc;
    std::string old_text;
    std::vector<std::string> names;
    std::vector<uint32_t> lengths;
    size_t p = 0;
    if ((rc = sd_bam_parse_header(raw, in_path, old_text, names, lengths, &p, err, errlen)) != SD_OK) return rc;
    std::vector<size_t> recoff;
    if ((rc = sd_bam_index_records(raw, p, (int32_t)names.size(), threads, recoff, in_path, err, errlen)) != SD_OK) return rc;
    if (recoff.size() != n_records)
        return wfail(err, errlen, SD_ERR_ARG, "%s holds %zu records, %llu decisions given", in_path, recoff.size(), (unsigned long long)n_records);
    const uint8_t *R = raw.data();

    // ---- RD:Z texts of the retained multimappers (the reference's multimapList, filter.py:170 and its resets)
    std::vector<std::string> rd_text(n_records ? 1 : 0);
    std::vector<uint32_t> rd_slot;
    bool any2 = false;
    for (uint64_t i = 0; i < n_records; i++) any2 |= keep[i] == 2;
    if (any2) {
        if (!state) return wfail(err, errlen, SD_ERR_ARG, "sd_rewrite_bam: retained multimappers need the state array");
        rd_slot.assign(n_records, 0);
        rd_text.assign(1, std::string());
        std::vector<uint64_t> mlist;
        const uint8_t *prev = nullptr;
        uint32_t prev_len = 0;
        bool buffer_nonempty = false;
        long long open_retained = -1;
        auto flush = [&]() {
            if (open_retained >= 0) {
                std::string t;
                for (size_t k = 0; k < mlist.size(); k++) {
                    const uint8_t *r = R + recoff[mlist[k]];
                    const int32_t ref_id = (int32_t)g32(r), pos = (int32_t)g32(r + 4);
                    const uint32_t l_name = r[8], n_cig = g16(r + 12);
                    uint32_t reflen = 0;
                    for (uint32_t c = 0; c < n_cig; c++) {
                        const uint32_t op = g32(r + 32 + l_name + 4 * c);
                        if ((0x18Du >> (op & 15u)) & 1u) reflen += op >> 4;
                    }
                    if (reflen == 0) reflen = 1;
                    char tmp[64];
                    snprintf(tmp, sizeof tmp, ":%d-%d", pos, pos + (int32_t)reflen);
                    if (k) t += ' ';
                    t += (ref_id >= 0 && (size_t)ref_id < names.size()) ? names[(size_t)ref_id] : std::string("*");
                    t += tmp;
                }
                rd_slot[(size_t)open_retained] = (uint32_t)rd_text.size();
                rd_text.push_back(t);
            }
            open_retained = -1;
            buffer_nonempty = false;
            mlist.clear();
        };
        for (uint64_t i = 0; i < n_records; i++) {
            const uint32_t st = state[i] & 3u;
            if (st == 0) continue;
            const uint8_t *r = R + recoff[i];
            const uint8_t *name = r + 32;
            const uint32_t l_name = r[8];
            if (st == 2) {
                if (prev && !(prev_len == l_name && memcmp(prev, name, l_name) == 0)) flush();   // filter.py:115-133
                mlist.push_back(i);
            } else if (buffer_nonempty) flush();                                                 // filter.py:176-187
            prev = name;
            prev_len = l_name;
            if (st == 2 && (state[i] & 4u)) buffer_nonempty = true;
            if (keep[i] == 2) open_retained = (long long)i;
        }
        flush();                                                                                 // filter.py:197-199
    }

    // ---- kept records, in output order
    std::vector<uint32_t> order;
    order.reserve(n_records);
    for (uint64_t i = 0; i < n_records; i++)
        if (keep[i]) order.push_back((uint32_t)i);
    if (sort && order.size() > 1) {
        // samtools sort: (tid, pos, reverse strand), unmapped tid = -1 last (as unsigned), ties in input order.  The keys are
        // pulled out of the records once (a comparator that follows two pointers into gigabytes of records misses the cache
        // on every comparison); with the record index as the last key component every key is distinct, so runs sorted on
        // their own threads and merged pairwise give exactly the stable order.
        struct Key {
            uint64_t a, b;      // a = tid << 32 | pos with the sign bit flipped; b = reverse strand << 32 | record index
            bool operator<(const Key &o) const { return a != o.a ? a < o.a : b < o.b; }
        };
        const size_t m = order.size();
        std::vector<Key> keys(m), spare(m);
        const size_t n_runs = std::max<size_t>(1, std::min<size_t>((size_t)threads, m / 512));
        std::vector<size_t> cut(n_runs + 1);
        for (size_t c = 0; c <= n_runs; c++) cut[c] = m * c / n_runs;
        pfor(n_runs, threads, [&](size_t c) {
            for (size_t k = cut[c]; k < cut[c + 1]; k++) {
                const uint8_t *r = R + recoff[order[k]];
                keys[k].a = ((uint64_t)g32(r) << 32) | (uint64_t)(g32(r + 4) ^ 0x80000000u);
                keys[k].b = ((uint64_t)((g16(r + 14) >> 4) & 1u) << 32) | (uint64_t)order[k];
            }
            std::sort(keys.begin() + (ptrdiff_t)cut[c], keys.begin() + (ptrdiff_t)cut[c + 1]);
        });
        Key *src = keys.data(), *dst = spare.data();
        for (size_t width = 1; width < n_runs; width *= 2) {
            const size_t n_pairs = (n_runs + 2 * width - 1) / (2 * width);
            pfor(n_pairs, threads, [&](size_t q) {
                const size_t lo = cut[std::min(n_runs, 2 * width * q)], mid = cut[std::min(n_runs, 2 * width * q + width)],
                             hi = cut[std::min(n_runs, 2 * width * (q + 1))];
                std::merge(src + lo, src + mid, src + mid, src + hi, dst + lo);
            });
            std::swap(src, dst);
        }
        for (size_t k = 0; k < m; k++) order[k] = (uint32_t)(src[k].b & 0xFFFFFFFFu);
    }

    // ---- modified copies of the retained multimappers
    std::vector<std::vector<uint8_t>> owned;
    std::vector<Span> spans;
    spans.reserve(order.size() + 1);
    std::vector<uint8_t> head;
    {
        const size_t lt = strlen(header_text);
        head.insert(head.end(), {'B', 'A', 'M', 1});
        uint8_t b4[4];
        p32(b4, (uint32_t)lt);
        head.insert(head.end(), b4, b4 + 4);
        head.insert(head.end(), header_text, header_text + lt);
        p32(b4, (uint32_t)names.size());
        head.insert(head.end(), b4, b4 + 4);
        for (size_t k = 0; k < names.size(); k++) {
            p32(b4, (uint32_t)names[k].size() + 1);
            head.insert(head.end(), b4, b4 + 4);
            head.insert(head.end(), names[k].begin(), names[k].end());
            head.push_back(0);
            p32(b4, lengths[k]);
            head.insert(head.end(), b4, b4 + 4);
        }
    }
    spans.push_back({head.data(), head.size()});
    owned.reserve(64);
    size_t n_mod = 0;
    for (uint32_t i : order) n_mod += keep[i] == 2;
    owned.resize(n_mod);
    size_t mi = 0;
    for (uint32_t i : order) {
        const uint8_t *r = R + recoff[i];
        const uint32_t bs = g32(r - 4);
        if (keep[i] != 2) {
            spans.push_back({r - 4, (size_t)bs + 4});
            continue;
        }
        const uint32_t l_name = r[8], n_cig = g16(r + 12);
        const int32_t l_seq = (int32_t)g32(r + 16);
        const uint8_t *aux = r + 32 + l_name + 4 * (size_t)n_cig + ((size_t)l_seq + 1) / 2 + (size_t)l_seq, *end = r + bs;
        std::vector<uint8_t> &o = owned[mi++];
        o.assign(r - 4, aux);
        for (const uint8_t *a = aux; a < end;) {   // drop an existing RD tag (pysam set_tag replaces it)
            const size_t fs = aux_field_size(a, end);
            if (!fs) return wfail(err, errlen, SD_ERR_FORMAT, "%s: malformed tag in record %u", in_path, i);
            if (!(a[0] == 'R' && a[1] == 'D')) o.insert(o.end(), a, a + fs);
            a += fs;
        }
        const std::string &t = rd_text[rd_slot.empty() ? 0 : rd_slot[i]];
        o.push_back('R'); o.push_back('D'); o.push_back('Z');
        o.insert(o.end(), t.begin(), t.end());
        o.push_back(0);
        p32(o.data(), (uint32_t)o.size() - 4);
        p16(o.data() + 4 + 14, g16(o.data() + 4 + 14) & ~0x900u);   // is_secondary = is_supplementary = False
        spans.push_back({o.data(), o.size()});
    }

    // ---- BGZF: the output stream is cut into 0xFF00-byte payloads, compressed independently
    std::vector<size_t> start(spans.size() + 1, 0);
    for (size_t k = 0; k < spans.size(); k++) start[k + 1] = start[k] + spans[k].n;
    const size_t total = start.back();
    const size_t payload = 0xFF00;
    const size_t n_blocks = (total + payload - 1) / payload;
    std::vector<std::vector<uint8_t>> blocks(n_blocks);
    std::vector<uint32_t> block_size(n_blocks, 0);
    std::atomic<int> bad(0);
    // the output file takes the blocks in order while later ones are still being compressed: whoever finishes a block and
    // finds the file free writes every block that is ready (and releases its buffer); what is left goes out after the loop
    FILE *fh = fopen(out_path, "wb");
    if (!fh) return wfail(err, errlen, SD_ERR_IO, "cannot create %s", out_path);
    struct OutputGuard {            // any early return closes the file and takes the partial output away
        FILE *&f;
        const char *path;
        bool keep = false;
        ~OutputGuard() {
            if (f) fclose(f);
            if (!keep) remove(path);
        }
    } guard{fh, out_path};
    std::unique_ptr<std::atomic<uint8_t>[]> ready(new std::atomic<uint8_t>[n_blocks ? n_blocks : 1]);
    for (size_t b = 0; b < n_blocks; b++) ready[b].store(0, std::memory_order_relaxed);
    std::mutex file_mutex;
    size_t next_write = 0;          // guarded by file_mutex
    bool write_failed = false;      // guarded by file_mutex
    auto write_ready = [&]() {      // call with file_mutex held
        while (next_write < n_blocks && ready[next_write].load(std::memory_order_acquire)) {
            std::vector<uint8_t> &blk = blocks[next_write];
            if (!write_failed && fwrite(blk.data(), 1, blk.size(), fh) != blk.size()) write_failed = true;
            std::vector<uint8_t>().swap(blk);
            next_write++;
        }
    };
    const char *env_deflate = getenv("SD_DEFLATE");
    const bool use_own_deflate = !(env_deflate && strcmp(env_deflate, "zlib") == 0);
    pfor(n_blocks, threads, [&](size_t b) {
        const size_t lo = b * payload, hi = std::min(total, lo + payload), len = hi - lo;
        std::vector<uint8_t> in(len);
        size_t k = (size_t)(std::upper_bound(start.begin(), start.end(), lo) - start.begin()) - 1, w = 0, at = lo;
        while (w < len) {
            const size_t off = at - start[k], take = std::min(spans[k].n - off, len - w);
            memcpy(in.data() + w, spans[k].p + off, take);
            w += take;
            at += take;
            k++;
        }
        std::vector<uint8_t> &out = blocks[b];
        out.resize(0x10000 + 64);
        // the block compressor of sd_deflate.cpp; zlib when it declines (or with SD_DEFLATE=zlib, level 0)
        size_t clen = (use_own_deflate && level > 0) ? (size_t)sd_deflate_raw(in.data(), len, out.data() + 18, 0x10000 - 18 - 8, level) : 0;
        if (!clen) {
            z_stream zs;
            memset(&zs, 0, sizeof zs);
            if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) { bad = 1; return; }
            zs.next_in = in.data();
            zs.avail_in = (uInt)len;
            zs.next_out = out.data() + 18;
            zs.avail_out = (uInt)(0x10000 - 18 - 8);
            const int zr = deflate(&zs, Z_FINISH);
            clen = zs.total_out;
            deflateEnd(&zs);
            if (zr != Z_STREAM_END) { bad = 1; return; }
        }
        const size_t bsize = 18 + clen + 8;
        static const uint8_t hd[12] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0};
        memcpy(out.data(), hd, 12);
        out[12] = 66; out[13] = 67; out[14] = 2; out[15] = 0;
        p16(out.data() + 16, (uint32_t)bsize - 1);
        p32(out.data() + 18 + clen, sd_crc32(in.data(), len));
        p32(out.data() + 18 + clen + 4, (uint32_t)len);
        out.resize(bsize);
        block_size[b] = (uint32_t)bsize;
        ready[b].store(1, std::memory_order_release);
        while (ready[b].load(std::memory_order_relaxed) && file_mutex.try_lock()) {
            const size_t before = next_write;
            write_ready();
            const bool advanced = next_write != before;
            file_mutex.unlock();
            if (!advanced) break;       // not this block's turn yet; a block marked ready meanwhile is picked up by the next pass
        }
    });
    {
        std::lock_guard<std::mutex> lk(file_mutex);
        if (!bad) write_ready();
    }
    if (bad || write_failed || next_write != n_blocks) {
        return bad ? wfail(err, errlen, SD_ERR_LIMIT, "BGZF compression failed (block does not fit 64 KiB)")
                   : wfail(err, errlen, SD_ERR_IO, "short write on %s", out_path);
    }
    if (bai_path) {
        // compressed start of every block (one more: the EOF block) -> virtual offset of a position in the record stream
        std::vector<uint64_t> coff(n_blocks + 1, 0);
        for (size_t b = 0; b < n_blocks; b++) coff[b + 1] = coff[b] + block_size[b];
        auto voff = [&](size_t u) { return (coff[u / payload] << 16) | (uint64_t)(u % payload); };
        std::vector<BaiRef> refs(names.size());
        uint64_t n_no_coor = 0;
        int32_t last_ref = -1;
        int64_t last_pos = -1;
        for (size_t k = 0; k < order.size(); k++) {
            const Span &sp = spans[k + 1];              // spans[0] is the header
            const uint8_t *r = sp.p + 4;
            const int32_t ref_id = (int32_t)g32(r), pos = (int32_t)g32(r + 4);
            const uint32_t flag = g16(r + 14);
            if (ref_id < 0 || (size_t)ref_id >= refs.size()) { n_no_coor++; continue; }
            if (n_no_coor || ref_id < last_ref || (ref_id == last_ref && pos < last_pos))
                return wfail(err, errlen, SD_ERR_ARG, "%s: records are not coordinate sorted, cannot index", out_path);
            last_ref = ref_id;
            last_pos = pos;
            const uint32_t l_name = r[8], n_cig = g16(r + 12);
            int64_t reflen = 0;
            if (!(flag & 4))
                for (uint32_t c = 0; c < n_cig; c++) {
                    const uint32_t op = g32(r + 32 + l_name + 4 * c);
                    if ((0x18Du >> (op & 15u)) & 1u) reflen += op >> 4;
                }
            const int64_t beg = pos < 0 ? 0 : pos, end = beg + (reflen > 0 ? reflen : 1);
            const uint64_t v0 = voff(start[k + 1]), v1 = voff(start[k + 2]);
            BaiRef &R2 = refs[(size_t)ref_id];
            if (!R2.used) { R2.used = true; R2.off_beg = v0; }
            R2.off_end = v1;
            if (flag & 4) R2.n_unmapped++; else R2.n_mapped++;
            const uint32_t bin = reg2bin(beg, end);
            if (R2.has_last && R2.last_bin == bin) {
                R2.last_chunks->back().second = v1;   // consecutive records of one bin: one chunk
            } else {
                R2.last_chunks = &R2.bins[bin];       // (map nodes stay where they are)
                R2.last_chunks->push_back({v0, v1});
            }
            R2.has_last = true;
            R2.last_bin = bin;
            const size_t w0 = (size_t)(beg >> 14), w1 = (size_t)((end - 1) >> 14);
            if (R2.linear.size() < w1 + 1) R2.linear.resize(w1 + 1, UINT64_MAX);
            for (size_t w = w0; w <= w1; w++)
                if (R2.linear[w] == UINT64_MAX) R2.linear[w] = v0;
        }
        std::vector<uint8_t> bai = {'B', 'A', 'I', 1};
        put32(bai, (uint32_t)refs.size());
        for (BaiRef &R2 : refs) {
            if (!R2.used) { put32(bai, 0); put32(bai, 0); continue; }
            put32(bai, (uint32_t)R2.bins.size() + 1);
            for (auto &bn : R2.bins) {
                put32(bai, bn.first);
                put32(bai, (uint32_t)bn.second.size());
                for (auto &ch : bn.second) { put64(bai, ch.first); put64(bai, ch.second); }
            }
            put32(bai, 37450);                          // htslib's metadata pseudo-bin
            put32(bai, 2);
            put64(bai, R2.off_beg); put64(bai, R2.off_end);
            put64(bai, R2.n_mapped); put64(bai, R2.n_unmapped);
            for (size_t w = R2.linear.size(); w-- > 0;)  // windows without a record start where the next one does
                if (R2.linear[w] == UINT64_MAX) R2.linear[w] = (w + 1 < R2.linear.size()) ? R2.linear[w + 1] : 0;
            put32(bai, (uint32_t)R2.linear.size());
            for (uint64_t v : R2.linear) put64(bai, v);
        }
        put64(bai, n_no_coor);
        FILE *fb = fopen(bai_path, "wb");
        if (!fb) return wfail(err, errlen, SD_ERR_IO, "cannot create %s", bai_path);
        const bool ok = fwrite(bai.data(), 1, bai.size(), fb) == bai.size();
        if (fclose(fb) != 0 || !ok) return wfail(err, errlen, SD_ERR_IO, "short write on %s", bai_path);
    }
    static const uint8_t eof_block[28] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 66, 67, 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    const bool eof_ok = fwrite(eof_block, 1, sizeof eof_block, fh) == sizeof eof_block;
    const int close_rc = fclose(fh);
    fh = nullptr;
    if (close_rc != 0 || !eof_ok) return wfail(err, errlen, SD_ERR_IO, "short write on %s", out_path);
    guard.keep = true;
    if (n_written) *n_written = order.size();
    return SD_OK;
}

extern "C" int sd_rewrite_bam(const char *in_path, const char *out_path, const char *header_text, const uint8_t *keep,
                              const uint8_t *state, uint64_t n_records, int sort, int level, int threads, uint64_t *n_written,
                              char *err, int errlen) {
    return sd_rewrite_bam_indexed(in_path, out_path, nullptr, header_text, keep, state, n_records, sort, level, threads, n_written, err,
                                  errlen);
}
