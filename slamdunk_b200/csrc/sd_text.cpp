// sd_text.cpp -- host-side writer of the two conversion bedgraphs (reference: dunks/tcounter.py:277-285, 315-326).
//
// The values are integers from the device; the only floating-point step is rate = tc * 100.0 / cov, printed the way
// Python prints a float (repr: shortest digits that round-trip, fixed notation for 1e-4 <= |x| < 1e16, otherwise
// d.ddde[+-]XX), so the files are byte-identical to the reference's `print(..., conversionBedGraph[position])`.
#include <charconv>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "slamdunk_b200.h"

// Python's repr(float) for finite x.  buf must hold 32 bytes.  Returns the length.
extern "C" int sd_format_repr(double x, char *buf) {
    char tmp[48];
    auto r = std::to_chars(tmp, tmp + sizeof tmp, x, std::chars_format::scientific);   // shortest round-trip digits
    *r.ptr = 0;
    // tmp = [-]d[.ddd]e[+-]XX
    const char *p = tmp;
    char *o = buf;
    if (*p == '-') *o++ = *p++;
    if (!(*p >= '0' && *p <= '9')) {   // inf / nan (not produced by this path)
        strcpy(o, p);
        return (int)(o - buf + strlen(p));
    }
    char digits[24];
    int nd = 0;
    for (; *p && *p != 'e'; p++)
        if (*p != '.') digits[nd++] = *p;
    const int exp10 = atoi(p + 1);
    while (nd > 1 && digits[nd - 1] == '0') nd--;
    const int decpt = exp10 + 1;   // position of the decimal point relative to the digit string
    if (nd == 1 && digits[0] == '0') {
        memcpy(o, "0.0", 3);
        o += 3;
    } else if (decpt > 16 || decpt <= -4) {   // float_repr_style 'r': exponent form
        *o++ = digits[0];
        if (nd > 1) {
            *o++ = '.';
            memcpy(o, digits + 1, (size_t)nd - 1);
            o += nd - 1;
        }
        o += sprintf(o, "e%c%02d", exp10 < 0 ? '-' : '+', exp10 < 0 ? -exp10 : exp10);
    } else if (decpt <= 0) {
        *o++ = '0';
        *o++ = '.';
        for (int k = 0; k < -decpt; k++) *o++ = '0';
        memcpy(o, digits, (size_t)nd);
        o += nd;
    } else if (decpt >= nd) {
        memcpy(o, digits, (size_t)nd);
        o += nd;
        for (int k = nd; k < decpt; k++) *o++ = '0';
        *o++ = '.';
        *o++ = '0';
    } else {
        memcpy(o, digits, (size_t)decpt);
        o += decpt;
        *o++ = '.';
        memcpy(o, digits + decpt, (size_t)(nd - decpt));
        o += nd - decpt;
    }
    *o = 0;
    return (int)(o - buf);
}

namespace {
struct Fd {
    FILE *fh;
    explicit Fd(const char *path) : fh(fopen(path, "w")) {}
    bool put(const std::string &s) { return s.empty() || fwrite(s.data(), 1, s.size(), fh) == s.size(); }
    int close() {
        const int rc = fh ? fclose(fh) : 0;
        fh = nullptr;
        return rc;
    }
    ~Fd() {
        if (fh) fclose(fh);
    }
};

inline char *put_int(char *o, long long v) { return std::to_chars(o, o + 24, v).ptr; }
}  // namespace

// One line "chrom pos pos+1 rate" per covered T (A on '-') position of every BED row that had reads, in the
// reference's dict order: BED order, ascending position, a position keeps the place of its first insertion
// (the row that owns it = the first row in BED order that prints it; only rows that share positions -- overlapping
// same-strand UTRs -- need that table).  Rows are cut into chunks, `threads` workers (0 = all cores) format them and every
// chunk is appended to the files in its turn, so the text held in memory is two buffers per worker.
extern "C" int sd_write_bedgraphs(const char *path_plus, const char *path_minus, uint32_t n_utr, const char *const *row_chrom,
                                  const int64_t *row_first_pos, const uint8_t *row_strand, const uint8_t *row_has_reads,
                                  const uint64_t *row_pos_off, const uint32_t *row_pos_len, const uint64_t *row_gstart,
                                  const uint32_t *ref_lo, const uint32_t *ref_hi, const uint32_t *ref_nmask, const int32_t *pos_cov,
                                  const int32_t *pos_tc, uint64_t n_positions, int threads) {
    if (!path_plus || !path_minus || (n_utr && (!ref_lo || !ref_hi || !ref_nmask))) return SD_ERR_ARG;
    Fd out[2] = {Fd(path_plus), Fd(path_minus)};
    if (!out[0].fh || !out[1].fh) return SD_ERR_IO;
    if (threads <= 0) threads = (int)std::max(1u, std::thread::hardware_concurrency());

    auto active = [&](uint32_t i) { return row_has_reads[i] && row_pos_len[i] != 0 && row_pos_off[i] != UINT64_MAX; };
    // rows that share positions with another row: sort the active rows by offset and look at neighbours
    std::vector<uint32_t> rows;
    for (uint32_t i = 0; i < n_utr; i++) {
        if (!active(i)) continue;
        if (strlen(row_chrom[i]) > 400) return SD_ERR_LIMIT;
        if (row_pos_off[i] + row_pos_len[i] > n_positions) return SD_ERR_ARG;
        rows.push_back(i);
    }
    std::vector<uint8_t> shares(n_utr, 0);
    {
        std::vector<uint32_t> by_off(rows);
        std::sort(by_off.begin(), by_off.end(), [&](uint32_t x, uint32_t y) { return row_pos_off[x] < row_pos_off[y]; });
        uint64_t reach = 0;     // furthest end among the rows before by_off[k], and the row that has it
        uint32_t reach_row = 0;
        for (size_t k = 0; k < by_off.size(); k++) {
            const uint32_t i = by_off[k];
            const uint64_t end = row_pos_off[i] + row_pos_len[i];
            if (k && row_pos_off[i] < reach) shares[i] = shares[reach_row] = 1;
            if (!k || end > reach) {
                reach = end;
                reach_row = i;
            }
        }
    }
    bool any_shared = false;
    for (uint32_t i : rows) any_shared |= shares[i] != 0;

    // the positions of row i that get a line, ascending (tcounter.py:277-285: covered, and the upper-cased FASTA base is T on
    // '+' rows / A on '-' rows): walked 32 positions at a time over the "is that base" bits of the reference planes, so that
    // the coverage array is only looked at under a T / A
    auto for_each_printed = [&](uint32_t i, auto &&f) {
        const uint64_t g0 = row_gstart[i], g_end = g0 + row_pos_len[i], off = row_pos_off[i];
        const bool minus = row_strand[i] != 0;
        for (uint64_t g = g0; g < g_end;) {
            const uint64_t w = g >> 5, word_end = (w + 1) << 5;
            uint32_t bits = (minus ? ~ref_lo[w] & ~ref_hi[w] : ref_lo[w] & ref_hi[w]) & ~ref_nmask[w];
            bits &= 0xFFFFFFFFu << (g & 31u);
            if (word_end > g_end) bits &= 0xFFFFFFFFu >> (word_end - g_end);
            while (bits) {
                const uint32_t p = (uint32_t)((w << 5) + (uint64_t)__builtin_ctz(bits) - g0);
                bits &= bits - 1u;
                if (pos_cov[off + p] > 0) f(p);
            }
            g = word_end;
        }
    };
    // owner[idx] = first row in BED order that prints position idx (only filled under rows that share positions)
    std::vector<uint32_t> owner;
    if (any_shared) {
        owner.assign((size_t)n_positions, UINT32_MAX);
        for (uint32_t i : rows) {
            if (!shares[i]) continue;
            for_each_printed(i, [&](uint32_t p) {
                if (owner[row_pos_off[i] + p] == UINT32_MAX) owner[row_pos_off[i] + p] = i;
            });
        }
    }

    // chunks of consecutive rows holding ~128k positions each, claimed in order; a worker formats its chunk into its own
    // two buffers (reused from chunk to chunk) and appends them to the files when every earlier chunk has been written
    std::vector<uint32_t> cut(1, 0);
    {
        uint64_t acc = 0;
        for (uint32_t k = 0; k < rows.size(); k++) {
            acc += row_pos_len[rows[k]];
            if (acc >= (128u << 10) || k + 1 == rows.size()) {
                cut.push_back(k + 1);
                acc = 0;
            }
        }
    }
    const size_t n_chunks = cut.size() - 1;
    std::atomic<size_t> next(0);
    std::mutex turn_mutex;              // guards `written` (chunks appended to the files so far) and `failed`
    std::condition_variable turn_cv;
    size_t written = 0;
    bool failed = false;
    auto work = [&]() {
        std::string text[2];
        char line[512], rate[40];
        for (size_t c; (c = next.fetch_add(1, std::memory_order_relaxed)) < n_chunks;) {
            text[0].clear();
            text[1].clear();
            for (uint32_t k = cut[c]; k < cut[c + 1]; k++) {
                const uint32_t i = rows[k];
                std::string &dst = text[row_strand[i] ? 1 : 0];
                const size_t name_len = strlen(row_chrom[i]);
                memcpy(line, row_chrom[i], name_len);
                line[name_len] = ' ';
                const bool shared = shares[i] != 0;
                for_each_printed(i, [&](uint32_t p) {
                    const uint64_t idx = row_pos_off[i] + p;
                    if (shared && owner[idx] != i) return;
                    const int32_t cov = pos_cov[idx], tc = pos_tc[idx];
                    int rl;
                    if (tc == 0) { memcpy(rate, "0.0", 4); rl = 3; }
                    else rl = sd_format_repr((double)tc * 100.0 / (double)cov, rate);   // tcounter.py:252
                    char *o = put_int(line + name_len + 1, row_first_pos[i] + p);
                    *o++ = ' ';
                    o = put_int(o, row_first_pos[i] + p + 1);
                    *o++ = ' ';
                    memcpy(o, rate, (size_t)rl);
                    o += rl;
                    *o++ = '\n';
                    dst.append(line, (size_t)(o - line));
                });
            }
            std::unique_lock<std::mutex> lk(turn_mutex);
            turn_cv.wait(lk, [&] { return written == c; });
            if (!out[0].put(text[0]) || !out[1].put(text[1])) failed = true;
            written = c + 1;
            lk.unlock();
            turn_cv.notify_all();
        }
    };
    std::vector<std::thread> pool;
    const size_t n_workers = std::min((size_t)threads, (n_chunks + 1) / 2);
    for (size_t t = 1; t < n_workers; t++) pool.emplace_back(work);
    work();
    for (auto &t : pool) t.join();
    const bool closed = (out[0].close() == 0) & (out[1].close() == 0);
    return (closed && !failed) ? SD_OK : SD_ERR_IO;
}


// The per-UTR rows of the tcount table (SlamSeqInterval.__repr__, slamseq/SlamSeqFile.py:99-100, filled in by
// tcounter.py:251-307): 16 tab-separated columns, every value printed as Python's str() prints it -- the two rates are
// the int 0 or a float (repr), the last two columns always -1.0.  counters = [n][5] in SD_C_* order.  The text comes back
// in a malloc'ed buffer (*text, *text_len) that the caller releases with sd_free_text.
extern "C" int sd_format_tcount_rows(uint32_t n_utr, const char *const *row_chrom, const int64_t *row_start, const int64_t *row_stop,
                                     const char *const *row_name, const uint8_t *row_strand, const int64_t *counters,
                                     const int64_t *tcontent, int64_t read_number, char **text, uint64_t *text_len) {
    if (!text || !text_len || (n_utr && (!row_chrom || !row_start || !row_stop || !row_name || !row_strand || !counters || !tcontent)))
        return SD_ERR_ARG;
    std::string out;
    out.reserve((size_t)n_utr * 96 + 16);
    char num[40];
    auto put_i = [&](long long v) { out.append(num, (size_t)(put_int(num, v) - num)); };
    for (uint32_t i = 0; i < n_utr; i++) {
        const int64_t *c = counters + (size_t)i * SD_NCOUNTER;
        const int64_t reads = c[SD_C_READS];
        out.append(row_chrom[i]);
        out.push_back('\t');
        put_i(row_start[i]);
        out.push_back('\t');
        put_i(row_stop[i]);
        out.push_back('\t');
        out.append(row_name[i]);
        out.push_back('\t');
        put_i(row_stop[i] - row_start[i]);
        out.push_back('\t');
        out.push_back(row_strand[i] ? '-' : '+');
        out.push_back('\t');
        if (reads > 0) {                                    // tcounter.py:251
            const int64_t cov = c[SD_C_COV_ON_T], conv = c[SD_C_CONV_ON_T];
            if (cov > 0) out.append(num, (size_t)sd_format_repr((double)conv / (double)cov, num));     // tcounter.py:301-303
            else out.push_back('0');
            out.push_back('\t');
            if (read_number > 0) out.append(num, (size_t)sd_format_repr((double)reads * 1000000.0 / (double)read_number, num));   // :295-297
            else out.push_back('0');
            out.push_back('\t');
            put_i(tcontent[i]);
            out.push_back('\t');
            put_i(cov);
            out.push_back('\t');
            put_i(conv);
            out.push_back('\t');
            put_i(reads);
            out.push_back('\t');
            put_i(c[SD_C_TCREADS]);
            out.push_back('\t');
            put_i(c[SD_C_MULTIMAP]);
        } else {                                            // the default row: ints, Tcontent kept (tcounter.py:167, 307)
            out.append("0\t0\t");
            put_i(tcontent[i]);
            out.append("\t0\t0\t0\t0\t0");
        }
        out.append("\t-1.0\t-1.0\n");
    }
    char *buf = (char *)malloc(out.size() + 1);
    if (!buf) return SD_ERR_NOMEM;
    memcpy(buf, out.data(), out.size());
    buf[out.size()] = 0;
    *text = buf;
    *text_len = out.size();
    return SD_OK;
}

extern "C" void sd_free_text(char *text) { free(text); }

// Run-length lines of one genome-wide track of one chromosome (tcounter.py:417-476): a line is printed whenever the value
// changes, holding the run that just ended; the first run is the implicit "0 from position 0".
extern "C" int sd_write_track(const char *path, const char *chrom, int kind, uint64_t n_runs, const uint32_t *h_pos,
                              const double *h_val, const uint8_t *h_isfloat) {
    if (!path || !chrom || (n_runs && (!h_pos || !h_val))) return SD_ERR_ARG;
    if (kind == SD_TRACK_RATE && n_runs && !h_isfloat) return SD_ERR_ARG;
    FILE *fh = fopen(path, "ab");
    if (!fh) return SD_ERR_IO;
    std::string out;
    out.reserve(1 << 20);
    const size_t lc = strlen(chrom);
    uint32_t prev_pos = 0;
    double prev_val = 0.0;
    bool prev_float = false;
    char num[64];
    for (uint64_t j = 0; j < n_runs; j++) {
        out.append(chrom, lc);
        out.push_back('\t');
        out.append(num, (size_t)(std::to_chars(num, num + sizeof num, prev_pos + 1u).ptr - num));
        out.push_back('\t');
        out.append(num, (size_t)(std::to_chars(num, num + sizeof num, h_pos[j] + 1u).ptr - num));
        out.push_back('\t');
        if (kind == SD_TRACK_RATE && prev_float) {
            sd_format_repr(prev_val, num);
            out.append(num);
        } else {
            out.append(num, (size_t)(std::to_chars(num, num + sizeof num, (long long)prev_val).ptr - num));
        }
        out.push_back('\n');
        prev_pos = h_pos[j];
        prev_val = h_val[j];
        prev_float = kind == SD_TRACK_RATE && h_isfloat[j];
        if (out.size() > (1u << 20) - 256) {
            if (fwrite(out.data(), 1, out.size(), fh) != out.size()) { fclose(fh); return SD_ERR_IO; }
            out.clear();
        }
    }
    const bool ok = fwrite(out.data(), 1, out.size(), fh) == out.size();
    return (fclose(fh) == 0 && ok) ? SD_OK : SD_ERR_IO;
}
