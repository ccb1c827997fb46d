// sd_text.cpp -- host-side writer of the two conversion bedgraphs (reference: dunks/tcounter.py:277-285, 315-326).
//
// The values are integers from the device; the only floating-point step is rate = tc * 100.0 / cov, printed the way
// Python prints a float (repr: shortest digits that round-trip, fixed notation for 1e-4 <= |x| < 1e16, otherwise
// d.ddde[+-]XX), so the files are byte-identical to the reference's `print(..., conversionBedGraph[position])`.
#include <charconv>
#include <stdio.h>
#include <string.h>

#include <string>
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
struct Out {
    FILE *fh;
    std::vector<char> buf;
    explicit Out(const char *path) : fh(fopen(path, "w")) { buf.reserve(1 << 20); }
    void flush() {
        if (fh && !buf.empty()) fwrite(buf.data(), 1, buf.size(), fh);
        buf.clear();
    }
    void put(const char *s, size_t n) {
        buf.insert(buf.end(), s, s + n);
        if (buf.size() > (1 << 20) - 256) flush();
    }
    ~Out() {
        flush();
        if (fh) fclose(fh);
    }
};
int put_int(char *o, long long v) { return sprintf(o, "%lld", v); }
}  // namespace

// One line "chrom pos pos+1 rate" per covered T (A on '-') position of every BED row that had reads, in the
// reference's dict order: BED order, ascending position, a position keeps the place of its first insertion.
extern "C" int sd_write_bedgraphs(const char *path_plus, const char *path_minus, uint32_t n_utr, const char *const *row_chrom,
                                  const int64_t *row_first_pos, const uint8_t *row_strand, const uint8_t *row_has_reads,
                                  const uint64_t *row_pos_off, const uint32_t *row_pos_len, const uint64_t *row_gstart,
                                  const uint32_t *isx_plus, const uint32_t *isx_minus, const int32_t *pos_cov,
                                  const int32_t *pos_tc, uint64_t n_positions) {
    if (!path_plus || !path_minus) return SD_ERR_ARG;
    Out out[2] = {Out(path_plus), Out(path_minus)};
    if (!out[0].fh || !out[1].fh) return SD_ERR_IO;
    std::vector<uint8_t> emitted((size_t)n_positions, 0);
    char line[512], rate[40];
    for (uint32_t i = 0; i < n_utr; i++) {
        if (!row_has_reads[i] || row_pos_len[i] == 0 || row_pos_off[i] == UINT64_MAX) continue;
        const int s = row_strand[i] ? 1 : 0;
        const uint32_t *isx = s ? isx_minus : isx_plus;
        const size_t name_len = strlen(row_chrom[i]);
        if (name_len > 400) return SD_ERR_LIMIT;
        for (uint32_t p = 0; p < row_pos_len[i]; p++) {
            const uint64_t idx = row_pos_off[i] + p;
            const int32_t cov = pos_cov[idx];
            if (cov <= 0 || emitted[idx]) continue;
            const uint64_t g = row_gstart[i] + p;
            if (!((isx[g >> 5] >> (g & 31)) & 1u)) continue;
            emitted[idx] = 1;
            const int32_t tc = pos_tc[idx];
            int rl;
            if (tc == 0) { memcpy(rate, "0.0", 4); rl = 3; }
            else rl = sd_format_repr((double)tc * 100.0 / (double)cov, rate);   // tcounter.py:252
            char *o = line;
            memcpy(o, row_chrom[i], name_len);
            o += name_len;
            *o++ = ' ';
            o += put_int(o, row_first_pos[i] + p);
            *o++ = ' ';
            o += put_int(o, row_first_pos[i] + p + 1);
            *o++ = ' ';
            memcpy(o, rate, (size_t)rl);
            o += rl;
            *o++ = '\n';
            out[s].put(line, (size_t)(o - line));
        }
    }
    return SD_OK;
}


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
