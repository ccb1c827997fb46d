// sd_deflate.cpp -- raw DEFLATE (RFC 1951) compressor for BGZF payloads: one block of at most 65 535 bytes in, one
// complete DEFLATE stream out (a single dynamic-Huffman block, or a stored block when that is smaller).
//
// The BAM writer (sd_bamwrite.cpp) compresses every 0xFF00-byte payload on its own, so there is no window to carry over
// and no streaming state: positions fit 16 bits, the hash chains live in two small tables on the stack, and the Huffman
// codes are built once per payload from the exact symbol counts.  Match search: 4-byte hash, chains of bounded depth,
// one step of lazy evaluation at the higher levels (take the match at i + 1 when it is longer).  BAM records repeat
// their tag names, flags and quality runs at short distances, which is what the 4-byte hash finds first.
#include <stdint.h>
#include <string.h>

#include <algorithm>

#include "slamdunk_b200.h"

namespace {

constexpr int kHashBits = 15;
constexpr uint16_t kNil = 0xFFFF;
constexpr int kMinMatch = 4, kMaxMatch = 258, kMaxDist = 32768;

const uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
const uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
const uint8_t kPreOrder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

struct SymTables {
    uint8_t len_sym[259];      // match length -> length symbol - 257
    uint8_t dist_sym_lo[513];  // distance 1..512 -> distance symbol
    SymTables() {
        for (int s = 0; s < 29; s++)
            for (int l = kLenBase[s]; l <= (s == 28 ? 258 : kLenBase[s] + (1 << kLenExtra[s]) - 1) && l <= 258; l++) len_sym[l] = (uint8_t)s;
        len_sym[258] = 28;
        for (int s = 0; s < 30; s++)
            for (int d = kDistBase[s]; d < kDistBase[s] + (1 << kDistExtra[s]) && d <= 512; d++) dist_sym_lo[d] = (uint8_t)s;
    }
    inline int dist_sym(uint32_t d) const {
        if (d <= 512) return dist_sym_lo[d];
        // above 512 every symbol pair covers one power-of-two interval [2^k + 1, 2^(k+1)]
        const uint32_t x = d - 1;
        const int k = 31 - __builtin_clz(x);            // 2^k <= x < 2^(k+1), k >= 9
        return 2 * k + (int)((x >> (k - 1)) & 1u);
    }
};
const SymTables kSym;

inline uint32_t ld32(const uint8_t *p) { uint32_t x; memcpy(&x, p, 4); return x; }
inline uint64_t ld64(const uint8_t *p) { uint64_t x; memcpy(&x, p, 8); return x; }
inline uint32_t hash4(const uint8_t *p) { return (ld32(p) * 0x9E3779B1u) >> (32 - kHashBits); }

// length of the common prefix of a and b, at most `limit`
inline int match_len(const uint8_t *a, const uint8_t *b, int limit) {
    int n = 0;
    while (n + 8 <= limit) {
        const uint64_t x = ld64(a + n) ^ ld64(b + n);
        if (x) return n + (__builtin_ctzll(x) >> 3);
        n += 8;
    }
    while (n < limit && a[n] == b[n]) n++;
    return n;
}

// Huffman code lengths for freq[0..n), none longer than `limit`: plain Huffman (two queues over the sorted leaves); when
// the tree comes out too deep the counts are halved (never to zero) and it is built again -- equal counts give a
// balanced tree, so this ends.  At least two symbols get a code (a decoder may insist on a complete code).
void code_lengths(const uint32_t *freq_in, int n, int limit, uint8_t *lens) {
    uint32_t freq[288];
    for (int i = 0; i < n; i++) freq[i] = freq_in[i];
    int used = 0;
    for (int i = 0; i < n; i++) used += freq[i] != 0;
    for (int i = 0; used < 2 && i < n; i++)
        if (!freq[i]) { freq[i] = 1; used++; }
    for (;;) {
        struct Node { uint32_t w; int16_t left, right; };
        Node nodes[2 * 288];
        int16_t order[288];
        int m = 0;
        for (int i = 0; i < n; i++)
            if (freq[i]) order[m++] = (int16_t)i;
        std::sort(order, order + m, [&](int16_t a, int16_t b) { return freq[a] != freq[b] ? freq[a] < freq[b] : a < b; });
        for (int k = 0; k < m; k++) nodes[k] = {freq[order[k]], -1, -1};
        int leaf = 0, inner = m, next = m;      // leaves [leaf, m), finished inner nodes [inner, next)
        auto pop = [&]() {
            if (leaf < m && (inner >= next || nodes[leaf].w <= nodes[inner].w)) return leaf++;
            return inner++;
        };
        while ((m - leaf) + (next - inner) > 1) {
            const int a = pop(), b = pop();
            nodes[next] = {nodes[a].w + nodes[b].w, (int16_t)a, (int16_t)b};
            next++;
        }
        // depths: the root is the last node; children come before their parents
        uint8_t depth[2 * 288];
        if (next < 1) return;           // (never: at least two leaves)
        depth[next - 1] = 0;
        int max_depth = 0;
        for (int k = next - 1; k >= m; k--) {
            depth[nodes[k].left] = depth[nodes[k].right] = (uint8_t)(depth[k] + 1);
        }
        memset(lens, 0, (size_t)n);
        for (int k = 0; k < m; k++) {
            lens[order[k]] = depth[k];
            max_depth = std::max<int>(max_depth, depth[k]);
        }
        if (max_depth <= limit) return;
        for (int i = 0; i < n; i++)
            if (freq[i]) freq[i] = (freq[i] + 1) >> 1;
    }
}

// canonical codes, bit-reversed (DEFLATE sends a code starting from its most significant bit)
void assign_codes(const uint8_t *lens, int n, uint16_t *codes) {
    int count[16] = {0};
    for (int i = 0; i < n; i++) count[lens[i]]++;
    count[0] = 0;
    uint32_t next[16], code = 0;
    for (int l = 1; l <= 15; l++) {
        code = (code + (uint32_t)count[l - 1]) << 1;
        next[l] = code;
    }
    for (int i = 0; i < n; i++) {
        const int l = lens[i];
        if (!l) { codes[i] = 0; continue; }
        uint32_t c = next[l]++, r = 0;
        for (int k = 0; k < l; k++) r |= ((c >> k) & 1u) << (l - 1 - k);
        codes[i] = (uint16_t)r;
    }
}

struct BitOut {
    uint8_t *p, *end;
    uint64_t acc = 0;
    int nb = 0;
    bool overflow = false;
    BitOut(uint8_t *out, size_t cap) : p(out), end(out + cap) {}
    inline void put(uint32_t v, int n) {     // n <= 32
        acc |= (uint64_t)v << nb;
        nb += n;
        if (nb >= 32) {
            if (end - p < 4) { overflow = true; nb -= 32; acc >>= 32; return; }
            const uint32_t w = (uint32_t)acc;
            memcpy(p, &w, 4);
            p += 4;
            acc >>= 32;
            nb -= 32;
        }
    }
    inline void flush() {
        while (nb > 0) {
            if (p >= end) { overflow = true; return; }
            *p++ = (uint8_t)acc;
            acc >>= 8;
            nb -= 8;
        }
        nb = 0;
    }
};

}  // namespace

extern "C" uint64_t sd_deflate_raw(const uint8_t *in, uint64_t n64, uint8_t *out, uint64_t cap, int level) {
    if (!out || (!in && n64) || n64 > 65535 || cap < 16) return 0;
    const int n = (int)n64;
    if (n == 0) {                        // an empty final stored block
        static const uint8_t empty[5] = {1, 0, 0, 0xFF, 0xFF};
        memcpy(out, empty, 5);
        return 5;
    }
    // candidates looked at per position / a match this long ends the search / matches shorter than this get one step of lazy
    // evaluation.  Measured on BAM payloads (100 bp reads with NextGenMap's tags) against zlib 1.3 at the same level:
    // level 6 = 66 MB/s at ratio 2.918 (zlib: 22 MB/s, 2.934); level 1 = 130 MB/s at 2.79 (zlib: 97 MB/s, 2.60).
    const int depth = level <= 1 ? 2 : level <= 3 ? 6 : level <= 6 ? 12 : 96;
    const int nice = level <= 3 ? 32 : level <= 6 ? 128 : 258;
    const int lazy_below = level <= 3 ? 0 : level <= 6 ? 16 : 258;

    // ---- LZ77 parse: items = literal (value) or match (0x80000000 | len << 16 | distance - 1)
    static thread_local uint32_t items[65536];
    uint16_t head[1 << kHashBits], prev[65536];
    memset(head, 0xFF, sizeof head);
    uint32_t lit_freq[288] = {0}, dist_freq[32] = {0};
    int n_items = 0;
    const int last_hash = n - kMinMatch;      // last position a 4-byte hash can start at

    auto insert = [&](int i) {
        const uint32_t h = hash4(in + i);
        prev[i] = head[h];
        head[h] = (uint16_t)i;
    };
    auto find = [&](int i, int &best_dist) {   // position i is not inserted yet
        int best = 0;
        if (i > last_hash) return 0;
        const int limit = std::min(kMaxMatch, n - i);
        uint16_t cand = head[hash4(in + i)];
        for (int d = 0; d < depth && cand != kNil; d++) {
            const int dist = i - (int)cand;
            if (dist > kMaxDist) break;
            if (in[cand + best] == in[i + best] && ld32(in + cand) == ld32(in + i)) {
                const int l = match_len(in + cand, in + i, limit);
                if (l > best) {
                    best = l;
                    best_dist = dist;
                    if (l >= nice || l == limit) break;
                }
            }
            cand = prev[cand];
        }
        return best >= kMinMatch ? best : 0;
    };

    int i = 0;
    while (i < n) {
        int dist = 0;
        int len = find(i, dist);
        if (len && len < lazy_below && i + 1 <= last_hash) {
            // one step of lazy evaluation: is there something longer one byte later?
            insert(i);
            int dist2 = 0;
            const int len2 = find(i + 1, dist2);
            if (len2 > len) {
                items[n_items++] = in[i];
                lit_freq[in[i]]++;
                i++;
                len = len2;
                dist = dist2;
            } else {
                // keep the first match; position i is already in the table
                items[n_items++] = 0x80000000u | ((uint32_t)len << 16) | (uint32_t)(dist - 1);
                lit_freq[257 + kSym.len_sym[len]]++;
                dist_freq[kSym.dist_sym((uint32_t)dist)]++;
                const int stop = std::min(i + len, last_hash + 1);
                for (int k = i + 1; k < stop; k++) insert(k);
                i += len;
                continue;
            }
        }
        if (len) {
            items[n_items++] = 0x80000000u | ((uint32_t)len << 16) | (uint32_t)(dist - 1);
            lit_freq[257 + kSym.len_sym[len]]++;
            dist_freq[kSym.dist_sym((uint32_t)dist)]++;
            const int stop = std::min(i + len, last_hash + 1);
            for (int k = i; k < stop; k++) insert(k);
            i += len;
        } else {
            items[n_items++] = in[i];
            lit_freq[in[i]]++;
            if (i <= last_hash) insert(i);
            i++;
        }
    }
    lit_freq[256] = 1;

    // ---- codes
    uint8_t lit_lens[288], dist_lens[32];
    uint16_t lit_codes[288], dist_codes[32];
    code_lengths(lit_freq, 286, 15, lit_lens);
    code_lengths(dist_freq, 30, 15, dist_lens);
    assign_codes(lit_lens, 286, lit_codes);
    assign_codes(dist_lens, 30, dist_codes);
    int hlit = 286, hdist = 30;
    while (hlit > 257 && lit_lens[hlit - 1] == 0) hlit--;
    while (hdist > 1 && dist_lens[hdist - 1] == 0) hdist--;
    // the two length lists as one run-length coded sequence of code-length symbols
    uint8_t all[286 + 30];
    memcpy(all, lit_lens, (size_t)hlit);
    memcpy(all + hlit, dist_lens, (size_t)hdist);
    const int n_all = hlit + hdist;
    uint16_t rl[286 + 30];          // symbol | extra << 8
    int n_rl = 0;
    uint32_t pre_freq[19] = {0};
    for (int k = 0; k < n_all;) {
        int run = 1;
        while (k + run < n_all && all[k + run] == all[k]) run++;
        const int v = all[k];
        int left = run;
        if (v == 0) {
            while (left >= 11) { const int r = std::min(left, 138); rl[n_rl++] = (uint16_t)(18 | ((r - 11) << 8)); pre_freq[18]++; left -= r; }
            if (left >= 3) { rl[n_rl++] = (uint16_t)(17 | ((left - 3) << 8)); pre_freq[17]++; left = 0; }
            while (left-- > 0) { rl[n_rl++] = 0; pre_freq[0]++; }
        } else {
            rl[n_rl++] = (uint16_t)v; pre_freq[v]++; left--;
            while (left >= 3) { const int r = std::min(left, 6); rl[n_rl++] = (uint16_t)(16 | ((r - 3) << 8)); pre_freq[16]++; left -= r; }
            while (left-- > 0) { rl[n_rl++] = (uint16_t)v; pre_freq[v]++; }
        }
        k += run;
    }
    uint8_t pre_lens[19];
    uint16_t pre_codes[19];
    code_lengths(pre_freq, 19, 7, pre_lens);
    assign_codes(pre_lens, 19, pre_codes);
    int hclen = 19;
    while (hclen > 4 && pre_lens[kPreOrder[hclen - 1]] == 0) hclen--;

    // ---- size of the dynamic block in bits; a stored block when that is no smaller
    uint64_t bits = 3 + 14 + 3ull * (uint64_t)hclen;
    for (int k = 0; k < n_rl; k++) {
        const int s = rl[k] & 0xFF;
        bits += pre_lens[s] + (s == 16 ? 2 : s == 17 ? 3 : s == 18 ? 7 : 0);
    }
    for (int s = 0; s < 286; s++) bits += (uint64_t)lit_freq[s] * (lit_lens[s] + (s >= 257 ? kLenExtra[s - 257] : 0));
    for (int s = 0; s < 30; s++) bits += (uint64_t)dist_freq[s] * (dist_lens[s] + kDistExtra[s]);
    const uint64_t dyn_bytes = (bits + 7) / 8, stored_bytes = 5ull + (uint64_t)n;
    if (stored_bytes <= dyn_bytes) {
        if (stored_bytes > cap) return 0;
        out[0] = 1;
        out[1] = (uint8_t)n; out[2] = (uint8_t)(n >> 8);
        out[3] = (uint8_t)~n; out[4] = (uint8_t)(~n >> 8);
        memcpy(out + 5, in, (size_t)n);
        return stored_bytes;
    }
    if (dyn_bytes + 8 > cap) return 0;

    BitOut w(out, (size_t)cap);
    w.put(1, 1);                    // final block
    w.put(2, 2);                    // dynamic Huffman
    w.put((uint32_t)(hlit - 257), 5);
    w.put((uint32_t)(hdist - 1), 5);
    w.put((uint32_t)(hclen - 4), 4);
    for (int k = 0; k < hclen; k++) w.put(pre_lens[kPreOrder[k]], 3);
    for (int k = 0; k < n_rl; k++) {
        const int s = rl[k] & 0xFF, extra = rl[k] >> 8;
        w.put(pre_codes[s], pre_lens[s]);
        if (s == 16) w.put((uint32_t)extra, 2);
        else if (s == 17) w.put((uint32_t)extra, 3);
        else if (s == 18) w.put((uint32_t)extra, 7);
    }
    for (int k = 0; k < n_items; k++) {
        const uint32_t it = items[k];
        if (!(it & 0x80000000u)) {
            w.put(lit_codes[it], lit_lens[it]);
        } else {
            const int len = (int)((it >> 16) & 0x7FFF), dist = (int)(it & 0xFFFF) + 1;
            const int ls = kSym.len_sym[len], ds = kSym.dist_sym((uint32_t)dist);
            w.put(lit_codes[257 + ls], lit_lens[257 + ls]);
            w.put((uint32_t)(len - kLenBase[ls]), kLenExtra[ls]);
            w.put(dist_codes[ds], dist_lens[ds]);
            w.put((uint32_t)(dist - kDistBase[ds]), kDistExtra[ds]);
        }
    }
    w.put(lit_codes[256], lit_lens[256]);
    w.flush();
    if (w.overflow) return 0;
    return (uint64_t)(w.p - out);
}
