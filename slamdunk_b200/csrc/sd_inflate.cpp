// sd_inflate.cpp -- raw DEFLATE (RFC 1951) decoder for BGZF blocks: whole block in, whole block out.
//
// Every BGZF block of a BAM is an independent DEFLATE stream of at most 64 KiB with its CRC32 and size in the
// trailer, so the decoder needs no streaming state, no window and no checksum of its own: sd_bam_inflate_file
// (sd_decode.cpp) checks the CRC of what comes out and falls back to zlib's inflate for any block this decoder turns
// down (it returns 0 rather than guess on anything unusual).  What makes it faster than a general-purpose inflate:
// a 64-bit bit buffer refilled with one unaligned load, 11-bit first-level tables whose entries carry literal /
// length base / extra-bit count ready to use, up to three literals per refill, 8-byte match copies.
#include <stdint.h>
#include <string.h>

#include "slamdunk_b200.h"

namespace {

constexpr int kLitBits = 11;    // first-level index width of the literal/length table
constexpr int kDistBits = 8;    // ... of the distance table
constexpr int kPreBits = 7;     // code-length code: at most 7 bits, no second level
constexpr int kLitCap = (1 << kLitBits) + 2048;
constexpr int kDistCap = (1 << kDistBits) + 1024;

// table entry: [31] literal  [30] end of block  [29] second-level pointer  [28] invalid
//              [27:12] literal byte / base value / second-level start index
//              [11:8]  number of extra bits (or second-level index width)
//              [7:0]   code bits to drop at this level
constexpr uint32_t kLit = 1u << 31, kEob = 1u << 30, kSub = 1u << 29, kBad = 1u << 28;

struct Tables {
    uint32_t lit[kLitCap];
    uint32_t dist[kDistCap];
    uint32_t pre[1 << kPreBits];
};

const uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
const uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
const uint8_t kPreOrder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

inline uint32_t lit_entry(uint32_t sym) {
    if (sym < 256) return kLit | (sym << 12);
    if (sym == 256) return kEob;
    if (sym > 285) return kBad;     // 286, 287 take part in the fixed code but never occur
    return ((uint32_t)kLenBase[sym - 257] << 12) | ((uint32_t)kLenExtra[sym - 257] << 8);
}
inline uint32_t dist_entry(uint32_t sym) {
    if (sym > 29) return kBad;
    return ((uint32_t)kDistBase[sym] << 12) | ((uint32_t)kDistExtra[sym] << 8);
}
inline uint32_t pre_entry(uint32_t sym) { return sym << 12; }

inline uint32_t reverse_bits(uint32_t code, int len) {
    uint32_t r = 0;
    for (int i = 0; i < len; i++) r |= ((code >> i) & 1u) << (len - 1 - i);
    return r;
}

// Canonical Huffman code (lens[0..n), 0 = unused, at most 15 bits) -> decode table read with the NEXT bits of the
// stream (DEFLATE packs codes starting from their most significant bit, so entries sit at the bit-reversed code).
// Only complete codes are taken, plus the one incomplete form DEFLATE permits in practice: a single code of one bit
// (a block with one distance code, or none).  Anything else -> false, and the caller leaves the block to zlib.
template <class EntryFn>
bool build_table(uint32_t *table, int cap, int root, const uint8_t *lens, int n, EntryFn entry_of) {
    int count[16] = {0};
    for (int i = 0; i < n; i++) count[lens[i]]++;
    if (count[0] == n) {            // no code at all (a block of literals only has no distance code): every lookup is invalid
        for (int k = 0; k < (1 << root); k++) table[k] = kBad | 1u;
        return true;
    }
    int left = 1, max_len = 0, used = 0;
    for (int len = 1; len <= 15; len++) {
        left = (left << 1) - count[len];
        if (left < 0) return false;                     // over-subscribed
        if (count[len]) max_len = len;
        used += count[len];
    }
    const bool single = used == 1 && count[1] == 1;
    if (left != 0 && !single) return false;             // incomplete
    uint32_t next_code[16];
    {
        uint32_t code = 0;
        for (int len = 1; len <= 15; len++) {
            code = (code + (uint32_t)count[len - 1] * (len > 1)) << 1;
            next_code[len] = code;
        }
    }
    const uint32_t root_mask = (1u << root) - 1u;
    for (int k = 0; k < (1 << root); k++) table[k] = kBad | 1u;
    // codes longer than the first level: width of each prefix's second level = its longest code - root
    uint8_t sub_len[1 << kLitBits];
    if (max_len > root) memset(sub_len, 0, (size_t)1 << root);
    // first pass: assign codes in canonical order (by length, then by symbol), remember each symbol's reversed code
    uint16_t rev_code[320];
    {
        uint32_t nc[16];
        memcpy(nc, next_code, sizeof nc);
        for (int i = 0; i < n; i++) {
            const int len = lens[i];
            if (!len) continue;
            const uint32_t rev = reverse_bits(nc[len]++, len);
            rev_code[i] = (uint16_t)rev;
            if (len > root) {
                uint8_t &m = sub_len[rev & root_mask];
                if ((int)m < len - root) m = (uint8_t)(len - root);
            }
        }
    }
    int next_free = 1 << root;
    if (max_len > root) {
        for (uint32_t pfx = 0; pfx <= root_mask; pfx++) {
            if (!sub_len[pfx]) continue;
            const int width = sub_len[pfx];
            if (next_free + (1 << width) > cap) return false;
            table[pfx] = kSub | ((uint32_t)next_free << 12) | ((uint32_t)width << 8) | (uint32_t)root;
            for (int k = 0; k < (1 << width); k++) table[next_free + k] = kBad | 1u;
            next_free += 1 << width;
        }
    }
    for (int i = 0; i < n; i++) {
        const int len = lens[i];
        if (!len) continue;
        const uint32_t rev = rev_code[i];
        const uint32_t e = entry_of((uint32_t)i);
        if (len <= root) {
            for (uint32_t k = rev; k <= root_mask; k += 1u << len) table[k] = e | (uint32_t)len;
        } else {
            const uint32_t head = table[rev & root_mask];
            if (!(head & kSub)) return false;           // a shorter code owns this prefix: not a prefix code
            const uint32_t start = (head >> 12) & 0xFFFFu, width = (head >> 8) & 15u;
            for (uint32_t k = rev >> root; k < (1u << width); k += 1u << (len - root)) table[start + k] = e | (uint32_t)(len - root);
        }
    }
    return true;
}

struct Fixed {
    uint32_t lit[kLitCap];
    uint32_t dist[kDistCap];
    bool ok;
    Fixed() {
        uint8_t l[288], d[32];
        for (int i = 0; i < 144; i++) l[i] = 8;
        for (int i = 144; i < 256; i++) l[i] = 9;
        for (int i = 256; i < 280; i++) l[i] = 7;
        for (int i = 280; i < 288; i++) l[i] = 8;
        for (int i = 0; i < 32; i++) d[i] = 5;
        ok = build_table(lit, kLitCap, kLitBits, l, 288, lit_entry) && build_table(dist, kDistCap, kDistBits, d, 32, dist_entry);
    }
};
const Fixed kFixed;

struct Bits {
    const uint8_t *in, *in_end;
    uint64_t buf = 0;
    int cnt = 0;            // valid bits in buf; goes negative when the stream is read past its end
    Bits(const uint8_t *p, size_t n) : in(p), in_end(p + n) {}
    inline void refill() {
        if (in_end - in >= 8) {
            uint64_t x;
            memcpy(&x, in, 8);
            buf |= x << cnt;
            in += (63 - cnt) >> 3;
            cnt |= 56;
        } else {
            while (cnt <= 56 && in < in_end) {
                buf |= (uint64_t)*in++ << cnt;
                cnt += 8;
            }
        }
    }
    inline void refill_fast() {     // the caller knows that 8 bytes can be read
        uint64_t x;
        memcpy(&x, in, 8);
        buf |= x << cnt;
        in += (63 - cnt) >> 3;
        cnt |= 56;
    }
    inline uint32_t peek(int n) const { return (uint32_t)(buf & ((1ull << n) - 1ull)); }
    inline void drop(int n) {
        buf >>= n;
        cnt -= n;
    }
    inline uint32_t take(int n) {
        const uint32_t v = peek(n);
        drop(n);
        return v;
    }
};

// next symbol of a two-level table: returns the entry, its code bits already dropped
inline uint32_t decode(Bits &b, const uint32_t *table, int root) {
    uint32_t e = table[b.peek(root)];
    if (e & kSub) {
        b.drop((int)(e & 0xFFu));
        e = table[((e >> 12) & 0xFFFFu) + b.peek((int)((e >> 8) & 15u))];
    }
    b.drop((int)(e & 0xFFu));
    return e;
}

}  // namespace

extern "C" int sd_inflate_raw(const uint8_t *in, uint64_t in_len, uint8_t *out, uint64_t out_len) {
    if ((!in && in_len) || (!out && out_len) || !kFixed.ok) return 0;
    Bits b(in, (size_t)in_len);
    uint8_t *const out_begin = out, *const out_end = out + out_len;
    Tables T;
    for (;;) {
        b.refill();
        const uint32_t final_block = b.take(1), type = b.take(2);
        if (b.cnt < 0) return 0;
        const uint32_t *lit = nullptr, *dist = nullptr;
        if (type == 0) {            // stored: back to a byte boundary, LEN, ~LEN, bytes
            b.drop(b.cnt & 7);
            const uint8_t *p = b.in - (b.cnt >> 3);     // bytes still in the bit buffer have not been used
            if (b.in_end - p < 4) return 0;
            const uint32_t len = p[0] | ((uint32_t)p[1] << 8), nlen = p[2] | ((uint32_t)p[3] << 8);
            p += 4;
            if ((len ^ nlen) != 0xFFFFu || (uint64_t)(b.in_end - p) < len || (uint64_t)(out_end - out) < len) return 0;
            memcpy(out, p, len);
            out += len;
            b.in = p + len;
            b.buf = 0;
            b.cnt = 0;
            if (final_block) break;
            continue;
        } else if (type == 1) {
            lit = kFixed.lit;
            dist = kFixed.dist;
        } else if (type == 2) {
            const uint32_t hlit = b.take(5) + 257, hdist = b.take(5) + 1, hclen = b.take(4) + 4;
            if (hlit > 286 || hdist > 30) return 0;
            uint8_t pre_lens[19] = {0};
            for (uint32_t i = 0; i < hclen; i++) {
                if (b.cnt < 3) b.refill();
                pre_lens[kPreOrder[i]] = (uint8_t)b.take(3);
            }
            if (b.cnt < 0 || !build_table(T.pre, 1 << kPreBits, kPreBits, pre_lens, 19, pre_entry)) return 0;
            uint8_t lens[286 + 30 + 138];
            uint32_t i = 0;
            while (i < hlit + hdist) {
                b.refill();
                const uint32_t e = T.pre[b.peek(kPreBits)];
                if (e & kBad) return 0;
                b.drop((int)(e & 0xFFu));
                const uint32_t sym = (e >> 12) & 0xFFFFu;
                if (sym < 16) {
                    lens[i++] = (uint8_t)sym;
                } else {
                    uint32_t rep, val = 0;
                    if (sym == 16) {
                        if (i == 0) return 0;
                        val = lens[i - 1];
                        rep = 3 + b.take(2);
                    } else if (sym == 17) {
                        rep = 3 + b.take(3);
                    } else {
                        rep = 11 + b.take(7);
                    }
                    if (i + rep > hlit + hdist) return 0;
                    memset(lens + i, (int)val, rep);
                    i += rep;
                }
                if (b.cnt < 0) return 0;
            }
            if (lens[256] == 0) return 0;       // no end-of-block code
            if (!build_table(T.lit, kLitCap, kLitBits, lens, (int)hlit, lit_entry)) return 0;
            if (!build_table(T.dist, kDistCap, kDistBits, lens + hlit, (int)hdist, dist_entry)) return 0;
            lit = T.lit;
            dist = T.dist;
        } else {
            return 0;
        }

        // ---- symbols of the block
        for (;;) {
            uint32_t e;
            // fast loop: input for two 8-byte refills, room for the longest match plus the slack of the 8-byte copies
            while (b.in_end - b.in >= 16 && out_end - out >= 258 + 16) {
                b.refill_fast();                            // >= 56 bits
                e = lit[b.peek(kLitBits)];
                // literals whose code fits the first level (<= 11 bits each): three of them and the symbol after fit one refill
                if (e & kLit) {
                    b.drop((int)(e & 0xFFu));
                    *out++ = (uint8_t)(e >> 12);
                    e = lit[b.peek(kLitBits)];
                    if (e & kLit) {
                        b.drop((int)(e & 0xFFu));
                        *out++ = (uint8_t)(e >> 12);
                        e = lit[b.peek(kLitBits)];
                        if (e & kLit) {
                            b.drop((int)(e & 0xFFu));
                            *out++ = (uint8_t)(e >> 12);
                            e = lit[b.peek(kLitBits)];
                        }
                    }
                }
                if (e & kSub) {                             // second level: codes of 12 to 15 bits
                    b.drop((int)(e & 0xFFu));
                    e = lit[((e >> 12) & 0xFFFFu) + b.peek((int)((e >> 8) & 15u))];
                }
                b.drop((int)(e & 0xFFu));                   // <= 33 + 15 bits so far
                if (e & kLit) {
                    *out++ = (uint8_t)(e >> 12);
                    continue;
                }
                if (e & (kEob | kBad)) goto block_done;
                {
                    const uint32_t len = ((e >> 12) & 0xFFFFu) + b.take((int)((e >> 8) & 15u));      // <= 53 bits so far
                    b.refill_fast();                        // the first refill took at most 7 of the 16 bytes
                    uint32_t d = dist[b.peek(kDistBits)];
                    if (d & kSub) {
                        b.drop((int)(d & 0xFFu));
                        d = dist[((d >> 12) & 0xFFFFu) + b.peek((int)((d >> 8) & 15u))];
                    }
                    b.drop((int)(d & 0xFFu));
                    if (d & kBad) return 0;
                    const uint32_t distance = ((d >> 12) & 0xFFFFu) + b.take((int)((d >> 8) & 15u));
                    if (distance > (uint32_t)(out - out_begin)) return 0;
                    const uint8_t *src = out - distance;
                    uint8_t *dst = out;
                    out += len;
                    if (distance >= 8) {
                        memcpy(dst, src, 8);
                        memcpy(dst + 8, src + 8, 8);
                        if (len > 16) {
                            dst += 16;
                            src += 16;
                            do {
                                memcpy(dst, src, 8);
                                dst += 8;
                                src += 8;
                            } while (dst < out);
                        }
                    } else if (distance == 1) {             // a run of one byte
                        const uint64_t rep = 0x0101010101010101ULL * *src;
                        do {
                            memcpy(dst, &rep, 8);
                            dst += 8;
                        } while (dst < out);
                    } else {
                        // period 2..7: every 8-byte copy from the start of the pattern lays down `distance` good bytes (the
                        // rest is overwritten by the next one)
                        do {
                            memcpy(dst, src, 8);
                            dst += distance;
                        } while (dst < out);
                    }
                }
            }
            // careful loop: near the end of the input or of the output
            b.refill();
            e = decode(b, lit, kLitBits);
            if (b.cnt < 0) return 0;
            if (e & kLit) {
                if (out >= out_end) return 0;
                *out++ = (uint8_t)(e >> 12);
                continue;
            }
            if (e & (kEob | kBad)) goto block_done;
            {
                const uint32_t len = ((e >> 12) & 0xFFFFu) + b.take((int)((e >> 8) & 15u));
                b.refill();
                const uint32_t d = decode(b, dist, kDistBits);
                if (d & (kBad | kLit | kEob)) return 0;
                const uint32_t distance = ((d >> 12) & 0xFFFFu) + b.take((int)((d >> 8) & 15u));
                if (b.cnt < 0 || distance > (uint32_t)(out - out_begin) || len > (uint64_t)(out_end - out)) return 0;
                const uint8_t *src = out - distance;
                for (uint32_t k = 0; k < len; k++) out[k] = src[k];
                out += len;
            }
            continue;
        block_done:
            if (e & kBad) return 0;
            break;
        }
        if (b.cnt < 0) return 0;
        if (final_block) break;
    }
    return out == out_end ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------------------------
// CRC-32 (the gzip / BGZF checksum, reflected polynomial 0xEDB88320) by carry-less multiplication: four 128-bit lanes
// are folded forward 64 bytes at a time, then into one lane, then reduced to 32 bits (Barrett) -- the scheme of Intel's
// "Fast CRC Computation for Generic Polynomials Using PCLMULQDQ".  Only used where the CPU has PCLMULQDQ; the tail that
// is not a multiple of 16 bytes, short inputs and every other CPU go through zlib's crc32.
// ---------------------------------------------------------------------------------------------------------------
#include <immintrin.h>
#include <zlib.h>

namespace {
// one lane 128 bits forward, then the next 16 bytes on top
__attribute__((target("pclmul,sse4.1"))) inline __m128i crc32_fold(__m128i acc, __m128i next, __m128i k) {
    const __m128i lo = _mm_clmulepi64_si128(acc, k, 0x00), hi = _mm_clmulepi64_si128(acc, k, 0x11);
    return _mm_xor_si128(_mm_xor_si128(hi, lo), next);
}
// constants: x^(4*128+32), x^(4*128-32), x^(128+32), x^(128-32), x^64 mod P (bit-reflected), then P and floor(x^64 / P)
__attribute__((target("pclmul,sse4.1"))) uint32_t crc32_clmul(const uint8_t *p, size_t n, uint32_t state) {
    // n >= 64 and a multiple of 16; state = running CRC register (already inverted)
    const __m128i k1k2 = _mm_set_epi64x(0x01c6e41596LL, 0x0154442bd4LL);
    const __m128i k3k4 = _mm_set_epi64x(0x00ccaa009eLL, 0x01751997d0LL);
    const __m128i k5 = _mm_set_epi64x(0, 0x0163cd6124LL);
    const __m128i poly = _mm_set_epi64x(0x01f7011641LL, 0x01db710641LL);
    const __m128i low32 = _mm_set_epi32(0, ~0, 0, ~0);
    __m128i x1 = _mm_loadu_si128((const __m128i *)(p + 0)), x2 = _mm_loadu_si128((const __m128i *)(p + 16));
    __m128i x3 = _mm_loadu_si128((const __m128i *)(p + 32)), x4 = _mm_loadu_si128((const __m128i *)(p + 48));
    x1 = _mm_xor_si128(x1, _mm_cvtsi32_si128((int)state));
    p += 64;
    n -= 64;
    while (n >= 64) {
        const __m128i a1 = _mm_clmulepi64_si128(x1, k1k2, 0x00), a2 = _mm_clmulepi64_si128(x2, k1k2, 0x00);
        const __m128i a3 = _mm_clmulepi64_si128(x3, k1k2, 0x00), a4 = _mm_clmulepi64_si128(x4, k1k2, 0x00);
        x1 = _mm_clmulepi64_si128(x1, k1k2, 0x11);
        x2 = _mm_clmulepi64_si128(x2, k1k2, 0x11);
        x3 = _mm_clmulepi64_si128(x3, k1k2, 0x11);
        x4 = _mm_clmulepi64_si128(x4, k1k2, 0x11);
        x1 = _mm_xor_si128(_mm_xor_si128(x1, a1), _mm_loadu_si128((const __m128i *)(p + 0)));
        x2 = _mm_xor_si128(_mm_xor_si128(x2, a2), _mm_loadu_si128((const __m128i *)(p + 16)));
        x3 = _mm_xor_si128(_mm_xor_si128(x3, a3), _mm_loadu_si128((const __m128i *)(p + 32)));
        x4 = _mm_xor_si128(_mm_xor_si128(x4, a4), _mm_loadu_si128((const __m128i *)(p + 48)));
        p += 64;
        n -= 64;
    }
    // four lanes into one
    x1 = crc32_fold(x1, x2, k3k4);
    x1 = crc32_fold(x1, x3, k3k4);
    x1 = crc32_fold(x1, x4, k3k4);
    while (n >= 16) {
        x1 = crc32_fold(x1, _mm_loadu_si128((const __m128i *)p), k3k4);
        p += 16;
        n -= 16;
    }
    // 128 -> 64 bits
    __m128i t = _mm_clmulepi64_si128(x1, k3k4, 0x10);
    x1 = _mm_xor_si128(_mm_srli_si128(x1, 8), t);
    t = _mm_srli_si128(x1, 4);
    x1 = _mm_and_si128(x1, low32);
    x1 = _mm_xor_si128(_mm_clmulepi64_si128(x1, k5, 0x00), t);
    // Barrett reduction 64 -> 32 bits
    t = _mm_and_si128(x1, low32);
    t = _mm_clmulepi64_si128(t, poly, 0x10);
    t = _mm_and_si128(t, low32);
    t = _mm_clmulepi64_si128(t, poly, 0x00);
    x1 = _mm_xor_si128(x1, t);
    return (uint32_t)_mm_extract_epi32(x1, 1);
}
}  // namespace

extern "C" uint32_t sd_crc32(const uint8_t *p, uint64_t n) {
    static const bool have_clmul = __builtin_cpu_supports("pclmul") && __builtin_cpu_supports("sse4.1");
    uint32_t crc = 0;       // zlib's convention: the value after the final inversion
    if (have_clmul && n >= 64) {
        const size_t body = (size_t)n & ~(size_t)15;
        crc = ~crc32_clmul(p, body, ~crc);
        p += body;
        n -= body;
    }
    while (n) {             // zlib takes 32-bit lengths
        const uInt step = (uInt)(n > 0x40000000ull ? 0x40000000ull : n);
        crc = (uint32_t)crc32(crc, p, step);
        p += step;
        n -= step;
    }
    return crc;
}
