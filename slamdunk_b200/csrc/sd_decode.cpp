// sd_decode.cpp -- host-side decode for the count path: BGZF/BAM -> columnar batch, FASTA -> bit-planes.
//
// Takes over what htslib does for the reference behind pysam (AlignmentFile iteration,
// FastaFile.fetch; call sites slamseq/SlamSeqFile.py:407-441, dunks/tcounter.py:127,181,
// dunks/filter.py:225-256).  Written against the SAM/BAM specification; C++17 threads + zlib.
// The decode cost is reported separately from the counting kernels (sd_hbatch.seconds_*).
#include <cuda_runtime.h>
#include <fcntl.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "sd_bamio.h"
#include "slamdunk_b200.h"

namespace {

double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int fail(char *err, int errlen, int code, const char *fmt, ...) {
    if (err && errlen > 0) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(err, (size_t)errlen, fmt, ap);
        va_end(ap);
    }
    return code;
}

// Page-locked when a CUDA device is present (async H2D), plain aligned memory otherwise
// (CPU-only decode tests).
// Page-locked host memory for the decoded columns.  Pinning pages is slow (a few GB/s, single threaded), about as long as
// inflating the file on all cores, so one arena sized from the file is pinned on a helper thread WHILE the BGZF blocks are
// inflated; the columns are then carved out of it (anything that does not fit gets its own allocation).  A freed arena is
// parked in a one-slot process-wide cache: `count` decodes one BAM after the other (12 samples in cfg3), and every file
// after the first gets its pinned memory for nothing.
struct ArenaCache {
    std::mutex mu;
    uint8_t *p = nullptr;
    size_t cap = 0;
    int dev = -1;
    bool take(size_t want, int device, uint8_t **out, size_t *out_cap) {
        std::lock_guard<std::mutex> g(mu);
        if (!p || cap < want || dev != device) return false;
        *out = p;
        *out_cap = cap;
        p = nullptr;
        cap = 0;
        return true;
    }
    void give(uint8_t *q, size_t qcap, int device) {
        uint8_t *drop = q;
        {
            std::lock_guard<std::mutex> g(mu);
            if (!p || qcap > cap) {
                drop = p;
                p = q;
                cap = qcap;
                dev = device;
            }
        }
        if (drop) cudaFreeHost(drop);
    }
};
static ArenaCache &arena_cache() {
    static ArenaCache *c = new ArenaCache();   // never destroyed: the CUDA context may be gone at exit
    return *c;
}

struct HostAlloc {
    std::vector<std::pair<void *, bool>> blocks;
    uint8_t *arena = nullptr;
    size_t arena_cap = 0, arena_used = 0;
    int arena_dev = 0;
    std::thread arena_thread;

    static bool have_device() {
        static int have_dev = -1;
        if (have_dev < 0) {
            int n = 0;
            have_dev = (cudaGetDeviceCount(&n) == cudaSuccess && n > 0) ? 1 : 0;
            cudaGetLastError();
        }
        return have_dev == 1;
    }
    void reserve_async(size_t bytes) {
        if (!have_device() || bytes == 0 || arena_thread.joinable() || arena) return;
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return; }
        const size_t cap = (bytes + 4095) & ~(size_t)4095;
        arena_dev = dev;
        if (arena_cache().take(cap, dev, &arena, &arena_cap)) return;
        arena_thread = std::thread([this, dev, cap]() {
            void *p = nullptr;
            if (cudaSetDevice(dev) == cudaSuccess && cudaHostAlloc(&p, cap, cudaHostAllocDefault) == cudaSuccess) {
                arena = static_cast<uint8_t *>(p);
                arena_cap = cap;
            } else {
                cudaGetLastError();
            }
        });
    }
    void *get(size_t bytes) {
        if (arena_thread.joinable()) arena_thread.join();
        if (bytes == 0) bytes = 16;
        bytes = (bytes + 255) & ~(size_t)255;
        if (arena && arena_used + bytes <= arena_cap) {
            void *p = arena + arena_used;
            arena_used += bytes;
            return p;
        }
        void *p = nullptr;
        bool pinned = false;
        if (have_device() && cudaHostAlloc(&p, bytes, cudaHostAllocDefault) == cudaSuccess) pinned = true;
        else {
            cudaGetLastError();
            if (posix_memalign(&p, 256, bytes) != 0) return nullptr;
        }
        blocks.push_back({p, pinned});
        return p;
    }
    ~HostAlloc() {
        if (arena_thread.joinable()) arena_thread.join();
        if (arena) arena_cache().give(arena, arena_cap, arena_dev);
        for (auto &b : blocks) {
            if (b.second) cudaFreeHost(b.first);
            else free(b.first);
        }
    }
};

template <class F>
void parallel_for(size_t n, int threads, F f) {
    if (threads <= 1 || n < 2) {
        for (size_t i = 0; i < n; i++) f(i);
        return;
    }
    std::atomic<size_t> next(0);
    const size_t grain = std::max<size_t>(1, n / ((size_t)threads * 16));
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++)
        pool.emplace_back([&]() {
            for (;;) {
                const size_t b = next.fetch_add(grain);
                if (b >= n) break;
                const size_t e = std::min(n, b + grain);
                for (size_t i = b; i < e; i++) f(i);
            }
        });
    for (auto &th : pool) th.join();
}

// the same over ranges [b, e): per-range state (a local table folded into the shared result once per range)
template <class F>
void parallel_ranges(size_t n, int threads, F f) {
    const size_t grain = std::max<size_t>(1, n / ((size_t)std::max(1, threads) * 16));
    parallel_for((n + grain - 1) / grain, threads, [&](size_t c) { f(c * grain, std::min(n, (c + 1) * grain)); });
}

inline uint16_t rd16(const uint8_t *p) { return (uint16_t)(p[0] | (p[1] << 8)); }
inline uint32_t rd32(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
inline int32_t rdi32(const uint8_t *p) { return (int32_t)rd32(p); }

struct AuxInfo {
    float xi;
    uint32_t nm;
    const char *mp;  // MP:Z payload or null
    bool has_mp;
    bool has_xi, has_nm;
};

// Walk the BAM optional fields once.
bool parse_aux(const uint8_t *p, const uint8_t *end, AuxInfo &a) {
    a.xi = NAN;
    a.nm = 0;
    a.mp = nullptr;
    a.has_mp = false;
    a.has_xi = a.has_nm = false;
    while (p + 3 <= end) {
        const uint8_t t0 = p[0], t1 = p[1], ty = p[2];
        p += 3;
        double num = 0;
        bool is_num = false;
        switch (ty) {
            case 'A': case 'c': case 'C': {
                if (p + 1 > end) return false;
                num = (ty == 'c') ? (double)(int8_t)p[0] : (double)p[0];
                is_num = ty != 'A';
                p += 1;
                break;
            }
            case 's': case 'S': {
                if (p + 2 > end) return false;
                num = (ty == 's') ? (double)(int16_t)rd16(p) : (double)rd16(p);
                is_num = true;
                p += 2;
                break;
            }
            case 'i': case 'I': {
                if (p + 4 > end) return false;
                num = (ty == 'i') ? (double)rdi32(p) : (double)rd32(p);
                is_num = true;
                p += 4;
                break;
            }
            case 'f': {
                if (p + 4 > end) return false;
                float f;
                memcpy(&f, p, 4);
                if (t0 == 'X' && t1 == 'I') { a.xi = f; a.has_xi = true; }
                num = f;
                p += 4;
                break;
            }
            case 'Z': case 'H': {
                const uint8_t *z = (const uint8_t *)memchr(p, 0, (size_t)(end - p));
                if (!z) return false;
                if (ty == 'Z' && t0 == 'M' && t1 == 'P') {
                    a.mp = (const char *)p;
                    a.has_mp = true;
                }
                p = z + 1;
                break;
            }
            case 'B': {
                if (p + 5 > end) return false;
                const uint8_t sub = p[0];
                const uint32_t n = rd32(p + 1);
                size_t sz = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : 4;
                p += 5 + (size_t)n * sz;
                if (p > end) return false;
                break;
            }
            default:
                return false;
        }
        if (is_num) {
            if (t0 == 'N' && t1 == 'M') { a.nm = num < 0 ? 0u : (num > 65535.0 ? 65535u : (uint32_t)num); a.has_nm = true; }
            if (t0 == 'X' && t1 == 'I') { a.xi = (float)num; a.has_xi = true; }
        }
    }
    return true;
}

// One decimal number of an MP entry: plain digits are read directly; anything else strtol accepts (leading blanks, a
// sign, more than 18 digits) goes through strtol itself, so the accepted texts and their values are strtol's.
inline bool mp_number(const char *&s, long &v) {
    const char *p = s;
    if (*p >= '0' && *p <= '9') {
        long x = 0;
        int digits = 0;
        do {
            x = x * 10 + (*p - '0');
            p++;
            digits++;
        } while (*p >= '0' && *p <= '9' && digits < 18);
        if (!(*p >= '0' && *p <= '9')) {
            v = x;
            s = p;
            return true;
        }
    }
    char *e;
    v = strtol(s, &e, 10);
    if (e == s) return false;
    s = e;
    return true;
}

// MP:Z is "conv:readPos:refPos,..." (slamseq/SlamSeqFile.py:321-324).  Calls f(readPos-1, refPos-1)
// for entries whose conv equals `want`.  Returns false on malformed text.
template <class F>
bool mp_for_each(const char *s, int want, F f) {
    while (*s) {
        long v[3];
        for (int k = 0; k < 3; k++) {
            if (!mp_number(s, v[k])) return false;
            if (k < 2) {
                if (*s != ':') return false;
                s++;
            }
        }
        if (v[0] == want) f(v[1] - 1, v[2] - 1);
        if (*s == ',') s++;
        else if (*s) return false;
    }
    return true;
}

// Every entry: f(conv, readPos-1, refPos-1).
template <class F>
bool mp_for_all(const char *s, F f) {
    while (*s) {
        long v[3];
        for (int k = 0; k < 3; k++) {
            if (!mp_number(s, v[k])) return false;
            if (k < 2) {
                if (*s != ':') return false;
                s++;
            }
        }
        f(v[0], v[1] - 1, v[2] - 1);
        if (*s == ',') s++;
        else if (*s) return false;
    }
    return true;
}

struct SliceLut {
    uint32_t v[256];
    SliceLut() {
        for (int b = 0; b < 256; b++) {
            const int hi = b >> 4, lo = b & 15;  // BAM: first base in the high nibble
            uint32_t w = 0;
            for (int p = 0; p < 4; p++) w |= (uint32_t)(((hi >> p) & 1) | (((lo >> p) & 1) << 1)) << (8 * p);
            v[b] = w;
        }
    }
};
const SliceLut kSlice;

// ---- word-at-a-time packers of the per-base columns (sixteen bases / eight qualities per step instead of one) ----
inline uint64_t ld64(const uint8_t *p) { uint64_t x; memcpy(&x, p, 8); return x; }
// up to 8 bytes, the rest zero
inline uint64_t ld_tail(const uint8_t *p, uint32_t nbytes) { uint64_t x = 0; memcpy(&x, p, nbytes); return x; }
// BAM packs the first base of a byte in the HIGH nibble: after this, nibble i of the word is base i
inline uint64_t nibble_swap(uint64_t x) { return ((x & 0x0F0F0F0F0F0F0F0FULL) << 4) | ((x >> 4) & 0x0F0F0F0F0F0F0F0FULL); }
// bit 4i of x (i = 0..15) -> bit i
inline uint32_t gather_every_4th(uint64_t x) {
    x &= 0x1111111111111111ULL;
    x = (x | (x >> 3)) & 0x0303030303030303ULL;
    x = (x | (x >> 6)) & 0x000F000F000F000FULL;
    x = (x | (x >> 12)) & 0x000000FF000000FFULL;
    x = (x | (x >> 24)) & 0xFFFFULL;
    return (uint32_t)x;
}
// sixteen nibbles (base i in nibble i) -> per-nibble flags at bit 4i: low / high bit of the 2-bit base code A=0 C=1 G=2 T=3,
// anything that is not exactly one of C (2), G (4), T (8) coded A
inline void code_flags(uint64_t x, uint64_t &lo, uint64_t &hi) {
    const uint64_t m = 0x1111111111111111ULL;
    const uint64_t b0 = x & m, b1 = (x >> 1) & m, b2 = (x >> 2) & m, b3 = (x >> 3) & m;
    const uint64_t n0 = b0 ^ m, n1 = b1 ^ m, n2 = b2 ^ m, n3 = b3 ^ m;
    const uint64_t onlyC = n0 & b1 & n2 & n3, onlyG = n0 & n1 & b2 & n3, onlyT = n0 & n1 & n2 & b3;
    lo = onlyC | onlyT;
    hi = onlyG | onlyT;
}
// sixteen nibbles -> sixteen 2-bit codes, base i at bits 2i
inline uint32_t codes16(uint64_t x) {
    uint64_t lo, hi;
    code_flags(x, lo, hi);
    uint64_t c = lo | (hi << 1);
    c = (c | (c >> 2)) & 0x0F0F0F0F0F0F0F0FULL;
    c = (c | (c >> 4)) & 0x00FF00FF00FF00FFULL;
    c = (c | (c >> 8)) & 0x0000FFFF0000FFFFULL;
    c = (c | (c >> 16)) & 0x00000000FFFFFFFFULL;
    return (uint32_t)c;
}
// eight dictionary codes, one per byte -> 8 x bits (bits = 2 or 4), quality j at bits j * bits
inline uint32_t squeeze8(uint64_t w, uint32_t bits) {
    if (bits == 2) {
        w = (w | (w >> 6)) & 0x000F000F000F000FULL;
        w = (w | (w >> 12)) & 0x000000FF000000FFULL;
        w = (w | (w >> 24)) & 0xFFFFULL;
    } else {
        w = (w | (w >> 4)) & 0x00FF00FF00FF00FFULL;
        w = (w | (w >> 8)) & 0x0000FFFF0000FFFFULL;
        w = (w | (w >> 16)) & 0xFFFFFFFFULL;
    }
    return (uint32_t)w;
}
// little-endian bit stream: put() appends the low nbits (<= 32) of v; the bytes come out as if every field had been
// OR-ed into place one by one
struct BitWriter {
    uint8_t *p;
    uint64_t acc = 0;
    uint32_t nb = 0;
    explicit BitWriter(uint8_t *dst) : p(dst) {}
    inline void put(uint32_t v, uint32_t nbits) {
        acc |= (uint64_t)v << nb;
        nb += nbits;
        if (nb >= 64) {
            memcpy(p, &acc, 8);
            p += 8;
            nb -= 64;
            acc = nb ? (uint64_t)v >> (nbits - nb) : 0;
        }
    }
    // whole bytes written so far go out, a partly filled last byte too; returns the end of the written bytes
    inline uint8_t *finish() {
        const uint32_t bytes = (nb + 7) / 8;
        memcpy(p, &acc, bytes);
        p += bytes;
        acc = 0;
        nb = 0;
        return p;
    }
};
// l_seq BAM bases -> 2-bit codes appended to the stream
inline void put_seq_codes(BitWriter &w, const uint8_t *seq, uint32_t l_seq) {
    uint32_t j = 0;
    for (; j + 16 <= l_seq; j += 16) w.put(codes16(nibble_swap(ld64(seq + (j >> 1)))), 32);
    if (j < l_seq) {
        const uint32_t rem = l_seq - j;     // 1..15 bases; an odd length leaves a pad nibble that the mask drops
        w.put(codes16(nibble_swap(ld_tail(seq + (j >> 1), (rem + 1) / 2))) & ((1u << (2 * rem)) - 1u), 2 * rem);
    }
}
// l_seq qualities -> dictionary codes of `bits` bits appended to the stream
inline void put_qual_codes(BitWriter &w, const uint8_t *qual, uint32_t l_seq, const uint8_t *code, uint32_t bits) {
    uint32_t j = 0;
    for (; j + 8 <= l_seq; j += 8) {
        const uint8_t *q = qual + j;
        const uint64_t c = (uint64_t)code[q[0]] | ((uint64_t)code[q[1]] << 8) | ((uint64_t)code[q[2]] << 16) | ((uint64_t)code[q[3]] << 24) |
                           ((uint64_t)code[q[4]] << 32) | ((uint64_t)code[q[5]] << 40) | ((uint64_t)code[q[6]] << 48) | ((uint64_t)code[q[7]] << 56);
        w.put(squeeze8(c, bits), 8 * bits);
    }
    if (j < l_seq) {
        uint64_t c = 0;
        for (uint32_t k = 0; j + k < l_seq; k++) c |= (uint64_t)code[qual[j + k]] << (8 * k);
        w.put(squeeze8(c, bits), (l_seq - j) * bits);
    }
}

struct Owner {
    HostAlloc mem;
    std::vector<std::string> names;
    std::vector<char *> name_ptrs;
    std::vector<uint32_t> lengths;
    std::string header;
};

}  // namespace

namespace {
// an input file (BAM, FASTA) as one read-only span: mapped when the path can be mapped, read otherwise
struct FileSpan {
    const char *p = nullptr;
    size_t n = 0;
    void *map = nullptr;
    std::vector<char> buf;
    bool open(const char *path) {
        const int fd = ::open(path, O_RDONLY);
        if (fd < 0) return false;
        struct stat st;
        if (fstat(fd, &st) == 0 && S_ISREG(st.st_mode) && st.st_size > 0) {
            void *m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
            if (m != MAP_FAILED) {
                madvise(m, (size_t)st.st_size, MADV_SEQUENTIAL);
                map = m;
                p = (const char *)m;
                n = (size_t)st.st_size;
                ::close(fd);
                return true;
            }
        }
        bool ok = true;     // pipes, empty files, file systems without mmap
        char chunk[1 << 16];
        for (;;) {
            const ssize_t got = ::read(fd, chunk, sizeof chunk);
            if (got < 0) { ok = false; break; }
            if (got == 0) break;
            buf.insert(buf.end(), chunk, chunk + got);
        }
        ::close(fd);
        p = buf.data();
        n = buf.size();
        return ok;
    }
    const uint8_t *data() const { return (const uint8_t *)p; }
    size_t size() const { return n; }
    void release() {
        if (map) munmap(map, n);
        map = nullptr;
        std::vector<char>().swap(buf);
        p = nullptr;
    }
    ~FileSpan() {
        if (map) munmap(map, n);
    }
};
}  // namespace

int sd_bam_inflate_file(const char *path, int threads, RawBuf &raw, size_t *compressed_bytes, char *err, int errlen) {
    if (threads <= 0) threads = (int)std::max(1u, std::thread::hardware_concurrency());
    // ---- the file: mapped, so that the inflating threads read it straight from the page cache
    FileSpan file;
    if (!file.open(path)) return fail(err, errlen, SD_ERR_IO, "cannot open %s", path);

    // ---- BGZF block table
    struct Blk { size_t coff, clen, uoff, ulen; };
    std::vector<Blk> blks;
    size_t off = 0, utotal = 0;
    while (off < file.size()) {
        const uint8_t *h = file.data() + off;
        if (off + 18 > file.size() || h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4))
            return fail(err, errlen, SD_ERR_FORMAT, "%s: not a BGZF block at offset %zu", path, off);
        const uint32_t xlen = rd16(h + 10);
        int bsize = -1;
        size_t x = 12;
        while (x + 4 <= 12 + (size_t)xlen) {
            const uint32_t slen = rd16(h + x + 2);
            if (h[x] == 66 && h[x + 1] == 67 && slen == 2) bsize = rd16(h + x + 4);
            x += 4 + slen;
        }
        if (bsize < 0 || off + (size_t)bsize + 1 > file.size())
            return fail(err, errlen, SD_ERR_FORMAT, "%s: truncated BGZF block at offset %zu", path, off);
        const size_t total = (size_t)bsize + 1;
        const uint32_t isize = rd32(h + total - 4);
        blks.push_back({off + 12 + xlen, total - 12 - xlen - 8, utotal, isize});
        utotal += isize;
        off += total;
    }
    raw.reset(utotal);
    if (!raw.data()) return fail(err, errlen, SD_ERR_NOMEM, "out of host memory");
    std::atomic<int> bad(0);
    std::atomic<size_t> zlib_blocks(0);
    const char *env_inflate = getenv("SD_INFLATE");                 // SD_INFLATE=zlib: every block through zlib (A/B timing, tests)
    const bool use_own_inflate = !(env_inflate && strcmp(env_inflate, "zlib") == 0);
    parallel_for(blks.size(), threads, [&](size_t i) {
        const Blk &b = blks[i];
        if (b.ulen == 0) return;
        const uint32_t crc = rd32(file.data() + b.coff + b.clen);
        // the block decoder of sd_inflate.cpp first; zlib for whatever it turns down or gets wrong by the block's CRC
        if (use_own_inflate && sd_inflate_raw(file.data() + b.coff, b.clen, raw.data() + b.uoff, b.ulen) &&
            sd_crc32(raw.data() + b.uoff, b.ulen) == crc)
            return;
        z_stream zs;
        memset(&zs, 0, sizeof zs);
        if (inflateInit2(&zs, -15) != Z_OK) { bad = 1; return; }
        zs.next_in = const_cast<Bytef *>(file.data() + b.coff);
        zs.avail_in = (uInt)b.clen;
        zs.next_out = raw.data() + b.uoff;
        zs.avail_out = (uInt)b.ulen;
        const int rc = inflate(&zs, Z_FINISH);
        if (rc != Z_STREAM_END || zs.avail_out != 0) bad = 1;
        inflateEnd(&zs);
        if ((uint32_t)crc32(crc32(0L, Z_NULL, 0), raw.data() + b.uoff, (uInt)b.ulen) != crc) bad = 1;
        if (use_own_inflate) zlib_blocks.fetch_add(1, std::memory_order_relaxed);
    });
    if (bad) return fail(err, errlen, SD_ERR_FORMAT, "%s: corrupt BGZF block", path);
    if (getenv("SD_DECODE_TIMING")) fprintf(stderr, "[sd_bam_inflate_file] %zu blocks, %zu left to zlib\n", blks.size(), zlib_blocks.load());
    const size_t compressed = file.size();
    file.release();
    if (compressed_bytes) *compressed_bytes = compressed;
    return SD_OK;
}

int sd_bam_parse_header(const RawBuf &raw, const char *path, std::string &text, std::vector<std::string> &names,
                        std::vector<uint32_t> &lengths, size_t *after, char *err, int errlen) {
    if (raw.size() < 12 || memcmp(raw.data(), "BAM\1", 4) != 0) return fail(err, errlen, SD_ERR_FORMAT, "%s: not a BAM file", path);
    size_t p = 4;
    const int32_t l_text = rdi32(raw.data() + p);
    p += 4;
    if (l_text < 0 || p + (size_t)l_text + 4 > raw.size()) return fail(err, errlen, SD_ERR_FORMAT, "%s: bad header length", path);
    text.assign((const char *)raw.data() + p, strnlen((const char *)raw.data() + p, (size_t)l_text));
    p += (size_t)l_text;
    const int32_t n_ref = rdi32(raw.data() + p);
    p += 4;
    if (n_ref < 0) return fail(err, errlen, SD_ERR_FORMAT, "%s: bad reference count", path);
    names.clear();
    lengths.clear();
    for (int32_t i = 0; i < n_ref; i++) {
        if (p + 4 > raw.size()) return fail(err, errlen, SD_ERR_FORMAT, "%s: truncated reference table", path);
        const int32_t l_name = rdi32(raw.data() + p);
        if (l_name < 1 || p + 4 + (size_t)l_name + 4 > raw.size()) return fail(err, errlen, SD_ERR_FORMAT, "%s: truncated reference table", path);
        names.emplace_back((const char *)raw.data() + p + 4, (size_t)l_name - 1);
        lengths.push_back(rd32(raw.data() + p + 4 + l_name));
        p += 8 + (size_t)l_name;
    }
    *after = p;
    return SD_OK;
}

int sd_bam_index_records(const RawBuf &raw, size_t begin, int32_t n_ref, int threads, std::vector<size_t> &recoff,
                         const char *path, char *err, int errlen) {
    if (threads <= 0) threads = (int)std::max(1u, std::thread::hardware_concurrency());
    // BAM records carry no sync marker, so finding them is a dependent chain of block_size reads.  The chain is cut
    // into chunks: each chunk guesses its first record by a plausibility test that must hold for a run of consecutive
    // records, walks from there, and the pieces are accepted only if every chunk ends exactly where the next one
    // started -- which proves the guesses were on the true chain.  Otherwise the plain serial walk is used.
    const size_t rec_begin = begin;
    size_t p = begin;
    recoff.clear();
    {
        const uint8_t *R = raw.data();
        const size_t N = raw.size();
        auto plausible = [&](size_t at) -> long long {   // -> block_size if a record could start at `at`, else -1
            if (at + 36 > N) return -1;
            const int32_t bs = rdi32(R + at);
            if (bs < 34 || (size_t)bs > (64u << 20) || at + 4 + (size_t)bs > N) return -1;
            const uint8_t *r = R + at + 4;
            const int32_t ref_id = rdi32(r), pos = rdi32(r + 4), l_seq = rdi32(r + 16), nref = rdi32(r + 20);
            const uint32_t l_name = r[8], n_cig = rd16(r + 12);
            if (ref_id < -1 || ref_id >= n_ref || pos < -1 || l_seq < 0 || nref < -1 || nref >= n_ref || l_name < 1) return -1;
            const size_t need = 32 + (size_t)l_name + 4 * (size_t)n_cig + ((size_t)l_seq + 1) / 2 + (size_t)l_seq;
            if (need > (size_t)bs) return -1;
            if (r[32 + l_name - 1] != 0) return -1;
            for (uint32_t k = 0; k + 1 < l_name; k++)
                if (r[32 + k] < 33 || r[32 + k] > 126) return -1;
            for (uint32_t k = 0; k < n_cig; k++)
                if ((rd32(r + 32 + l_name + 4 * k) & 15u) > 8u) return -1;
            return bs;
        };
        const size_t span = N - rec_begin;
        const char *par_min_env = getenv("SD_DECODE_PAR_MIN");   // bytes of records below which the serial walk is used (tests lower it)
        const size_t par_min = par_min_env ? (size_t)atoll(par_min_env) : (size_t)(8u << 20);
        size_t n_chunks = (threads > 1 && span > par_min && span > 4096) ? (size_t)threads * 2 : 1;
        bool ok = n_chunks > 1;
        if (ok) {
            const size_t chunk = span / n_chunks;
            std::vector<size_t> first(n_chunks, 0), last(n_chunks, 0);
            std::vector<std::vector<size_t>> part(n_chunks);
            std::atomic<int> failed(0);
            parallel_for(n_chunks, threads, [&](size_t c) {
                const size_t lo = rec_begin + c * chunk, hi = (c + 1 == n_chunks) ? N : rec_begin + (c + 1) * chunk;
                size_t at = lo;
                if (c > 0) {   // search the first offset from which 6 consecutive records look right
                    bool found = false;
                    for (; at < hi && at + 36 <= N; at++) {
                        size_t q = at;
                        int good = 0;
                        while (good < 6) {
                            const long long bs = plausible(q);
                            if (bs < 0) break;
                            good++;
                            q += 4 + (size_t)bs;
                            if (q == N) { good = 6; break; }
                        }
                        if (good == 6) { found = true; break; }
                    }
                    if (!found) { failed = 1; return; }
                }
                first[c] = at;
                std::vector<size_t> &v = part[c];
                v.reserve((hi - lo) / 150 + 16);
                while (at < hi) {   // records that START inside this chunk
                    if (at + 4 > N) { failed = 1; return; }
                    const int32_t bs = rdi32(R + at);
                    if (bs < 32 || at + 4 + (size_t)bs > N) { failed = 1; return; }
                    v.push_back(at + 4);
                    at += 4 + (size_t)bs;
                }
                last[c] = at;
            });
            ok = !failed;
            for (size_t c = 0; ok && c + 1 < n_chunks; c++) ok = (last[c] == first[c + 1]);
            ok = ok && last[n_chunks - 1] == N;
            if (ok) {
                size_t total = 0;
                for (auto &v : part) total += v.size();
                recoff.reserve(total);
                for (auto &v : part) recoff.insert(recoff.end(), v.begin(), v.end());
            }
        }
        if (!ok) {
            recoff.clear();
            while (p + 4 <= N) {
                const int32_t bs = rdi32(R + p);
                if (bs < 32 || p + 4 + (size_t)bs > N) return fail(err, errlen, SD_ERR_FORMAT, "%s: truncated alignment record", path);
                recoff.push_back(p + 4);
                p += 4 + (size_t)bs;
            }
        }
    }
    return SD_OK;
}

extern "C" void sd_hbatch_free(sd_hbatch *hb) {
    if (!hb) return;
    delete (Owner *)hb->opaque;
    delete hb;
}

extern "C" int sd_decode_bam(const char *path, const char *const *fasta_names, uint32_t n_fasta, int options,
                             uint32_t tile_reads, int threads, sd_hbatch **out, char *err, int errlen) {
    if (!path || !out) return fail(err, errlen, SD_ERR_ARG, "sd_decode_bam: null argument");
    *out = nullptr;
    const bool mp_all = (options & SD_DECODE_MP_ALL) != 0;   // every MP entry, two words each (sd_read_stats)
    const bool want_mp = (options & SD_DECODE_MP) != 0 || mp_all, group = (options & SD_DECODE_GROUP) != 0;
    const uint64_t mp_entry_bytes = mp_all ? 8 : 4;
    const bool seq2 = (options & SD_DECODE_SEQ2) != 0;      // 2-bit SEQ column (count path only)
    const bool want_wire = (options & SD_DECODE_WIRE) != 0;  // also the packed form sd_count_wire sends over PCIe
    if (seq2 && mp_all) return fail(err, errlen, SD_ERR_ARG, "SD_DECODE_SEQ2 and SD_DECODE_MP_ALL exclude each other (sd_read_stats needs the 4-plane SEQ column)");
    const uint64_t group_bytes = seq2 ? 8 : 16;
    const bool raw_qual = (options & SD_DECODE_RAW_QUAL) != 0;
    if (tile_reads != 0 && tile_reads != 32) return fail(err, errlen, SD_ERR_ARG, "tile_reads must be 32 (or 0 = default)");
    tile_reads = 32;
    if (threads <= 0) threads = (int)std::max(1u, std::thread::hardware_concurrency());
    const double t0 = now_s();
    const bool timing = getenv("SD_DECODE_TIMING") != nullptr;
    double tp = t0;
    auto lap = [&](const char *what) {
        if (!timing) return;
        const double t = now_s();
        fprintf(stderr, "[sd_decode_bam] %-14s %.4f s\n", what, t - tp);
        tp = t;
    };

    Owner *own = new Owner();
    sd_hbatch *hb = new sd_hbatch();
    memset(hb, 0, sizeof *hb);
    hb->opaque = own;
    {   // pin the column arena while the file inflates: the columns take about as much room as three times the file
        FILE *probe = fopen(path, "rb");
        if (probe) {
            fseek(probe, 0, SEEK_END);
            const long fsz = ftell(probe);
            fclose(probe);
            if (fsz > 0) own->mem.reserve_async(3 * (size_t)fsz + (1u << 20));
        }
    }
    RawBuf raw;
    size_t compressed = 0;
    {
        const int rc = sd_bam_inflate_file(path, threads, raw, &compressed, err, errlen);
        if (rc != SD_OK) { sd_hbatch_free(hb); return rc; }
    }
    const double t1 = now_s();
    lap("inflate");

    // ---- BAM header
    auto bail = [&](int code, const char *msg) {
        sd_hbatch_free(hb);
        return fail(err, errlen, code, "%s: %s", path, msg);
    };
    size_t p = 0;
    {
        const int rc = sd_bam_parse_header(raw, path, own->header, own->names, own->lengths, &p, err, errlen);
        if (rc != SD_OK) { sd_hbatch_free(hb); return rc; }
    }
    const int32_t n_ref = (int32_t)own->names.size();
    std::unordered_map<std::string, uint32_t> fasta_idx;
    for (uint32_t i = 0; i < n_fasta; i++) fasta_idx.emplace(fasta_names[i], i);
    std::vector<uint32_t> ref_map((size_t)n_ref, SD_REF_NONE);
    for (int32_t i = 0; i < n_ref; i++) {
        if (n_fasta == 0) {   // no FASTA table: chromosome index = BAM reference id (the filter's coordinate space)
            if ((uint32_t)i < SD_REF_NONE) ref_map[(size_t)i] = (uint32_t)i;
        } else {
            auto it = fasta_idx.find(own->names[(size_t)i]);
            if (it != fasta_idx.end()) ref_map[(size_t)i] = it->second;
        }
    }
    for (auto &s_ : own->names) own->name_ptrs.push_back(const_cast<char *>(s_.c_str()));
    hb->header_text = const_cast<char *>(own->header.c_str());
    hb->n_ref = (uint32_t)n_ref;
    hb->ref_names = own->name_ptrs.data();
    hb->ref_lengths = own->lengths.data();

    // ---- record index
    std::vector<size_t> recoff;
    {
        const int rc = sd_bam_index_records(raw, p, n_ref, threads, recoff, path, err, errlen);
        if (rc != SD_OK) { sd_hbatch_free(hb); return rc; }
    }
    const size_t n = recoff.size();
    lap("record index");
    const size_t n_tiles = (n + tile_reads - 1) / tile_reads;
    if (n_tiles > 0xFFFFFFF0ull) return bail(SD_ERR_LIMIT, "too many records for one batch");

    sd_read_rec *rec = (sd_read_rec *)own->mem.get(std::max<size_t>(1, n_tiles) * tile_reads * sizeof(sd_read_rec));
    uint64_t *name_hash = (uint64_t *)own->mem.get(std::max<size_t>(1, n) * sizeof(uint64_t));
    sd_tile *tiles = (sd_tile *)own->mem.get((n_tiles + 1) * sizeof(sd_tile));
    if (!rec || !name_hash || !tiles) return bail(SD_ERR_NOMEM, "out of host memory");
    std::vector<const char *> mp_ptr(n, nullptr);
    std::atomic<int> fmt_err(0), limit_err(0);
    std::atomic<uint32_t> max_span(0);
    std::atomic<uint64_t> no_xi(0), no_nm(0);   // mapped records without the tag: read.get_tag raises KeyError there (filter.py:246,249)
    std::atomic<uint64_t> qual_seen[4];   // which of the 256 quality values occur (sizes the transfer coding of the column)
    for (auto &w : qual_seen) w = 0;

    lap("alloc");
    // ---- pass A: fixed-size fields, tags, sizes
    parallel_ranges(n, threads, [&](size_t range_b, size_t range_e) {
      uint8_t qual_tab[256];      // quality values met in this range
      memset(qual_tab, 0, sizeof qual_tab);
      for (size_t i = range_b; i < range_e; i++) {
        const uint8_t *r = raw.data() + recoff[i];
        const uint32_t bs = rd32(r - 4);
        const int32_t ref_id = rdi32(r), pos = rdi32(r + 4);
        const uint32_t l_name = r[8], mapq = r[9], n_cig = rd16(r + 12), flag = rd16(r + 14);
        const int32_t l_seq = rdi32(r + 16);
        const uint8_t *name = r + 32;
        const uint8_t *cig = name + l_name;
        const uint8_t *seq = cig + 4 * (size_t)n_cig;
        const uint8_t *qual = seq + ((size_t)(l_seq > 0 ? l_seq : 0) + 1) / 2;
        const uint8_t *aux = qual + (size_t)(l_seq > 0 ? l_seq : 0);
        const uint8_t *end = r + bs;
        if (l_seq < 0 || aux > end) { fmt_err = 1; return; }
        if (l_seq > 65535) { limit_err = 1; return; }
        AuxInfo a;
        if (!parse_aux(aux, end, a)) { fmt_err = 1; return; }
        uint32_t f = 0;
        if (flag & 0x10) f |= SD_F_REVERSE;
        if ((flag & 0x4) || n_cig == 0 || ref_id < 0) f |= SD_F_UNMAPPED;
        if (flag & 0x100) f |= SD_F_SECONDARY;
        if (flag & 0x800) f |= SD_F_SUPPLEMENTARY;
        if (a.has_mp) f |= SD_F_HAS_MP;
        if (!(f & SD_F_UNMAPPED)) {
            if (!a.has_xi) no_xi.fetch_add(1, std::memory_order_relaxed);
            if (!a.has_nm) no_nm.fetch_add(1, std::memory_order_relaxed);
        }
        uint32_t ref = SD_REF_NONE;
        if (ref_id >= 0 && ref_id < n_ref) ref = ref_map[(size_t)ref_id];
        if (ref == SD_REF_NONE) f |= SD_F_SKIP;
        // query-name run boundaries (multimapper grouping, dunks/filter.py:115)
        uint64_t h = 1469598103934665603ULL;
        for (uint32_t k = 0; k + 1 < l_name; k++) { h ^= name[k]; h *= 1099511628211ULL; }
        name_hash[i] = h;
        bool newname = true;
        if (i > 0) {
            const uint8_t *pr = raw.data() + recoff[i - 1];
            newname = !(pr[8] == l_name && memcmp(pr + 32, name, l_name) == 0);
        }
        if (newname) f |= SD_F_NEWNAME;
        uint32_t nmp = 0;
        if (want_mp && a.has_mp) {
            bool range_ok = true, ok;
            if (mp_all)
                ok = mp_for_all(a.mp, [&](long conv, long rp, long fp) {
                    nmp++;
                    if (rp < 0 || rp > 65535 || fp < 0 || fp > 65535 || conv < 0 || conv > 24) range_ok = false;
                });
            else
                ok = mp_for_each(a.mp, (flag & 0x10) ? 2 : 16, [&](long rp, long fp) {
                    nmp++;
                    if (rp < 0 || rp > 65535 || fp < 0 || fp > 65535) range_ok = false;
                });
            if (!ok) { fmt_err = 1; return; }
            if (!range_ok || nmp > 65535) { limit_err = 1; return; }
            mp_ptr[i] = a.mp;
        }
        if (!(f & (SD_F_UNMAPPED | SD_F_SKIP))) {   // reference span of the alignment: sizes the count kernel's register window
            uint32_t span = 0;
            for (uint32_t k = 0; k < n_cig; k++) {
                const uint32_t c = rd32(cig + 4 * (size_t)k);
                if ((0x18Du >> (c & 15u)) & 1u) span += c >> 4;
            }
            if (span <= (uint32_t)l_seq + 64u) {
                uint32_t cur = max_span.load(std::memory_order_relaxed);
                while (span > cur && !max_span.compare_exchange_weak(cur, span, std::memory_order_relaxed)) {}
            }
        }
        if (!raw_qual)
            for (int32_t k = 0; k < l_seq; k++) qual_tab[qual[k]] = 1;
        sd_read_rec &o = rec[i];
        o.pos = pos;
        o.ref_mapq_flag = (ref & 0xFFFFu) | (mapq << 16) | (f << 24);
        o.xi = a.xi;
        o.lseq_ncig = (uint32_t)l_seq | (n_cig << 16);
        o.nm_nmp = a.nm | (nmp << 16);
      }
      if (!raw_qual) {
          uint64_t seen[4] = {0, 0, 0, 0};
          for (int v = 0; v < 256; v++)
              if (qual_tab[v]) seen[v >> 6] |= 1ull << (v & 63);
          for (int w = 0; w < 4; w++)
              if (seen[w] & ~qual_seen[w].load(std::memory_order_relaxed)) qual_seen[w].fetch_or(seen[w], std::memory_order_relaxed);
      }
    });
    if (fmt_err) return bail(SD_ERR_FORMAT, "malformed alignment record or tag");
    if (limit_err) return bail(SD_ERR_LIMIT, "read longer than 65535 bases (or MP position out of range)");
    for (size_t i = n; i < n_tiles * tile_reads; i++) {
        rec[i].pos = 0;
        rec[i].ref_mapq_flag = (uint32_t)SD_REF_NONE | ((uint32_t)(SD_F_PAD | SD_F_SKIP | SD_F_UNMAPPED) << 24);
        rec[i].xi = 0.f;
        rec[i].lseq_ncig = 0;
        rec[i].nm_nmp = 0;
    }

    lap("pass A");
    // ---- optional grouping: reads that are one plain match block first (plus strand before minus strand), the
    // rest after, each group in file order -- counting is order independent, and homogeneous tiles let whole warps
    // take the kernel's lean path.  order[slot] = record index in the file.
    uint32_t *order = (uint32_t *)own->mem.get(std::max<size_t>(1, n) * sizeof(uint32_t));
    if (!order) return bail(SD_ERR_NOMEM, "out of host memory");
    for (size_t i = 0; i < n; i++) order[i] = (uint32_t)i;
    if (group && n > 0) {
        auto cls = [&](size_t i) -> uint32_t {
            const uint8_t *r = raw.data() + recoff[i];
            const uint32_t n_cig = rd16(r + 12), flag = rd16(r + 14);
            bool simple = false;
            if (n_cig == 1) {
                const uint32_t op = rd32(r + 32 + r[8]) & 15u;
                simple = (op == 0 || op == 7 || op == 8);
            }
            return (simple ? 0u : 2u) | ((flag & 0x10) ? 1u : 0u);
        };
        // stable counting sort over four classes, chunked so that the class scan and the scatter run on all threads
        const size_t n_chunks = (size_t)std::max(1, threads) * 4;
        const size_t chunk = (n + n_chunks - 1) / n_chunks;
        std::vector<size_t> cnt(n_chunks * 4, 0);
        std::vector<uint8_t> cl(n);
        parallel_for(n_chunks, threads, [&](size_t c) {
            const size_t b = c * chunk, e = std::min(n, b + chunk);
            for (size_t i = b; i < e; i++) {
                cl[i] = (uint8_t)cls(i);
                cnt[c * 4 + cl[i]]++;
            }
        });
        std::vector<size_t> start(n_chunks * 4, 0);
        size_t run = 0;
        for (int k = 0; k < 4; k++)
            for (size_t c = 0; c < n_chunks; c++) {
                start[c * 4 + k] = run;
                run += cnt[c * 4 + k];
            }
        parallel_for(n_chunks, threads, [&](size_t c) {
            const size_t b = c * chunk, e = std::min(n, b + chunk);
            size_t pos[4] = {start[c * 4 + 0], start[c * 4 + 1], start[c * 4 + 2], start[c * 4 + 3]};
            for (size_t i = b; i < e; i++) order[pos[cl[i]]++] = (uint32_t)i;
        });
        RawBuf tmp_buf(n * sizeof(sd_read_rec));
        sd_read_rec *tmp = (sd_read_rec *)tmp_buf.data();
        if (!tmp) return bail(SD_ERR_NOMEM, "out of host memory");
        parallel_ranges(n, threads, [&](size_t b, size_t e) { memcpy(tmp + b, rec + b, (e - b) * sizeof(sd_read_rec)); });
        parallel_for(n, threads, [&](size_t k) { rec[k] = tmp[order[k]]; });
    }

    lap("grouping");
    // ---- tile offsets
    auto up16 = [](uint64_t x) { return (x + 15ull) & ~15ull; };
    uint32_t max_lseq = 0, mt_seq = 0, mt_qual = 0, mt_cig = 0, mt_mp = 0;
    {
        // sizes of every tile's blocks on all threads (tiles[t] holds sizes for a moment), then one running sum over the tiles
        std::vector<uint32_t> tile_lseq_of(n_tiles, 0);
        parallel_for(n_tiles, threads, [&](size_t t) {
            uint64_t bq = 0, bc = 0, bm = 0;
            uint32_t tile_lseq = 0;
            const size_t e = std::min(n, (t + 1) * (size_t)tile_reads);
            for (size_t i = t * (size_t)tile_reads; i < e; i++) {
                const uint32_t ls = rec[i].lseq_ncig & 0xFFFFu;
                tile_lseq = std::max(tile_lseq, ls);
                bq += ls;
                bc += 4ull * (rec[i].lseq_ncig >> 16);
                bm += mp_entry_bytes * (rec[i].nm_nmp >> 16);
            }
            tile_lseq_of[t] = tile_lseq;
            // group-major SEQ block: every row slot of the tile, the longest read's groups
            tiles[t] = {group_bytes * tile_reads * ((tile_lseq + 31u) >> 5), up16(bq), up16(bc), up16(bm)};
        });
        sd_tile cur = {0, 0, 0, 0};
        for (size_t t = 0; t < n_tiles; t++) {
            const sd_tile sz = tiles[t];
            tiles[t] = cur;
            max_lseq = std::max(max_lseq, tile_lseq_of[t]);
            mt_seq = std::max<uint32_t>(mt_seq, (uint32_t)sz.seq_off);
            mt_qual = std::max<uint32_t>(mt_qual, (uint32_t)sz.qual_off);
            mt_cig = std::max<uint32_t>(mt_cig, (uint32_t)sz.cig_off);
            mt_mp = std::max<uint32_t>(mt_mp, (uint32_t)sz.mp_off);
            cur.seq_off += sz.seq_off; cur.qual_off += sz.qual_off; cur.cig_off += sz.cig_off; cur.mp_off += sz.mp_off;
        }
        tiles[n_tiles] = cur;
    }
    const sd_tile tot = tiles[n_tiles];
    uint32_t *seqc = (uint32_t *)own->mem.get(tot.seq_off + 16);
    // transfer coding of the base qualities: 2- or 4-bit dictionary codes when the file has few distinct values
    uint32_t qual_bits = 8;
    uint8_t qual_dict[16] = {0}, qual_code[256] = {0};
    if (!raw_qual) {
        int distinct = 0;
        for (int v = 0; v < 256; v++)
            if ((qual_seen[v >> 6].load() >> (v & 63)) & 1ull) {
                if (distinct < 16) { qual_dict[distinct] = (uint8_t)v; qual_code[v] = (uint8_t)distinct; }
                distinct++;
            }
        if (distinct <= 4) qual_bits = 2;
        else if (distinct <= 16) qual_bits = 4;
    }
    uint8_t *qualc = (uint8_t *)own->mem.get(tot.qual_off * qual_bits / 8 + 16);
    uint32_t *cigc = (uint32_t *)own->mem.get(tot.cig_off + 16);
    uint32_t *mpc = want_mp ? (uint32_t *)own->mem.get(tot.mp_off + 16) : nullptr;
    if (!seqc || !qualc || !cigc || (want_mp && !mpc)) return bail(SD_ERR_NOMEM, "out of host memory");

    lap("tiles+alloc");
    // ---- pass B: variable-length columns, one tile at a time
    parallel_for(n_tiles, threads, [&](size_t t) {
        uint8_t *ps = (uint8_t *)seqc + tiles[t].seq_off;   // group-major block: slot (g, row) at 16 * (g * tile_reads + row)
        memset(ps, 0, (size_t)(tiles[t + 1].seq_off - tiles[t].seq_off));
        uint8_t *pq = qualc + tiles[t].qual_off * qual_bits / 8;   // tile offsets are in bases, a multiple of 16: byte aligned at any coding
        BitWriter wq(pq);                                           // coded form: one bit stream per tile
        uint8_t *pc = (uint8_t *)cigc + tiles[t].cig_off;
        uint8_t *pm = mpc ? (uint8_t *)mpc + tiles[t].mp_off : nullptr;
        const size_t e = std::min(n, (t + 1) * (size_t)tile_reads);
        for (size_t k = t * (size_t)tile_reads; k < e; k++) {
            const size_t i = order[k];
            const uint8_t *r = raw.data() + recoff[i];
            const uint32_t l_name = r[8], n_cig = rd16(r + 12), flag = rd16(r + 14);
            const uint32_t l_seq = rd32(r + 16);
            const uint8_t *cig = r + 32 + l_name;
            const uint8_t *seq = cig + 4 * (size_t)n_cig;
            const uint8_t *qual = seq + ((size_t)l_seq + 1) / 2;
            memcpy(pc, cig, 4 * (size_t)n_cig);
            pc += 4 * (size_t)n_cig;
            if (qual_bits == 8) {
                memcpy(wq.p, qual, l_seq);
                wq.p += l_seq;
            } else {
                put_qual_codes(wq, qual, l_seq, qual_code, qual_bits);
            }
            const uint32_t nbytes = (l_seq + 1) / 2;
            const size_t row = k - t * (size_t)tile_reads;
            for (uint32_t g = 0; g < (l_seq + 31) / 32; g++) {   // one group = 32 bases = 16 packed bytes -> 2 code words / 4 plane words
                const uint32_t have = std::min(16u, nbytes - 16 * g);          // bytes of this group present in the record
                uint64_t x[2];
                if (have == 16) {
                    x[0] = ld64(seq + 16 * g);
                    x[1] = ld64(seq + 16 * g + 8);
                } else {
                    x[0] = ld_tail(seq + 16 * g, std::min(8u, have));
                    x[1] = have > 8 ? ld_tail(seq + 16 * g + 8, have - 8) : 0;
                }
                x[0] = nibble_swap(x[0]);
                x[1] = nibble_swap(x[1]);
                if (32 * g + 2 * have > l_seq) {   // odd length: ignore the pad nibble (the last one present)
                    const uint32_t nib = 2 * have - 1;
                    x[nib >> 4] &= ~(0xFULL << (4 * (nib & 15)));
                }
                if (seq2) {   // A=0 C=1 G=2 T=3, anything else A
                    uint64_t lo0, hi0, lo1, hi1;
                    code_flags(x[0], lo0, hi0);
                    code_flags(x[1], lo1, hi1);
                    const uint32_t two[2] = {gather_every_4th(lo0) | (gather_every_4th(lo1) << 16),
                                             gather_every_4th(hi0) | (gather_every_4th(hi1) << 16)};
                    memcpy(ps + 8 * ((size_t)g * tile_reads + row), two, 8);
                } else {      // bit-planes of the BAM nibble
                    uint32_t pl[4];
                    for (int b = 0; b < 4; b++) pl[b] = gather_every_4th(x[0] >> b) | (gather_every_4th(x[1] >> b) << 16);
                    memcpy(ps + 16 * ((size_t)g * tile_reads + row), pl, 16);
                }
            }
            if (pm && mp_ptr[i] && mp_all) {
                mp_for_all(mp_ptr[i], [&](long conv, long rp, long fp) {
                    const uint32_t v[2] = {(uint32_t)rp | ((uint32_t)fp << 16), (uint32_t)conv};
                    memcpy(pm, v, 8);
                    pm += 8;
                });
            } else if (pm && mp_ptr[i]) {
                mp_for_each(mp_ptr[i], (flag & 0x10) ? 2 : 16, [&](long rp, long fp) {
                    const uint32_t v = (uint32_t)rp | ((uint32_t)fp << 16);
                    memcpy(pm, &v, 4);
                    pm += 4;
                });
            }
        }
        // zero the 16-byte tile padding so the columns are deterministic
        pq = wq.finish();
        memset(pq, 0, (size_t)((qualc + tiles[t + 1].qual_off * qual_bits / 8) - pq));
        memset(pc, 0, (size_t)(((uint8_t *)cigc + tiles[t + 1].cig_off) - pc));
        if (pm) memset(pm, 0, (size_t)(((uint8_t *)mpc + tiles[t + 1].mp_off) - pm));
    });
    lap("pass B");

    // ---- optional: the wire form of the same batch (sd_wire, include/slamdunk_b200.h) -- every record field in its own column,
    // SEQ and the base qualities as back-to-back rows of 2-bit base codes / dictionary codes straight from the BAM records, no
    // l_seq / CIGAR for tiles of plain-match reads of one length.  Same layout as sd_wire_plan + sd_wire_fill make on the device.
    memset(&hb->wire, 0, sizeof hb->wire);
    if (want_wire && n_tiles > 0) {
        if (mp_all) return bail(SD_ERR_ARG, "SD_DECODE_WIRE and SD_DECODE_MP_ALL exclude each other");
        const size_t slots = n_tiles * (size_t)tile_reads;
        sd_wire_tile *wt = (sd_wire_tile *)own->mem.get((n_tiles + 1) * sizeof(sd_wire_tile));
        uint16_t *pos16 = (uint16_t *)own->mem.get(slots * 2), *mqf = (uint16_t *)own->mem.get(slots * 2);
        uint16_t *nm16 = (uint16_t *)own->mem.get(slots * 2), *nmp16 = want_mp ? (uint16_t *)own->mem.get(slots * 2) : nullptr;
        if (!wt || !pos16 || !mqf || !nm16 || (want_mp && !nmp16)) return bail(SD_ERR_NOMEM, "out of host memory");
        // tile shapes and sizes
        std::atomic<uint32_t> w_max_cig(0);
        parallel_for(n_tiles, threads, [&](size_t t) {
            const sd_read_rec *r = rec + t * (size_t)tile_reads;
            const uint32_t *cg = cigc + tiles[t].cig_off / 4;
            uint32_t lseq_u = 0, ref_u = SD_REF_NONE, flags = 0, sum_c = 0, bases = 0;
            int32_t pmin = 0x7FFFFFFF, pmax = (int32_t)0x80000000;
            bool first = true, plain = true, same_ref = true;
            for (uint32_t k = 0; k < tile_reads; k++) {
                const uint32_t fl = r[k].ref_mapq_flag >> 24, ref = r[k].ref_mapq_flag & 0xFFFFu;
                const uint32_t lseq = r[k].lseq_ncig & 0xFFFFu, ncig = r[k].lseq_ncig >> 16;
                if (!(fl & SD_F_PAD)) {
                    if (first) { lseq_u = lseq; ref_u = ref; first = false; }
                    if (!(lseq == lseq_u && ncig == 1u && cg[sum_c] == (lseq << 4))) plain = false;
                    if (ref != ref_u) same_ref = false;
                    pmin = std::min(pmin, r[k].pos);
                    pmax = std::max(pmax, r[k].pos);
                    bases += lseq;
                }
                sum_c += ncig;
            }
            if (!plain) flags |= SD_WT_SHAPE;
            if (!same_ref) flags |= SD_WT_REFS;
            if (!first && (long long)pmax - (long long)pmin > 65535ll) flags |= SD_WT_POS32;
            uint32_t var_words = ((flags & SD_WT_POS32) ? 32u : 0u) + ((flags & SD_WT_REFS) ? 16u : 0u) + ((flags & SD_WT_SHAPE) ? 32u + sum_c : 0u);
            var_words = (var_words + 3u) & ~3u;
            sd_wire_tile w;
            memset(&w, 0, sizeof w);
            w.base_off = ((uint64_t)bases + 63ull) & ~63ull;      // sizes for now, offsets after the scan below
            w.var_off = var_words;
            w.mp_off = want_mp ? (uint32_t)((tiles[t + 1].mp_off - tiles[t].mp_off) / 4u) : 0u;
            w.base_pos = first ? 0 : pmin;
            w.ref = (uint16_t)ref_u;
            w.flags = (uint16_t)flags;
            w.lseq = (uint16_t)lseq_u;
            wt[t] = w;
            if (flags & SD_WT_SHAPE) {
                const uint32_t cb = (4u * sum_c + 15u) & ~15u;
                uint32_t cur = w_max_cig.load(std::memory_order_relaxed);
                while (cb > cur && !w_max_cig.compare_exchange_weak(cur, cb, std::memory_order_relaxed)) {}
            }
        });
        lap("wire:shapes");
        uint64_t run_b = 0, run_v = 0, run_m = 0;
        for (size_t t = 0; t < n_tiles; t++) {
            const uint64_t nb = wt[t].base_off, nv = wt[t].var_off, nm_ = wt[t].mp_off;
            wt[t].base_off = run_b; wt[t].var_off = (uint32_t)run_v; wt[t].mp_off = (uint32_t)run_m;
            run_b += nb; run_v += nv; run_m += nm_;
        }
        if (run_v >> 32 || run_m >> 32) return bail(SD_ERR_LIMIT, "batch too large for the wire form's 32-bit word offsets");
        memset(&wt[n_tiles], 0, sizeof(sd_wire_tile));
        wt[n_tiles].base_off = run_b; wt[n_tiles].var_off = (uint32_t)run_v; wt[n_tiles].mp_off = (uint32_t)run_m;
        uint8_t *wseq = (uint8_t *)own->mem.get(run_b / 4 + 16), *wqual = (uint8_t *)own->mem.get(run_b * qual_bits / 8 + 16);
        uint32_t *wvar = (uint32_t *)own->mem.get((run_v + 4) * 4);
        if (!wseq || !wqual || !wvar) return bail(SD_ERR_NOMEM, "out of host memory");
        // XI dictionary: a file has a few thousand distinct identities (1 - mismatches / length); beyond 65536 the floats travel
        std::vector<uint32_t> xi_vals;
        bool xi_small = true;
        {
            const size_t n_chunks = (size_t)std::max(1, threads) * 4, chunk = (n + n_chunks - 1) / n_chunks;
            std::vector<std::vector<uint32_t>> part(n_chunks);
            std::atomic<int> big(0);
            parallel_for(n_chunks, threads, [&](size_t c) {
                const size_t b = c * chunk, e = std::min(n, b + chunk);
                std::vector<uint32_t> &v = part[c];
                for (size_t i = b; i < e && !big.load(std::memory_order_relaxed); i++) {
                    uint32_t x;
                    memcpy(&x, &rec[i].xi, 4);
                    v.push_back(x);
                    if (v.size() >= (1u << 20)) {
                        std::sort(v.begin(), v.end());
                        v.erase(std::unique(v.begin(), v.end()), v.end());
                        if (v.size() > 65536) big = 1;
                    }
                }
                std::sort(v.begin(), v.end());
                v.erase(std::unique(v.begin(), v.end()), v.end());
            });
            if (!big) {
                for (auto &v : part) xi_vals.insert(xi_vals.end(), v.begin(), v.end());
                std::sort(xi_vals.begin(), xi_vals.end());
                xi_vals.erase(std::unique(xi_vals.begin(), xi_vals.end()), xi_vals.end());
            }
            { uint32_t zero = 0; if (!big && !std::binary_search(xi_vals.begin(), xi_vals.end(), zero)) xi_vals.insert(std::lower_bound(xi_vals.begin(), xi_vals.end(), zero), zero); }
            xi_small = !big && xi_vals.size() <= 65536;
        }
        lap("wire:xi dict");
        uint16_t *xi16 = xi_small ? (uint16_t *)own->mem.get(slots * 2) : nullptr;
        float *xi32 = xi_small ? nullptr : (float *)own->mem.get(slots * 4);
        float *xi_dict = xi_small ? (float *)own->mem.get(std::max<size_t>(1, xi_vals.size()) * 4) : nullptr;
        if ((xi_small && (!xi16 || !xi_dict)) || (!xi_small && !xi32)) return bail(SD_ERR_NOMEM, "out of host memory");
        if (xi_small) memcpy(xi_dict, xi_vals.data(), xi_vals.size() * 4);
        // columns, one tile per task (rows of a tile share bytes of the per-base columns, tiles do not: they start at multiples of 64 bases)
        parallel_for(n_tiles, threads, [&](size_t t) {
            const sd_read_rec *r = rec + t * (size_t)tile_reads;
            const sd_wire_tile h = wt[t];
            const size_t s0 = t * (size_t)tile_reads;
            uint32_t v = h.var_off;
            if (h.flags & SD_WT_POS32) v += 32u;
            uint16_t *refs = nullptr;
            if (h.flags & SD_WT_REFS) { refs = (uint16_t *)(wvar + v); v += 16u; }
            uint32_t *shape = nullptr, *ops = nullptr;
            if (h.flags & SD_WT_SHAPE) { shape = wvar + v; v += 32u; ops = wvar + v; }
            const uint32_t *cg = cigc + tiles[t].cig_off / 4;
            uint32_t sum_c = 0;
            // bit writers of the two per-base columns; the tile starts byte aligned
            BitWriter ws(wseq + h.base_off / 4);
            uint64_t tile_bases = 0;
            for (uint32_t k = 0; k < tile_reads; k++) {
                const uint32_t fl = r[k].ref_mapq_flag >> 24, ref = r[k].ref_mapq_flag & 0xFFFFu, mapq = (r[k].ref_mapq_flag >> 16) & 0xFFu;
                const uint32_t lseq = r[k].lseq_ncig & 0xFFFFu, ncig = r[k].lseq_ncig >> 16;
                const bool pad = (fl & SD_F_PAD) != 0;
                pos16[s0 + k] = (pad || (h.flags & SD_WT_POS32)) ? (uint16_t)0 : (uint16_t)(r[k].pos - h.base_pos);
                mqf[s0 + k] = (uint16_t)(mapq | (fl << 8));
                uint32_t xb;
                memcpy(&xb, &r[k].xi, 4);
                if (xi16) xi16[s0 + k] = (uint16_t)(std::lower_bound(xi_vals.begin(), xi_vals.end(), pad ? 0u : xb) - xi_vals.begin());
                else xi32[s0 + k] = r[k].xi;
                nm16[s0 + k] = (uint16_t)(r[k].nm_nmp & 0xFFFFu);
                if (nmp16) nmp16[s0 + k] = (uint16_t)(r[k].nm_nmp >> 16);
                if (h.flags & SD_WT_POS32) wvar[h.var_off + k] = (uint32_t)r[k].pos;
                if (refs) refs[k] = (uint16_t)ref;
                if (shape) {
                    shape[k] = pad ? 0u : r[k].lseq_ncig;
                    for (uint32_t i = 0; i < ncig; i++) ops[sum_c + i] = cg[sum_c + i];
                }
                sum_c += ncig;
                if (pad) continue;
                const size_t i = order[s0 + k];
                const uint8_t *rr = raw.data() + recoff[i];
                const uint8_t *seq = rr + 32 + rr[8] + 4 * (size_t)rd16(rr + 12);
                put_seq_codes(ws, seq, lseq);       // A=0 C=1 G=2 T=3, anything else A
                tile_bases += lseq;
            }
            // the tile's qualities: the batch column already holds them as the same back-to-back stream of bytes / dictionary
            // codes (pass B), only the tile starts differ (multiples of 16 bases there, of 64 here) -- both byte aligned
            const size_t qual_bytes = (size_t)((tile_bases * qual_bits + 7) / 8);
            uint8_t *pq = wqual + h.base_off * qual_bits / 8;
            memcpy(pq, qualc + tiles[t].qual_off * qual_bits / 8, qual_bytes);
            pq += qual_bytes;
            // flush the writer and zero the padding up to the next tile (64 bases = 16 / 16 / 32 / 64 bytes)
            const uint8_t *es = wseq + wt[t + 1].base_off / 4, *eq = wqual + wt[t + 1].base_off * qual_bits / 8;
            uint8_t *ps = ws.finish();
            memset(ps, 0, (size_t)(es - ps));
            memset(pq, 0, (size_t)(eq - pq));
            if (shape)      // pad words of the var block
                for (uint32_t i = (uint32_t)(ops - wvar) + sum_c; i < wt[t + 1].var_off; i++) wvar[i] = 0u;
        });
        memset(wseq + run_b / 4, 0, 16);
        memset(wqual + run_b * qual_bits / 8, 0, 16);
        memset(wvar + run_v, 0, 16);
        sd_wire &W = hb->wire;
        W.n_reads = n;
        W.n_tiles = (uint32_t)n_tiles;
        W.tile_reads = tile_reads;
        W.tiles = wt;
        W.pos16 = pos16; W.mqf = mqf; W.xi16 = xi16; W.nm16 = nm16; W.nmp16 = nmp16;
        W.xi32 = xi32;
        W.seq = wseq; W.qual = wqual;
        W.var = wvar;
        W.mp = mpc;     // same blocks as the batch's mp column: the tile offsets are its tile table's, in words
        W.xi_dict = xi_dict;
        W.n_xi_dict = xi_small ? (uint32_t)xi_vals.size() : 0u;
        W.qual_bits = qual_bits;
        memcpy(W.qual_dict, qual_dict, 16);
        W.max_lseq = max_lseq;
        W.max_span = max_span.load();
        W.max_tile_cig = w_max_cig.load();
        W.max_tile_mp = want_mp ? mt_mp : 0u;
        lap("wire");
    }

    hb->batch.n_reads = n;
    hb->batch.n_tiles = (uint32_t)n_tiles;
    hb->batch.tile_reads = tile_reads;
    hb->batch.rec = rec;
    hb->batch.seq = seqc;
    hb->batch.qual = qualc;
    hb->batch.cigar = cigc;
    hb->batch.mp = mpc;
    hb->batch.tiles = tiles;
    hb->batch.max_lseq = max_lseq;
    hb->batch.max_tile_seq = mt_seq;
    hb->batch.max_tile_qual = mt_qual;
    hb->batch.max_tile_cig = mt_cig;
    hb->batch.max_tile_mp = mt_mp;
    hb->batch.max_span = max_span.load();
    hb->batch.qual_bits = qual_bits;
    hb->batch.seq_bits = seq2 ? 2 : 4;
    memcpy(hb->batch.qual_dict, qual_dict, 16);
    hb->name_hash = name_hash;
    hb->order = order;
    hb->seconds_inflate = t1 - t0;
    hb->seconds_decode = now_s() - t1;      // record decode incl. the wire form when it was asked for
    hb->compressed_bytes = compressed;
    hb->inflated_bytes = raw.size();
    hb->n_mapped_no_xi = no_xi.load();
    hb->n_mapped_no_nm = no_nm.load();
    *out = hb;
    return SD_OK;
}

// ------------------------------------------------------------------------------------------
// FASTA -> bit-planes
// ------------------------------------------------------------------------------------------
namespace {
struct GenomeOwner {
    HostAlloc mem;
    std::vector<std::string> names;
    std::vector<char *> name_ptrs;
    std::vector<uint32_t> len, off;
};
}  // namespace

extern "C" void sd_hgenome_free(sd_hgenome *g) {
    if (!g) return;
    delete (GenomeOwner *)g->opaque;
    delete g;
}

namespace {
// byte -> 0..3 = A C G T (either case, tcounter.py:181 upper-cases the slice), 4 = any other base letter (N, IUPAC, ...),
// 5 = white space inside sequence lines (skipped)
struct BaseLut {
    uint8_t v[256];
    BaseLut() {
        memset(v, 4, sizeof v);
        v[(uint8_t)'\n'] = v[(uint8_t)'\r'] = v[(uint8_t)' '] = v[(uint8_t)'\t'] = 5;
        const char *acgt = "ACGT";
        for (int k = 0; k < 4; k++) v[(uint8_t)acgt[k]] = v[(uint8_t)(acgt[k] | 0x20)] = (uint8_t)k;
    }
};
const BaseLut kBase;
}  // namespace

extern "C" int sd_load_fasta(const char *path, int threads, sd_hgenome **out, char *err, int errlen) {
    if (!path || !out) return fail(err, errlen, SD_ERR_ARG, "sd_load_fasta: null argument");
    *out = nullptr;
    if (threads <= 0) threads = (int)std::max(1u, std::thread::hardware_concurrency());
    const double t0 = now_s();
    FileSpan file;
    if (!file.open(path)) return fail(err, errlen, SD_ERR_IO, "cannot read %s", path);
    const char *txt = file.p;
    const size_t n = file.n;

    // pass 1: the header lines ('>' at the start of a line); a record's body runs to the next header
    struct Rec { size_t body, end; uint64_t len; };
    std::vector<Rec> recs;
    GenomeOwner *own = new GenomeOwner();
    for (size_t i = 0; i < n;) {
        const char *gt = (const char *)memchr(txt + i, '>', n - i);
        if (!gt) break;
        const size_t k = (size_t)(gt - txt);
        if (k != 0 && txt[k - 1] != '\n') {     // a '>' inside a line (descriptions may hold them)
            i = k + 1;
            continue;
        }
        const char *nl = (const char *)memchr(gt, '\n', n - k);
        const size_t e = nl ? (size_t)(nl - txt) : n;
        size_t ne = k + 1;
        while (ne < e && txt[ne] != ' ' && txt[ne] != '\t' && txt[ne] != '\r') ne++;
        own->names.emplace_back(txt + k + 1, ne - k - 1);
        if (!recs.empty()) recs.back().end = k;
        recs.push_back({std::min(n, e + 1), n, 0});
        i = e + 1;
    }
    if (recs.empty()) { delete own; return fail(err, errlen, SD_ERR_FORMAT, "%s: no FASTA records", path); }

    // the bodies cut into pieces of about 1 MB of text, so that one long chromosome is spread over all threads
    struct Piece { uint32_t rec; size_t b, e; uint64_t bases, first; };
    std::vector<Piece> pieces;
    const size_t piece_bytes = 1u << 20;
    for (size_t r = 0; r < recs.size(); r++)
        for (size_t b = recs[r].body; b < recs[r].end || b == recs[r].body; b += piece_bytes) {
            pieces.push_back({(uint32_t)r, b, std::min(recs[r].end, b + piece_bytes), 0, 0});
            if (b + piece_bytes >= recs[r].end) break;
        }
    parallel_for(pieces.size(), threads, [&](size_t k) {
        uint64_t L = 0;
        const char *q = txt + pieces[k].b, *qe = txt + pieces[k].e;
        for (; q < qe; q++) L += (uint64_t)((*q != '\n') & (*q != '\r') & (*q != ' ') & (*q != '\t'));
        pieces[k].bases = L;
    });
    for (auto &pc : pieces) {
        pc.first = recs[pc.rec].len;
        recs[pc.rec].len += pc.bases;
    }
    uint64_t bits = 32;  // leading pad so that no chromosome starts at linear position 0
    for (auto &r : recs) {
        if (r.len > 0xFFFFFFF0ull) { delete own; return fail(err, errlen, SD_ERR_LIMIT, "chromosome too long"); }
        own->off.push_back((uint32_t)bits);
        own->len.push_back((uint32_t)r.len);
        bits = ((bits + r.len + 31ull) & ~31ull) + 32ull;
        if (bits + 4096ull > 0xFFFFFFFFull) { delete own; return fail(err, errlen, SD_ERR_LIMIT, "genome exceeds the 32-bit linear coordinate space"); }
    }
    const uint64_t n_words = bits / 32;
    uint32_t *lo = (uint32_t *)own->mem.get(n_words * 4), *hi = (uint32_t *)own->mem.get(n_words * 4),
             *nm = (uint32_t *)own->mem.get(n_words * 4);
    if (!lo || !hi || !nm) { delete own; return fail(err, errlen, SD_ERR_NOMEM, "out of host memory"); }
    parallel_ranges((size_t)n_words, threads, [&](size_t b, size_t e) {
        memset(lo + b, 0, (e - b) * 4);
        memset(hi + b, 0, (e - b) * 4);
        memset(nm + b, 0xFF, (e - b) * 4);     // N / pad until a base says otherwise
    });
    // pass 2: encode.  A piece starts at base `first` of its chromosome, in general in the middle of a 32-base word: the
    // word a piece shares with its neighbour is merged with atomic OR / AND, every other word is written whole.
    parallel_for(pieces.size(), threads, [&](size_t k) {
        const Piece &pc = pieces[k];
        if (!pc.bases) return;
        const uint64_t g0 = (uint64_t)own->off[pc.rec] + pc.first;
        uint64_t word = g0 >> 5;
        uint32_t nb = (uint32_t)(g0 & 31u);
        uint32_t wl = 0, wh = 0, wa = 0;        // low / high bit of the base code, "is one of ACGT"
        bool shared = nb != 0;                  // the first word may hold the previous piece's last bases
        auto flush = [&](bool merge) {
            if (merge) {
                __atomic_fetch_or(&lo[word], wl, __ATOMIC_RELAXED);
                __atomic_fetch_or(&hi[word], wh, __ATOMIC_RELAXED);
                __atomic_fetch_and(&nm[word], ~wa, __ATOMIC_RELAXED);
            } else {
                lo[word] = wl;
                hi[word] = wh;
                nm[word] = ~wa;
            }
            wl = wh = wa = 0;
            nb = 0;
            word++;
            shared = false;
        };
        const uint8_t *q = (const uint8_t *)txt + pc.b, *qe = (const uint8_t *)txt + pc.e;
        for (; q < qe; q++) {
            const uint32_t v = kBase.v[*q];
            if (v == 5) continue;
            wl |= (v & 1u) << nb;
            wh |= ((v >> 1) & 1u) << nb;
            wa |= (uint32_t)(v < 4) << nb;
            if (++nb == 32) flush(shared);
        }
        if (nb) flush(true);                    // the last, partly filled word: the next piece (or nobody) has the rest
    });
    for (auto &s : own->names) own->name_ptrs.push_back(const_cast<char *>(s.c_str()));
    sd_hgenome *G = new sd_hgenome();
    memset(G, 0, sizeof *G);
    G->n_chrom = (uint32_t)recs.size();
    G->names = own->name_ptrs.data();
    G->chrom_len = own->len.data();
    G->chrom_off = own->off.data();
    G->n_words = n_words;
    G->lo = lo; G->hi = hi; G->nmask = nm;
    G->seconds = now_s() - t0;
    G->opaque = own;
    *out = G;
    return SD_OK;
}
