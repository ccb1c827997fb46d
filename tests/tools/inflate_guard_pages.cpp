// sd_inflate_raw with its input and output placed right before PROT_NONE pages: round trips, then 60 000 truncated / bit-flipped / wrongly sized streams -- any access past either buffer faults
// (test infrastructure; g++ -O2 -std=c++17 -Iinclude tests/tools/inflate_guard_pages.cpp slamdunk_b200/csrc/sd_inflate.cpp slamdunk_b200/csrc/sd_deflate.cpp -lz)
// input and output placed right before PROT_NONE pages: any read/write past either buffer faults
#include <sys/mman.h>
#include <zlib.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>
#include "slamdunk_b200.h"
struct Guarded { uint8_t *base; size_t pages; uint8_t *end;
    Guarded(size_t cap) { pages = (cap + 4095) / 4096 + 1; base = (uint8_t *)mmap(0, (pages + 1) * 4096, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        mprotect(base + pages * 4096, 4096, PROT_NONE); end = base + pages * 4096; }
    uint8_t *place(size_t n) { return end - n; } };
int main() {
    std::mt19937_64 rng(3);
    Guarded gin(1 << 17), gout(1 << 17);
    size_t runs = 0, oks = 0;
    for (int iter = 0; iter < 300; iter++) {
        size_t n = rng() % 66000;
        std::vector<uint8_t> data(n);
        int kind = rng() % 4;
        for (size_t i = 0; i < n; i++) data[i] = kind == 0 ? (uint8_t)rng() : kind == 1 ? "ACGT"[rng() % 4] : kind == 2 ? (uint8_t)(i / 7) : (uint8_t)((rng() % 16) ? 'F' : rng());
        std::vector<uint8_t> comp(compressBound(n) + 64);
        z_stream zs; memset(&zs, 0, sizeof zs); deflateInit2(&zs, (int)(rng() % 10), Z_DEFLATED, -15, 8, (int)(rng() % 5));
        zs.next_in = data.data(); zs.avail_in = n; zs.next_out = comp.data(); zs.avail_out = comp.size(); deflate(&zs, Z_FINISH); size_t cl = zs.total_out; deflateEnd(&zs);
        uint8_t *in = gin.place(cl); memcpy(in, comp.data(), cl);
        uint8_t *out = gout.place(n);
        if (!sd_inflate_raw(in, cl, out, n) || memcmp(out, data.data(), n)) { printf("round trip failed n=%zu\n", n); return 1; }
        for (int k = 0; k < 200; k++) {         // corrupt / truncate / wrong sizes
            size_t cl2 = cl; if (rng() % 4 == 0 && cl > 1) cl2 = cl - 1 - rng() % (cl - 1);
            uint8_t *in2 = gin.place(cl2); memcpy(in2, comp.data(), cl2);
            int flips = rng() % 4; for (int f = 0; f < flips && cl2; f++) in2[rng() % cl2] ^= (uint8_t)(1u << (rng() % 8));
            size_t n2 = (rng() % 3 == 0) ? rng() % 66000 : n;
            oks += sd_inflate_raw(in2, cl2, gout.place(n2), n2); runs++;
        }
    }
    printf("guarded runs %zu (ok %zu): no fault\n", runs, oks);
}
