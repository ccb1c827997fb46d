// sd_deflate_raw the same way: inputs of every size 0..299 and random sizes up to 65 535, output buffers both roomy and too small, every stream inflated back
// (test infrastructure; g++ -O2 -std=c++17 -Iinclude tests/tools/deflate_guard_pages.cpp slamdunk_b200/csrc/sd_inflate.cpp slamdunk_b200/csrc/sd_deflate.cpp -lz)
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
    std::mt19937_64 rng(7);
    Guarded gin(1 << 17), gout(1 << 17);
    std::vector<uint8_t> back(70000);
    size_t runs = 0, declined = 0;
    for (int iter = 0; iter < 4000; iter++) {
        size_t n = iter < 300 ? iter : rng() % 65536;
        uint8_t *in = gin.place(n);
        int kind = rng() % 5;
        for (size_t i = 0; i < n; i++) in[i] = kind == 0 ? (uint8_t)rng() : kind == 1 ? "ACGT"[rng() % 4] : kind == 2 ? (uint8_t)(i / 7) : kind == 3 ? (uint8_t)((rng() % 16) ? 'F' : rng()) : (uint8_t)(i % 5);
        int level = 1 + rng() % 9;
        size_t cap = (rng() % 4 == 0) ? 16 + rng() % (n + 64) : n + 1024;
        uint8_t *out = gout.place(cap);
        uint64_t c = sd_deflate_raw(in, n, out, cap, level);
        runs++;
        if (!c) { declined++; continue; }
        if (c > cap) { printf("wrote more than cap\n"); return 1; }
        if (!sd_inflate_raw(out, c, back.data(), n) || memcmp(back.data(), in, n)) { printf("round trip failed n=%zu level=%d\n", n, level); return 1; }
    }
    printf("guarded deflate runs %zu (declined %zu): no fault, all round trips ok\n", runs, declined);
}
