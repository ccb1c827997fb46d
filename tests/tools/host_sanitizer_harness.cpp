// Host code under AddressSanitizer + UndefinedBehaviorSanitizer (test infrastructure; run by hand in the build container):
//   g++ -O1 -g -fsanitize=address,undefined -std=c++17 -I/usr/local/cuda/include -Iinclude -Islamdunk_b200/csrc \
//       tests/tools/host_sanitizer_harness.cpp slamdunk_b200/csrc/sd_{decode,inflate,deflate,bamwrite,text}.cpp \
//       -o /tmp/san -L/usr/local/cuda/lib64 -lcudart -lz -lpthread
//   ASAN_OPTIONS=detect_leaks=0:protect_shadow_gap=0 /tmp/san ref.fa reads.bam [more.bam ...]
// sd_load_fasta on 1 and 3 threads, sd_decode_bam with every option set the package uses on 1 and 3 threads,
// sd_rewrite_bam_indexed at levels 1 / 6 / 9 (the own inflate, deflate and CRC underneath all of them).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <string>
#include "slamdunk_b200.h"
int main(int argc, char **argv) {
    // argv: fasta bam [bam...]
    char err[512];
    sd_hgenome *g = nullptr;
    for (int th : {1, 3}) {
        if (sd_load_fasta(argv[1], th, &g, err, sizeof err) != SD_OK) { printf("fasta: %s\n", err); return 1; }
        if (th == 1) sd_hgenome_free(g);
    }
    std::vector<const char *> names(g->names, g->names + g->n_chrom);
    for (int a = 2; a < argc; a++) {
        const int opts[] = {SD_DECODE_MP | SD_DECODE_GROUP | SD_DECODE_SEQ2 | SD_DECODE_WIRE, SD_DECODE_GROUP | SD_DECODE_SEQ2 | SD_DECODE_WIRE,
                            SD_DECODE_MP | SD_DECODE_GROUP, SD_DECODE_MP | SD_DECODE_MP_ALL, SD_DECODE_RAW_QUAL | SD_DECODE_SEQ2, 0};
        uint64_t n_rec = 0;
        for (int o : opts)
            for (int th : {1, 3}) {
                sd_hbatch *hb = nullptr;
                int rc = sd_decode_bam(argv[a], names.data(), (uint32_t)names.size(), o, 0, th, &hb, err, sizeof err);
                if (rc != SD_OK) { printf("decode %s opts %d: %s\n", argv[a], o, err); return 1; }
                n_rec = hb->batch.n_reads;
                sd_hbatch_free(hb);
            }
        std::vector<uint8_t> keep(n_rec ? n_rec : 1, 1);
        uint64_t nw = 0;
        sd_hbatch *hb = nullptr;
        sd_decode_bam(argv[a], nullptr, 0, 0, 0, 1, &hb, err, sizeof err);
        std::string text = hb->header_text ? hb->header_text : "";
        sd_hbatch_free(hb);
        for (int level : {1, 6, 9}) {
            int rc = sd_rewrite_bam_indexed(argv[a], "/tmp/sd_san_out.bam", "/tmp/sd_san_out.bam.bai", text.c_str(), keep.data(), nullptr, n_rec, 1, level, 2, &nw, err, sizeof err);
            if (rc != SD_OK) { printf("rewrite %s: %s\n", argv[a], err); return 1; }
        }
        printf("%s: %llu records ok\n", argv[a], (unsigned long long)n_rec);
    }
    sd_hgenome_free(g);
    return 0;
}
