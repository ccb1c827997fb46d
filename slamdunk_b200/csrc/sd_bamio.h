// Internal BAM I/O helpers shared by the decoder and the BAM writer.
#pragma once
#include <stdint.h>
#include <stdlib.h>

#include <string>
#include <vector>

// Byte buffer without the zero fill of std::vector (every byte is overwritten anyway).
struct RawBuf {
    uint8_t *p = nullptr;
    size_t n = 0;
    RawBuf() {}
    explicit RawBuf(size_t bytes) : p((uint8_t *)malloc(bytes ? bytes : 1)), n(bytes) {}
    RawBuf(const RawBuf &) = delete;
    RawBuf &operator=(const RawBuf &) = delete;
    ~RawBuf() { free(p); }
    void reset(size_t bytes) { free(p); p = (uint8_t *)malloc(bytes ? bytes : 1); n = bytes; }
    void release() { free(p); p = nullptr; n = 0; }
    uint8_t *data() const { return p; }
    size_t size() const { return n; }
};

// Read a BGZF file and inflate all blocks on `threads` threads (CRC checked).
int sd_bam_inflate_file(const char *path, int threads, RawBuf &raw, size_t *compressed_bytes, char *err, int errlen);
// Parse magic, header text and reference table; *after = offset of the first alignment record.
int sd_bam_parse_header(const RawBuf &raw, const char *path, std::string &text, std::vector<std::string> &names,
                        std::vector<uint32_t> &lengths, size_t *after, char *err, int errlen);
// Offsets (just past block_size) of all alignment records starting at `begin` (parallel, verified; see sd_decode.cpp).
int sd_bam_index_records(const RawBuf &raw, size_t begin, int32_t n_ref, int threads, std::vector<size_t> &recoff,
                         const char *path, char *err, int errlen);
