// lib_file.cu -- the flat on-disk / wire container of include/fhe_b200_file.h (host code only: no kernels).
#include <stdio.h>
#include <string.h>

#include <vector>

#include "../../include/fhe_b200_file.h"
#include "common.cuh"

using namespace fhe;

namespace {
const char MAGIC[8] = {'F', 'H', 'E', 'B', '2', '0', '0', '\0'};
constexpr size_t HEADER = 96;

u64 fnv1a(const unsigned char *p, size_t n, u64 h = 0xcbf29ce484222325ull) {
    for (size_t i = 0; i < n; i++) {
        h ^= p[i];
        h *= 0x100000001b3ull;
    }
    return h;
}
void put32(unsigned char *b, u32 v) { for (int i = 0; i < 4; i++) b[i] = (unsigned char)(v >> (8 * i)); }
void put64(unsigned char *b, u64 v) { for (int i = 0; i < 8; i++) b[i] = (unsigned char)(v >> (8 * i)); }
u32 get32(const unsigned char *b) { u32 v = 0; for (int i = 0; i < 4; i++) v |= (u32)b[i] << (8 * i); return v; }
u64 get64(const unsigned char *b) { u64 v = 0; for (int i = 0; i < 8; i++) v |= (u64)b[i] << (8 * i); return v; }

struct File {
    FILE *f = nullptr;
    ~File() { if (f) fclose(f); }
};
}  // namespace

extern "C" {
uint64_t fhe_file_payload_bytes(const fhe_file_info *info) {
    if (info == nullptr) return 0;
    const u64 words = info->count * info->words_per_object;
    if (info->count != 0 && words / info->count != info->words_per_object) return 0;  // overflow
    switch (info->encoding) {
        case FHE_ENC_U64: return words * 8;
        case FHE_ENC_U32: return (info->q != 0 && info->q <= (1ull << 32)) ? words * 4 : 0;
        case FHE_ENC_PACKED:
            if (info->bits < 1 || info->bits > 32 || info->words_per_object % 32 != 0) return 0;
            if (info->q == 0 || (info->bits < 64 && info->q > (1ull << info->bits))) return 0;
            return words / 32 * info->bits * 4;
        default: return 0;
    }
}
int fhe_file_write(const char *path, fhe_file_info *info, const void *payload) {
    FHE_REQUIRE(path && info, "fhe_file_write: null pointer");
    FHE_REQUIRE(info->kind >= FHE_FILE_RQ && info->kind <= FHE_FILE_GLEV_RQ, "fhe_file_write: unknown kind");
    const u64 bytes = fhe_file_payload_bytes(info);
    FHE_REQUIRE(bytes != 0 || info->count == 0, "fhe_file_write: inconsistent header fields for this encoding");
    FHE_REQUIRE(payload != nullptr || bytes == 0, "fhe_file_write: null payload");
    info->version = 1;
    info->payload_bytes = bytes;
    info->checksum = fnv1a(static_cast<const unsigned char *>(payload), (size_t)bytes);
    unsigned char h[HEADER] = {0};
    memcpy(h, MAGIC, 8);
    put32(h + 8, info->version); put32(h + 12, info->kind); put32(h + 16, info->encoding); put32(h + 20, info->bits);
    put64(h + 24, info->q); put64(h + 32, info->n); put64(h + 40, info->k); put64(h + 48, info->l);
    put64(h + 56, info->count); put64(h + 64, info->words_per_object); put64(h + 72, info->payload_bytes);
    put64(h + 80, info->checksum);
    File fp;
    fp.f = fopen(path, "wb");
    FHE_REQUIRE(fp.f != nullptr, std::string("fhe_file_write: cannot open ") + path);
    FHE_REQUIRE(fwrite(h, 1, HEADER, fp.f) == HEADER, "fhe_file_write: short write (header)");
    FHE_REQUIRE(bytes == 0 || fwrite(payload, 1, (size_t)bytes, fp.f) == (size_t)bytes, "fhe_file_write: short write (payload)");
    return 0;
}
static int read_header(FILE *f, fhe_file_info *info) {
    unsigned char h[HEADER];
    FHE_REQUIRE(fread(h, 1, HEADER, f) == HEADER, "fhe_file: truncated header");
    FHE_REQUIRE(memcmp(h, MAGIC, 8) == 0, "fhe_file: bad magic");
    info->version = get32(h + 8); info->kind = get32(h + 12); info->encoding = get32(h + 16); info->bits = get32(h + 20);
    info->q = get64(h + 24); info->n = get64(h + 32); info->k = get64(h + 40); info->l = get64(h + 48);
    info->count = get64(h + 56); info->words_per_object = get64(h + 64); info->payload_bytes = get64(h + 72);
    info->checksum = get64(h + 80);
    FHE_REQUIRE(info->version == 1, "fhe_file: unsupported version");
    FHE_REQUIRE(info->kind >= FHE_FILE_RQ && info->kind <= FHE_FILE_GLEV_RQ, "fhe_file: unknown kind");
    FHE_REQUIRE(fhe_file_payload_bytes(info) == info->payload_bytes, "fhe_file: header fields do not match the payload size");
    return 0;
}
int fhe_file_read_info(const char *path, fhe_file_info *info) {
    FHE_REQUIRE(path && info, "fhe_file_read_info: null pointer");
    File fp;
    fp.f = fopen(path, "rb");
    FHE_REQUIRE(fp.f != nullptr, std::string("fhe_file_read_info: cannot open ") + path);
    int rc = read_header(fp.f, info);
    if (rc) return rc;
    FHE_REQUIRE(fseek(fp.f, 0, SEEK_END) == 0, "fhe_file: seek failed");
    const long end = ftell(fp.f);
    FHE_REQUIRE(end >= 0 && (u64)end == HEADER + info->payload_bytes, "fhe_file: file length does not match the header");
    return 0;
}
int fhe_file_read_payload(const char *path, void *payload, size_t capacity) {
    FHE_REQUIRE(path, "fhe_file_read_payload: null path");
    File fp;
    fp.f = fopen(path, "rb");
    FHE_REQUIRE(fp.f != nullptr, std::string("fhe_file_read_payload: cannot open ") + path);
    fhe_file_info info;
    int rc = read_header(fp.f, &info);
    if (rc) return rc;
    FHE_REQUIRE(capacity >= info.payload_bytes, "fhe_file_read_payload: buffer too small");
    FHE_REQUIRE(payload != nullptr || info.payload_bytes == 0, "fhe_file_read_payload: null payload");
    FHE_REQUIRE(info.payload_bytes == 0 || fread(payload, 1, (size_t)info.payload_bytes, fp.f) == (size_t)info.payload_bytes,
                "fhe_file: truncated payload");
    FHE_REQUIRE(fnv1a(static_cast<const unsigned char *>(payload), (size_t)info.payload_bytes) == info.checksum,
                "fhe_file: checksum mismatch (corrupted payload)");
    return 0;
}
}
