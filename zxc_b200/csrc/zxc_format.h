/*
 * zxc_format.h -- wire-format v8 primitives shared by the host C code.
 *
 * Host-side restatement of the small L1 layer of the reference (SURVEY.md
 * section 8 row D9 / component 4,5,11): constants, little-endian access,
 * header CRCs, rapidhash-based checksum, header/footer/SEK read+write.
 * Pure integer code, O(headers); the per-byte work lives in zxc_gpu.cu.
 *
 * Follows: src/lib/zxc_internal.h:331-547 (constants), :1188-1214 (hash8/16),
 * :1353-1393 (checksum fold, global hash); src/lib/zxc_common.c:534-680
 * (headers/footer), :850-926 (bounds); src/lib/vendors/rapidhash.h (V3).
 */
#ifndef ZXC_B200_FORMAT_H
#define ZXC_B200_FORMAT_H

#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include "zxc_constants.h"
#include "zxc_error.h"

#ifdef __cplusplus
extern "C" {
#endif

#define ZXF_MAGIC 0x9CB02EF5u
#define ZXF_DICT_MAGIC 0x9CB0D1C7u
#define ZXF_VERSION 8
#define ZXF_DICT_VERSION 1
#define ZXF_BLOCK_HDR 8
#define ZXF_BLOCK_CKS 4
#define ZXF_SUB_HDR 12
#define ZXF_LIT_SLACK 32
#define ZXF_PAD 32
#define ZXF_TAIL_PAD (ZXF_PAD * 66)
#define ZXF_BLOCK_OVERHEAD 68
#define ZXF_SEEK_ENTRY 4
#define ZXF_FLAG_CHECKSUM 0x80u
#define ZXF_FLAG_DICT 0x40u
#define ZXF_MIN_MATCH 5

enum { ZXF_BT_RAW = 0, ZXF_BT_GLO = 1, ZXF_BT_GHI = 2, ZXF_BT_SEK = 254, ZXF_BT_EOF = 255 };

static inline uint32_t zxf_le16(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }
static inline uint32_t zxf_le32(const uint8_t* p) {
    uint32_t v;
    memcpy(&v, p, 4);
    return v; /* x86-64 / aarch64 LE hosts only (the B200 boxes) */
}
static inline uint64_t zxf_le64(const uint8_t* p) {
    uint64_t v;
    memcpy(&v, p, 8);
    return v;
}
static inline void zxf_st16(uint8_t* p, uint32_t v) {
    p[0] = (uint8_t)v;
    p[1] = (uint8_t)(v >> 8);
}
static inline void zxf_st32(uint8_t* p, uint32_t v) { memcpy(p, &v, 4); }
static inline void zxf_st64(uint8_t* p, uint64_t v) { memcpy(p, &v, 8); }

static inline int zxf_valid_block_size(size_t bs) {
    return bs >= ZXC_BLOCK_SIZE_MIN && bs <= ZXC_BLOCK_SIZE_MAX && (bs & (bs - 1)) == 0;
}
static inline unsigned zxf_log2(uint32_t v) { return 31u - (unsigned)__builtin_clz(v); }
/* smallest valid block size >= n (zxc_block_size_ceil) */
static inline size_t zxf_block_size_ceil(size_t n) {
    size_t bs = ZXC_BLOCK_SIZE_MIN;
    while (bs < n && bs < ZXC_BLOCK_SIZE_MAX) bs <<= 1;
    return bs;
}

uint8_t zxf_hash8(const uint8_t* hdr8);
uint16_t zxf_hash16(const uint8_t* hdr16);
uint64_t zxf_rapidhash(const void* key, size_t len, uint64_t seed);
uint32_t zxf_checksum(const void* p, size_t len);
uint32_t zxf_checksum_seed(const void* p, size_t len, uint32_t seed);
static inline uint32_t zxf_hash_combine(uint32_t h, uint32_t blk) { return ((h << 1) | (h >> 31)) ^ blk; }

typedef struct {
    size_t block_size;
    int has_checksum;
    uint32_t dict_id;
} zxf_file_header_t;

int zxf_write_file_header(uint8_t* dst, size_t cap, size_t block_size, int has_checksum,
                          uint32_t dict_id);
/* want_block_size == 0 skips the block-size-code range check (reference passes NULL there) */
int zxf_read_file_header(const uint8_t* src, size_t n, zxf_file_header_t* out, int want_block_size);
int zxf_write_block_header(uint8_t* dst, size_t cap, uint8_t type, uint32_t comp_size);
/* returns ZXC_OK / SRC_TOO_SMALL / BAD_HEADER; fills type and comp_size */
int zxf_read_block_header(const uint8_t* src, size_t n, uint8_t* type, uint32_t* comp_size);
int zxf_write_footer(uint8_t* dst, size_t cap, uint64_t src_size, uint32_t global_hash, int checksum);

/* is the footer's size reachable by an archive of comp_size bytes (zxc_dispatch.c:1019-1024) */
static inline int zxf_dsize_plausible(uint64_t dsize, size_t block_size, size_t comp_size) {
    const uint64_t need = dsize / block_size + (dsize % block_size != 0);
    return need <= (uint64_t)(comp_size / ZXF_BLOCK_HDR);
}

#ifdef __cplusplus
}
#endif
#endif
