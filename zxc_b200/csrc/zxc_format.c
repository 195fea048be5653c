/*
 * zxc_format.c -- host-side wire-format primitives (see zxc_format.h).
 */
#include "zxc_format.h"

/* ---- header CRCs: one xorshift round over the LE words (zxc_internal.h:1188-1214) ---- */
static inline uint64_t xorshift_round(uint64_t h) {
    h ^= h << 13;
    h ^= h >> 7;
    h ^= h << 17;
    return h;
}

uint8_t zxf_hash8(const uint8_t* p) {
    const uint64_t h = xorshift_round(zxf_le64(p) ^ 0x9E3779B97F4A7C15ull);
    return (uint8_t)((h >> 32) ^ h);
}

uint16_t zxf_hash16(const uint8_t* p) {
    const uint64_t h = xorshift_round(zxf_le64(p) ^ zxf_le64(p + 8) ^ 0xD2D84A61D2D84A61ull);
    const uint32_t r = (uint32_t)((h >> 32) ^ h);
    return (uint16_t)((r >> 16) ^ r);
}

/* ---- rapidhash V3 (vendors/rapidhash.h:130-345), default secrets ---- */
static const uint64_t k_secret[8] = {0x2d358dccaa6c78a5ull, 0x8bb84b93962eacc9ull,
                                     0x4b33a62ed433d4a3ull, 0x4d5a2da51de1aa47ull,
                                     0xa0761d6478bd642full, 0xe7037ed1a0b428dbull,
                                     0x90ed1765281c388cull, 0xaaaaaaaaaaaaaaaaull};

static inline uint64_t mul_fold(uint64_t a, uint64_t b) {
    const __uint128_t r = (__uint128_t)a * b;
    return (uint64_t)r ^ (uint64_t)(r >> 64);
}

uint64_t zxf_rapidhash(const void* key, size_t len, uint64_t seed) {
    const uint8_t* p = (const uint8_t*)key;
    uint64_t a = 0, b = 0;
    size_t rem = len;
    seed ^= mul_fold(seed ^ k_secret[2], k_secret[1]);
    if (len <= 16) {
        if (len >= 8) {
            seed ^= len;
            a = zxf_le64(p);
            b = zxf_le64(p + len - 8);
        } else if (len >= 4) {
            seed ^= len;
            a = zxf_le32(p);
            b = zxf_le32(p + len - 4);
        } else if (len > 0) {
            a = ((uint64_t)p[0] << 45) | p[len - 1];
            b = p[len >> 1];
        }
    } else {
        if (len > 112) {
            /* seven independent accumulators, one per 16-byte lane of a 112-byte stripe */
            uint64_t acc[7] = {seed, seed, seed, seed, seed, seed, seed};
            do {
                for (int k = 0; k < 7; k++)
                    acc[k] = mul_fold(zxf_le64(p + 16 * k) ^ k_secret[k],
                                      zxf_le64(p + 16 * k + 8) ^ acc[k]);
                p += 112;
                rem -= 112;
            } while (rem > 112);
            seed = (acc[0] ^ acc[1] ^ acc[6]) ^ ((acc[2] ^ acc[3]) ^ (acc[4] ^ acc[5]));
        }
        /* up to six chained 16-byte mixes; the last 16 bytes go to the finaliser */
        static const uint8_t tail_secret[6] = {2, 2, 1, 1, 2, 1};
        for (unsigned k = 0; k < 6 && rem > 16u * (k + 1); k++)
            seed = mul_fold(zxf_le64(p + 16 * k) ^ k_secret[tail_secret[k]],
                            zxf_le64(p + 16 * k + 8) ^ seed);
        a = zxf_le64(p + rem - 16) ^ rem;
        b = zxf_le64(p + rem - 8);
    }
    a ^= k_secret[1];
    b ^= seed;
    {
        const __uint128_t r = (__uint128_t)a * b;
        a = (uint64_t)r;
        b = (uint64_t)(r >> 64);
    }
    return mul_fold(a ^ k_secret[7], b ^ k_secret[1] ^ rem);
}

uint32_t zxf_checksum(const void* p, size_t len) {
    const uint64_t h = zxf_rapidhash(p, len, 0);
    return (uint32_t)(h ^ (h >> 32));
}

uint32_t zxf_checksum_seed(const void* p, size_t len, uint32_t seed) {
    const uint64_t h = zxf_rapidhash(p, len, seed);
    return (uint32_t)(h ^ (h >> 32));
}

/* ---- file header (zxc_common.c:534-603) ---- */
int zxf_write_file_header(uint8_t* dst, size_t cap, size_t block_size, int has_checksum,
                          uint32_t dict_id) {
    if (cap < ZXC_FILE_HEADER_SIZE) return ZXC_ERROR_DST_TOO_SMALL;
    memset(dst, 0, ZXC_FILE_HEADER_SIZE);
    zxf_st32(dst, ZXF_MAGIC);
    dst[4] = ZXF_VERSION;
    dst[5] = (uint8_t)zxf_log2((uint32_t)block_size);
    dst[6] = (uint8_t)((has_checksum ? ZXF_FLAG_CHECKSUM : 0) | (dict_id ? ZXF_FLAG_DICT : 0));
    if (dict_id) zxf_st32(dst + 7, dict_id);
    zxf_st16(dst + 14, zxf_hash16(dst));
    return ZXC_FILE_HEADER_SIZE;
}

int zxf_read_file_header(const uint8_t* src, size_t n, zxf_file_header_t* out, int want_block_size) {
    if (n < ZXC_FILE_HEADER_SIZE) return ZXC_ERROR_SRC_TOO_SMALL;
    if (zxf_le32(src) != ZXF_MAGIC) return ZXC_ERROR_BAD_MAGIC;
    if (src[4] != ZXF_VERSION) return ZXC_ERROR_BAD_VERSION;
    uint8_t tmp[ZXC_FILE_HEADER_SIZE];
    memcpy(tmp, src, sizeof tmp);
    tmp[14] = tmp[15] = 0;
    if (zxf_le16(src + 14) != zxf_hash16(tmp) || (src[6] & 0x0Fu) != 0) return ZXC_ERROR_BAD_HEADER;
    out->block_size = 0;
    if (want_block_size) {
        if (src[5] < ZXC_BLOCK_SIZE_MIN_LOG2 || src[5] > ZXC_BLOCK_SIZE_MAX_LOG2)
            return ZXC_ERROR_BAD_BLOCK_SIZE;
        out->block_size = (size_t)1 << src[5];
    }
    out->has_checksum = (src[6] & ZXF_FLAG_CHECKSUM) ? 1 : 0;
    out->dict_id = (src[6] & ZXF_FLAG_DICT) ? zxf_le32(src + 7) : 0;
    return ZXC_OK;
}

/* ---- block header (zxc_common.c:612-660) ---- */
int zxf_write_block_header(uint8_t* dst, size_t cap, uint8_t type, uint32_t comp_size) {
    if (cap < ZXF_BLOCK_HDR) return ZXC_ERROR_DST_TOO_SMALL;
    dst[0] = type;
    dst[1] = 0;
    dst[2] = 0;
    zxf_st32(dst + 3, comp_size);
    dst[7] = 0;
    dst[7] = zxf_hash8(dst);
    return ZXF_BLOCK_HDR;
}

int zxf_read_block_header(const uint8_t* src, size_t n, uint8_t* type, uint32_t* comp_size) {
    if (n < ZXF_BLOCK_HDR) return ZXC_ERROR_SRC_TOO_SMALL;
    uint8_t tmp[ZXF_BLOCK_HDR];
    memcpy(tmp, src, sizeof tmp);
    tmp[7] = 0;
    if (src[7] != zxf_hash8(tmp)) return ZXC_ERROR_BAD_HEADER;
    *type = src[0];
    *comp_size = zxf_le32(src + 3);
    return ZXC_OK;
}

/* ---- footer (zxc_common.c:667-680) ---- */
int zxf_write_footer(uint8_t* dst, size_t cap, uint64_t src_size, uint32_t global_hash, int checksum) {
    if (cap < ZXC_FILE_FOOTER_SIZE) return ZXC_ERROR_DST_TOO_SMALL;
    zxf_st64(dst, src_size);
    zxf_st32(dst + 8, checksum ? global_hash : 0);
    return ZXC_FILE_FOOTER_SIZE;
}
