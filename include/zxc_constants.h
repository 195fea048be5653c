/*
 * zxc_constants.h -- version, block-size limits, level names.
 *
 * Replaces: reference include/zxc_constants.h:30-131.  Values are ABI / wire
 * format facts (library 0.13.3, wire format v8) and must not drift.
 */
#ifndef ZXC_CONSTANTS_H
#define ZXC_CONSTANTS_H

#define ZXC_VERSION_MAJOR 0
#define ZXC_VERSION_MINOR 13
#define ZXC_VERSION_PATCH 3

#define ZXC_STR_HELPER(x) #x
#define ZXC_STR(x) ZXC_STR_HELPER(x)
#define ZXC_LIB_VERSION_STR \
    ZXC_STR(ZXC_VERSION_MAJOR) "." ZXC_STR(ZXC_VERSION_MINOR) "." ZXC_STR(ZXC_VERSION_PATCH)

/* block_size = 1 << code, code in [12, 21] (4 KiB .. 2 MiB); default 512 KiB */
#define ZXC_BLOCK_SIZE_MIN_LOG2 12
#define ZXC_BLOCK_SIZE_MAX_LOG2 21
#define ZXC_BLOCK_SIZE_MIN (1U << ZXC_BLOCK_SIZE_MIN_LOG2)
#define ZXC_BLOCK_SIZE_MAX (1U << ZXC_BLOCK_SIZE_MAX_LOG2)
#define ZXC_BLOCK_SIZE_DEFAULT (512 * 1024)

/* dictionaries: content <= 65535 bytes, .zxd header 16 bytes, 128-byte packed code lengths */
#define ZXC_DICT_SIZE_MAX ((1U << 16) - 1U)
#define ZXC_DICT_HEADER_SIZE 16
#define ZXC_HUF_TABLE_SIZE 128

#define ZXC_MAX_THREADS 512

#define ZXC_FILE_HEADER_SIZE 16
#define ZXC_FILE_FOOTER_SIZE 12

typedef enum {
    ZXC_LEVEL_FASTEST = 1,
    ZXC_LEVEL_FAST = 2,
    ZXC_LEVEL_DEFAULT = 3,
    ZXC_LEVEL_BALANCED = 4,
    ZXC_LEVEL_COMPACT = 5,
    ZXC_LEVEL_DENSITY = 6,
    ZXC_LEVEL_ULTRA = 7
} zxc_compression_level_t;

#endif /* ZXC_CONSTANTS_H */
