/*
 * zxc_opts.h -- option structs (layout is ABI).
 *
 * Replaces: reference include/zxc_opts.h:43-44 (callback), :58-80 (compress
 * opts), :86-95 (decompress opts), :104-112 (size getters).  Zero-initialised
 * structs and NULL both mean "defaults" (level 3, 512 KiB blocks, no checksum,
 * not seekable, no dictionary).
 */
#ifndef ZXC_OPTS_H
#define ZXC_OPTS_H

#include <stddef.h>
#include <stdint.h>

#include "zxc_export.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef void (*zxc_progress_callback_t)(uint64_t bytes_processed, uint64_t bytes_total,
                                        const void* user_data);

typedef struct {
    int n_threads;        /* ignored by the buffer API (as in the reference) */
    int level;            /* 1..7, 0 = default (3); > 7 clamps to 7 */
    size_t block_size;    /* power of two in [4 KiB, 2 MiB]; 0 = 512 KiB */
    int checksum_enabled; /* per-block rapidhash fold + global rotate-xor hash */
    int seekable;         /* append a SEK table (one u32 per block) */
    const void* dict;     /* dictionary content, re-prepended to every block */
    size_t dict_size;     /* <= ZXC_DICT_SIZE_MAX */
    const void* dict_huf; /* 128-byte shared code-length table, part of dict_id */
    zxc_progress_callback_t progress_cb;
    void* user_data;
} zxc_compress_opts_t;

typedef struct {
    int n_threads;
    int checksum_enabled; /* verify only if the frame header flag is also set */
    const void* dict;
    size_t dict_size;
    const void* dict_huf;
    zxc_progress_callback_t progress_cb;
    void* user_data;
} zxc_decompress_opts_t;

ZXC_EXPORT size_t zxc_compress_opts_size(void);
ZXC_EXPORT size_t zxc_decompress_opts_size(void);

#ifdef __cplusplus
}
#endif
#endif /* ZXC_OPTS_H */
