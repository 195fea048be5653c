/* zxc_frame.h -- host frame walker / SEK parser (see zxc_frame.c). */
#ifndef ZXC_B200_FRAME_H
#define ZXC_B200_FRAME_H

#include <stddef.h>
#include <stdint.h>

#include "zxc_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

enum { ZXW_END_EOF = 0, ZXW_END_BAD_HEADER = 1, ZXW_END_RAN_OFF = 2 };

typedef struct {
    zxc_b200_job_t* jobs; /* dst_off = i * block_size, dst_cap left 0 */
    size_t n_jobs;
    uint32_t block_size;
    int has_checksum;
    uint32_t dict_id;
    int end;              /* how the block stream ended (ZXW_END_*) */
    uint64_t footer_size; /* last 12 bytes of the buffer, whatever they are */
    uint32_t footer_hash;
    uint32_t global_hash; /* rotate-xor fold of the stored per-block checksums */
} zxw_walk_t;

int zxw_walk(const uint8_t* src, size_t src_size, zxw_walk_t* w);
void zxw_free(zxw_walk_t* w);

/* abstract positional read: returns ZXC_OK when exactly len bytes were delivered */
typedef int (*zxw_fetch_fn)(void* ctx, void* dst, size_t len, uint64_t offset);

typedef struct {
    uint32_t num_blocks;
    uint32_t block_size;
    int has_checksum;
    uint32_t dict_id;
    uint64_t total;
    uint32_t* comp_sizes;   /* [num_blocks] on-disk block sizes */
    uint64_t* comp_offsets; /* [num_blocks+1] prefix sums, first = 16 */
} zxw_seek_t;

int zxw_seek_parse(zxw_fetch_fn fetch, void* fctx, uint64_t size, zxw_seek_t* s);
void zxw_seek_free(zxw_seek_t* s);

#ifdef __cplusplus
}
#endif
#endif
