/*
 * zxc_seekable.h -- random access over frames that carry a SEK table.
 *
 * This is the primary GPU entry: zxc_seekable_decompress_range[_mt] turns the
 * covered block span into a device job table and decodes all of it in one
 * launch (one warp per block), where the reference forks one CPU thread per
 * stripe of blocks.
 *
 * Reference interface replaced (file:line in /root/reference):
 *   zxc_seekable_open               include/zxc_seekable.h:88    src/lib/zxc_seekable.c:409
 *   zxc_reader_t / open_reader      include/zxc_seekable.h:105-140 src/lib/zxc_seekable.c:430-554
 *   getters                         include/zxc_seekable.h:148-176 src/lib/zxc_seekable.c:561-600
 *   zxc_seekable_decompress_range   include/zxc_seekable.h:191   src/lib/zxc_seekable.c:695
 *   ..._range_mt                    include/zxc_seekable.h:214   src/lib/zxc_seekable.c:999
 *   zxc_seekable_set_dict           include/zxc_seekable.h:243   src/lib/zxc_seekable.c:1144
 *   zxc_write_seek_table / size     include/zxc_seekable.h:262-276 src/lib/zxc_seekable.c:172-214
 */
#ifndef ZXC_SEEKABLE_H
#define ZXC_SEEKABLE_H

#include <stddef.h>
#include <stdint.h>

#include "zxc_export.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct zxc_seekable_s zxc_seekable;

/* Borrow `src` (must outlive the handle); NULL if there is no valid SEK table. */
ZXC_EXPORT zxc_seekable* zxc_seekable_open(const void* src, const size_t src_size);

/* Positional reader for archives that do not live in memory.  read_at must
 * return `len` on success and be callable from several threads for _range_mt. */
typedef struct {
    int64_t (*read_at)(void* ctx, void* dst, size_t len, uint64_t offset);
    void* ctx;
    uint64_t size;
} zxc_reader_t;

ZXC_EXPORT zxc_seekable* zxc_seekable_open_reader(const zxc_reader_t* r);

ZXC_EXPORT uint32_t zxc_seekable_get_num_blocks(const zxc_seekable* s);
ZXC_EXPORT uint64_t zxc_seekable_get_decompressed_size(const zxc_seekable* s);
ZXC_EXPORT uint32_t zxc_seekable_get_block_comp_size(const zxc_seekable* s,
                                                     const uint32_t block_idx);
ZXC_EXPORT uint32_t zxc_seekable_get_block_decomp_size(const zxc_seekable* s,
                                                       const uint32_t block_idx);

/* Decode bytes [offset, offset+len) of the original data into dst.  Returns len
 * or a negative zxc_error_t.  Per-block checksums are not verified here (the
 * reference does not either, zxc_seekable.c:707). */
ZXC_EXPORT int64_t zxc_seekable_decompress_range(zxc_seekable* s, void* dst,
                                                 const size_t dst_capacity, const uint64_t offset,
                                                 const size_t len);

/* Same result; n_threads only sizes the host-side staging fan-out, the decode
 * itself is always one GPU launch over all covered blocks. */
ZXC_EXPORT int64_t zxc_seekable_decompress_range_mt(zxc_seekable* s, void* dst,
                                                    const size_t dst_capacity,
                                                    const uint64_t offset, const size_t len,
                                                    int n_threads);

ZXC_EXPORT void zxc_seekable_free(zxc_seekable* s);

/* Attach the dictionary (copied) that the frame header's dict_id names. */
ZXC_EXPORT int zxc_seekable_set_dict(zxc_seekable* s, const void* dict, size_t dict_size,
                                     const void* dict_huf);

/* SEK block writer: 8-byte block header (type 254) + one u32le per block. */
ZXC_EXPORT int64_t zxc_write_seek_table(uint8_t* dst, const size_t dst_capacity,
                                        const uint32_t* comp_sizes, const uint32_t num_blocks);
ZXC_EXPORT size_t zxc_seek_table_size(const uint32_t num_blocks);

#ifdef __cplusplus
}
#endif
#endif /* ZXC_SEEKABLE_H */
