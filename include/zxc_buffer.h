/*
 * zxc_buffer.h -- one-shot frame API, frameless block API, reusable contexts.
 *
 * Drop-in declarations for the reference's include/zxc_buffer.h.  Every codec
 * entry point below is backed by sm_100a CUDA kernels (zxc_b200/csrc); there is
 * no CPU codec behind them.  All pointers are HOST pointers owned by the
 * caller, exactly as in the reference; device-resident entry points are the
 * additive ones in zxc_b200.h.
 *
 * Reference interface replaced (file:line in /root/reference):
 *   info getters            include/zxc_buffer.h:60-84   src/lib/zxc_common.c:930-1017
 *   zxc_compress_bound      include/zxc_buffer.h:97      src/lib/zxc_common.c:850
 *   zxc_compress            include/zxc_buffer.h:119     src/lib/zxc_dispatch.c:658
 *   zxc_decompress          include/zxc_buffer.h:140     src/lib/zxc_dispatch.c:842
 *   zxc_decompress_inplace* include/zxc_buffer.h:169-214 src/lib/zxc_dispatch.c:1118-1190
 *   zxc_get_decompressed_size / zxc_get_dict_id          src/lib/zxc_dispatch.c:1203-1241
 *   block API               include/zxc_buffer.h:283-420 src/lib/zxc_dispatch.c:1627-1858
 *   ctx API                 include/zxc_buffer.h:440-540 src/lib/zxc_dispatch.c:1260-1601
 *   static-workspace API    include/zxc_buffer.h:560-608 src/lib/zxc_dispatch.c:1871-1965
 */
#ifndef ZXC_BUFFER_H
#define ZXC_BUFFER_H

#include <stddef.h>
#include <stdint.h>

#include "zxc_export.h"
#include "zxc_opts.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- library info ---- */
ZXC_EXPORT int zxc_min_level(void);
ZXC_EXPORT int zxc_max_level(void);
ZXC_EXPORT int zxc_default_level(void);
ZXC_EXPORT const char* zxc_version_string(void);

/* ---- whole-frame, host buffers ---- */

/* Worst-case frame size for input_size bytes at any level / block size; 0 on overflow. */
ZXC_EXPORT uint64_t zxc_compress_bound(const size_t input_size);

/* Encode src into a complete v8 frame (header, blocks, EOF, optional SEK, footer).
 * Returns bytes written or a negative zxc_error_t. */
ZXC_EXPORT int64_t zxc_compress(const void* src, const size_t src_size, void* dst,
                                const size_t dst_capacity, const zxc_compress_opts_t* opts);

/* Decode a complete frame.  dst_capacity may equal the decoded size exactly.
 * Returns bytes produced or a negative zxc_error_t; the first failing block in
 * stream order decides the code. */
ZXC_EXPORT int64_t zxc_decompress(const void* src, const size_t src_size, void* dst,
                                  const size_t dst_capacity, const zxc_decompress_opts_t* opts);

/* Single-buffer decode: the frame sits flush-right in `buffer`. */
ZXC_EXPORT size_t zxc_decompress_inplace_bound(const void* src, const size_t src_size);
ZXC_EXPORT int64_t zxc_decompress_inplace(void* buffer, const size_t buffer_capacity,
                                          const size_t comp_size,
                                          const zxc_decompress_opts_t* opts);

/* Footer / header probes; 0 when the buffer is not a plausible frame. */
ZXC_EXPORT uint64_t zxc_get_decompressed_size(const void* src, const size_t src_size);
ZXC_EXPORT uint32_t zxc_get_dict_id(const void* src, size_t src_size);

/* ---- opaque contexts ---- */
typedef struct zxc_cctx_s zxc_cctx;
typedef struct zxc_dctx_s zxc_dctx;

/* ---- frameless single-block API (8-byte block header + payload [+ checksum]) ---- */
ZXC_EXPORT uint64_t zxc_compress_block_bound(size_t input_size);
ZXC_EXPORT uint64_t zxc_decompress_block_bound(const size_t uncompressed_size);
ZXC_EXPORT int64_t zxc_compress_block(zxc_cctx* cctx, const void* src, size_t src_size, void* dst,
                                      size_t dst_capacity, const zxc_compress_opts_t* opts);
ZXC_EXPORT int64_t zxc_decompress_block(zxc_dctx* dctx, const void* src, size_t src_size, void* dst,
                                        size_t dst_capacity, const zxc_decompress_opts_t* opts);
ZXC_EXPORT int64_t zxc_decompress_block_safe(zxc_dctx* dctx, const void* src, const size_t src_size,
                                             void* dst, const size_t dst_capacity,
                                             const zxc_decompress_opts_t* opts);
ZXC_EXPORT uint64_t zxc_estimate_cctx_size(size_t src_size, int level);

/* ---- reusable contexts (hold device scratch + a stream between calls) ---- */
ZXC_EXPORT zxc_cctx* zxc_create_cctx(const zxc_compress_opts_t* opts);
ZXC_EXPORT void zxc_free_cctx(zxc_cctx* cctx);
ZXC_EXPORT int64_t zxc_compress_cctx(zxc_cctx* cctx, const void* src, size_t src_size, void* dst,
                                     size_t dst_capacity, const zxc_compress_opts_t* opts);
ZXC_EXPORT zxc_dctx* zxc_create_dctx(void);
ZXC_EXPORT void zxc_free_dctx(zxc_dctx* dctx);
ZXC_EXPORT int64_t zxc_decompress_dctx(zxc_dctx* dctx, const void* src, size_t src_size, void* dst,
                                       size_t dst_capacity, const zxc_decompress_opts_t* opts);

/* ---- caller-provided workspace variants ---- */
ZXC_EXPORT size_t zxc_static_cctx_workspace_size(const size_t block_size, const int level);
ZXC_EXPORT zxc_cctx* zxc_init_static_cctx(void* workspace, const size_t workspace_size,
                                          const zxc_compress_opts_t* opts);
ZXC_EXPORT size_t zxc_static_dctx_workspace_size(const size_t block_size);
ZXC_EXPORT zxc_dctx* zxc_init_static_dctx(void* workspace, const size_t workspace_size,
                                          const size_t block_size);

#ifdef __cplusplus
}
#endif
#endif /* ZXC_BUFFER_H */
