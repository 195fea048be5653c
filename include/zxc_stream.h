/*
 * zxc_stream.h -- FILE*-based streaming engine.
 *
 * OUT OF HOT-PATH SCOPE (CPU I/O pipeline).  The
 * symbols exist so that binaries linked against the reference's libzxc.so.4
 * still resolve; zxc_stream_decompress / zxc_stream_get_decompressed_size /
 * zxc_seekable_open_file are thin host readers over the GPU buffer path,
 * zxc_stream_compress reports ZXC_B200_ERROR_UNSUPPORTED until the encode
 * kernel covers it.
 *
 * Reference interface: include/zxc_stream.h:60-120, src/lib/zxc_driver.c:1035-1251.
 */
#ifndef ZXC_STREAM_H
#define ZXC_STREAM_H

#include <stdint.h>
#include <stdio.h>

#include "zxc_export.h"
#include "zxc_opts.h"
#include "zxc_seekable.h"

#ifdef __cplusplus
extern "C" {
#endif

ZXC_EXPORT int64_t zxc_stream_compress(FILE* f_in, FILE* f_out, const zxc_compress_opts_t* opts);
ZXC_EXPORT int64_t zxc_stream_decompress(FILE* f_in, FILE* f_out,
                                         const zxc_decompress_opts_t* opts);
ZXC_EXPORT int64_t zxc_stream_get_decompressed_size(FILE* f_in);
ZXC_EXPORT zxc_seekable* zxc_seekable_open_file(FILE* f);

#ifdef __cplusplus
}
#endif
#endif /* ZXC_STREAM_H */
