/*
 * zxc_pstream.h -- push-streaming (caller-driven) API.
 *
 * OUT OF HOT-PATH SCOPE (one block at a time,
 * latency-shaped).  Declared and exported for ABI completeness; create()
 * returns NULL and the step functions report ZXC_B200_ERROR_UNSUPPORTED.
 *
 * Reference interface: include/zxc_pstream.h:70-296, src/lib/zxc_pstream.c.
 */
#ifndef ZXC_PSTREAM_H
#define ZXC_PSTREAM_H

#include <stddef.h>
#include <stdint.h>

#include "zxc_export.h"
#include "zxc_opts.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    const void* src;
    size_t size;
    size_t pos;
} zxc_inbuf_t;

typedef struct {
    void* dst;
    size_t size;
    size_t pos;
} zxc_outbuf_t;

typedef struct zxc_cstream_s zxc_cstream;
typedef struct zxc_dstream_s zxc_dstream;

ZXC_EXPORT zxc_cstream* zxc_cstream_create(const zxc_compress_opts_t* opts);
ZXC_EXPORT void zxc_cstream_free(zxc_cstream* cs);
ZXC_EXPORT int64_t zxc_cstream_compress(zxc_cstream* cs, zxc_outbuf_t* out, zxc_inbuf_t* in);
ZXC_EXPORT int64_t zxc_cstream_end(zxc_cstream* cs, zxc_outbuf_t* out);
ZXC_EXPORT size_t zxc_cstream_in_size(const zxc_cstream* cs);
ZXC_EXPORT size_t zxc_cstream_out_size(const zxc_cstream* cs);

ZXC_EXPORT zxc_dstream* zxc_dstream_create(const zxc_decompress_opts_t* opts);
ZXC_EXPORT void zxc_dstream_free(zxc_dstream* ds);
ZXC_EXPORT int64_t zxc_dstream_decompress(zxc_dstream* ds, zxc_outbuf_t* out, zxc_inbuf_t* in);
ZXC_EXPORT int zxc_dstream_finished(const zxc_dstream* ds);
ZXC_EXPORT size_t zxc_dstream_in_size(const zxc_dstream* ds);
ZXC_EXPORT size_t zxc_dstream_out_size(const zxc_dstream* ds);

#ifdef __cplusplus
}
#endif
#endif /* ZXC_PSTREAM_H */
