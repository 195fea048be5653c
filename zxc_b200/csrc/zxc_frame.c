/*
 * zxc_frame.c -- host-side frame walker: frame bytes -> device job table.
 *
 * This is row D9 of SURVEY.md section 8: everything in the reference's frame
 * drivers that is O(blocks) and pure header arithmetic stays on the host; the
 * O(bytes) work is one GPU launch over the resulting jobs.
 *
 * Follows: zxc_decompress_frame (src/lib/zxc_dispatch.c:858-1005) for the
 * sequential walk and its check order; zxc_seekable_parse
 * (src/lib/zxc_seekable.c:270-396) for the SEK table.
 */
#include "zxc_frame.h"

#include <stdlib.h>

#include "zxc_format.h"

/* Sequential header walk.  Fills w->jobs (malloc'ed, caller frees). */
int zxw_walk(const uint8_t* src, size_t src_size, zxw_walk_t* w) {
    memset(w, 0, sizeof *w);
    zxf_file_header_t fh;
    const int hrc = zxf_read_file_header(src, src_size, &fh, 1);
    if (hrc != ZXC_OK) return hrc;
    w->block_size = (uint32_t)fh.block_size;
    w->has_checksum = fh.has_checksum;
    w->dict_id = fh.dict_id;
    if (src_size >= ZXC_FILE_FOOTER_SIZE) {
        w->footer_size = zxf_le64(src + src_size - ZXC_FILE_FOOTER_SIZE);
        w->footer_hash = zxf_le32(src + src_size - ZXC_FILE_FOOTER_SIZE + 8);
    }
    size_t cap = 64, n = 0;
    zxc_b200_job_t* jobs = (zxc_b200_job_t*)malloc(cap * sizeof *jobs);
    if (!jobs) return ZXC_ERROR_MEMORY;
    const size_t trailer = fh.has_checksum ? ZXF_BLOCK_CKS : 0;
    size_t ip = ZXC_FILE_HEADER_SIZE;
    uint32_t ghash = 0;
    w->end = ZXW_END_RAN_OFF;
    /* The walk is a chain of dependent reads -- a header says where the next one is -- and the frame is cold in the
     * caches: ~170 ns per block, 11 ms for the 65 536 blocks of a 4 GiB frame, all of it ahead of the first H2D copy.
     * When the tail of the buffer looks like a seek table its entries predict where the headers are, so they are
     * prefetched a few dozen blocks ahead.  Only a hint: every header is still read, CRC-checked and followed in
     * order exactly as before, a wrong or forged table merely prefetches the wrong lines. */
    const uint8_t* hint_ent = NULL;
    uint64_t hint_n = 0, hint_idx = 0, hint_off = ZXC_FILE_HEADER_SIZE;
    if (w->footer_size > 0 && fh.block_size > 0) {
        const uint64_t nb = (w->footer_size + fh.block_size - 1) / fh.block_size;
        const uint64_t sek_total = ZXF_BLOCK_HDR + nb * ZXF_SEEK_ENTRY;
        if (nb <= 0xFFFFFFFFull && sek_total + ZXC_FILE_FOOTER_SIZE + ZXC_FILE_HEADER_SIZE <= src_size) {
            const uint8_t* sek = src + (src_size - ZXC_FILE_FOOTER_SIZE - sek_total);
            if (sek[0] == ZXF_BT_SEK && zxf_le32(sek + 3) == (uint32_t)(nb * ZXF_SEEK_ENTRY)) {
                hint_ent = sek + ZXF_BLOCK_HDR;
                hint_n = nb;
            }
        }
    }
    while (ip < src_size) {
        const size_t rem = src_size - ip;
        while (hint_idx < hint_n && hint_idx < n + 32) {
            if (hint_off + ZXF_BLOCK_HDR <= src_size) {
                __builtin_prefetch(src + hint_off - 4); /* the previous block's checksum trailer */
                __builtin_prefetch(src + hint_off + ZXF_BLOCK_HDR - 1);
            }
            hint_off += zxf_le32(hint_ent + ZXF_SEEK_ENTRY * hint_idx);
            hint_idx++;
        }
        uint8_t type;
        uint32_t comp;
        if (zxf_read_block_header(src + ip, rem, &type, &comp) != ZXC_OK) {
            w->end = ZXW_END_BAD_HEADER; /* any header failure reads as BAD_HEADER (:916-919) */
            break;
        }
        if (type == ZXF_BT_EOF) {
            w->end = comp == 0 ? ZXW_END_EOF : ZXW_END_BAD_HEADER;
            break;
        }
        if (n == cap) {
            cap *= 2;
            zxc_b200_job_t* nj = (zxc_b200_job_t*)realloc(jobs, cap * sizeof *jobs);
            if (!nj) {
                free(jobs);
                return ZXC_ERROR_MEMORY;
            }
            jobs = nj;
        }
        const uint64_t on_disk = (uint64_t)ZXF_BLOCK_HDR + comp + trailer;
        jobs[n].src_off = ip;
        jobs[n].src_len = (uint32_t)(on_disk < rem ? on_disk : (rem > 0xFFFFFFFFull ? 0xFFFFFFFFull : rem));
        jobs[n].dst_off = (uint64_t)n * fh.block_size;
        jobs[n].dst_cap = 0; /* filled by the caller once the capacity is known */
        n++;
        if (fh.has_checksum && on_disk <= rem)
            ghash = zxf_hash_combine(ghash, zxf_le32(src + ip + ZXF_BLOCK_HDR + comp));
        if (on_disk >= rem) { /* next header would start at/after the end: the loop just ends (:912) */
            ip = src_size;
            break;
        }
        ip += (size_t)on_disk;
    }
    w->jobs = jobs;
    w->n_jobs = n;
    w->global_hash = ghash;
    return ZXC_OK;
}

void zxw_free(zxw_walk_t* w) {
    free(w->jobs);
    w->jobs = NULL;
    w->n_jobs = 0;
}

/* SEK table location + validation over an abstract reader. */
int zxw_seek_parse(zxw_fetch_fn fetch, void* fctx, uint64_t size, zxw_seek_t* s) {
    memset(s, 0, sizeof *s);
    if (size < ZXC_FILE_HEADER_SIZE + 2 * ZXF_BLOCK_HDR + ZXC_FILE_FOOTER_SIZE) return ZXC_ERROR_SRC_TOO_SMALL;
    uint8_t hdr[ZXC_FILE_HEADER_SIZE], ftr[ZXC_FILE_FOOTER_SIZE];
    if (fetch(fctx, hdr, sizeof hdr, 0) != ZXC_OK) return ZXC_ERROR_IO;
    zxf_file_header_t fh;
    const int hrc = zxf_read_file_header(hdr, sizeof hdr, &fh, 1);
    if (hrc != ZXC_OK) return hrc;
    if (fetch(fctx, ftr, sizeof ftr, size - ZXC_FILE_FOOTER_SIZE) != ZXC_OK) return ZXC_ERROR_IO;
    const uint64_t total = zxf_le64(ftr);
    if (total == 0) return ZXC_ERROR_CORRUPT_DATA;
    const uint64_t nb = (total + fh.block_size - 1) / fh.block_size;
    if (nb > 0xFFFFFFFFull) return ZXC_ERROR_CORRUPT_DATA;
    const uint64_t sek_total = ZXF_BLOCK_HDR + nb * ZXF_SEEK_ENTRY;
    if (sek_total + ZXC_FILE_FOOTER_SIZE > size) return ZXC_ERROR_CORRUPT_DATA;
    const uint64_t sek_pos = size - ZXC_FILE_FOOTER_SIZE - sek_total;
    uint8_t* table = (uint8_t*)malloc((size_t)sek_total);
    if (!table) return ZXC_ERROR_MEMORY;
    int rc = ZXC_ERROR_CORRUPT_DATA;
    uint32_t* comp = NULL;
    uint64_t* offs = NULL;
    if (fetch(fctx, table, (size_t)sek_total, sek_pos) != ZXC_OK) {
        rc = ZXC_ERROR_IO;
        goto fail;
    }
    uint8_t type;
    uint32_t csz;
    if (zxf_read_block_header(table, (size_t)sek_total, &type, &csz) != ZXC_OK || type != ZXF_BT_SEK ||
        csz != (uint32_t)(nb * ZXF_SEEK_ENTRY))
        goto fail;
    comp = (uint32_t*)malloc((size_t)nb * sizeof *comp);
    offs = (uint64_t*)malloc(((size_t)nb + 1) * sizeof *offs);
    if (!comp || !offs) {
        rc = ZXC_ERROR_MEMORY;
        goto fail;
    }
    uint64_t acc = ZXC_FILE_HEADER_SIZE;
    for (uint64_t i = 0; i < nb; i++) {
        const uint32_t c = zxf_le32(table + ZXF_BLOCK_HDR + 4 * i);
        if (c < ZXF_BLOCK_HDR || c > size) goto fail;
        comp[i] = c;
        offs[i] = acc;
        acc += c;
        if (acc > size) goto fail;
    }
    offs[nb] = acc;
    if (acc != sek_pos - ZXF_BLOCK_HDR) goto fail; /* prefix sum must land on the EOF block */
    uint8_t eof[ZXF_BLOCK_HDR];
    if (fetch(fctx, eof, sizeof eof, acc) != ZXC_OK) {
        rc = ZXC_ERROR_IO;
        goto fail;
    }
    if (zxf_read_block_header(eof, sizeof eof, &type, &csz) != ZXC_OK || type != ZXF_BT_EOF) goto fail;
    free(table);
    s->num_blocks = (uint32_t)nb;
    s->block_size = (uint32_t)fh.block_size;
    s->has_checksum = fh.has_checksum;
    s->dict_id = fh.dict_id;
    s->total = total;
    s->comp_sizes = comp;
    s->comp_offsets = offs;
    return ZXC_OK;
fail:
    free(table);
    free(comp);
    free(offs);
    return rc;
}

void zxw_seek_free(zxw_seek_t* s) {
    free(s->comp_sizes);
    free(s->comp_offsets);
    memset(s, 0, sizeof *s);
}
