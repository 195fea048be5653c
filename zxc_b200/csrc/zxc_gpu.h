/*
 * zxc_gpu.h -- internal C interface between the host C code (zxc_api.c,
 * zxc_frame.c) and the CUDA translation unit (zxc_gpu.cu).  Plain C types only.
 */
#ifndef ZXC_B200_GPU_H
#define ZXC_B200_GPU_H

#include <stddef.h>
#include <stdint.h>

#include "zxc_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* A device context: one CUDA stream plus growable device / pinned buffers that
 * survive between calls.  The buffer API is stateless and thread-safe in the
 * reference (docs/API.md:1530-1538), so contexts come from a mutex-guarded
 * free list instead of a global singleton stream. */
typedef struct zxg_ctx zxg_ctx;

/* ZXC_OK, or ZXC_B200_ERROR_NO_DEVICE (message printed once to stderr). */
int zxg_init(void);
zxg_ctx* zxg_acquire(void);
void zxg_release(zxg_ctx* c);
void zxg_destroy(zxg_ctx* c); /* for contexts owned by a zxc_dctx / zxc_seekable */
zxg_ctx* zxg_create(void);

/* Named device buffers of a context, grown on demand (never shrunk). */
enum { ZXG_BUF_IN = 0, ZXG_BUF_OUT, ZXG_BUF_JOBS, ZXG_BUF_STATUS, ZXG_BUF_DICT, ZXG_BUF_SCRATCH,
       ZXG_BUF_AUX, ZXG_BUF_COUNT };
void* zxg_buffer(zxg_ctx* c, int which, size_t bytes); /* NULL on allocation failure */

/* Host <-> device copies on the context's stream.  Pageable host memory is
 * staged through the context's pinned bounce buffers in chunks so the copy and
 * the host memcpy overlap. */
int zxg_h2d(zxg_ctx* c, void* d_dst, const void* h_src, size_t bytes);
int zxg_d2h(zxg_ctx* c, void* h_dst, const void* d_src, size_t bytes);
int zxg_sync(zxg_ctx* c);
void* zxg_stream(zxg_ctx* c);

/* Whole decode step for host-resident input: upload jobs, launch, fetch statuses.
 * h_status receives n_jobs entries. */
int zxg_decode_jobs(zxg_ctx* c, const void* d_src, void* d_dst, const zxc_b200_job_t* h_jobs,
                    uint32_t n_jobs, int32_t* h_status, const void* h_dict, uint32_t dict_size,
                    const void* h_dict_huf, uint32_t block_size, int verify_checksums);

/* Encode src into the frame body (data blocks back to back); see zxc_gpu.cu. */
int zxg_encode_body(zxg_ctx* c, const uint8_t* h_src, uint64_t src_size, uint32_t block_size, int level,
                    int checksum, uint32_t n_blocks, uint8_t* h_body, uint64_t body_cap, uint32_t* h_sizes,
                    uint64_t* body_size, const void* h_dict, uint32_t dict_size, const uint8_t* h_dict_huf_lens);

/* Device selection for the calling thread (multi-device fork-join in zxc_api.c): current device, device count,
 * cudaSetDevice.  zxg_acquire() hands out a context of the calling thread's current device. */
int zxg_current_device(void);
int zxg_device_count(void);
int zxg_set_device(int dev); /* ZXC_OK or ZXC_B200_ERROR_CUDA */

/* 1 when the pointer is page-locked (cudaHostAlloc / cudaHostRegister / managed) */
int zxg_host_pinned(const void* p);

/* Frame decode with H2D / decode / D2H overlapped over chunks of whole blocks; both host
 * buffers must be page-locked.  Job offsets are relative to h_src / h_dst; h_status gets
 * n_jobs entries.  The output is copied back even for failing jobs (caller decides). */
int zxg_decode_pipelined(zxg_ctx* c, const uint8_t* h_src, uint64_t src_lo, uint64_t src_hi, uint8_t* h_dst,
                         uint64_t produced, const zxc_b200_job_t* h_jobs, uint32_t n_jobs, int32_t* h_status,
                         const void* h_dict, uint32_t dict_size, const void* h_dict_huf, uint32_t block_size,
                         int verify_checksums);

/* Frame decode for ordinary (pageable) host memory, staged through the context's pinned bounce buffers with
 * H2D / decode / D2H and the host copies overlapped.  Source bytes come from h_src (absolute offsets, like the
 * jobs' src_off) or, when `fetch` is given, from fetch(fetch_ctx, dst, len, offset) (ZXC_OK or a negative code):
 * the reader path of zxc_seekable (include/zxc_seekable.h:96-140).  Decoded bytes [clip_lo, clip_hi) (job dst
 * coordinates) land at h_dst. */
typedef int (*zxg_fetch_fn)(void* ctx, void* dst, size_t len, uint64_t off);
int zxg_decode_staged(zxg_ctx* c, const uint8_t* h_src, zxg_fetch_fn fetch, void* fetch_ctx, uint64_t src_lo,
                      uint64_t src_hi, uint8_t* h_dst, uint64_t clip_lo, uint64_t clip_hi, const zxc_b200_job_t* h_jobs,
                      uint32_t n_jobs, int32_t* h_status, const void* h_dict, uint32_t dict_size, const void* h_dict_huf,
                      uint32_t block_size, int verify_checksums);

#ifdef __cplusplus
}
#endif
#endif
