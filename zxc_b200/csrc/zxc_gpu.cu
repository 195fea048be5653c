/*
 * zxc_gpu.cu -- the single CUDA translation unit behind libzxc: kernels (included .cuh files) and
 * the thin extern "C" shim the host C code calls (zxc_gpu.h) -- device bring-up, contexts, staging
 * copies, launches.
 *
 * Decode launches (launch_decode below):
 *   zxc_decode2_kernel  (zxc_decode2.cuh) one CTA per block of <= 64 KiB, window in shared memory,
 *                       cp.async.bulk in and out; takes GLO / GHI blocks with raw sections and RAW blocks
 *   zxc_decode_kernel   (zxc_decode.cuh)  one warp per block, any block size / section encoding;
 *                       runs second over whatever the first kernel deferred, or alone for block
 *                       sizes above 64 KiB and for checksum-verifying decodes
 * Encode: zxc_encode.cuh (levels 1-5), zxc_encode_opt.cuh (levels 6-7).
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#include "zxc_error.h"
#include "zxc_gpu.h"

#include "zxc_decode.cuh"
#include "zxc_decode2.cuh"
#include "zxc_encode.cuh"

/* ========================================================================= */
/* host side: device bring-up, contexts, copies, launches                    */
/* ========================================================================= */
static pthread_once_t g_once = PTHREAD_ONCE_INIT;
static int g_init_rc = ZXC_B200_ERROR_NO_DEVICE;
static int g_sm_count = 0;
static unsigned long long g_launches = 0;
static pthread_mutex_t g_pool_mu = PTHREAD_MUTEX_INITIALIZER;

#define PIN_CHUNK ((size_t)32 << 20)

/* Pageable host buffers are staged through pinned bounce buffers; one thread's memcpy (~10 GB/s)
 * would be far below PCIe Gen5, so large copies are split over a few helper threads. */
#define STAGE_THREADS 6
struct stage_job { void* d; const void* s; size_t n; };
static void* stage_worker(void* p) {
    const stage_job* j = (const stage_job*)p;
    memcpy(j->d, j->s, j->n);
    return NULL;
}
static void parallel_memcpy(void* dst, const void* src, size_t n) {
    if (n < ((size_t)4 << 20)) {
        memcpy(dst, src, n);
        return;
    }
    pthread_t th[STAGE_THREADS];
    stage_job jobs[STAGE_THREADS];
    const size_t per = ((n / STAGE_THREADS) + 4095) & ~(size_t)4095;
    int started = 0;
    size_t off = 0;
    for (int t = 0; t < STAGE_THREADS && off < n; t++) {
        const size_t len = (t == STAGE_THREADS - 1 || off + per > n) ? n - off : per;
        jobs[t].d = (u8*)dst + off;
        jobs[t].s = (const u8*)src + off;
        jobs[t].n = len;
        off += len;
        if (t == 0) continue; /* the calling thread takes the first slice */
        if (pthread_create(&th[started], NULL, stage_worker, &jobs[t]) != 0) {
            memcpy(jobs[t].d, jobs[t].s, jobs[t].n);
            continue;
        }
        started++;
    }
    memcpy(jobs[0].d, jobs[0].s, jobs[0].n);
    for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
}

struct zxg_ctx {
    cudaStream_t stream;
    void* buf[ZXG_BUF_COUNT];
    size_t cap[ZXG_BUF_COUNT];
    void* pin[2];
    cudaEvent_t pin_ev[2];
    unsigned long long* counter; /* [0] work counter, [1..2] reduce output */
    cudaStream_t s_h2d, s_d2h;   /* copy engines for the pipelined frame path (lazily created) */
    struct zxg_ctx* next;
    int device;
};
static zxg_ctx* g_free_list = NULL;

static void init_once(void) {
    int n = 0;
    const cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        fprintf(stderr,
                "libzxc (B200 build): no usable CUDA device (%s); this library has no CPU codec, "
                "codec entry points return ZXC_B200_ERROR_NO_DEVICE\n",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
        g_init_rc = ZXC_B200_ERROR_NO_DEVICE;
        return;
    }
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
        g_init_rc = ZXC_B200_ERROR_CUDA;
        return;
    }
    g_sm_count = prop.multiProcessorCount;
    g_init_rc = ZXC_OK;
}

extern "C" int zxg_init(void) {
    pthread_once(&g_once, init_once);
    return g_init_rc;
}

extern "C" int zxc_b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

extern "C" uint64_t zxc_b200_launch_count(void) { return g_launches; }

extern "C" zxg_ctx* zxg_create(void) {
    if (zxg_init() != ZXC_OK) return NULL;
    zxg_ctx* c = (zxg_ctx*)calloc(1, sizeof(zxg_ctx));
    if (!c) return NULL;
    cudaGetDevice(&c->device);
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMalloc((void**)&c->counter, 4 * sizeof(unsigned long long)) != cudaSuccess) {
        free(c);
        return NULL;
    }
    return c;
}

extern "C" void zxg_destroy(zxg_ctx* c) {
    if (!c) return;
    cudaStreamSynchronize(c->stream);
    for (int i = 0; i < ZXG_BUF_COUNT; i++)
        if (c->buf[i]) cudaFree(c->buf[i]);
    for (int i = 0; i < 2; i++) {
        if (c->pin[i]) cudaFreeHost(c->pin[i]);
        if (c->pin_ev[i]) cudaEventDestroy(c->pin_ev[i]);
    }
    cudaFree(c->counter);
    if (c->s_h2d) cudaStreamDestroy(c->s_h2d);
    if (c->s_d2h) cudaStreamDestroy(c->s_d2h);
    cudaStreamDestroy(c->stream);
    free(c);
}

extern "C" zxg_ctx* zxg_acquire(void) {
    if (zxg_init() != ZXC_OK) return NULL;
    int dev = 0;
    cudaGetDevice(&dev);
    pthread_mutex_lock(&g_pool_mu);
    zxg_ctx** pp = &g_free_list;
    while (*pp && (*pp)->device != dev) pp = &(*pp)->next;
    zxg_ctx* c = *pp;
    if (c) *pp = c->next;
    pthread_mutex_unlock(&g_pool_mu);
    if (c) {
        c->next = NULL;
        return c;
    }
    return zxg_create();
}

extern "C" void zxg_release(zxg_ctx* c) {
    if (!c) return;
    pthread_mutex_lock(&g_pool_mu);
    c->next = g_free_list;
    g_free_list = c;
    pthread_mutex_unlock(&g_pool_mu);
}

extern "C" void* zxg_buffer(zxg_ctx* c, int which, size_t bytes) {
    if (bytes == 0) bytes = 16;
    if (c->cap[which] >= bytes) return c->buf[which];
    if (c->buf[which]) {
        cudaStreamSynchronize(c->stream);
        cudaFree(c->buf[which]);
        c->buf[which] = NULL;
        c->cap[which] = 0;
    }
    size_t want = bytes + (bytes >> 3) + 256; /* slack against regrowth */
    void* p = NULL;
    if (cudaMalloc(&p, want) != cudaSuccess) {
        want = bytes;
        if (cudaMalloc(&p, want) != cudaSuccess) return NULL;
    }
    c->buf[which] = p;
    c->cap[which] = want;
    return p;
}

extern "C" void* zxg_stream(zxg_ctx* c) { return (void*)c->stream; }

extern "C" int zxg_sync(zxg_ctx* c) {
    return cudaStreamSynchronize(c->stream) == cudaSuccess ? ZXC_OK : ZXC_B200_ERROR_CUDA;
}

static int host_is_pinned(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

static int ensure_pins(zxg_ctx* c) {
    for (int i = 0; i < 2; i++) {
        if (!c->pin[i]) {
            if (cudaMallocHost(&c->pin[i], PIN_CHUNK) != cudaSuccess) return ZXC_ERROR_MEMORY;
            if (cudaEventCreateWithFlags(&c->pin_ev[i], cudaEventDisableTiming) != cudaSuccess)
                return ZXC_B200_ERROR_CUDA;
        }
    }
    return ZXC_OK;
}

extern "C" int zxg_h2d(zxg_ctx* c, void* d_dst, const void* h_src, size_t bytes) {
    if (bytes == 0) return ZXC_OK;
    if (host_is_pinned(h_src) || bytes <= (64u << 10)) {
        return cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, c->stream) == cudaSuccess
                   ? ZXC_OK : ZXC_B200_ERROR_CUDA;
    }
    const int rc = ensure_pins(c);
    if (rc != ZXC_OK) return rc;
    size_t done = 0;
    int slot = 0;
    while (done < bytes) {
        const size_t n = bytes - done < PIN_CHUNK ? bytes - done : PIN_CHUNK;
        cudaEventSynchronize(c->pin_ev[slot]); /* previous use of this bounce buffer */
        parallel_memcpy(c->pin[slot], (const u8*)h_src + done, n);
        if (cudaMemcpyAsync((u8*)d_dst + done, c->pin[slot], n, cudaMemcpyHostToDevice, c->stream) != cudaSuccess)
            return ZXC_B200_ERROR_CUDA;
        cudaEventRecord(c->pin_ev[slot], c->stream);
        done += n;
        slot ^= 1;
    }
    return ZXC_OK;
}

extern "C" int zxg_d2h(zxg_ctx* c, void* h_dst, const void* d_src, size_t bytes) {
    if (bytes == 0) return ZXC_OK;
    if (host_is_pinned(h_dst)) {
        if (cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess)
            return ZXC_B200_ERROR_CUDA;
        return zxg_sync(c);
    }
    const int rc = ensure_pins(c);
    if (rc != ZXC_OK) return rc;
    /* double buffer: while chunk k is copied out of its bounce buffer, chunk k+1 is in flight */
    const size_t nchunks = (bytes + PIN_CHUNK - 1) / PIN_CHUNK;
    for (size_t k = 0; k <= nchunks; k++) {
        if (k < nchunks) { /* issue chunk k into slot k&1 (its previous contents were drained at k-1) */
            const size_t off = k * PIN_CHUNK;
            const size_t n = bytes - off < PIN_CHUNK ? bytes - off : PIN_CHUNK;
            if (cudaMemcpyAsync(c->pin[k & 1], (const u8*)d_src + off, n, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess)
                return ZXC_B200_ERROR_CUDA;
            cudaEventRecord(c->pin_ev[k & 1], c->stream);
        }
        if (k > 0) { /* drain chunk k-1 */
            const size_t off = (k - 1) * PIN_CHUNK;
            const size_t n = bytes - off < PIN_CHUNK ? bytes - off : PIN_CHUNK;
            if (cudaEventSynchronize(c->pin_ev[(k - 1) & 1]) != cudaSuccess) return ZXC_B200_ERROR_CUDA;
            parallel_memcpy((u8*)h_dst + off, c->pin[(k - 1) & 1], n);
        }
    }
    return ZXC_OK;
}

/* resident decode CTAs per SM: CTAS_PER_SM unless ZXC_B200_DECODE_CTAS (1..CTAS_PER_SM) says fewer --
 * a tuning knob: fewer warps keep fewer 64 KiB output windows alive in L2 */
static u32 decode_ctas_per_sm(void) {
    static int cached = 0;
    if (cached == 0) {
        const char* e = getenv("ZXC_B200_DECODE_CTAS");
        const int v = e ? atoi(e) : 0;
        cached = (v >= 1 && v <= (int)CTAS_PER_SM) ? v : (int)CTAS_PER_SM;
    }
    return (u32)cached;
}

static int grid_for(u32 n_jobs) {
    const u32 ctas_needed = (n_jobs + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
    const u32 resident = (u32)(g_sm_count > 0 ? g_sm_count : 148) * decode_ctas_per_sm();
    return (int)(ctas_needed < resident ? ctas_needed : resident);
}

/* per-warp scratch for expanded literal sections; 256 bytes of lead-in so word loads may start below it */
static u32 scratch_stride_for(u32 block_size) { return scr_stride(block_size); }

/* ---- block-cooperative kernel: launch geometry by block size ---------------------------------- */
struct D2Config {
    u32 win, gap, rcap, threads, smem, ctas_per_sm, spill_stride;
};
static int d2_enabled(void) {
    static int cached = -1;
    if (cached < 0) {
        const char* e = getenv("ZXC_B200_DECODE_V2");
        cached = (e && e[0] == '0') ? 0 : 1;
    }
    return cached;
}
/* 1 when blocks of `block_size` decoded bytes go through zxc_decode2_kernel */
static int d2_config(u32 block_size, D2Config* c) {
    if (!d2_enabled() || block_size == 0 || block_size > Z2_WIN_MAX) return 0;
    const u32 win = (block_size + Z2_GROUP - 1) & ~(Z2_GROUP - 1);
    const u32 gap = win < 4096u ? win : 4096u;
    const u32 nseq_max = block_size / 5u + 16u + 4u; /* zxc_common.c:142-144 */
    const u32 fixed = d2_off_rec(win, gap) + 4u * 8u;
    const u32 sm_total = 233472u; /* 228 KB per SM, 1 KB reserved per resident CTA */
    u32 n = sm_total / (fixed + nseq_max * 8u + 1024u);
    u32 rcap = nseq_max, spill = 0;
    if (n < 2) { /* two blocks per SM: the tail of very long sequence lists lives in global memory */
        n = 2;
        rcap = (sm_total / 2u - 1024u - fixed) / 8u;
        spill = nseq_max + 4u;
    }
    const u32 threads = win >= 32768u ? 256u : win >= 16384u ? 128u : 64u;
    if (n > 32u) n = 32u;
    if (n * threads > 2048u) n = 2048u / threads;
    c->win = win;
    c->gap = gap;
    c->rcap = rcap;
    c->threads = threads;
    c->smem = d2_smem_bytes(win, gap, rcap);
    c->ctas_per_sm = n;
    c->spill_stride = spill;
    return 1;
}
static size_t d2_spill_bytes(const D2Config* c) {
    return (size_t)(g_sm_count > 0 ? g_sm_count : 148) * c->ctas_per_sm * c->spill_stride * sizeof(z2_rec_t);
}

#define SCRATCH_TAIL (sizeof(unsigned long long) * 4)
#define DEFER_CAP (1u << 16) /* listed deferred jobs; beyond that the second launch scans the status array */

extern "C" size_t zxc_b200_decode_scratch_size(uint32_t block_size) {
    if (zxg_init() != ZXC_OK) return 0;
    const size_t warps = (size_t)g_sm_count * CTAS_PER_SM * WARPS_PER_CTA;
    size_t n = warps * scratch_stride_for(block_size);
    D2Config c;
    if (d2_config(block_size, &c)) n += d2_spill_bytes(&c) + 256 + (size_t)DEFER_CAP * 4 + 256;
    return n + SCRATCH_TAIL;
}

/* d_counter: two 64-bit work counters (zeroed here) */
static int launch_decode(const void* d_src, void* d_dst, const zxc_b200_job_t* d_jobs, u32 n_jobs,
                         i32* d_status, const void* d_dict, u32 dict_size, const void* d_dict_huf,
                         void* d_scratch, size_t scratch_size, u32 block_size, int verify,
                         unsigned long long* d_counter, cudaStream_t st) {
    if (n_jobs == 0) return ZXC_OK;
    DecodeParams P;
    P.src = (const u8*)d_src;
    P.dst = (u8*)d_dst;
    P.jobs = d_jobs;
    P.status = d_status;
    P.dict = (const u8*)d_dict;
    P.dict_huf = (const u8*)d_dict_huf;
    P.scratch = (u8*)d_scratch;
    P.counter = d_counter;
    P.n_jobs = n_jobs;
    P.dict_size = d_dict ? dict_size : 0;
    P.scratch_stride = scratch_stride_for(block_size);
    P.flags = verify ? FLAG_VERIFY : 0;
    P.block_cap = block_size;
    P.defer_list = NULL;
    P.defer_count = NULL;
    P.defer_cap = 0;
    const int grid = grid_for(n_jobs);
    const size_t warp_scratch = (size_t)grid * WARPS_PER_CTA * P.scratch_stride;
    if (warp_scratch > scratch_size) return ZXC_ERROR_MEMORY;
    if (cudaMemsetAsync(d_counter, 0, 3 * sizeof(unsigned long long), st) != cudaSuccess) return ZXC_B200_ERROR_CUDA;
    D2Config c;
    if (!verify && d2_config(block_size, &c)) {
        /* launch 1: one CTA per block, window in shared memory; launch 2: whatever it deferred */
        static int attr_done_smem = 0;
        if (attr_done_smem < (int)c.smem) {
            if (cudaFuncSetAttribute(zxc_decode2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 233472 - 1024) !=
                cudaSuccess)
                return ZXC_B200_ERROR_CUDA;
            attr_done_smem = 233472;
        }
        Decode2Params Q;
        Q.src = P.src;
        Q.dst = P.dst;
        Q.jobs = d_jobs;
        Q.status = d_status;
        Q.dict = P.dict;
        Q.counter = d_counter;
        Q.spill = NULL;
        Q.n_jobs = n_jobs;
        Q.dict_size = P.dict_size;
        Q.win = c.win;
        Q.gap = c.gap;
        Q.rcap = c.rcap;
        Q.spill_stride = c.spill_stride;
        const u32 resident = (u32)(g_sm_count > 0 ? g_sm_count : 148) * c.ctas_per_sm;
        const u32 grid2 = n_jobs < resident ? n_jobs : resident;
        size_t off = (warp_scratch + 255) & ~(size_t)255;
        if (c.spill_stride) {
            if (off + (size_t)grid2 * c.spill_stride * sizeof(z2_rec_t) > scratch_size) return ZXC_ERROR_MEMORY;
            Q.spill = (z2_rec_t*)((u8*)d_scratch + off);
            off = (off + d2_spill_bytes(&c) + 255) & ~(size_t)255;
        }
        /* deferred-job list behind the spill area; its counter is the third work counter */
        Q.defer_count = (u32*)(d_counter + 2);
        Q.defer_cap = off + (size_t)DEFER_CAP * 4 <= scratch_size ? DEFER_CAP : 0u;
        Q.defer_list = (u32*)((u8*)d_scratch + off);
        P.defer_list = Q.defer_list;
        P.defer_count = Q.defer_count;
        P.defer_cap = Q.defer_cap;
        zxc_decode2_kernel<<<grid2, c.threads, c.smem, st>>>(Q);
        __atomic_add_fetch(&g_launches, 1, __ATOMIC_RELAXED);
        if (cudaGetLastError() != cudaSuccess) return ZXC_B200_ERROR_CUDA;
        P.flags |= FLAG_DEFERRED;
        P.counter = d_counter + 1;
    }
    zxc_decode_kernel<<<grid, CTA_THREADS, DECODE_SMEM_BYTES, st>>>(P);
    __atomic_add_fetch(&g_launches, 1, __ATOMIC_RELAXED);
    return cudaGetLastError() == cudaSuccess ? ZXC_OK : ZXC_B200_ERROR_CUDA;
}

/* device scratch one launch over n_jobs blocks needs (without the counter tail) */
static size_t launch_scratch_bytes(u32 n_jobs, u32 block_size) {
    size_t n = (size_t)grid_for(n_jobs) * WARPS_PER_CTA * scratch_stride_for(block_size);
    D2Config c;
    if (d2_config(block_size, &c)) {
        n = (n + 255) & ~(size_t)255;
        if (c.spill_stride) n = (n + d2_spill_bytes(&c) + 255) & ~(size_t)255;
        n += (size_t)DEFER_CAP * 4 + 256;
    }
    return n;
}

extern "C" int zxc_b200_decode_blocks(const void* d_src, void* d_dst, const zxc_b200_job_t* d_jobs,
                                      uint32_t n_jobs, int32_t* d_status, const void* d_dict,
                                      uint32_t dict_size, const void* d_dict_huf, void* d_scratch,
                                      size_t scratch_size, uint32_t block_size, int verify_checksums,
                                      void* stream) {
    const int rc = zxg_init();
    if (rc != ZXC_OK) return rc;
    if (!d_src || !d_dst || !d_jobs || !d_status || !d_scratch) return ZXC_ERROR_NULL_INPUT;
    if (scratch_size < SCRATCH_TAIL) return ZXC_ERROR_MEMORY;
    /* the work counters live in the last 32 bytes of the caller's scratch */
    const size_t usable = (scratch_size - SCRATCH_TAIL) & ~(size_t)7;
    unsigned long long* counter = (unsigned long long*)((u8*)d_scratch + usable);
    return launch_decode(d_src, d_dst, d_jobs, n_jobs, d_status, d_dict, dict_size, d_dict_huf,
                         d_scratch, usable, block_size, verify_checksums, counter, (cudaStream_t)stream);
}

extern "C" int64_t zxc_b200_reduce_status(const int32_t* d_status, const zxc_b200_job_t* d_jobs,
                                          uint32_t n_jobs, void* stream) {
    const int rc = zxg_init();
    if (rc != ZXC_OK) return rc;
    if (n_jobs == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long* d_out = NULL;
    if (cudaMalloc((void**)&d_out, 16) != cudaSuccess) return ZXC_ERROR_MEMORY;
    const unsigned long long init[2] = {~0ull, 0ull};
    cudaMemcpyAsync(d_out, init, 16, cudaMemcpyHostToDevice, st);
    zxc_reduce_kernel<<<(n_jobs + 255) / 256, 256, 0, st>>>(d_status, d_jobs, n_jobs, d_out);
    __atomic_add_fetch(&g_launches, 1, __ATOMIC_RELAXED);
    unsigned long long h[2] = {0, 0};
    cudaMemcpyAsync(h, d_out, 16, cudaMemcpyDeviceToHost, st);
    const cudaError_t e = cudaStreamSynchronize(st);
    int64_t ret;
    if (e != cudaSuccess) {
        ret = ZXC_B200_ERROR_CUDA;
    } else if (h[0] != ~0ull) {
        i32 s = 0;
        cudaMemcpy(&s, d_status + h[0], 4, cudaMemcpyDeviceToHost);
        ret = s < 0 ? s : ZXC_ERROR_CORRUPT_DATA;
    } else {
        ret = (int64_t)h[1];
    }
    cudaFree(d_out);
    return ret;
}

extern "C" int zxg_decode_jobs(zxg_ctx* c, const void* d_src, void* d_dst, const zxc_b200_job_t* h_jobs,
                               uint32_t n_jobs, int32_t* h_status, const void* h_dict, uint32_t dict_size,
                               const void* h_dict_huf, uint32_t block_size, int verify_checksums) {
    if (n_jobs == 0) return ZXC_OK;
    zxc_b200_job_t* d_jobs = (zxc_b200_job_t*)zxg_buffer(c, ZXG_BUF_JOBS, (size_t)n_jobs * sizeof(zxc_b200_job_t));
    i32* d_status = (i32*)zxg_buffer(c, ZXG_BUF_STATUS, (size_t)n_jobs * sizeof(i32));
    const size_t scratch_size = launch_scratch_bytes(n_jobs, block_size);
    void* d_scratch = zxg_buffer(c, ZXG_BUF_SCRATCH, scratch_size);
    if (!d_jobs || !d_status || !d_scratch) return ZXC_ERROR_MEMORY;
    u8* d_dict = NULL;
    u8* d_huf = NULL;
    if (h_dict && dict_size) {
        d_dict = (u8*)zxg_buffer(c, ZXG_BUF_DICT, (size_t)dict_size + 128);
        if (!d_dict) return ZXC_ERROR_MEMORY;
        int rc = zxg_h2d(c, d_dict, h_dict, dict_size);
        if (rc == ZXC_OK && h_dict_huf) {
            d_huf = d_dict + dict_size;
            rc = zxg_h2d(c, d_huf, h_dict_huf, 128);
        }
        if (rc != ZXC_OK) return rc;
    }
    int rc = zxg_h2d(c, d_jobs, h_jobs, (size_t)n_jobs * sizeof(zxc_b200_job_t));
    if (rc != ZXC_OK) return rc;
    rc = launch_decode(d_src, d_dst, d_jobs, n_jobs, d_status, d_dict, dict_size, d_huf, d_scratch,
                       scratch_size, block_size, verify_checksums, c->counter, c->stream);
    if (rc != ZXC_OK) return rc;
    if (cudaMemcpyAsync(h_status, d_status, (size_t)n_jobs * sizeof(i32), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess)
        return ZXC_B200_ERROR_CUDA;
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) {
        fprintf(stderr, "libzxc (B200 build): decode kernel failed: %s\n", cudaGetErrorString(cudaGetLastError()));
        return ZXC_B200_ERROR_CUDA;
    }
    return ZXC_OK;
}


/* ------------------------------------------------------------------------- */
/* Pipelined frame decode for page-locked host buffers: the frame is cut into  */
/* chunks of whole blocks; chunk k's H2D copy, chunk k-1's decode and chunk     */
/* k-2's D2H copy run concurrently on three streams (both PCIe directions and   */
/* the SMs busy at once).  Pageable buffers take the staged path instead.       */
/* ------------------------------------------------------------------------- */
extern "C" int zxg_host_pinned(const void* p) { return host_is_pinned(p); }

extern "C" int zxg_decode_pipelined(zxg_ctx* c, const uint8_t* h_src, uint64_t src_lo, uint64_t src_hi,
                                    uint8_t* h_dst, uint64_t produced, const zxc_b200_job_t* h_jobs,
                                    uint32_t n_jobs, int32_t* h_status, const void* h_dict, uint32_t dict_size,
                                    const void* h_dict_huf, uint32_t block_size, int verify_checksums) {
    if (n_jobs == 0) return ZXC_OK;
    if (!c->s_h2d && cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking) != cudaSuccess) return ZXC_B200_ERROR_CUDA;
    if (!c->s_d2h && cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking) != cudaSuccess) return ZXC_B200_ERROR_CUDA;
    u8* d_in = (u8*)zxg_buffer(c, ZXG_BUF_IN, (size_t)(src_hi - src_lo) + 16);
    u8* d_out = (u8*)zxg_buffer(c, ZXG_BUF_OUT, (size_t)produced + 16);
    zxc_b200_job_t* d_jobs = (zxc_b200_job_t*)zxg_buffer(c, ZXG_BUF_JOBS, (size_t)n_jobs * sizeof(zxc_b200_job_t));
    i32* d_status = (i32*)zxg_buffer(c, ZXG_BUF_STATUS, (size_t)n_jobs * sizeof(i32));
    const size_t scratch_size = launch_scratch_bytes(n_jobs, block_size);
    void* d_scratch = zxg_buffer(c, ZXG_BUF_SCRATCH, scratch_size);
    if (!d_in || !d_out || !d_jobs || !d_status || !d_scratch) return ZXC_ERROR_MEMORY;
    u8* d_dict = NULL;
    u8* d_huf = NULL;
    if (h_dict && dict_size) {
        d_dict = (u8*)zxg_buffer(c, ZXG_BUF_DICT, (size_t)dict_size + 128);
        if (!d_dict) return ZXC_ERROR_MEMORY;
        int rc = zxg_h2d(c, d_dict, h_dict, dict_size);
        if (rc == ZXC_OK && h_dict_huf) {
            d_huf = d_dict + dict_size;
            rc = zxg_h2d(c, d_huf, h_dict_huf, 128);
        }
        if (rc != ZXC_OK) return rc;
    }
    if (cudaMemcpyAsync(d_jobs, h_jobs, (size_t)n_jobs * sizeof(zxc_b200_job_t), cudaMemcpyHostToDevice, c->stream) != cudaSuccess)
        return ZXC_B200_ERROR_CUDA;

    const uint64_t chunk_target = (uint64_t)64 << 20; /* decoded bytes per pipeline stage */
    cudaEvent_t ev_in, ev_dec;
    int rc = ZXC_OK;
    uint32_t j0 = 0;
    while (j0 < n_jobs && rc == ZXC_OK) {
        uint32_t j1 = j0;
        uint64_t acc = 0;
        while (j1 < n_jobs && acc < chunk_target) acc += h_jobs[j1++].dst_cap;
        const uint64_t s0 = h_jobs[j0].src_off, s1 = h_jobs[j1 - 1].src_off + h_jobs[j1 - 1].src_len;
        const uint64_t o0 = h_jobs[j0].dst_off, o1 = h_jobs[j1 - 1].dst_off + h_jobs[j1 - 1].dst_cap;
        if (cudaEventCreateWithFlags(&ev_in, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&ev_dec, cudaEventDisableTiming) != cudaSuccess) {
            rc = ZXC_B200_ERROR_CUDA;
            break;
        }
        if (cudaMemcpyAsync(d_in + (s0 - src_lo), h_src + s0, (size_t)(s1 - s0), cudaMemcpyHostToDevice, c->s_h2d) != cudaSuccess)
            rc = ZXC_B200_ERROR_CUDA;
        cudaEventRecord(ev_in, c->s_h2d);
        cudaStreamWaitEvent(c->stream, ev_in, 0);
        if (rc == ZXC_OK)
            rc = launch_decode(d_in - src_lo, d_out, d_jobs + j0, j1 - j0, d_status + j0, d_dict, dict_size, d_huf,
                               d_scratch, scratch_size, block_size, verify_checksums, c->counter, c->stream);
        cudaEventRecord(ev_dec, c->stream);
        cudaStreamWaitEvent(c->s_d2h, ev_dec, 0);
        if (rc == ZXC_OK &&
            cudaMemcpyAsync(h_dst + o0, d_out + o0, (size_t)(o1 - o0), cudaMemcpyDeviceToHost, c->s_d2h) != cudaSuccess)
            rc = ZXC_B200_ERROR_CUDA;
        cudaEventDestroy(ev_in); /* deferred by the runtime until the events complete */
        cudaEventDestroy(ev_dec);
        j0 = j1;
    }
    if (rc == ZXC_OK &&
        cudaMemcpyAsync(h_status, d_status, (size_t)n_jobs * sizeof(i32), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess)
        rc = ZXC_B200_ERROR_CUDA;
    const cudaError_t e1 = cudaStreamSynchronize(c->stream);
    const cudaError_t e2 = cudaStreamSynchronize(c->s_d2h);
    const cudaError_t e3 = cudaStreamSynchronize(c->s_h2d);
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) {
        fprintf(stderr, "libzxc (B200 build): pipelined decode failed: %s\n", cudaGetErrorString(cudaGetLastError()));
        return ZXC_B200_ERROR_CUDA;
    }
    return rc;
}


/* ------------------------------------------------------------------------- */
/* frame body encode: blocks -> per-block slots -> compacted body             */
/* ------------------------------------------------------------------------- */
static u32 enc_staging_stride(u32 bs) { return ((bs + 8u + 68u + 4u) + 255u) & ~255u; }

/* Encodes src into the frame BODY (all data blocks back to back) in h_body.  h_sizes receives
 * n_blocks on-disk block sizes.  Returns ZXC_OK, or ZXC_ERROR_DST_TOO_SMALL when the body does
 * not fit body_cap (then *body_size holds the size that would have been needed). */
extern "C" int zxg_encode_body(zxg_ctx* c, const uint8_t* h_src, uint64_t src_size, uint32_t block_size, int level,
                               int checksum, uint32_t n_blocks, uint8_t* h_body, uint64_t body_cap,
                               uint32_t* h_sizes, uint64_t* body_size, const void* h_dict, uint32_t dict_size,
                               const uint8_t* h_dict_huf_lens) {
    *body_size = 0;
    if (n_blocks == 0) return ZXC_OK;
    const u32 sstride = enc_staging_stride(block_size);
    const size_t wstride = enc_layout(block_size, level).total;
    const u32 ctas_needed = (n_blocks + ENC_WARPS_PER_CTA - 1) / ENC_WARPS_PER_CTA;
    const u32 resident = (u32)(g_sm_count > 0 ? g_sm_count : 148) * ENC_CTAS_PER_SM;
    const u32 grid = ctas_needed < resident ? ctas_needed : resident;
    u8* d_src = (u8*)zxg_buffer(c, ZXG_BUF_IN, (size_t)src_size + 64);
    u8* d_stage = (u8*)zxg_buffer(c, ZXG_BUF_OUT, (size_t)n_blocks * sstride);
    u8* d_scratch = (u8*)zxg_buffer(c, ZXG_BUF_SCRATCH, (size_t)grid * ENC_WARPS_PER_CTA * wstride);
    u32* d_sizes = (u32*)zxg_buffer(c, ZXG_BUF_STATUS, (size_t)n_blocks * 4);
    unsigned long long* d_offs = (unsigned long long*)zxg_buffer(c, ZXG_BUF_JOBS, (size_t)n_blocks * 8);
    if (!d_src || !d_stage || !d_scratch || !d_sizes || !d_offs) return ZXC_ERROR_MEMORY;
    int rc = zxg_h2d(c, d_src, h_src, (size_t)src_size);
    if (rc != ZXC_OK) return rc;
    if (cudaMemsetAsync(d_src + src_size, 0, 64, c->stream) != cudaSuccess) return ZXC_B200_ERROR_CUDA;
    EncodeParams P;
    P.src = d_src;
    P.staging = d_stage;
    P.out_size = d_sizes;
    P.scratch = d_scratch;
    P.counter = c->counter;
    P.dict = NULL;
    P.seed_head = NULL;
    P.seed_chain = NULL;
    P.dict_huf_lens = NULL;
    if (h_dict && dict_size) {
        /* dictionary + its seeded tables: [dict (padded)] [head 128 KB] [chain 128 KB] */
        const size_t dpad = ((size_t)dict_size + 16 + 255) & ~(size_t)255;
        const size_t dtot = dpad + (size_t)ENC_HASH_SIZE * 4 + (size_t)ENC_WINDOW * 2 + 256;
        u8* d_dict = (u8*)zxg_buffer(c, ZXG_BUF_DICT, dtot);
        if (!d_dict) return ZXC_ERROR_MEMORY;
        if (cudaMemsetAsync(d_dict, 0, dtot, c->stream) != cudaSuccess) return ZXC_B200_ERROR_CUDA;
        rc = zxg_h2d(c, d_dict, h_dict, dict_size);
        if (rc != ZXC_OK) return rc;
        P.dict = d_dict;
        P.seed_head = (const u32*)(d_dict + dpad);
        P.seed_chain = (const unsigned short*)(d_dict + dpad + (size_t)ENC_HASH_SIZE * 4);
        if (h_dict_huf_lens && level >= 6) { /* the shared literal table, one length per byte */
            u8* d_lens = d_dict + dpad + (size_t)ENC_HASH_SIZE * 4 + (size_t)ENC_WINDOW * 2;
            rc = zxg_h2d(c, d_lens, h_dict_huf_lens, 256);
            if (rc != ZXC_OK) return rc;
            P.dict_huf_lens = d_lens;
        }
        zxc_seed_kernel<<<1, 32, 0, c->stream>>>(d_dict, dict_size, (u32)level, (u32*)P.seed_head, (unsigned short*)P.seed_chain);
        __atomic_add_fetch(&g_launches, 1, __ATOMIC_RELAXED);
    }
    P.src_size = src_size;
    P.scratch_stride = wstride;
    P.block_size = block_size;
    P.n_blocks = n_blocks;
    P.staging_stride = sstride;
    P.level = (u32)level;
    P.checksum = checksum ? 1u : 0u;
    P.dict_size = P.dict ? dict_size : 0;
    if (cudaMemsetAsync(c->counter, 0, sizeof(unsigned long long), c->stream) != cudaSuccess) return ZXC_B200_ERROR_CUDA;
    if (level >= 6) zxc_encode_kernel<true><<<grid, ENC_CTA_THREADS, 0, c->stream>>>(P);
    else zxc_encode_kernel<false><<<grid, ENC_CTA_THREADS, 0, c->stream>>>(P);
    __atomic_add_fetch(&g_launches, 1, __ATOMIC_RELAXED);
    if (cudaMemcpyAsync(h_sizes, d_sizes, (size_t)n_blocks * 4, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
        cudaStreamSynchronize(c->stream) != cudaSuccess) {
        fprintf(stderr, "libzxc (B200 build): encode kernel failed: %s\n", cudaGetErrorString(cudaGetLastError()));
        return ZXC_B200_ERROR_CUDA;
    }
    unsigned long long* h_offs = (unsigned long long*)malloc((size_t)n_blocks * 8);
    if (!h_offs) return ZXC_ERROR_MEMORY;
    uint64_t acc = 0;
    for (u32 i = 0; i < n_blocks; i++) {
        h_offs[i] = acc;
        acc += h_sizes[i];
    }
    *body_size = acc;
    if (acc > body_cap) {
        free(h_offs);
        return ZXC_ERROR_DST_TOO_SMALL;
    }
    /* the input buffer is no longer needed: reuse it for the compacted body when it is big enough */
    u8* d_body = (u8*)zxg_buffer(c, ZXG_BUF_AUX, (size_t)acc + 16);
    rc = d_body ? zxg_h2d(c, d_offs, h_offs, (size_t)n_blocks * 8) : ZXC_ERROR_MEMORY;
    if (rc == ZXC_OK) {
        const u32 cgrid = (n_blocks + 7) / 8 < 148u * 8u ? (n_blocks + 7) / 8 : 148u * 8u;
        zxc_compact_kernel<<<cgrid, 256, 0, c->stream>>>(d_stage, sstride, d_offs, d_sizes, d_body, n_blocks);
        __atomic_add_fetch(&g_launches, 1, __ATOMIC_RELAXED);
        rc = zxg_d2h(c, h_body, d_body, (size_t)acc);
        if (rc == ZXC_OK) rc = zxg_sync(c);
    }
    free(h_offs);
    return rc;
}
