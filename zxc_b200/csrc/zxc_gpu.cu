/*
 * zxc_gpu.cu -- the single CUDA translation unit behind libzxc: kernels (included .cuh files) and
 * the thin extern "C" shim the host C code calls (zxc_gpu.h) -- device bring-up, contexts, staging
 * copies, launches.
 *
 * Decode launches (launch_decode below):
 *   zxc_decode_kernel   (zxc_decode.cuh)  one warp per block, any block size / section encoding: every launch by
 *                       default (instances: sequence-centric or output-centric body, with / without dictionary)
 *   zxc_decode2_kernel  (zxc_decode2.cuh) one CTA per block of <= 64 KiB, window in shared memory, cp.async.bulk in
 *                       and out; only with ZXC_B200_DECODE_V2=1 (it measured 10x slower); what
 *                       it defers (entropy-coded sections, checksum verification) goes to zxc_decode_kernel
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

/* ------------------------------------------------------------------------- */
/* Host copy pool.  Pageable caller buffers are staged through pinned bounce  */
/* buffers; one thread's memcpy (~10 GB/s) is far below a PCIe Gen5 link, so   */
/* copies are cut into slices that a few persistent threads claim.  There is   */
/* one pool per device, created on first use, its threads bound to the CPUs of  */
/* the device's NUMA node (/sys/bus/pci/devices/<bdf>/numa_node), which is also */
/* where the bounce buffers are placed: on the 8-GPU hosts GPUs 0-3 hang off    */
/* node 0 and 4-7 off node 1, and unplaced staging memory made the 8-rank       */
/* host-to-host figure of round 1 fall below the 4-rank one.                    */
/* ------------------------------------------------------------------------- */
#include <sched.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <time.h>
#include <unistd.h>

#define POOL_MAX_DEV 16
#define POOL_MAX_THREADS 32
#define POOL_SLICE ((size_t)512 << 10)

extern "C" void zxh_stream_copy(void* dst, const void* src, size_t n); /* zxc_hostcopy.c */

struct copy_pool {
    pthread_mutex_t job_mu; /* one job at a time */
    pthread_mutex_t mu;
    pthread_cond_t cv_go, cv_done;
    pthread_t th[POOL_MAX_THREADS];
    int n_threads, started;
    int numa_node; /* -1 unknown */
    cpu_set_t cpus;
    int have_cpus;
    /* current job */
    u8* d;
    const u8* s;
    size_t bytes;
    size_t next; /* next slice offset (atomic) */
    unsigned long gen;
    int busy;    /* workers still inside the current job */
};
static copy_pool g_pools[POOL_MAX_DEV];
static pthread_mutex_t g_pools_mu = PTHREAD_MUTEX_INITIALIZER;

static int device_numa_node(int dev) {
    char bdf[32];
    if (cudaDeviceGetPCIBusId(bdf, sizeof bdf, dev) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    for (char* p = bdf; *p; p++)
        if (*p >= 'A' && *p <= 'F') *p = (char)(*p - 'A' + 'a');
    char path[128];
    snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bdf);
    FILE* f = fopen(path, "r");
    if (!f) return -1;
    int node = -1;
    if (fscanf(f, "%d", &node) != 1) node = -1;
    fclose(f);
    return node;
}

static int node_cpu_set(int node, cpu_set_t* set) {
    char path[128];
    snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", node);
    FILE* f = fopen(path, "r");
    if (!f) return 0;
    CPU_ZERO(set);
    int a, b, any = 0;
    for (;;) {
        if (fscanf(f, "%d", &a) != 1) break;
        b = a;
        int ch = fgetc(f);
        if (ch == '-') {
            if (fscanf(f, "%d", &b) != 1) break;
            ch = fgetc(f);
        }
        for (int c = a; c <= b && c < CPU_SETSIZE; c++) {
            CPU_SET(c, set);
            any = 1;
        }
        if (ch != ',') break;
    }
    fclose(f);
    /* stay inside what the process may use */
    cpu_set_t allowed;
    if (any && sched_getaffinity(0, sizeof allowed, &allowed) == 0) {
        CPU_AND(set, set, &allowed);
        any = CPU_COUNT(set) > 0;
    }
    return any;
}

static void pool_run_slices(copy_pool* p) {
    for (;;) {
        const size_t off = __atomic_fetch_add(&p->next, POOL_SLICE, __ATOMIC_RELAXED);
        if (off >= p->bytes) break;
        const size_t n = p->bytes - off < POOL_SLICE ? p->bytes - off : POOL_SLICE;
        zxh_stream_copy(p->d + off, p->s + off, n);
    }
}

static void* pool_worker(void* arg) {
    copy_pool* p = (copy_pool*)arg;
    if (p->have_cpus) pthread_setaffinity_np(pthread_self(), sizeof p->cpus, &p->cpus);
    unsigned long seen = 0;
    pthread_mutex_lock(&p->mu);
    for (;;) {
        while (p->gen == seen) pthread_cond_wait(&p->cv_go, &p->mu);
        seen = p->gen;
        pthread_mutex_unlock(&p->mu);
        pool_run_slices(p);
        pthread_mutex_lock(&p->mu);
        if (--p->busy == 0) pthread_cond_signal(&p->cv_done);
    }
    return NULL;
}

static copy_pool* pool_for_device(int dev) {
    if (dev < 0 || dev >= POOL_MAX_DEV) dev = 0;
    copy_pool* p = &g_pools[dev];
    if (__atomic_load_n(&p->started, __ATOMIC_ACQUIRE)) return p;
    pthread_mutex_lock(&g_pools_mu);
    if (!p->started) {
        pthread_mutex_init(&p->job_mu, NULL);
        pthread_mutex_init(&p->mu, NULL);
        pthread_cond_init(&p->cv_go, NULL);
        pthread_cond_init(&p->cv_done, NULL);
        const int async_pool = dev >= POOL_MAX_DEV / 2;
        p->numa_node = device_numa_node(async_pool ? dev - POOL_MAX_DEV / 2 : dev);
        p->have_cpus = p->numa_node >= 0 && node_cpu_set(p->numa_node, &p->cpus);
        long ncpu = p->have_cpus ? CPU_COUNT(&p->cpus) : sysconf(_SC_NPROCESSORS_ONLN);
        const char* e = getenv("ZXC_B200_COPY_THREADS");
        /* per pool an eighth of the node's CPUs: measured on the 2 x 64-thread hosts, more copy threads do not help --
         * the staged path is bound by host memory traffic (9 bytes moved per 2 decoded), and PCIe DMA slows down
         * when 32 threads compete with it (H2D 65 -> 44 ms, D2H 73 -> 49 ms per 2 GiB going from 32 to 12 threads) */
        int want = e ? atoi(e) : (int)(ncpu / 8);
        if (want < 2) want = 2;
        if (want > POOL_MAX_THREADS) want = POOL_MAX_THREADS;
        p->n_threads = 0;
        for (int t = 0; t < want - (async_pool ? 0 : 1); t++) /* a synchronous job counts its caller as a member */
            if (pthread_create(&p->th[p->n_threads], NULL, pool_worker, p) == 0) p->n_threads++;
        __atomic_store_n(&p->started, 1, __ATOMIC_RELEASE);
    }
    pthread_mutex_unlock(&g_pools_mu);
    return p;
}

/* asynchronous variant on the device's second pool (workers only): pool_copy_begin returns at once,
 * pool_copy_end waits for the copy -- lets the caller's drain overlap its next fill */
static copy_pool* pool_copy_begin(int dev, void* dst, const void* src, size_t n) {
    if (dev < 0 || dev >= POOL_MAX_DEV / 2) dev = 0;
    copy_pool* p = pool_for_device(dev + POOL_MAX_DEV / 2);
    if (p->n_threads == 0 || n < ((size_t)2 << 20)) {
        zxh_stream_copy(dst, src, n);
        return NULL;
    }
    pthread_mutex_lock(&p->job_mu);
    pthread_mutex_lock(&p->mu);
    p->d = (u8*)dst;
    p->s = (const u8*)src;
    p->bytes = n;
    p->next = 0;
    p->busy = p->n_threads;
    p->gen++;
    pthread_cond_broadcast(&p->cv_go);
    pthread_mutex_unlock(&p->mu);
    return p;
}
static void pool_copy_end(copy_pool* p) {
    if (!p) return;
    pthread_mutex_lock(&p->mu);
    while (p->busy) pthread_cond_wait(&p->cv_done, &p->mu);
    pthread_mutex_unlock(&p->mu);
    pthread_mutex_unlock(&p->job_mu);
}

static void pool_memcpy(int dev, void* dst, const void* src, size_t n) {
    if (n < ((size_t)2 << 20)) {
        zxh_stream_copy(dst, src, n);
        return;
    }
    if (dev < 0 || dev >= POOL_MAX_DEV / 2) dev = 0;
    copy_pool* p = pool_for_device(dev);
    pthread_mutex_lock(&p->job_mu);
    pthread_mutex_lock(&p->mu);
    p->d = (u8*)dst;
    p->s = (const u8*)src;
    p->bytes = n;
    p->next = 0;
    p->busy = p->n_threads;
    p->gen++;
    pthread_cond_broadcast(&p->cv_go);
    pthread_mutex_unlock(&p->mu);
    pool_run_slices(p);
    pthread_mutex_lock(&p->mu);
    while (p->busy) pthread_cond_wait(&p->cv_done, &p->mu);
    pthread_mutex_unlock(&p->mu);
    pthread_mutex_unlock(&p->job_mu);
}

/* page-locked memory on the device's NUMA node: anonymous mapping, preferred-node policy, touched,
 * then registered with CUDA.  Falls back to cudaMallocHost when the node is unknown. */
struct pinned_buf {
    void* p;
    size_t bytes;
    int mapped; /* 1: mmap + cudaHostRegister, 0: cudaMallocHost */
};
static int pinned_alloc(pinned_buf* b, size_t bytes, int node) {
    b->p = NULL;
    b->bytes = bytes;
    b->mapped = 0;
    /* the driver allocates in the calling task's context: a preferred-node policy around the call puts the pages
     * next to the GPU; cudaMallocHost memory also DMAs faster than a registered 4 KiB-page mapping (measured) */
    int policy_set = 0;
#ifdef SYS_set_mempolicy
    if (node >= 0 && node < 64) {
        unsigned long mask = 1ul << node;
        policy_set = syscall(SYS_set_mempolicy, 1 /* MPOL_PREFERRED */, &mask, sizeof(mask) * 8) == 0;
    }
#endif
    const cudaError_t e = cudaMallocHost(&b->p, bytes);
#ifdef SYS_set_mempolicy
    if (policy_set) syscall(SYS_set_mempolicy, 0 /* MPOL_DEFAULT */, NULL, 0);
#endif
    if (e != cudaSuccess) {
        b->p = NULL;
        return ZXC_ERROR_MEMORY;
    }
    return ZXC_OK;
}
static void pinned_free(pinned_buf* b) {
    if (!b->p) return;
    if (b->mapped) {
        cudaHostUnregister(b->p);
        munmap(b->p, b->bytes);
    } else {
        cudaFreeHost(b->p);
    }
    b->p = NULL;
}

#define STAGE_SLOTS 4
#define EV_RING 8

struct zxg_ctx {
    cudaStream_t stream;
    void* buf[ZXG_BUF_COUNT];
    size_t cap[ZXG_BUF_COUNT];
    pinned_buf pinb[2];
    void* pin[2];
    cudaEvent_t pin_ev[2];
    unsigned long long* counter; /* [0..1] work counters, [2] deferred-job counter */
    unsigned long long* reduce_out; /* device: [0] first bad index, [1] byte sum (zxc_b200_reduce_status) */
    cudaStream_t s_h2d, s_d2h;   /* copy engines for the pipelined frame paths (lazily created) */
    cudaStream_t s_dec[STAGE_SLOTS]; /* staged path: chunk decodes overlap (a 512-block launch is latency-bound) */
    cudaEvent_t ev_ring[EV_RING]; /* reused by the pipelines: no event is created per chunk */
    pinned_buf st_in[STAGE_SLOTS], st_out[STAGE_SLOTS]; /* staging for pageable callers (lazily allocated) */
    cudaEvent_t st_ev_in[STAGE_SLOTS], st_ev_dec[STAGE_SLOTS], st_ev_out[STAGE_SLOTS];
    int st_ready;
    struct zxg_ctx* next;
    int device;
    int numa_node;
    int sm_count;
    int trimmed; /* idle and already cut back by zxg_release */
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

extern "C" int zxg_current_device(void) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return dev;
}
extern "C" int zxg_device_count(void) { return zxc_b200_device_count(); }
extern "C" int zxg_set_device(int dev) {
    if (cudaSetDevice(dev) != cudaSuccess) {
        cudaGetLastError();
        return ZXC_B200_ERROR_CUDA;
    }
    return ZXC_OK;
}

extern "C" zxg_ctx* zxg_create(void) {
    if (zxg_init() != ZXC_OK) return NULL;
    zxg_ctx* c = (zxg_ctx*)calloc(1, sizeof(zxg_ctx));
    if (!c) return NULL;
    cudaGetDevice(&c->device);
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMalloc((void**)&c->counter, 32 * sizeof(unsigned long long)) != cudaSuccess) {
        free(c);
        return NULL;
    }
    c->reduce_out = c->counter + 28;
    c->numa_node = device_numa_node(c->device);
    int smc = 0;
    if (cudaDeviceGetAttribute(&smc, cudaDevAttrMultiProcessorCount, c->device) != cudaSuccess || smc <= 0) smc = g_sm_count;
    c->sm_count = smc;
    return c;
}

extern "C" void zxg_destroy(zxg_ctx* c) {
    if (!c) return;
    cudaStreamSynchronize(c->stream);
    for (int i = 0; i < ZXG_BUF_COUNT; i++)
        if (c->buf[i]) cudaFree(c->buf[i]);
    for (int i = 0; i < 2; i++) {
        pinned_free(&c->pinb[i]);
        if (c->pin_ev[i]) cudaEventDestroy(c->pin_ev[i]);
    }
    for (int i = 0; i < STAGE_SLOTS; i++) {
        pinned_free(&c->st_in[i]);
        pinned_free(&c->st_out[i]);
        if (c->st_ev_in[i]) cudaEventDestroy(c->st_ev_in[i]);
        if (c->st_ev_dec[i]) cudaEventDestroy(c->st_ev_dec[i]);
        if (c->st_ev_out[i]) cudaEventDestroy(c->st_ev_out[i]);
    }
    for (int i = 0; i < EV_RING; i++)
        if (c->ev_ring[i]) cudaEventDestroy(c->ev_ring[i]);
    cudaFree(c->counter);
    for (int i = 0; i < STAGE_SLOTS; i++)
        if (c->s_dec[i]) cudaStreamDestroy(c->s_dec[i]);
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

/* Gives back what an idle context holds beyond `keep` bytes per buffer (and its staging slots). */
static void ctx_trim(zxg_ctx* c, size_t keep) {
    cudaStreamSynchronize(c->stream);
    for (int i = 0; i < ZXG_BUF_COUNT; i++)
        if (c->buf[i] && c->cap[i] > keep) {
            cudaFree(c->buf[i]);
            c->buf[i] = NULL;
            c->cap[i] = 0;
        }
    for (int i = 0; i < STAGE_SLOTS; i++) {
        pinned_free(&c->st_in[i]);
        pinned_free(&c->st_out[i]);
    }
}

/* The free list keeps contexts warm: the two most recently released contexts of a device keep their buffers (a
 * caller that decodes frame after frame should not pay cudaMalloc each time); older idle ones -- left behind by a
 * burst of concurrent callers -- are trimmed to POOL_KEEP_BYTES per buffer so that one multi-GiB frame does not pin
 * several GiB of HBM per past thread for the life of the process. */
#define POOL_WARM_PER_DEVICE 2
#define POOL_KEEP_BYTES ((size_t)64 << 20)
extern "C" void zxg_release(zxg_ctx* c) {
    if (!c) return;
    pthread_mutex_lock(&g_pool_mu);
    c->next = g_free_list;
    g_free_list = c;
    int seen = 0;
    for (zxg_ctx* q = g_free_list; q; q = q->next) {
        if (q->device != c->device) continue;
        if (++seen > POOL_WARM_PER_DEVICE && !q->trimmed) {
            ctx_trim(q, POOL_KEEP_BYTES);
            q->trimmed = 1;
        }
    }
    c->trimmed = 0;
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
            if (pinned_alloc(&c->pinb[i], PIN_CHUNK, c->numa_node) != ZXC_OK) return ZXC_ERROR_MEMORY;
            c->pin[i] = c->pinb[i].p;
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
        pool_memcpy(c->device, c->pin[slot], (const u8*)h_src + done, n);
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
            pool_memcpy(c->device, (u8*)h_dst + off, c->pin[(k - 1) & 1], n);
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
        cached = (e && e[0] == '1') ? 1 : 0; /* off: measured slower than the warp-per-block kernel */
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
    {
        static int units_mode = -2; /* ZXC_B200_UNITS: 1 = always, 0 = never, unset = by measured rule */
        if (units_mode == -2) {
            const char* e = getenv("ZXC_B200_UNITS");
            units_mode = e ? (e[0] == '1' ? 1 : 0) : -1;
        }
        if (units_mode == 1) P.flags |= FLAG_UNITS_ON;
        if (units_mode == 0) P.flags |= FLAG_UNITS_OFF;
    }
    /* The output-centric body (zxc_decode_units.cuh) was the faster one for dictionary decodes of small blocks while the
     * sequence-centric body copied dictionary sources byte by byte (298 vs 152 GB/s on 4 KiB records).  Since that body
     * reads them as ordinary global sources and copies long items as balanced chunks it wins there too (369 vs 321 GB/s,
     * DESIGN.md), so the unit walk runs only on request (ZXC_B200_UNITS=1). */
    const bool units = (P.flags & FLAG_UNITS_ON) != 0;
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
        Q.trace = NULL;
        const int want_trace = getenv("ZXC_B200_D2_TRACE") != NULL; /* development: per-phase cycle counts */
        if (want_trace && cudaMalloc((void**)&Q.trace, (size_t)n_jobs * 128) == cudaSuccess)
            cudaMemsetAsync(Q.trace, 0, (size_t)n_jobs * 128, st);
        zxc_decode2_kernel<<<grid2, c.threads, c.smem, st>>>(Q);
        __atomic_add_fetch(&g_launches, 1, __ATOMIC_RELAXED);
        if (cudaGetLastError() != cudaSuccess) return ZXC_B200_ERROR_CUDA;
        if (Q.trace) {
            unsigned long long* h = (unsigned long long*)malloc((size_t)n_jobs * 128);
            cudaStreamSynchronize(st);
            cudaMemcpy(h, Q.trace, (size_t)n_jobs * 128, cudaMemcpyDeviceToHost);
            double acc[6] = {0, 0, 0, 0, 0, 0};
            u32 cnt = 0;
            for (u32 i = 0; i < n_jobs; i++) {
                const unsigned long long* t = h + (size_t)i * 8;
                if (!t[6] || !t[5]) continue; /* deferred / raw / failed jobs leave the later stamps empty */
                for (int q = 0; q < 6; q++) acc[q] += (double)(t[q + 1] - t[q]);
                cnt++;
            }
            if (cnt)
                fprintf(stderr, "d2 trace (%u blocks, mean cycles): header+issue %.0f | load wait %.0f | extras %.0f | records %.0f | "
                                "words %.0f | drain %.0f\n", cnt, acc[0] / cnt, acc[1] / cnt, acc[2] / cnt, acc[3] / cnt,
                        acc[4] / cnt, acc[5] / cnt);
            double w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (u32 i = 0; i < n_jobs; i++) {
                const unsigned long long* t = h + (size_t)i * 8;
                if (!t[6] || !t[5]) continue;
                for (int q = 0; q < 8; q++) w[q] += (double)h[(size_t)n_jobs * 8 + (size_t)i * 8 + q];
            }
            if (cnt)
                fprintf(stderr, "d2 trace warp 0 per block: steps %.1f rounds %.1f sleeps %.1f | cycles: guard %.0f plan %.0f rounds %.0f "
                                "sleep %.0f finish %.0f\n", w[7] / cnt, w[5] / cnt, w[6] / cnt, w[0] / cnt, w[1] / cnt, w[2] / cnt,
                        w[3] / cnt, w[4] / cnt);
            free(h);
            cudaFree(Q.trace);
        }
        P.flags |= FLAG_DEFERRED;
        P.counter = d_counter + 1;
    }
    /* the dictionary-free instance carries neither the dictionary pointer nor its source classification (zxc_decode.cuh) */
    const bool has_dict = P.dict != NULL && P.dict_size != 0;
    if (P.flags & FLAG_DEFERRED) {
        if (has_dict) zxc_decode_kernel<false, true, true><<<grid, CTA_THREADS, DECODE_SMEM_BYTES, st>>>(P);
        else zxc_decode_kernel<false, true, false><<<grid, CTA_THREADS, DECODE_SMEM_BYTES, st>>>(P);
    } else if (units) {
        if (has_dict) zxc_decode_kernel<true, false, true><<<grid, CTA_THREADS, DECODE_SMEM_BYTES, st>>>(P);
        else zxc_decode_kernel<true, false, false><<<grid, CTA_THREADS, DECODE_SMEM_BYTES, st>>>(P);
    } else {
        if (has_dict) zxc_decode_kernel<false, false, true><<<grid, CTA_THREADS, DECODE_SMEM_BYTES, st>>>(P);
        else zxc_decode_kernel<false, false, false><<<grid, CTA_THREADS, DECODE_SMEM_BYTES, st>>>(P);
    }
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

/* one 16-byte result slot per device, allocated on first use and kept: no cudaMalloc / cudaFree per call */
static unsigned long long* g_reduce_dev[POOL_MAX_DEV];
static pthread_mutex_t g_reduce_mu = PTHREAD_MUTEX_INITIALIZER;

extern "C" int64_t zxc_b200_reduce_status(const int32_t* d_status, const zxc_b200_job_t* d_jobs,
                                          uint32_t n_jobs, void* stream) {
    const int rc = zxg_init();
    if (rc != ZXC_OK) return rc;
    if (n_jobs == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= POOL_MAX_DEV) dev = 0;
    pthread_mutex_lock(&g_reduce_mu);
    if (!g_reduce_dev[dev] && cudaMalloc((void**)&g_reduce_dev[dev], 16) != cudaSuccess) {
        g_reduce_dev[dev] = NULL;
        pthread_mutex_unlock(&g_reduce_mu);
        return ZXC_ERROR_MEMORY;
    }
    unsigned long long* d_out = g_reduce_dev[dev];
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
    pthread_mutex_unlock(&g_reduce_mu);
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
    /* the job table goes up chunk by chunk, just ahead of each launch: one copy of the whole table (24 MB for a million
     * 4 KiB records, staged synchronously by the driver when the table is pageable) would sit in front of the first
     * H2D of payload */
    const uint64_t chunk_target = (uint64_t)64 << 20; /* decoded bytes per pipeline stage */
    for (int i = 0; i < EV_RING; i++)
        if (!c->ev_ring[i] && cudaEventCreateWithFlags(&c->ev_ring[i], cudaEventDisableTiming) != cudaSuccess)
            return ZXC_B200_ERROR_CUDA;
    int rc = ZXC_OK;
    uint32_t j0 = 0, chunk_no = 0;
    while (j0 < n_jobs && rc == ZXC_OK) {
        /* events are recycled: a wait refers to the record that preceded it, so reuse is safe */
        const cudaEvent_t ev_in = c->ev_ring[(2u * chunk_no) % EV_RING], ev_dec = c->ev_ring[(2u * chunk_no + 1u) % EV_RING];
        chunk_no++;
        uint32_t j1 = j0;
        uint64_t acc = 0;
        while (j1 < n_jobs && acc < chunk_target) acc += h_jobs[j1++].dst_cap;
        const uint64_t s0 = h_jobs[j0].src_off, s1 = h_jobs[j1 - 1].src_off + h_jobs[j1 - 1].src_len;
        const uint64_t o0 = h_jobs[j0].dst_off, o1 = h_jobs[j1 - 1].dst_off + h_jobs[j1 - 1].dst_cap;
        if (cudaMemcpyAsync(d_in + (s0 - src_lo), h_src + s0, (size_t)(s1 - s0), cudaMemcpyHostToDevice, c->s_h2d) != cudaSuccess)
            rc = ZXC_B200_ERROR_CUDA;
        cudaEventRecord(ev_in, c->s_h2d);
        if (cudaMemcpyAsync(d_jobs + j0, h_jobs + j0, (size_t)(j1 - j0) * sizeof(zxc_b200_job_t), cudaMemcpyHostToDevice,
                            c->stream) != cudaSuccess)
            rc = ZXC_B200_ERROR_CUDA;
        cudaStreamWaitEvent(c->stream, ev_in, 0);
        if (rc == ZXC_OK)
            rc = launch_decode(d_in - src_lo, d_out, d_jobs + j0, j1 - j0, d_status + j0, d_dict, dict_size, d_huf,
                               d_scratch, scratch_size, block_size, verify_checksums, c->counter, c->stream);
        cudaEventRecord(ev_dec, c->stream);
        cudaStreamWaitEvent(c->s_d2h, ev_dec, 0);
        if (rc == ZXC_OK &&
            cudaMemcpyAsync(h_dst + o0, d_out + o0, (size_t)(o1 - o0), cudaMemcpyDeviceToHost, c->s_d2h) != cudaSuccess)
            rc = ZXC_B200_ERROR_CUDA;
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
/* Staged frame decode for ordinary (pageable) caller memory: the same three     */
/* streams as zxg_decode_pipelined, with pinned bounce buffers either side.  The  */
/* host thread fills input slot k (copy pool, or the caller's read_at), queues    */
/* H2D / decode / D2H for chunk k, then drains output slot k-2 into the caller's   */
/* buffer while the GPU works -- so PCIe in, the SMs, PCIe out and the host copies */
/* all overlap.  Decoded coordinates: job dst offsets; [clip_lo, clip_hi) of them  */
/* is what the caller wants at h_dst (ranges start and end inside blocks).         */
/* ------------------------------------------------------------------------- */
#define STAGE_OUT ((size_t)32 << 20)
#define STAGE_IN (STAGE_OUT + ((size_t)1 << 20))

static int staged_ready(zxg_ctx* c) {
    if (c->st_ready) return ZXC_OK;
    for (int i = 0; i < STAGE_SLOTS; i++) {
        if (!c->st_in[i].p && pinned_alloc(&c->st_in[i], STAGE_IN, c->numa_node) != ZXC_OK) return ZXC_ERROR_MEMORY;
        if (!c->st_out[i].p && pinned_alloc(&c->st_out[i], STAGE_OUT, c->numa_node) != ZXC_OK) return ZXC_ERROR_MEMORY;
        if (!c->st_ev_in[i] && cudaEventCreateWithFlags(&c->st_ev_in[i], cudaEventDisableTiming) != cudaSuccess) return ZXC_B200_ERROR_CUDA;
        if (!c->st_ev_dec[i] && cudaEventCreateWithFlags(&c->st_ev_dec[i], cudaEventDisableTiming) != cudaSuccess) return ZXC_B200_ERROR_CUDA;
        if (!c->st_ev_out[i] && cudaEventCreateWithFlags(&c->st_ev_out[i], cudaEventDisableTiming) != cudaSuccess) return ZXC_B200_ERROR_CUDA;
    }
    if (!c->s_h2d && cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking) != cudaSuccess) return ZXC_B200_ERROR_CUDA;
    if (!c->s_d2h && cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking) != cudaSuccess) return ZXC_B200_ERROR_CUDA;
    for (int i = 0; i < STAGE_SLOTS; i++)
        if (!c->s_dec[i] && cudaStreamCreateWithFlags(&c->s_dec[i], cudaStreamNonBlocking) != cudaSuccess) return ZXC_B200_ERROR_CUDA;
    c->st_ready = 1;
    return ZXC_OK;
}

struct staged_chunk {
    uint32_t j0, j1;
    uint64_t s0, s1;   /* source byte range */
    uint64_t c0, c1;   /* clipped decoded range delivered to the caller */
};

extern "C" int zxg_decode_staged(zxg_ctx* c, const uint8_t* h_src, zxg_fetch_fn fetch, void* fetch_ctx, uint64_t src_lo,
                                 uint64_t src_hi, uint8_t* h_dst, uint64_t clip_lo, uint64_t clip_hi,
                                 const zxc_b200_job_t* h_jobs, uint32_t n_jobs, int32_t* h_status, const void* h_dict,
                                 uint32_t dict_size, const void* h_dict_huf, uint32_t block_size, int verify_checksums) {
    if (n_jobs == 0) return ZXC_OK;
    struct timespec tw0, tw1, tw2;
    clock_gettime(CLOCK_MONOTONIC, &tw0);
    int rc = staged_ready(c);
    if (rc != ZXC_OK) return rc;
    const uint64_t produced = h_jobs[n_jobs - 1].dst_off + h_jobs[n_jobs - 1].dst_cap;
    u8* d_in = (u8*)zxg_buffer(c, ZXG_BUF_IN, (size_t)(src_hi - src_lo) + 16);
    u8* d_out = (u8*)zxg_buffer(c, ZXG_BUF_OUT, (size_t)produced + 16);
    zxc_b200_job_t* d_jobs = (zxc_b200_job_t*)zxg_buffer(c, ZXG_BUF_JOBS, (size_t)n_jobs * sizeof(zxc_b200_job_t));
    i32* d_status = (i32*)zxg_buffer(c, ZXG_BUF_STATUS, (size_t)n_jobs * sizeof(i32));
    /* a chunk holds at most STAGE_OUT / (smallest decoded block) jobs; every slot decodes on its own stream with its
     * own scratch region and work counters, so up to STAGE_SLOTS chunk launches are resident together */
    uint32_t chunk_jobs_max = 1;
    {
        uint32_t j = 0;
        while (j < n_jobs) {
            uint32_t j1 = j;
            uint64_t ao = 0, ai = 0;
            while (j1 < n_jobs && ao + h_jobs[j1].dst_cap <= STAGE_OUT && ai + h_jobs[j1].src_len <= STAGE_IN) {
                ao += h_jobs[j1].dst_cap;
                ai += h_jobs[j1].src_len;
                j1++;
            }
            if (j1 == j) break;
            if (j1 - j > chunk_jobs_max) chunk_jobs_max = j1 - j;
            j = j1;
        }
    }
    const size_t scratch_size = (launch_scratch_bytes(chunk_jobs_max, block_size) + 255) & ~(size_t)255;
    u8* d_scratch = (u8*)zxg_buffer(c, ZXG_BUF_SCRATCH, scratch_size * STAGE_SLOTS);
    if (!d_in || !d_out || !d_jobs || !d_status || !d_scratch) return ZXC_ERROR_MEMORY;
    u8* d_dict = NULL;
    u8* d_huf = NULL;
    if (h_dict && dict_size) {
        d_dict = (u8*)zxg_buffer(c, ZXG_BUF_DICT, (size_t)dict_size + 128);
        if (!d_dict) return ZXC_ERROR_MEMORY;
        rc = zxg_h2d(c, d_dict, h_dict, dict_size);
        if (rc == ZXC_OK && h_dict_huf) {
            d_huf = d_dict + dict_size;
            rc = zxg_h2d(c, d_huf, h_dict_huf, 128);
        }
        if (rc != ZXC_OK) return rc;
    }
    if (cudaMemcpyAsync(d_jobs, h_jobs, (size_t)n_jobs * sizeof(zxc_b200_job_t), cudaMemcpyHostToDevice, c->stream) != cudaSuccess)
        return ZXC_B200_ERROR_CUDA;
    cudaStreamSynchronize(c->stream); /* h_jobs may be pageable: do not let the caller free it under the copy */

    staged_chunk ring[STAGE_SLOTS];
    uint32_t j0 = 0, issued = 0, drained = 0;
    copy_pool* drain_job = NULL; /* the drain in flight on the second pool ... */
    int drain_slot = -1;         /* ... and the output slot it is reading */
    const int trace = getenv("ZXC_B200_STAGE_TRACE") != NULL; /* development: where the host thread's time goes */
    double t_fill = 0, t_drain = 0, t_wait_in = 0, t_wait_out = 0, t_launch = 0;
    cudaEvent_t tev[STAGE_SLOTS][6]; /* trace only: H2D, decode, D2H brackets of the chunk in each slot */
    float g_h2d = 0, g_dec = 0, g_d2h = 0, g_lat = 0;
    if (trace)
        for (int i = 0; i < STAGE_SLOTS; i++)
            for (int q = 0; q < 6; q++) cudaEventCreate(&tev[i][q]);
    struct timespec ts0, ts1;
#define STG_T0() do { if (trace) clock_gettime(CLOCK_MONOTONIC, &ts0); } while (0)
#define STG_T1(acc) do { if (trace) { clock_gettime(CLOCK_MONOTONIC, &ts1); acc += (ts1.tv_sec - ts0.tv_sec) + 1e-9 * (ts1.tv_nsec - ts0.tv_nsec); } } while (0)
    while (rc == ZXC_OK && (drained < issued || j0 < n_jobs)) {
        /* one slot is always left to the drain in flight, so filling the next chunk never waits for it */
        if (j0 < n_jobs && issued - drained < STAGE_SLOTS - 1) { /* a slot is free: stage and queue the next chunk */
            const int slot = (int)(issued % STAGE_SLOTS);
            staged_chunk ch;
            ch.j0 = j0;
            uint32_t j1 = j0;
            uint64_t acc_out = 0, acc_in = 0;
            while (j1 < n_jobs && acc_out + h_jobs[j1].dst_cap <= STAGE_OUT && acc_in + h_jobs[j1].src_len <= STAGE_IN) {
                acc_out += h_jobs[j1].dst_cap;
                acc_in += h_jobs[j1].src_len;
                j1++;
            }
            if (j1 == j0) { /* a single block larger than a stage: cannot happen with <= 2 MiB blocks */
                rc = ZXC_ERROR_MEMORY;
                break;
            }
            ch.j1 = j1;
            ch.s0 = h_jobs[j0].src_off;
            ch.s1 = h_jobs[j1 - 1].src_off + h_jobs[j1 - 1].src_len;
            const uint64_t o0 = h_jobs[j0].dst_off, o1 = h_jobs[j1 - 1].dst_off + h_jobs[j1 - 1].dst_cap;
            ch.c0 = o0 > clip_lo ? o0 : clip_lo;
            ch.c1 = o1 < clip_hi ? o1 : clip_hi;
            if (ch.c1 < ch.c0) ch.c1 = ch.c0;
            if (drain_job && drain_slot == slot) { /* this chunk's D2H will overwrite what the drain is still reading */
                pool_copy_end(drain_job);
                drain_job = NULL;
            }
            STG_T0();
            cudaEventSynchronize(c->st_ev_in[slot]); /* the slot's previous H2D has left the bounce buffer */
            STG_T1(t_wait_in);
            u8* pin = (u8*)c->st_in[slot].p;
            STG_T0();
            if (fetch) {
                rc = fetch(fetch_ctx, pin, (size_t)(ch.s1 - ch.s0), ch.s0);
                if (rc != ZXC_OK) break;
            } else {
                pool_memcpy(c->device, pin, h_src + ch.s0, (size_t)(ch.s1 - ch.s0));
            }
            STG_T1(t_fill);
            STG_T0();
            if (trace) cudaEventRecord(tev[slot][0], c->s_h2d);
            if (cudaMemcpyAsync(d_in + (ch.s0 - src_lo), pin, (size_t)(ch.s1 - ch.s0), cudaMemcpyHostToDevice, c->s_h2d) != cudaSuccess) {
                rc = ZXC_B200_ERROR_CUDA;
                break;
            }
            if (trace) cudaEventRecord(tev[slot][1], c->s_h2d);
            cudaEventRecord(c->st_ev_in[slot], c->s_h2d);
            cudaStreamWaitEvent(c->s_dec[slot], c->st_ev_in[slot], 0);
            if (trace) cudaEventRecord(tev[slot][2], c->s_dec[slot]);
            rc = launch_decode(d_in - src_lo, d_out, d_jobs + j0, j1 - j0, d_status + j0, d_dict, dict_size, d_huf,
                               d_scratch + (size_t)slot * scratch_size, scratch_size, block_size, verify_checksums,
                               c->counter + 4 * slot, c->s_dec[slot]);
            if (rc != ZXC_OK) break;
            if (trace) cudaEventRecord(tev[slot][3], c->s_dec[slot]);
            cudaEventRecord(c->st_ev_dec[slot], c->s_dec[slot]);
            cudaStreamWaitEvent(c->s_d2h, c->st_ev_dec[slot], 0);
            if (trace) cudaEventRecord(tev[slot][4], c->s_d2h);
            if (ch.c1 > ch.c0 &&
                cudaMemcpyAsync(c->st_out[slot].p, d_out + ch.c0, (size_t)(ch.c1 - ch.c0), cudaMemcpyDeviceToHost, c->s_d2h) != cudaSuccess) {
                rc = ZXC_B200_ERROR_CUDA;
                break;
            }
            if (trace) cudaEventRecord(tev[slot][5], c->s_d2h);
            cudaEventRecord(c->st_ev_out[slot], c->s_d2h);
            ring[slot] = ch;
            j0 = j1;
            issued++;
            STG_T1(t_launch);
            if (j0 < n_jobs && issued - drained < STAGE_SLOTS - 1) continue; /* fill the pipeline before draining */
        }
        /* hand the oldest finished chunk to the caller while the GPU works on the younger ones */
        const int ds = (int)(drained % STAGE_SLOTS);
        STG_T0();
        if (cudaEventSynchronize(c->st_ev_out[ds]) != cudaSuccess) {
            rc = ZXC_B200_ERROR_CUDA;
            break;
        }
        STG_T1(t_wait_out);
        if (trace) {
            float a = 0, b = 0, d2 = 0, l = 0;
            cudaEventElapsedTime(&a, tev[ds][0], tev[ds][1]);
            cudaEventElapsedTime(&b, tev[ds][2], tev[ds][3]);
            cudaEventElapsedTime(&d2, tev[ds][4], tev[ds][5]);
            cudaEventElapsedTime(&l, tev[ds][0], tev[ds][5]);
            g_h2d += a; g_dec += b; g_d2h += d2; g_lat += l;
        }
        const staged_chunk& dc = ring[ds];
        STG_T0();
        pool_copy_end(drain_job); /* one drain at a time; the previous one overlapped the fill above */
        drain_job = NULL;
        if (dc.c1 > dc.c0) drain_job = pool_copy_begin(c->device, h_dst + (dc.c0 - clip_lo), c->st_out[ds].p, (size_t)(dc.c1 - dc.c0));
        drain_slot = ds;
        STG_T1(t_drain);
        drained++;
    }
    clock_gettime(CLOCK_MONOTONIC, &tw1);
    pool_copy_end(drain_job);
    clock_gettime(CLOCK_MONOTONIC, &tw2);
    if (trace)
        fprintf(stderr, "staged decode: %u chunks; host seconds: fill %.4f, enqueue %.4f, drain %.4f, waiting for H2D slot %.4f, "
                        "waiting for D2H %.4f; loop %.4f, last drain %.4f; device ms summed over chunks: H2D %.1f decode %.1f D2H %.1f, "
                        "H2D start to D2H end %.1f\n", issued, t_fill, t_launch, t_drain, t_wait_in, t_wait_out,
                (tw1.tv_sec - tw0.tv_sec) + 1e-9 * (tw1.tv_nsec - tw0.tv_nsec), (tw2.tv_sec - tw1.tv_sec) + 1e-9 * (tw2.tv_nsec - tw1.tv_nsec), g_h2d, g_dec, g_d2h, g_lat);
    if (trace)
        for (int i = 0; i < STAGE_SLOTS; i++)
            for (int q = 0; q < 6; q++) cudaEventDestroy(tev[i][q]);
    cudaError_t e0 = cudaSuccess;
    for (int i = 0; i < STAGE_SLOTS; i++) {
        const cudaError_t e = cudaStreamSynchronize(c->s_dec[i]);
        if (e != cudaSuccess) e0 = e;
    }
    if (rc == ZXC_OK &&
        cudaMemcpyAsync(h_status, d_status, (size_t)n_jobs * sizeof(i32), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess)
        rc = ZXC_B200_ERROR_CUDA;
    const cudaError_t e1 = cudaStreamSynchronize(c->stream);
    const cudaError_t e2 = cudaStreamSynchronize(c->s_d2h);
    const cudaError_t e3 = cudaStreamSynchronize(c->s_h2d);
    if (e0 != cudaSuccess || e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) {
        fprintf(stderr, "libzxc (B200 build): staged decode failed: %s\n", cudaGetErrorString(cudaGetLastError()));
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
