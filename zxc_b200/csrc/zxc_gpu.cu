/*
 * zxc_gpu.cu -- sm_100a kernels and the thin extern "C" shim behind libzxc.
 *
 * Decode: one warp per independent block.  A
 * warp takes a batch of 32 sequences at a time, lane = sequence:
 *   1. token / offset unpack               (GLO zxc_decompress.c:626-694, GHI :701-727)
 *   2. escape resolution over the extras   (varint, :51-88) -- the k escapes of
 *      the batch are walked once, uniformly, each lane keeps the value(s) whose
 *      ordinal (ballot + popc) is its own
 *   3. warp prefix sums -> literal source offset and output offset per lane
 *   4. bounds / offset validation, first failing lane = first failing sequence
 *   5. literal copies (independent), then match copies in dependency rounds:
 *      a match is ready when its source ends below the lowest pending match
 *      destination; long runs are copied by the whole warp, overlap (off < ml)
 *      is resolved with the period-`off` index instead of the reference's
 *      shuffle tables (:197-413).
 * Output is written exactly (no wild-copy overshoot), so none of the
 * reference's PAD/TAIL_PAD slack is needed on the destination; the wire-level
 * 32-byte literal slack rule stays normative (:1003).
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#include "zxc_error.h"
#include "zxc_gpu.h"

typedef uint8_t u8;
typedef uint32_t u32;
typedef uint64_t u64;
typedef int32_t i32;

#define FULL 0xFFFFFFFFu
#define WARPS_PER_CTA 4
#define CTA_THREADS (WARPS_PER_CTA * 32)

#define BT_RAW 0
#define BT_GLO 1
#define BT_GHI 2
#define BT_EOF 255

#define FLAG_VERIFY 1u

struct DecodeParams {
    const u8* src;
    u8* dst;
    const zxc_b200_job_t* jobs;
    i32* status;
    const u8* dict;
    const u8* dict_huf;
    u8* scratch;
    unsigned long long* counter;
    u32 n_jobs;
    u32 dict_size;
    u32 scratch_stride;
    u32 flags;
};

/* ------------------------------------------------------------------------- */
/* small device helpers                                                      */
/* ------------------------------------------------------------------------- */
__device__ __forceinline__ u32 ld16(const u8* p) { return (u32)p[0] | ((u32)p[1] << 8); }
__device__ __forceinline__ u32 ld32(const u8* p) {
    return (u32)p[0] | ((u32)p[1] << 8) | ((u32)p[2] << 16) | ((u32)p[3] << 24);
}
__device__ __forceinline__ u64 ld64(const u8* p) { return (u64)ld32(p) | ((u64)ld32(p + 4) << 32); }

__device__ __forceinline__ u32 warp_incl_scan(u32 v, u32 lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const u32 t = __shfl_up_sync(FULL, v, d);
        if (lane >= (u32)d) v += t;
    }
    return v;
}

/* warp-wide byte copy, non-overlapping */
__device__ __forceinline__ void warp_copy(u8* d, const u8* s, u32 n, u32 lane) {
    for (u32 k = lane; k < n; k += 32) d[k] = s[k];
}

/* prefix varint with the reference's exact failure behaviour (zxc_decompress.c:51-88):
 * value 0 and the cursor jams to `end`, except at/after end where it stays. */
__device__ __forceinline__ u32 read_varint(const u8* e, u32& pos, u32 end) {
    if (pos >= end) return 0;
    const u32 b0 = e[pos];
    if (b0 < 0x80) {
        pos += 1;
        return b0;
    }
    if (b0 < 0xC0) {
        if (pos + 1 >= end) {
            pos = end;
            return 0;
        }
        const u32 v = (b0 & 0x3F) | ((u32)e[pos + 1] << 6);
        pos += 2;
        return v;
    }
    if (b0 < 0xE0) {
        if (pos + 2 >= end) {
            pos = end;
            return 0;
        }
        const u32 v = (b0 & 0x1F) | ((u32)e[pos + 1] << 5) | ((u32)e[pos + 2] << 13);
        pos += 3;
        return v;
    }
    pos = end;
    return 0;
}

/* ------------------------------------------------------------------------- */
/* rapidhash V3 folded to 32 bits, warp-cooperative (vendors/rapidhash.h).    */
/* Lanes 0..6 own the seven stripe accumulators; the tail runs on lane 0.     */
/* ------------------------------------------------------------------------- */
__device__ __forceinline__ u64 mul_fold(u64 a, u64 b) { return (a * b) ^ __umul64hi(a, b); }

__device__ u32 warp_checksum(const u8* p, u32 len, u32 lane) {
    const u64 S0 = 0x2d358dccaa6c78a5ull, S1 = 0x8bb84b93962eacc9ull, S2 = 0x4b33a62ed433d4a3ull,
              S3 = 0x4d5a2da51de1aa47ull, S4 = 0xa0761d6478bd642full, S5 = 0xe7037ed1a0b428dbull,
              S6 = 0x90ed1765281c388cull, S7 = 0xaaaaaaaaaaaaaaaaull;
    u64 seed = 0;
    seed ^= mul_fold(seed ^ S2, S1);
    u64 a = 0, b = 0;
    u32 rem = len;
    if (len <= 16) {
        if (len >= 8) {
            seed ^= len;
            a = ld64(p);
            b = ld64(p + len - 8);
        } else if (len >= 4) {
            seed ^= len;
            a = ld32(p);
            b = ld32(p + len - 4);
        } else if (len > 0) {
            a = ((u64)p[0] << 45) | p[len - 1];
            b = p[len >> 1];
        }
    } else {
        if (len > 112) {
            const u64 sk = lane == 0 ? S0 : lane == 1 ? S1 : lane == 2 ? S2 : lane == 3 ? S3
                         : lane == 4 ? S4 : lane == 5 ? S5 : S6;
            u64 acc = seed;
            const u32 stripes = (len - 1) / 112; /* while (rem > 112) */
            if (lane < 7) {
                const u8* q = p + 16 * lane;
                for (u32 s = 0; s < stripes; s++, q += 112) acc = mul_fold(ld64(q) ^ sk, ld64(q + 8) ^ acc);
            }
            u64 x = (lane < 7) ? acc : 0;
#pragma unroll
            for (int d = 4; d >= 1; d >>= 1) x ^= __shfl_xor_sync(FULL, x, d);
            /* lanes 0..7 now hold the xor over their 8-group; group 0 covers accumulators 0..6 */
            seed = __shfl_sync(FULL, x, 0);
            p += (size_t)stripes * 112;
            rem -= stripes * 112;
        }
        const u64 ts[6] = {S2, S2, S1, S1, S2, S1};
#pragma unroll
        for (u32 k = 0; k < 6; k++)
            if (rem > 16u * (k + 1)) seed = mul_fold(ld64(p + 16 * k) ^ ts[k], ld64(p + 16 * k + 8) ^ seed);
        a = ld64(p + rem - 16) ^ rem;
        b = ld64(p + rem - 8);
    }
    a ^= S1;
    b ^= seed;
    const u64 lo = a * b, hi = __umul64hi(a, b);
    const u64 h = mul_fold(lo ^ S7, hi ^ S1 ^ rem);
    return (u32)(h ^ (h >> 32));
}

/* ------------------------------------------------------------------------- */
/* RLE literal section -> scratch (zxc_decompress.c:906-978)                  */
/* Token chain is serial; each token's run is written by the whole warp.      */
/* ------------------------------------------------------------------------- */
__device__ int rle_expand(const u8* r, u32 rsz, u8* w, u32 wsz, u32 lane) {
    u32 rp = 0, wp = 0;
    while (rp < rsz && wp < wsz) {
        const u32 t = r[rp++];
        if (!(t & 0x80)) {
            const u32 len = t + 1;
            if (wsz - wp < len || rsz - rp < len) return ZXC_ERROR_CORRUPT_DATA;
            for (u32 k = lane; k < len; k += 32) w[wp + k] = r[rp + k];
            wp += len;
            rp += len;
        } else {
            const u32 len = (t & 0x7F) + 4;
            if (wsz - wp < len || rp >= rsz) return ZXC_ERROR_CORRUPT_DATA;
            const u8 v = r[rp++];
            for (u32 k = lane; k < len; k += 32) w[wp + k] = v;
            wp += len;
        }
    }
    return wp == wsz ? ZXC_OK : ZXC_ERROR_CORRUPT_DATA;
}

/* byte at signed output position `pos` (negative = dictionary tail) */
__device__ __forceinline__ u8 window_byte(const u8* out, const u8* dict, u32 dict_size, i32 pos) {
    return pos >= 0 ? out[pos] : dict[(i32)dict_size + pos];
}

/* whole-warp match copy of n bytes to out[d..) from distance off */
__device__ __forceinline__ void warp_match_copy(u8* out, const u8* dict, u32 dict_size, u32 d, u32 off,
                                                u32 n, u32 lane) {
    if (off >= 32) {
        /* chunk c only reads bytes below its own start: earlier chunks are complete */
        for (u32 c = 0; c < n; c += 32) {
            const u32 k = c + lane;
            if (k < n) out[d + k] = window_byte(out, dict, dict_size, (i32)(d + k) - (i32)off);
            __syncwarp();
        }
    } else {
        /* period-`off` replication of the complete window [d-off, d) */
        for (u32 k = lane; k < n; k += 32)
            out[d + k] = window_byte(out, dict, dict_size, (i32)d - (i32)off + (i32)(k % off));
    }
}

#define LIT_SHORT 16u
#define MATCH_SHORT 32u

/* ------------------------------------------------------------------------- */
/* GLO / GHI block body.  Returns decoded bytes or a negative zxc_error_t.    */
/* ------------------------------------------------------------------------- */
__device__ int decode_lz_block(const u8* pay, u32 comp, bool ghi, u8* out, u32 cap, const u8* dict,
                               u32 dict_size, const u8* dict_huf, u8* scratch, u32 scratch_cap,
                               u32 lane) {
    if (comp < 12) return ZXC_ERROR_BAD_HEADER;
    const u32 n_seq = ld32(pay), n_lit = ld32(pay + 4);
    const u32 enc_lit = pay[8], enc_tok = pay[9], enc_off = pay[11];

    const u8* lit;      /* literal bytes, n_lit_avail of them */
    const u8* tok;      /* GLO tokens or GHI sequence words */
    const u8* offs = 0; /* GLO offsets */
    const u8* ext;
    u32 ext_end;        /* extras size */
    u32 n_lit_avail;

    if (!ghi) {
        const u32 desc = (enc_lit != 0 ? 4u : 0u) + (enc_tok == 2 ? 4u : 0u);
        if (comp < 12 + desc) return ZXC_ERROR_BAD_HEADER;
        u32 lit_comp = n_lit, tok_comp = n_seq;
        const u8* dp = pay + 12;
        if (enc_lit != 0) {
            lit_comp = ld32(dp);
            dp += 4;
        }
        if (enc_tok == 2) tok_comp = ld32(dp);
        if (enc_off > 1) return ZXC_ERROR_CORRUPT_DATA;
        const u8* p_data = pay + 12 + desc;
        const u32 avail = comp - 12 - desc;

        if (enc_lit == 2 || enc_lit == 3) {
            if (lit_comp > avail) return ZXC_ERROR_CORRUPT_DATA;
            if (n_lit != 0) {
                if (n_lit > cap) return ZXC_ERROR_DST_TOO_SMALL;
                if (enc_lit == 3 && !dict_huf) return ZXC_ERROR_DICT_REQUIRED;
                return ZXC_B200_ERROR_UNSUPPORTED; /* PivCo literal sections: SURVEY 8(f)-1 */
            }
            lit = p_data;
            n_lit_avail = 0;
        } else if (enc_lit == 1) {
            if (n_lit > 0) {
                if (n_lit > cap) return ZXC_ERROR_DST_TOO_SMALL;
                if (n_lit > scratch_cap) return ZXC_ERROR_CORRUPT_DATA; /* lit_buffer_cap, :914 */
                if (lit_comp > avail) return ZXC_ERROR_CORRUPT_DATA;
                const int rc = rle_expand(p_data, lit_comp, scratch, n_lit, lane);
                if (rc != ZXC_OK) return rc;
                __syncwarp();
                lit = scratch;
                n_lit_avail = n_lit;
            } else {
                lit = p_data;
                n_lit_avail = 0;
            }
        } else if (enc_lit == 0) {
            lit = p_data;
            n_lit_avail = lit_comp;
        } else {
            return ZXC_ERROR_CORRUPT_DATA;
        }
        const u64 sz_off = enc_off ? (u64)n_seq : (u64)n_seq * 2;
        const u64 consumed = (u64)lit_comp + tok_comp + sz_off;
        if (consumed > avail) return ZXC_ERROR_CORRUPT_DATA;
        if (avail - lit_comp < 32) return ZXC_ERROR_CORRUPT_DATA;
        if (enc_tok == 2) return ZXC_B200_ERROR_UNSUPPORTED; /* PivCo token sections: 8(f)-1 */
        if (enc_tok != 0) return ZXC_ERROR_CORRUPT_DATA;
        tok = p_data + lit_comp;
        offs = tok + tok_comp;
        ext = offs + (u32)sz_off;
        ext_end = avail - (u32)consumed;
    } else {
        if (enc_lit != 0 || enc_tok != 0) return ZXC_ERROR_CORRUPT_DATA;
        const u32 avail = comp - 12;
        const u64 consumed = (u64)n_lit + (u64)n_seq * 4;
        if (consumed > avail) return ZXC_ERROR_CORRUPT_DATA;
        if (avail - n_lit < 32) return ZXC_ERROR_CORRUPT_DATA;
        lit = pay + 12;
        n_lit_avail = n_lit;
        tok = lit + n_lit;
        ext = tok + (size_t)n_seq * 4;
        ext_end = avail - (u32)consumed;
    }

    const u32 esc = ghi ? 255u : 15u;
    const u32 lt_mask = (1u << lane) - 1u;
    u32 O = 0, L = 0, epos = 0;

    for (u32 base = 0; base < n_seq; base += 32) {
        const u32 i = base + lane;
        const bool valid = i < n_seq;
        u32 ll = 0, ml = 0, off = 1;
        if (valid) {
            if (!ghi) {
                const u32 t = tok[i];
                ll = t >> 4;
                ml = t & 15;
                off = (enc_off ? (u32)offs[i] : ld16(offs + 2 * (size_t)i)) + 1;
            } else {
                const u32 w = ld32(tok + 4 * (size_t)i);
                ll = w >> 24;
                ml = (w >> 16) & 0xFF;
                off = (w & 0xFFFF) + 1;
            }
        }
        const bool e_ll = valid && ll == esc, e_ml = valid && ml == esc;
        const u32 m_ll = __ballot_sync(FULL, e_ll), m_ml = __ballot_sync(FULL, e_ml);
        if (m_ll | m_ml) {
            const u32 ord_ll = __popc(m_ll & lt_mask) + __popc(m_ml & lt_mask);
            const u32 ord_ml = ord_ll + (e_ll ? 1u : 0u);
            const u32 k = __popc(m_ll) + __popc(m_ml);
            for (u32 s = 0; s < k; s++) {
                const u32 v = read_varint(ext, epos, ext_end);
                if (e_ll && s == ord_ll) ll += v;
                if (e_ml && s == ord_ml) ml += v;
            }
        }
        if (valid) ml += 5;
        const u32 tot = ll + ml;
        const u32 s_ll = warp_incl_scan(ll, lane);
        const u32 s_tot = warp_incl_scan(tot, lane);
        const u32 lit_start = L + s_ll - ll;
        const u32 out_start = O + s_tot - tot;
        const u32 mdst = out_start + ll;

        const bool ovf = valid && (lit_start + ll > n_lit_avail || out_start + tot > cap);
        const bool bad = valid && (mdst + dict_size < off);
        const u32 m_err = __ballot_sync(FULL, ovf || bad);
        if (m_err) {
            const int code = ovf ? ZXC_ERROR_OVERFLOW : ZXC_ERROR_BAD_OFFSET;
            return __shfl_sync(FULL, code, __ffs(m_err) - 1);
        }

        /* ---- literals: independent of every match ---- */
        if (valid && ll <= LIT_SHORT) {
            for (u32 k = 0; k < ll; k++) out[out_start + k] = lit[lit_start + k];
        }
        u32 m_long = __ballot_sync(FULL, valid && ll > LIT_SHORT);
        while (m_long) {
            const int j = __ffs(m_long) - 1;
            m_long &= m_long - 1;
            warp_copy(out + __shfl_sync(FULL, out_start, j), lit + __shfl_sync(FULL, lit_start, j),
                      __shfl_sync(FULL, ll, j), lane);
        }
        __syncwarp();

        /* ---- matches: dependency rounds ---- */
        u32 pending = __ballot_sync(FULL, valid);
        const i32 src_lo = (i32)mdst - (i32)off;
        const i32 src_end = min((i32)mdst, src_lo + (i32)ml);
        while (pending) {
            const int first = __ffs(pending) - 1;
            const i32 W = (i32)__shfl_sync(FULL, mdst, first);
            const bool ready = ((pending >> lane) & 1u) && ((int)lane == first || src_end <= W);
            if (ready && ml <= MATCH_SHORT) {
                for (u32 k = 0; k < ml; k++)
                    out[mdst + k] = window_byte(out, dict, dict_size, src_lo + (i32)k);
            }
            u32 m_lm = __ballot_sync(FULL, ready && ml > MATCH_SHORT);
            while (m_lm) {
                const int j = __ffs(m_lm) - 1;
                m_lm &= m_lm - 1;
                warp_match_copy(out, dict, dict_size, __shfl_sync(FULL, mdst, j),
                                __shfl_sync(FULL, off, j), __shfl_sync(FULL, ml, j), lane);
            }
            pending &= ~__ballot_sync(FULL, ready);
            __syncwarp();
        }

        O += __shfl_sync(FULL, s_tot, 31);
        L += __shfl_sync(FULL, s_ll, 31);
    }

    /* trailing literals (zxc_decompress.c:1198-1206) */
    const u32 rem = n_lit_avail - L;
    if (rem > cap - O) return ZXC_ERROR_OVERFLOW;
    warp_copy(out + O, lit + L, rem, lane);
    return (int)(O + rem);
}

/* zxc_decompress_chunk_wrapper_body (zxc_decompress.c:1646-1695) for one job */
__device__ int decode_job(const DecodeParams& P, const zxc_b200_job_t& job, u8* scratch, u32 lane) {
    const u8* blk = P.src + job.src_off;
    u8* out = P.dst + job.dst_off;
    if (job.src_len < 8) return ZXC_ERROR_SRC_TOO_SMALL;
    const u32 type = blk[0];
    const u32 comp = ld32(blk + 3);
    const bool verify = (P.flags & FLAG_VERIFY) != 0;
    if ((u64)job.src_len < 8ull + comp + (verify ? 4u : 0u)) return ZXC_ERROR_SRC_TOO_SMALL;
    const u8* data = blk + 8;
    if (verify) {
        if (ld32(data + comp) != warp_checksum(data, comp, lane)) return ZXC_ERROR_BAD_CHECKSUM;
    }
    switch (type) {
        case BT_GLO:
        case BT_GHI:
            return decode_lz_block(data, comp, type == BT_GHI, out, job.dst_cap, P.dict, P.dict_size,
                                   P.dict_huf, scratch, P.scratch_stride, lane);
        case BT_RAW:
            if (comp > job.dst_cap) return ZXC_ERROR_DST_TOO_SMALL;
            warp_copy(out, data, comp, lane);
            return (int)comp;
        case BT_EOF:
            return ZXC_ERROR_CORRUPT_DATA;
        default:
            return ZXC_ERROR_BAD_BLOCK_TYPE;
    }
}

__global__ void __launch_bounds__(CTA_THREADS) zxc_decode_kernel(const DecodeParams P) {
    const u32 lane = threadIdx.x & 31;
    const u32 gwarp = (blockIdx.x * CTA_THREADS + threadIdx.x) >> 5;
    u8* scratch = P.scratch + (size_t)gwarp * P.scratch_stride;
    for (;;) {
        unsigned long long j = 0;
        if (lane == 0) j = atomicAdd(P.counter, 1ull);
        j = __shfl_sync(FULL, j, 0);
        if (j >= P.n_jobs) break;
        const zxc_b200_job_t job = P.jobs[j];
        const int r = decode_job(P, job, scratch, lane);
        __syncwarp();
        if (lane == 0) P.status[j] = r;
    }
}

/* status reduce: first job whose result differs from its dst_cap */
__global__ void zxc_reduce_kernel(const i32* status, const zxc_b200_job_t* jobs, u32 n,
                                  unsigned long long* out /* [0]=first bad idx, [1]=sum */) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long bad = ~0ull, sum = 0;
    if (i < n) {
        const i32 s = status[i];
        if (s < 0 || (u32)s != jobs[i].dst_cap) bad = i;
        else sum = (u32)s;
    }
    for (int d = 16; d >= 1; d >>= 1) {
        const unsigned long long ob = __shfl_xor_sync(FULL, bad, d);
        bad = ob < bad ? ob : bad;
        sum += __shfl_xor_sync(FULL, sum, d);
    }
    if ((threadIdx.x & 31) == 0) {
        if (bad != ~0ull) atomicMin(&out[0], bad);
        if (sum) atomicAdd(&out[1], sum);
    }
}

/* ========================================================================= */
/* host side: device bring-up, contexts, copies, launches                    */
/* ========================================================================= */
static pthread_once_t g_once = PTHREAD_ONCE_INIT;
static int g_init_rc = ZXC_B200_ERROR_NO_DEVICE;
static int g_sm_count = 0;
static unsigned long long g_launches = 0;
static pthread_mutex_t g_pool_mu = PTHREAD_MUTEX_INITIALIZER;

#define PIN_CHUNK ((size_t)16 << 20)

struct zxg_ctx {
    cudaStream_t stream;
    void* buf[ZXG_BUF_COUNT];
    size_t cap[ZXG_BUF_COUNT];
    void* pin[2];
    cudaEvent_t pin_ev[2];
    unsigned long long* counter; /* [0] work counter, [1..2] reduce output */
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
        memcpy(c->pin[slot], (const u8*)h_src + done, n);
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
    /* two-deep pipeline: while chunk k is copied out of its bounce buffer, chunk k+1 is in flight */
    size_t issued = 0, drained = 0;
    size_t len[2] = {0, 0};
    int slot = 0;
    while (drained < bytes) {
        if (issued < bytes && len[slot] == 0) {
            const size_t n = bytes - issued < PIN_CHUNK ? bytes - issued : PIN_CHUNK;
            if (cudaMemcpyAsync(c->pin[slot], (const u8*)d_src + issued, n, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess)
                return ZXC_B200_ERROR_CUDA;
            cudaEventRecord(c->pin_ev[slot], c->stream);
            len[slot] = n;
            issued += n;
            if (issued < bytes && len[slot ^ 1] == 0) {
                slot ^= 1;
                continue;
            }
        }
        const int ds = (len[slot ^ 1] != 0 && (issued - len[slot] - len[slot ^ 1] == drained)) ? (slot ^ 1) : slot;
        /* drain the older outstanding chunk */
        if (cudaEventSynchronize(c->pin_ev[ds]) != cudaSuccess) return ZXC_B200_ERROR_CUDA;
        memcpy((u8*)h_dst + drained, c->pin[ds], len[ds]);
        drained += len[ds];
        len[ds] = 0;
        slot = ds;
    }
    return ZXC_OK;
}

static int grid_for(u32 n_jobs) {
    const u32 ctas_needed = (n_jobs + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
    const u32 resident = (u32)(g_sm_count > 0 ? g_sm_count : 148) * 4u; /* 16 warps per SM */
    return (int)(ctas_needed < resident ? ctas_needed : resident);
}

static u32 scratch_stride_for(u32 block_size) { return (block_size + 255u) & ~255u; }

extern "C" size_t zxc_b200_decode_scratch_size(uint32_t block_size) {
    if (zxg_init() != ZXC_OK) return 0;
    const size_t warps = (size_t)g_sm_count * 4u * WARPS_PER_CTA;
    return warps * scratch_stride_for(block_size) + sizeof(unsigned long long) * 4;
}

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
    const int grid = grid_for(n_jobs);
    if ((size_t)grid * WARPS_PER_CTA * P.scratch_stride > scratch_size) return ZXC_ERROR_MEMORY;
    if (cudaMemsetAsync(d_counter, 0, sizeof(unsigned long long), st) != cudaSuccess) return ZXC_B200_ERROR_CUDA;
    zxc_decode_kernel<<<grid, CTA_THREADS, 0, st>>>(P);
    __atomic_add_fetch(&g_launches, 1, __ATOMIC_RELAXED);
    return cudaGetLastError() == cudaSuccess ? ZXC_OK : ZXC_B200_ERROR_CUDA;
}

extern "C" int zxc_b200_decode_blocks(const void* d_src, void* d_dst, const zxc_b200_job_t* d_jobs,
                                      uint32_t n_jobs, int32_t* d_status, const void* d_dict,
                                      uint32_t dict_size, const void* d_dict_huf, void* d_scratch,
                                      size_t scratch_size, uint32_t block_size, int verify_checksums,
                                      void* stream) {
    const int rc = zxg_init();
    if (rc != ZXC_OK) return rc;
    if (!d_src || !d_dst || !d_jobs || !d_status || !d_scratch) return ZXC_ERROR_NULL_INPUT;
    if (scratch_size < 4 * sizeof(unsigned long long)) return ZXC_ERROR_MEMORY;
    /* the work counter lives in the last 32 bytes of the caller's scratch */
    const size_t usable = (scratch_size - 4 * sizeof(unsigned long long)) & ~(size_t)7;
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
    const size_t warps = (size_t)grid_for(n_jobs) * WARPS_PER_CTA;
    const size_t scratch_size = warps * scratch_stride_for(block_size);
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
