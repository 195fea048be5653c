/*
 * zxc_decode2.cuh -- block-cooperative decode kernel for sm_100a (device code only).
 *
 * One CTA decodes one independent block with the whole output
 * window in shared memory, so every match source is a shared-memory read and DRAM sees the
 * compressed block once and the decoded block once:
 *
 *   load     two cp.async.bulk copies on an mbarrier: the token/offset/extras sections into the low
 *            part of the window (free until the first output byte is written) and the literal
 *            stream behind the window's end, `gap` bytes above where its last byte will be consumed
 *            (a literal's output position never exceeds its staged position, so the stream is
 *            eaten from below while the output grows into it).
 *   phase 1  sequence-parallel, 16 sequences per thread: escapes counted, the extras section
 *            resolved by a segment-map scan (every thread walks a few bytes for each of the three
 *            possible entry phases, the maps are composed by a prefix scan; varint semantics of
 *            zxc_decompress.c:51-88 including the jam-to-end failure), prefix sums of ll and
 *            ll+ml, validation in the reference's order (overflow before bad offset, first failing
 *            sequence wins), one 8-byte record per sequence (zxc_decode2_core.h).
 *   phase 2  output-parallel: a warp claims 512-byte groups in order; lane = aligned output word.
 *            The covering sequence of every word comes from a bitmask of sequence ends (sequences
 *            are >= 5 bytes, so ends fall into distinct words) and a popcount.  A word is gathered
 *            from one, two or three regions (zxc_decode2_core.h) once the words it reads are
 *            complete (per-row bitmasks in shared memory); overlapped matches are folded onto the
 *            period in front of the match, so runs never chain.
 *   store    completed 8 KiB chunks leave as cp.async.bulk shared -> global while later groups
 *            are still being decoded.
 *
 * Blocks this kernel does not take (entropy-coded literal or token sections, blocks that do not
 * fit the window) are marked D2_DEFER in the status array and decoded by the warp-per-block kernel
 * (zxc_decode.cuh) in a second launch.
 */
#pragma once
#include "zxc_decode.cuh"
#include "zxc_decode2_core.h"

#define D2_DEFER ((i32)0x80000000)
#define D2_CHUNK 16u /* sequences per thread in phase 1 */

struct Decode2Params {
    const u8* src;
    u8* dst;
    const zxc_b200_job_t* jobs;
    i32* status;
    const u8* dict;
    unsigned long long* counter;
    z2_rec_t* spill; /* per-CTA record overflow in global memory (spill_stride records each) */
    u32* defer_list; /* job indices this kernel left to the general kernel */
    u32* defer_count;
    u32 defer_cap;
    unsigned long long* trace; /* development: 8 clock64() stamps per job, or NULL */
    u32 n_jobs;
    u32 dict_size;
    u32 win;  /* output window bytes: multiple of 512, <= 65536 */
    u32 gap;  /* literal staging gap: multiple of 512 */
    u32 rcap; /* records held in shared memory */
    u32 spill_stride;
};

/* shared-memory layout (bytes from the 128-aligned dynamic base) */
#define D2_OFF_MBAR 0u
#define D2_OFF_CTRL 16u    /* u32[28] control words */
#define D2_OFF_SCAN 128u   /* u64[40] scan scratch */
#define D2_OFF_GIDX 448u   /* u16[136] first sequence of every group */
#define D2_OFF_RC 768u     /* u32[16]: one bit per completed 128-byte row */
#define D2_OFF_ROWBAR 896u /* u64[512]: one mbarrier per row, for warps that have nothing ready */
#define D2_OFF_WIN 4992u   /* window: win + gap + 128 */
__host__ __device__ __forceinline__ u32 d2_off_rec(u32 win, u32 gap) { return D2_OFF_WIN + win + gap + 128u; }
__host__ __device__ __forceinline__ u32 d2_smem_bytes(u32 win, u32 gap, u32 rcap) {
    return d2_off_rec(win, gap) + (rcap + 4u) * 8u;
}

enum { C_CLAIM = 0, C_ERR, C_JOB, C_NVAL, C_SUML, C_SUMO, C_SUME, C_PAD, C_GDONE /* 4 words */ };

/* ---- PTX wrappers ------------------------------------------------------------------------------ */
__device__ __forceinline__ u32 d2_saddr(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void d2_mbar_init(u32 bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void d2_mbar_expect(u32 bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void d2_bulk_load(u32 sdst, const void* gsrc, u32 bytes, u32 bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sdst),
                 "l"(gsrc), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void d2_mbar_wait(u32 bar, u32 parity) {
    u32 ok = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void d2_mbar_arrive(u32 bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void d2_bulk_store(void* gdst, u32 ssrc, u32 bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc), "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void d2_bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void d2_bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
/* Completion masks are published with release semantics and read with acquire semantics at CTA scope.
 * ptxas lowers ld.acquire.cta.shared to a plain LDS (shared loads of a thread are performed in order), so
 * only the publishing side pays a MEMBAR.ALL.CTA. */
__device__ __forceinline__ void d2_release_fence() { asm volatile("fence.acq_rel.cta;" ::: "memory"); }
__device__ __forceinline__ u32 d2_ld_acquire(const volatile u32* p) {
    u32 v;
    asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"((u32)__cvta_generic_to_shared(const_cast<u32*>(p))) : "memory");
    return v;
}
__device__ __forceinline__ void d2_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void d2_prefetch_l2(const void* g, u32 bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(g), "r"(bytes) : "memory");
}

/* exclusive scan of v over the CTA in thread order; total = CTA sum.  tmp: u32[32] shared. */
__device__ __forceinline__ u32 d2_cta_scan(u32 v, u32* tmp, u32 lane, u32 wic, u32 nwarps, u32& total) {
    const u32 inc = warp_incl_scan(v, lane);
    if (lane == 31) tmp[wic] = inc;
    __syncthreads();
    u32 base = 0, tot = 0;
    for (u32 k = 0; k < nwarps; k++) {
        const u32 x = tmp[k];
        if (k < wic) base += x;
        tot += x;
    }
    __syncthreads();
    total = tot;
    return base + inc - v;
}

/* unaligned 4-byte gather from the window */
__device__ __forceinline__ u32 d2_gather(const u8* win, i32 s) {
    const u32 a = (u32)s & ~3u;
    const volatile u32* wp = reinterpret_cast<const volatile u32*>(win + a);
    const u32 w0 = wp[0], w1 = wp[1];
    return __byte_perm(w0, w1, 0x3210u + 0x1111u * ((u32)s & 3u));
}
/* Completion state as one warp sees it while it works on group `g4 / 4`:
 *   lead   rows [0, lead) are complete (contiguous prefix of the row bitmap; a lower bound)
 *   own    completion masks of the warp's own four rows (registers, identical in every lane)
 * A source word is usable when its row is below `lead`, or it is one of the warp's own finished words. */
struct D2Prog {
    u32 lead, g4;
    u32 own[Z2_ROWS];
};
__device__ __forceinline__ bool d2_word_ok(const D2Prog& S, u32 w) {
    const u32 r = w >> 5;
    if (r < S.lead) return true;
    if (r < S.g4) return false;
    const u32 q = r - S.g4;
    const u32 m = q == 0 ? S.own[0] : q == 1 ? S.own[1] : q == 2 ? S.own[2] : S.own[3];
    return ((m >> (w & 31u)) & 1u) != 0u;
}
/* are the words holding window bytes [s, s+4) complete? */
__device__ __forceinline__ bool d2_ready(const D2Prog& S, i32 s) {
    const u32 w0 = (u32)s >> 2, w1 = ((u32)s + 3u) >> 2;
    if ((w1 >> 5) < S.lead) return true;
    return d2_word_ok(S, w0) && d2_word_ok(S, w1);
}

/* RLE literal section (zxc_decompress.c:906-978) expanded inside the window buffer: the compressed stream
 * sits right-aligned at the end of the buffer, the literals grow from `wp` upwards behind the read cursor.
 * Returns ZXC_OK, ZXC_ERROR_CORRUPT_DATA as the reference would, or 1 when a burst of run tokens would
 * overrun bytes that have not been read yet (the block then goes to the general kernel). */
__device__ int d2_rle_inplace(u8* buf, u32 rp, u32 rend, u32 wp, u32 wend, u32 lane) {
    while (rp < rend && wp < wend) {
        const u32 t = buf[rp++];
        if (!(t & 0x80u)) {
            const u32 len = t + 1u;
            if (wend - wp < len || rend - rp < len) return ZXC_ERROR_CORRUPT_DATA;
            if (wp > rp) return 1;
            for (u32 c = 0; c < len; c += 32) { /* forward overlapping copy: destination at or below the source */
                const u32 k = c + lane;
                const u8 v = k < len ? buf[rp + k] : (u8)0;
                __syncwarp();
                if (k < len) buf[wp + k] = v;
                __syncwarp();
            }
            wp += len;
            rp += len;
        } else {
            const u32 len = (t & 0x7Fu) + 4u;
            if (wend - wp < len || rp >= rend) return ZXC_ERROR_CORRUPT_DATA;
            const u8 v = buf[rp++];
            if (wp + len > rp) return 1;
            for (u32 k = lane; k < len; k += 32) buf[wp + k] = v;
            wp += len;
        }
        __syncwarp();
    }
    return wp == wend ? ZXC_OK : ZXC_ERROR_CORRUPT_DATA;
}

struct D2Block {
    u8* win;            /* window base (shared) */
    const z2_rec_t* rs; /* records in shared memory */
    const z2_rec_t* rg; /* records in global memory (same indexing) */
    const u8* dict;
    u32 dict_size;
    u32 rcap;
    u32 n_seq;
    i32 total;   /* bytes the block produces */
    i32 lit_pos; /* window position of literal-stream byte 0 */
};

__device__ __forceinline__ z2_rec_t d2_ld_rec(const D2Block& B, u32 k) {
    z2_rec_t r;
    if (k < B.rcap) {
        const uint2 v = *reinterpret_cast<const uint2*>(B.rs + k);
        r.w0 = v.x;
        r.w1 = v.y;
    } else {
        const uint2 v = *reinterpret_cast<const uint2*>(B.rg + k);
        r.w0 = v.x;
        r.w1 = v.y;
    }
    return r;
}
__device__ __forceinline__ void d2_seq_pair(const D2Block& B, u32 idx, z2_seq_t& c, z2_seq_t& n) {
    c = z2_unpack(d2_ld_rec(B, idx));
    n = z2_unpack(d2_ld_rec(B, idx + 1));
    if (idx >= B.n_seq) c.md = Z2_MD_INF; /* trailing literals: never a match */
    if (idx + 1 >= B.n_seq) n.md = B.total;
}

/* byte-wise word (wrapped periods, off < 4, dictionary sources).  Returns false when a source
 * byte is not complete yet. */
__device__ __noinline__ bool d2_slow_word(const D2Block& B, const D2Prog S, u32 idx, i32 p, u32& out) {
    z2_seq_t c, n;
    d2_seq_pair(B, idx, c, n);
    u32 acc = 0;
#pragma unroll 1
    for (i32 b = 0; b < 4; b++) {
        const i32 q = p + b;
        if (q >= B.total) break;
        int is_match;
        const i32 s = z2_byte_source(q, c, n, B.lit_pos, &is_match);
        u32 v;
        if (!is_match) {
            v = B.win[s];
        } else if (s < 0) {
            v = B.dict[(i32)B.dict_size + s];
        } else if (s >= p) {
            v = (acc >> (8 * (s - p))) & 0xFFu;
        } else {
            if (!d2_word_ok(S, (u32)s >> 2)) return false;
            v = *reinterpret_cast<volatile u8*>(B.win + s);
        }
        acc |= v << (8 * b);
    }
    out = acc;
    return true;
}

/* ------------------------------------------------------------------------------------------------ */
__global__ void __launch_bounds__(256, 2) zxc_decode2_kernel(const Decode2Params P) {
    extern __shared__ __align__(128) u8 sm[];
    const u32 tid = threadIdx.x, lane = tid & 31u, wic = tid >> 5, T = blockDim.x, NW = T >> 5;
    u32* ctrl = reinterpret_cast<u32*>(sm + D2_OFF_CTRL);
    volatile u32* vctrl = ctrl;
    u32* scan32 = reinterpret_cast<u32*>(sm + D2_OFF_SCAN);
    u64* scan64 = reinterpret_cast<u64*>(sm + D2_OFF_SCAN);
    unsigned short* gidx = reinterpret_cast<unsigned short*>(sm + D2_OFF_GIDX);
    u32* rcw = reinterpret_cast<u32*>(sm + D2_OFF_RC);
    const u32 rowbar = d2_saddr(sm + D2_OFF_ROWBAR);
    u8* win = sm + D2_OFF_WIN;
    z2_rec_t* recs = reinterpret_cast<z2_rec_t*>(sm + d2_off_rec(P.win, P.gap));
    z2_rec_t* recg = P.spill ? P.spill + (size_t)blockIdx.x * P.spill_stride : recs;
    const u32 rcap = P.spill ? P.rcap : 0xFFFFFFFFu;
    const u32 rec_room = P.spill ? P.spill_stride : P.rcap;
    const u32 bar = d2_saddr(sm + D2_OFF_MBAR);
    const u32 KGUARD = P.gap / Z2_GROUP;
    const u32 rows_max = P.win / 128u;
    u32 phase = 0, row_par = 0; /* parity of the load barrier / of the row barriers' current phase */

    if (tid == 0) d2_mbar_init(bar, 1);
    for (u32 r = tid; r < rows_max; r += T) d2_mbar_init(rowbar + 8u * r, 1);
    __syncthreads();

    for (;;) {
        __syncthreads(); /* everyone is done with the previous job (and has read its index) */
        if (tid == 0) ctrl[C_JOB] = (u32)atomicAdd(P.counter, 1ull);
        __syncthreads();
        const u32 j = vctrl[C_JOB];
        if (j >= P.n_jobs) break;
#define D2_STAMP(i) do { if (P.trace && tid == 0) P.trace[(size_t)j * 8 + (i)] = (unsigned long long)clock64(); } while (0)
        D2_STAMP(0);
        const zxc_b200_job_t job = P.jobs[j];
        const u8* blk = P.src + job.src_off;
        u8* out = P.dst + job.dst_off;
        const u32 cap = job.dst_cap;
        if (tid == 32 && j + gridDim.x < P.n_jobs) { /* pull the block a CTA will want one round from now into L2 */
            const zxc_b200_job_t nj = P.jobs[j + gridDim.x];
            const u8* nb = P.src + nj.src_off;
            const u32 sh = (u32)(reinterpret_cast<uintptr_t>(nb) & 15u);
            d2_prefetch_l2(nb - sh, (sh + nj.src_len + 15u) & ~15u);
        }

        /* ---- header: anything unusual is left to the general kernel ---- */
        int verdict = 0; /* 0 = take it, 1 = raw copy, 2 = defer */
        u32 comp = 0, n_seq = 0, n_lit = 0, enc_off = 0, desc = 0, lit_comp = 0;
        bool ghi = false, rle = false;
        if (job.src_len < 8 + 12) {
            verdict = 2;
        } else {
            const u32 type = blk[0];
            comp = ld32(blk + 3);
            if ((u64)job.src_len < 8ull + comp) verdict = 2;
            else if (type == BT_RAW) verdict = (comp <= cap) ? 1 : 2;
            else if (type != BT_GLO && type != BT_GHI) verdict = 2;
            else if (comp < 12 || cap > P.win) verdict = 2;
            else {
                ghi = type == BT_GHI;
                const u8* pay = blk + 8;
                n_seq = ld32(pay);
                n_lit = ld32(pay + 4);
                const u32 enc_lit = pay[8], enc_tok = pay[9];
                enc_off = pay[11];
                lit_comp = n_lit;
                if (enc_lit == 1 && !ghi && comp >= 16 && n_lit > 0) { /* RLE literal section (:906-978) */
                    rle = true;
                    desc = 4;
                    lit_comp = ld32(pay + 12);
                } else if (enc_lit != 0) {
                    verdict = 2;
                }
                if (enc_tok != 0 || (!ghi && enc_off > 1)) verdict = 2;
            }
        }
        u32 s_bytes = 0, ext_len = 0, rle_at = 0;
        if (verdict == 0) {
            const u32 avail = comp - 12 - desc;
            const u64 seq_bytes = ghi ? (u64)n_seq * 4 : (u64)n_seq * (enc_off ? 2 : 3);
            const u64 consumed = (u64)lit_comp + seq_bytes;
            if (lit_comp > avail || consumed > avail || avail - lit_comp < 32) verdict = 2;
            else if (n_lit > cap || n_seq + 3 > rec_room || n_seq > 0xFFF0u) verdict = 2;
            else {
                s_bytes = avail - lit_comp; /* tokens + offsets + extras (+ padding) */
                ext_len = avail - (u32)consumed;
                /* room: [S | values | RLE stream] below the staged literals */
                const u32 lba = (P.win + P.gap - n_lit + 15u) & ~15u;
                const u64 low = (u64)s_bytes + 48 + 4ull * ext_len + 16;
                if (low > lba) verdict = 2;
                if (rle) { /* compressed stream right-aligned at the end of the buffer (16-byte granules) */
                    const u32 buf_end = P.win + P.gap + 128u;
                    if ((u64)lit_comp + 64 + low > buf_end) verdict = 2;
                    else rle_at = (buf_end - lit_comp - 32u) & ~15u;
                }
            }
        }
        if (verdict == 2) {
            if (tid == 0) {
                P.status[j] = D2_DEFER;
                const u32 slot = atomicAdd(P.defer_count, 1u);
                if (slot < P.defer_cap) P.defer_list[slot] = j;
            }
            continue;
        }
        if (verdict == 1) { /* RAW block: word-granular copy (zxc_decompress.c:1646-1695) */
            const u8* data = blk + 8;
            if (((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(data)) & 15u) == 0) {
                const uint4* s4 = reinterpret_cast<const uint4*>(data);
                uint4* d4 = reinterpret_cast<uint4*>(out);
                const u32 n16 = comp >> 4;
                for (u32 k = tid; k < n16; k += T) d4[k] = s4[k];
                for (u32 k = (n16 << 4) + tid; k < comp; k += T) out[k] = data[k];
            } else if ((reinterpret_cast<uintptr_t>(out) & 3u) == 0) {
                const u32 m = (u32)(reinterpret_cast<uintptr_t>(data) & 3u);
                const u32* wp = reinterpret_cast<const u32*>(data - m);
                u32* dw = reinterpret_cast<u32*>(out);
                const u32 nw = comp >> 2;
                for (u32 k = tid; k < nw; k += T) {
                    const u32 a = wp[k];
                    const u32 b = m ? wp[k + 1] : 0u;
                    dw[k] = __funnelshift_r(a, b, m * 8u);
                }
                for (u32 k = (nw << 2) + tid; k < comp; k += T) out[k] = data[k];
            } else {
                for (u32 k = tid; k < comp; k += T) out[k] = data[k];
            }
            if (tid == 0) P.status[j] = (i32)comp;
            continue;
        }

        /* ---- stage the block ---- */
        const u8* pay = blk + 8;
        const u8* g_lit = pay + 12 + desc;
        const u8* g_seq = g_lit + lit_comp;
        const u32 s_shift = (u32)(reinterpret_cast<uintptr_t>(g_seq) & 15u);
        const u32 g_lit_shift = (u32)(reinterpret_cast<uintptr_t>(g_lit) & 15u);
        const u32 l_shift = rle ? 0u : g_lit_shift; /* RLE: the expanded stream is written, not copied */
        const u32 lba = (P.win + P.gap - n_lit + 15u) & ~15u; /* staged literal copy starts here (16-aligned);
                                                                 rounded up: the gap must not shrink */
        const i32 lit_pos = (i32)(lba + l_shift);
        const u32 s_copy = (s_shift + s_bytes + 15u) & ~15u;
        const u32 l_copy = lit_comp ? ((g_lit_shift + lit_comp + 15u) & ~15u) : 0u;
        if (tid == 0) {
            d2_fence_async();
            d2_mbar_expect(bar, s_copy + l_copy);
            d2_bulk_load(d2_saddr(win), g_seq - s_shift, s_copy, bar);
            if (l_copy) d2_bulk_load(d2_saddr(win + (rle ? rle_at : lba)), g_lit - g_lit_shift, l_copy, bar);
        }
        /* control state while the copies fly */
        const u32 n_groups_max = (cap + Z2_GROUP - 1) / Z2_GROUP;
        for (u32 k = tid; k <= n_groups_max; k += T) gidx[k] = (unsigned short)n_seq;
        if (tid < 16) rcw[tid] = 0;
        if (tid < 4) ctrl[C_GDONE + tid] = 0;
        if (tid == 0) {
            ctrl[C_CLAIM] = 0;
            ctrl[C_ERR] = 0xFFFFFFFFu;
            ctrl[C_NVAL] = 0;
        }
        D2_STAMP(1);
        d2_mbar_wait(bar, phase);
        phase ^= 1u;
        __syncthreads();
        D2_STAMP(2);
        if (rle) { /* expand the literal stream to where raw literals would have been staged */
            if (wic == 0) {
                const u32 rp = rle_at + g_lit_shift;
                const int rc = d2_rle_inplace(win, rp, rp + lit_comp, (u32)lit_pos, (u32)lit_pos + n_lit, lane);
                if (lane == 0) ctrl[C_NVAL] = (u32)rc;
            }
            __syncthreads();
            const int rc = (int)vctrl[C_NVAL];
            __syncthreads();
            if (rc != ZXC_OK) {
                if (tid == 0) {
                    if (rc == 1) { /* in-place expansion not possible for this stream: general kernel */
                        P.status[j] = D2_DEFER;
                        const u32 slot = atomicAdd(P.defer_count, 1u);
                        if (slot < P.defer_cap) P.defer_list[slot] = j;
                    } else {
                        P.status[j] = rc;
                    }
                }
                continue;
            }
        }

        const u8* S = win + s_shift;
        const u8* S_off = ghi ? S : S + n_seq;
        const u8* S_ext = ghi ? S + 4u * n_seq : S_off + (enc_off ? n_seq : 2u * n_seq);
        u32* vals = reinterpret_cast<u32*>(win + ((s_shift + s_bytes + 3u + 16u) & ~3u));
        const u32 esc = ghi ? 255u : 15u;

        /* ---- phase 1a: varint values of the extras section ---- */
        u32 n_val = 0;
        if (ext_len) {
            const u32 seg = max(4u, (ext_len + T - 1) / T);
            const u32 nseg = (ext_len + seg - 1) / seg;
            const u32 lo = tid * seg, hi = min(ext_len, lo + seg);
            const u64 map = tid < nseg ? z2_seg_map(S_ext, lo, hi, ext_len) : Z2_MAP_ID;
            /* inclusive scan of maps over the CTA */
            u64 inc = map;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const u64 o = __shfl_up_sync(FULL, inc, d);
                if (lane >= (u32)d) inc = z2_map_compose(o, inc);
            }
            if (lane == 31) scan64[wic] = inc;
            __syncthreads();
            u64 pre = Z2_MAP_ID;
            for (u32 k = 0; k < wic; k++) pre = z2_map_compose(pre, scan64[k]);
            u64 all = pre;
            for (u32 k = wic; k < NW; k++) all = z2_map_compose(all, scan64[k]);
            __syncthreads();
            u64 excl = __shfl_up_sync(FULL, inc, 1);
            if (lane == 0) excl = Z2_MAP_ID;
            excl = z2_map_compose(pre, excl);
            n_val = z2_map_cnt(all, 0);
            if (tid < nseg) {
                const u32 ent = z2_map_exit(excl, 0);
                if (ent != 3u) z2_seg_values(S_ext, lo, hi, ext_len, ent, z2_map_cnt(excl, 0), vals);
            }
            __syncthreads();
        }

        D2_STAMP(3);
        /* ---- phase 1b: sequences -> records ---- */
        u32 carryE = 0, carryL = 0, carryO = 0;
        for (u32 base = 0; base < n_seq; base += T * D2_CHUNK) {
            const u32 i0 = base + tid * D2_CHUNK;
            u32 lm[D2_CHUNK];
            u32 nesc = 0;
#pragma unroll
            for (u32 k = 0; k < D2_CHUNK; k++) {
                const u32 i = i0 + k;
                u32 ll = 0, ml = 0;
                if (i < n_seq) {
                    if (!ghi) {
                        const u32 t = S[i];
                        ll = t >> 4;
                        ml = t & 15u;
                    } else {
                        ll = S[4u * i + 3];
                        ml = S[4u * i + 2];
                    }
                    nesc += (ll == esc ? 1u : 0u) + (ml == esc ? 1u : 0u);
                }
                lm[k] = ll | (ml << 16);
            }
            u32 totE;
            u32 ord = carryE + d2_cta_scan(nesc, scan32, lane, wic, NW, totE);
            u32 sl = 0, st = 0;
#pragma unroll
            for (u32 k = 0; k < D2_CHUNK; k++) {
                if (i0 + k < n_seq) {
                    u32 ll = lm[k] & 0xFFFFu, ml = lm[k] >> 16;
                    if (ll == esc) {
                        ll += ord < n_val ? vals[ord] : 0u;
                        ord++;
                    }
                    if (ml == esc) {
                        ml += ord < n_val ? vals[ord] : 0u;
                        ord++;
                    }
                    ml += 5u;
                    ll = min(ll, 0xFFFFu);
                    ml = min(ml, 0xFFFFu);
                    lm[k] = ll | (ml << 16);
                    sl += ll;
                    st += ll + ml;
                }
            }
            u32 totL, totO;
            u32 L = carryL + d2_cta_scan(sl, scan32, lane, wic, NW, totL);
            u32 O = carryO + d2_cta_scan(st, scan32, lane, wic, NW, totO);
#pragma unroll
            for (u32 k = 0; k < D2_CHUNK; k++) {
                const u32 i = i0 + k;
                if (i < n_seq) {
                    const u32 ll = lm[k] & 0xFFFFu, ml = lm[k] >> 16;
                    u32 off;
                    if (ghi) off = (u32)S[4u * i] | ((u32)S[4u * i + 1] << 8);
                    else if (enc_off) off = S_off[i];
                    else off = (u32)S_off[2u * i] | ((u32)S_off[2u * i + 1] << 8);
                    off += 1u;
                    const u32 md = O + ll, E = md + ml;
                    const bool ovf = (L + ll > n_lit) || (E > cap);
                    const bool bad = md + P.dict_size < off;
                    if (ovf || bad) atomicMin(&ctrl[C_ERR], (i << 1) | (ovf ? 0u : 1u));
                    const z2_rec_t r = z2_pack(E, md & 0xFFFFu, off, (O - L) & 0xFFFFu);
                    if (i < rcap) recs[i] = r;
                    else recg[i] = r;
                    const u32 g_lo = (O + Z2_GROUP - 1) / Z2_GROUP, g_hi = (min(E, cap) + Z2_GROUP - 1) / Z2_GROUP;
                    for (u32 g = g_lo; g < g_hi; g++) gidx[g] = (unsigned short)i;
                    O = E;
                    L += ll;
                }
            }
            carryE += totE;
            carryL = min(carryL + totL, 1u << 30);
            carryO = min(carryO + totO, 1u << 30);
        }
        /* trailing literals (zxc_decompress.c:1198-1206) as a virtual sequence, then a sentinel */
        i32 result;
        u32 total = 0;
        {
            const u32 err = (__syncthreads(), vctrl[C_ERR]);
            if (err != 0xFFFFFFFFu) {
                result = (err & 1u) ? ZXC_ERROR_BAD_OFFSET : ZXC_ERROR_OVERFLOW;
            } else {
                const u32 rem = n_lit - carryL; /* carryL <= n_lit: no sequence overflowed */
                if (rem > cap - carryO) result = ZXC_ERROR_OVERFLOW;
                else {
                    total = carryO + rem;
                    result = (i32)total;
                }
            }
        }
        if (result <= 0) {
            if (tid == 0) P.status[j] = result;
            __syncthreads();
            continue;
        }
        if (tid == 0) {
            const z2_rec_t v = z2_pack(total, total & 0xFFFFu, 1u, (carryO - carryL) & 0xFFFFu);
            const z2_rec_t s = z2_pack(0x10000u, total & 0xFFFFu, 1u, 0u);
            if (n_seq < rcap) recs[n_seq] = v; else recg[n_seq] = v;
            if (n_seq + 1 < rcap) recs[n_seq + 1] = s; else recg[n_seq + 1] = s;
            if (n_seq + 2 < rcap) recs[n_seq + 2] = s; else recg[n_seq + 2] = s;
        }
        __syncthreads();

        D2_STAMP(4);
        /* ---- phase 2: output words ---- */
        D2Block B;
        B.win = win;
        B.rs = recs;
        B.rg = recg;
        B.dict = P.dict;
        B.dict_size = P.dict_size;
        B.rcap = rcap;
        B.n_seq = n_seq;
        B.total = (i32)total;
        B.lit_pos = lit_pos;
        const u32 n_groups = (total + Z2_GROUP - 1) / Z2_GROUP;
        const bool bulk_ok = (reinterpret_cast<uintptr_t>(out) & 15u) == 0;
        const u32 total16 = total & ~15u;
        const u32 le_mask = (2u << lane) - 1u;

        const u32 n_rows = (total + 127u) / 128u;
        D2Prog Pg;
        Pg.lead = 0;
        /* rows [0, Pg.lead) complete: extend the prefix over the row bitmap (acquire loads) */
        auto refresh_lead = [&]() {
            while (Pg.lead < n_rows) {
                const u32 sh = Pg.lead & 31u;
                const u32 m = d2_ld_acquire(rcw + (Pg.lead >> 5)) >> sh;
                const u32 ones = (~m) ? (u32)(__ffs(~m) - 1) : 32u; /* zeros were shifted in above bit 31 - sh */
                Pg.lead += ones;
                if (ones == 0u || (Pg.lead & 31u) != 0u) break;
            }
        };
        /* nothing to do until row Pg.lead completes: sleep on its barrier instead of polling */
        auto sleep_on_lead = [&]() { d2_mbar_wait(rowbar + 8u * Pg.lead, row_par); };

        unsigned long long tw_guard = 0, tw_plan = 0, tw_round = 0, tw_sleep = 0, tw_fin = 0, n_round = 0, n_sleep = 0, n_step = 0;
        const bool tr = P.trace != NULL && wic == 0;
#define D2_T() (tr ? (unsigned long long)clock64() : 0ull)
        for (;;) {
            u32 g = 0;
            if (lane == 0) g = atomicAdd(&ctrl[C_CLAIM], 1u);
            g = __shfl_sync(FULL, g, 0);
            if (g >= n_groups) break;
            const unsigned long long tt0 = D2_T();
            n_step++;
            /* staged literals above are overwritten by this group's output: every group that may
             * still read them (<= g - KGUARD) must be complete */
            if (g >= KGUARD) {
                const u32 need = min(n_rows, Z2_ROWS * (g - KGUARD + 1));
                while (Pg.lead < need) {
                    refresh_lead();
                    if (Pg.lead < need) sleep_on_lead();
                }
            }
            const unsigned long long tt1 = D2_T();
            tw_guard += tt1 - tt0;
            const i32 p0 = (i32)(g * Z2_GROUP);
            const u32 i0 = gidx[g];
            Pg.g4 = g * Z2_ROWS;
            /* sequence ends inside the group -> bitmask of the words where the next sequence starts */
            u32 M0 = 0, M1 = 0, M2 = 0, M3 = 0;
#pragma unroll 1
            for (u32 jj = 0; jj < 4; jj++) {
                const u32 k = i0 + lane + 32u * jj;
                bool in = false;
                u32 cw = 0;
                if (k < n_seq) {
                    const i32 rel = (i32)(d2_ld_rec(B, k).w0 & 0xFFFFu) + 1 - p0;
                    in = rel < (i32)Z2_GROUP;
                    cw = (u32)(rel + 3) >> 2;
                }
                const u32 bit = (in && cw < 128u) ? (1u << (cw & 31u)) : 0u;
                const u32 q = cw >> 5;
                M0 |= __reduce_or_sync(FULL, q == 0 ? bit : 0u);
                M1 |= __reduce_or_sync(FULL, q == 1 ? bit : 0u);
                M2 |= __reduce_or_sync(FULL, q == 2 ? bit : 0u);
                M3 |= __reduce_or_sync(FULL, q == 3 ? bit : 0u);
                if (!__shfl_sync(FULL, in, 31)) break;
            }
            const u32 pre1 = __popc(M0), pre2 = pre1 + __popc(M1), pre3 = pre2 + __popc(M2);

            i32 sX[Z2_ROWS], sY[Z2_ROWS], sZ[Z2_ROWS];
            u32 meta[Z2_ROWS]; /* t | t2 << 4 | flags << 8 | idx << 16 */
            u32 pend = 0, published = 0;
#pragma unroll
            for (u32 r = 0; r < Z2_ROWS; r++) {
                const u32 Mr = r == 0 ? M0 : r == 1 ? M1 : r == 2 ? M2 : M3;
                const u32 pre = r == 0 ? 0u : r == 1 ? pre1 : r == 2 ? pre2 : pre3;
                const u32 idx = i0 + pre + __popc(Mr & le_mask);
                const i32 p = p0 + (i32)(4u * (lane + 32u * r));
                const bool active = p < (i32)total;
                z2_seq_t c, n;
                d2_seq_pair(B, active ? idx : n_seq, c, n);
                const z2_plan_t pl = z2_word_plan(p, c, n, lit_pos);
                sX[r] = pl.srcX;
                sY[r] = pl.srcY;
                sZ[r] = pl.srcZ;
                meta[r] = pl.t | (pl.t2 << 4) | (pl.flags << 8) | (idx << 16);
                if (active) pend |= 1u << r;
                Pg.own[r] = __ballot_sync(FULL, !active);
            }
            refresh_lead();
            unsigned long long tt2 = D2_T();
            tw_plan += tt2 - tt1;
            /* rounds: a word is gathered once the words it reads are complete */
            for (;;) {
                n_round++;
                u32 rdy = 0;
#pragma unroll
                for (u32 r = 0; r < Z2_ROWS; r++) {
                    const u32 fl = (meta[r] >> 8) & 0xFFu;
                    if (((pend >> r) & 1u) && !(fl & Z2_SLOW)) {
                        bool ok = true;
                        if (fl & 1u) ok = d2_ready(Pg, sX[r]);
                        if ((fl & 2u) && ok) ok = d2_ready(Pg, sY[r]);
                        if ((fl & 4u) && ok) ok = d2_ready(Pg, sZ[r]);
                        if (ok) rdy |= 1u << r;
                    }
                }
                u32 did = 0;
#pragma unroll
                for (u32 r = 0; r < Z2_ROWS; r++) {
                    const i32 p = p0 + (i32)(4u * (lane + 32u * r));
                    if ((rdy >> r) & 1u) {
                        const u32 t = meta[r] & 15u, t2 = (meta[r] >> 4) & 15u;
                        u32 v = d2_gather(win, sX[r]);
                        if (t < 4u) {
                            const u32 vy = d2_gather(win, sY[r]);
                            v = __byte_perm(v, vy, 0x3210u | ((0x4444u << (4u * t)) & 0xFFFFu));
                            if (t2 < 4u) {
                                const u32 vz = d2_gather(win, sZ[r]);
                                v = __byte_perm(v, vz, 0x3210u | ((0x4444u << (4u * t2)) & 0xFFFFu));
                            }
                        }
                        *reinterpret_cast<volatile u32*>(win + p) = v;
                        did |= 1u << r;
                    } else if (((pend >> r) & 1u) && ((meta[r] >> 8) & Z2_SLOW)) {
                        u32 v;
                        if (d2_slow_word(B, Pg, meta[r] >> 16, p, v)) {
                            *reinterpret_cast<volatile u32*>(win + p) = v;
                            did |= 1u << r;
                        }
                    }
                }
                pend &= ~did;
                bool progress = false;
                u32 full = 0;
#pragma unroll
                for (u32 r = 0; r < Z2_ROWS; r++) {
                    const u32 m = __ballot_sync(FULL, (did >> r) & 1u);
                    Pg.own[r] |= m;
                    progress |= m != 0u;
                    if (Pg.own[r] == FULL) full |= 1u << r;
                }
                __syncwarp(); /* the words of this round are visible to the whole warp (next round reads them) */
                const u32 fresh = full & ~published;
                if (fresh) { /* rows that just completed: bitmap (release), then wake the sleepers */
                    if (lane == 0) {
                        d2_release_fence();
                        atomicOr(rcw + (Pg.g4 >> 5), fresh << (Pg.g4 & 31u));
#pragma unroll
                        for (u32 r = 0; r < Z2_ROWS; r++)
                            if ((fresh >> r) & 1u) d2_mbar_arrive(rowbar + 8u * (Pg.g4 + r));
                    }
                    published |= fresh;
                }
                if (!__any_sync(FULL, pend != 0u)) break;
                if (!progress) {
                    const u32 before = Pg.lead;
                    refresh_lead();
                    const unsigned long long ts0 = D2_T();
                    if (Pg.lead == before && Pg.lead < Pg.g4) {
                        sleep_on_lead();
                        n_sleep++;
                    }
                    refresh_lead();
                    const unsigned long long ts1 = D2_T();
                    tw_sleep += ts1 - ts0;
                    tt2 += ts1 - ts0;
                }
            }
            const unsigned long long tt3 = D2_T();
            tw_round += tt3 - tt2;
            /* group complete: ship the 8 KiB chunk it may have completed */
            d2_fence_async();
            __syncwarp();
            if (lane == 0) {
                d2_release_fence();
                const u32 wi = g >> 5, bit = 1u << (g & 31u);
                const u32 old = atomicOr(&ctrl[C_GDONE + wi], bit);
                if (bulk_ok) {
                    const u32 ch = g >> 4; /* 16 groups per chunk */
                    const u32 first = ch << 4;
                    const u32 cnt = min(16u, n_groups - first);
                    const u32 cm = ((1u << cnt) - 1u) << (first & 31u);
                    if (((old | bit) & cm) == cm) {
                        const u32 b0 = first * Z2_GROUP;
                        const u32 b1 = min(b0 + 16u * Z2_GROUP, total16);
                        if (b1 > b0) {
                            d2_fence_async();
                            d2_bulk_store(out + b0, d2_saddr(win + b0), b1 - b0);
                        }
                    }
                }
            }
            tw_fin += D2_T() - tt3;
        }
        if (tr && lane == 0) { /* second half of the job's trace slots is not used by the phase stamps: reuse slot 7 and a side table */
            unsigned long long* q = P.trace + (size_t)P.n_jobs * 8 + (size_t)j * 8;
            q[0] = tw_guard; q[1] = tw_plan; q[2] = tw_round; q[3] = tw_sleep; q[4] = tw_fin; q[5] = n_round; q[6] = n_sleep; q[7] = n_step;
        }
        __syncthreads();
        D2_STAMP(5);
        if (bulk_ok) {
            for (u32 k = total16 + tid; k < total; k += T) out[k] = win[k];
        } else {
            for (u32 k = tid; k < total; k += T) out[k] = win[k];
        }
        /* every row barrier advances exactly one phase per decoded block: the rows this block did not have */
        for (u32 r = n_groups * Z2_ROWS + tid; r < rows_max; r += T) d2_mbar_arrive(rowbar + 8u * r);
        row_par ^= 1u;
        if (tid == 0) P.status[j] = result;
        d2_bulk_wait_read();
        __syncthreads();
        D2_STAMP(6);
    }
    d2_bulk_wait_all();
}
