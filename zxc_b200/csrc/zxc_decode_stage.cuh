/*
 * zxc-b200 -- the block's sequential input streams (GLO tokens and offsets, GHI sequence words, raw literals) staged
 * through shared memory by the bulk-copy engine (TMA, cp.async.bulk global -> shared on an mbarrier).
 *
 * A warp decodes one block and walks each of these sections strictly forwards (zxc_decompress.c:1028-1190 reads them
 * through four advancing pointers).  Per stream the warp owns a small ring of ST_SLOTS chunks; coordinate x of a
 * stream is its byte offset from the 16-byte boundary at or below the section start, so chunk c = [c*CH, (c+1)*CH)
 * is 16-byte aligned in global memory and lands in slot c % ST_SLOTS, and the byte at x is ring[x % (ST_SLOTS*CH)].
 * Lane 0 issues chunk c (mbarrier.arrive.expect_tx + cp.async.bulk) as soon as chunk c - ST_SLOTS lies wholly below
 * the cursor -- one to ST_SLOTS-1 chunks ahead of the reads -- and every lane waits on the slot's mbarrier before the
 * first read of a chunk.  The copies stop at the last 16-byte boundary inside the block's payload (nothing outside
 * the block is read); a batch that reaches past it takes the plain global loads, as does a stream that does not live
 * in the payload (Huffman-decoded tokens, RLE / Huffman literals sit in the warp's scratch).
 *
 * Nothing of this lives in registers across a batch -- the decode loop is register-bound, every value kept would
 * be a spill.  A stream's state is two words of the warp's staging area (which sits right behind its output ring,
 * so its address follows from the ring's): x_clip, and one packed word with the chunks issued (12 bits), the chunks
 * waited for (12 bits) and the parity the next wait on each slot asks for.  need() brings a stream up to date at
 * one point per batch; close() waits for what is still in flight when a block ends, early or not.
 *
 * Slot reuse needs no extra fence: the previous batch's shared loads have delivered their values and the warp has
 * passed a __syncwarp() before lane 0 overwrites the slot.
 */
#pragma once

#ifndef ZXC_STAGE
#define ZXC_STAGE 0 /* measured on B200 (profiles/r02g_variants.txt): 415 GB/s without, 241 GB/s with -- see DESIGN.md section 3c-ter */
#endif
#ifndef ZXC_STAGE_LIT
#define ZXC_STAGE_LIT 0
#endif
#ifndef ZXC_BULK_FLUSH
#define ZXC_BULK_FLUSH 0 /* the ring leaves the SM by cp.async.bulk shared -> global instead of LDS.128 / STG.128 */
#endif
#define ST_SLOTS 2u
#define ST_CH 256u                       /* token / offset chunk: 8 GLO batches of tokens, 4 of offsets, 2 GHI batches */
#define ST_RING (ST_SLOTS * ST_CH)
#define ST_LIT_SLOTS 4u
#define ST_LIT_CH 512u
#define ST_LIT_RING (ST_LIT_SLOTS * ST_LIT_CH)
#define ST_BAR_BYTES 64u                 /* 2 + 2 + 4 mbarriers */
#define ST_STATE_BYTES 32u               /* per stream: x_clip, packed counters */
#if ZXC_STAGE
#define STAGE_BYTES (2u * ST_RING + (ZXC_STAGE_LIT ? ST_LIT_RING : 0u) + ST_BAR_BYTES + ST_STATE_BYTES)
#else
#define STAGE_BYTES 0u
#endif

#ifdef __CUDACC__
__device__ __forceinline__ void st_mbar_init(u32 bar) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void st_init_fence() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void st_issue(u32 sdst, const u8* gsrc, u32 bytes, u32 bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sdst),
                 "l"(gsrc), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void st_wait(u32 bar, u32 parity) {
    u32 ok = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}
/* shared -> global bulk copy of this thread's bulk group; the source must stay untouched until st_store_wait() */
__device__ __forceinline__ void st_store_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void st_store(u8* gdst, u32 ssrc, u32 bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void st_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void st_store_wait() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ u32 lds8(u32 a) {
    u32 v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ u32 lds16(u32 a) {
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ u32 lds32(u32 a) {
    u32 v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ const u8* st_generic(u32 a) { /* generic pointer of a shared address */
    return reinterpret_cast<const u8*>(__cvta_shared_to_generic((size_t)a));
}
#else
/* CPU emulator (tests/simt): the copy happens when lane 0 issues it; the wait is a rendezvous of the warp, which puts
 * lane 0's copy in front of the other lanes' reads the way the mbarrier does */
/* ... and keeps the books the hardware keeps: one copy in flight per barrier at most, every wait on the phase its
 * parity names, nothing in flight when a block ends (tests/simt/simt_decode.cc checks and aborts) */
extern u32 simt_bar_issued[64], simt_bar_waited[64];
extern u32 simt_bar_base;
void simt_stage_fail(const char* what);
static inline u32 simt_bar_idx(u32 bar) { return ((bar - simt_bar_base) / 8u) & 63u; }
static inline void st_mbar_init(u32 bar) { simt_bar_issued[simt_bar_idx(bar)] = simt_bar_waited[simt_bar_idx(bar)] = 0; }
static inline void st_init_fence() {}
static inline void st_issue(u32 sdst, const u8* gsrc, u32 bytes, u32 bar) {
    const u32 b = simt_bar_idx(bar);
    if (simt_bar_issued[b] != simt_bar_waited[b]) simt_stage_fail("second copy on a barrier whose first was not waited for");
    if (bytes == 0 || (bytes & 15u) || (reinterpret_cast<uintptr_t>(gsrc) & 15u) || (sdst & 15u)) simt_stage_fail("bulk copy alignment");
    simt_bar_issued[b]++;
    memcpy(smem + sdst, gsrc, bytes);
}
static inline void st_wait(u32 bar, u32 parity) {
    __syncwarp();
    if (simt::g_warp->current == 0) {
        const u32 b = simt_bar_idx(bar);
        if (simt_bar_issued[b] != simt_bar_waited[b] + 1u) simt_stage_fail("wait without a copy in flight");
        if ((simt_bar_waited[b] & 1u) != parity) simt_stage_fail("wait on the wrong parity");
        simt_bar_waited[b]++;
    }
    __syncwarp();
}
/* shared -> global bulk copies are DEFERRED to the wait, the latest moment the hardware may take: a ring byte
 * overwritten or an output byte read before the wait shows up as a wrong result */
struct SimtStore { u8* g; u32 s, n; };
extern SimtStore simt_stores[64];
extern u32 simt_n_stores;
static inline void st_store_fence() {}
static inline void st_store(u8* gdst, u32 ssrc, u32 bytes) {
    if (bytes == 0 || (bytes & 15u) || (reinterpret_cast<uintptr_t>(gdst) & 15u) || (ssrc & 15u)) simt_stage_fail("bulk store alignment");
    if (simt_n_stores >= 64) simt_stage_fail("too many bulk stores in flight");
    simt_stores[simt_n_stores++] = SimtStore{gdst, ssrc, bytes};
}
static inline void st_store_commit() {}
static inline void st_store_wait() {
    for (u32 i = 0; i < simt_n_stores; i++) memcpy(simt_stores[i].g, smem + simt_stores[i].s, simt_stores[i].n);
    simt_n_stores = 0;
}
static inline u32 lds8(u32 a) { return smem[a]; }
static inline u32 lds16(u32 a) { unsigned short v; memcpy(&v, smem + a, 2); return v; }
static inline u32 lds32(u32 a) { u32 v; memcpy(&v, smem + a, 4); return v; }
static inline const u8* st_generic(u32 a) { return smem + a; }
#endif

struct StWindow { u32 ring_s, lo, hi; }; /* coordinates [lo, hi) are in the ring at ring_s + (x % its size) */

template <u32 SLOTS, u32 CH, u32 RING_OFF, u32 BAR_OFF, u32 STATE_OFF>
struct Stream {
    typedef StWindow Window;

    static __device__ __forceinline__ u32 pack(u32 issued, u32 waited, u32 phase) { return issued | (waited << 12) | (phase << 24); }
    static __device__ __forceinline__ void wait_upto(u32 a, u32& waited, u32 to, u32& phase) {
#pragma unroll 1
        for (; waited < to; waited++) {
            const u32 sl = waited % SLOTS;
            st_wait(a + BAR_OFF + 8u * sl, (phase >> sl) & 1u);
            phase ^= 1u << sl;
        }
    }
    /* start of a block: the stream is `bytes` long from p; copies end at the last 16-byte boundary at or below lim */
    static __device__ __forceinline__ void open(u32 a, const u8* p, u32 bytes, const u8* lim, bool on) {
        const u32 x0 = (u32)(reinterpret_cast<uintptr_t>(p) & 15u);
        const u8* g0 = p - x0;
        u32 clip = 0;
        if (on && lim > g0 && bytes >= 64u && bytes < 0xFF0u * CH) { /* 12-bit chunk counters */
            const u64 room = (u64)(lim - g0) & ~15ull;
            const u64 want = ((u64)x0 + bytes + 15ull) & ~15ull;
            clip = (u32)(want < room ? want : room);
        }
        const u32 st = lds32(a + STATE_OFF + 4u);
        __syncwarp(); /* every lane has read the word before any lane rewrites it */
        sts32<0>(a + STATE_OFF, clip);
        sts32<4>(a + STATE_OFF, pack(0u, 0u, st >> 24)); /* the barriers keep their phases from block to block */
    }
    /* A batch reads [x_lo, x_hi) (+- slop bytes that aligned word loads may touch): retire the chunks below it, issue
     * what the ring has room for, wait for what the range needs.  Returns what is resident. */
    static __device__ __forceinline__ Window need(u32 a, const u8* p, u32 x_lo, u32 x_hi, u32 slop, u32 lane) {
        Window w;
        w.ring_s = a + RING_OFF;
        w.lo = w.hi = 0;
        const u32 clip = lds32(a + STATE_OFF);
        if (clip) {
            const u32 st = lds32(a + STATE_OFF + 4u);
            __syncwarp(); /* every lane has read the word before any lane rewrites it */
            u32 issued = st & 0xFFFu, waited = (st >> 12) & 0xFFFu, phase = st >> 24;
            const u32 n = (clip + CH - 1u) / CH;
            const u32 first = min(n, (x_lo >= slop ? x_lo - slop : 0u) / CH); /* chunks below it are dead */
            wait_upto(a, waited, min(issued, first), phase);                 /* ... but land before their slot is reused */
            if (issued < first) issued = waited = first;                     /* never asked for (a giant run): skipped */
            const u32 may = min(n, first + SLOTS);
            if (issued < may) {
                if (lane == 0) {
                    const u8* g0 = p - (reinterpret_cast<uintptr_t>(p) & 15u);
#pragma unroll 1
                    for (u32 c = issued; c < may; c++)
                        st_issue(a + RING_OFF + (c % SLOTS) * CH, g0 + (size_t)c * CH, min(CH, clip - c * CH),
                                 a + BAR_OFF + 8u * (c % SLOTS));
                }
                issued = may;
            }
            wait_upto(a, waited, min(issued, (x_hi + slop + CH - 1u) / CH), phase);
            sts32<4>(a + STATE_OFF, pack(issued, waited, phase));
            w.lo = first * CH;
            w.hi = min(waited * CH, clip);
        }
        return w;
    }
    /* end of a block, early or not: nothing stays in flight */
    static __device__ __forceinline__ void close(u32 a) {
        if (lds32(a + STATE_OFF)) {
            const u32 st = lds32(a + STATE_OFF + 4u);
            __syncwarp();
            u32 issued = st & 0xFFFu, waited = (st >> 12) & 0xFFFu, phase = st >> 24;
            wait_upto(a, waited, issued, phase);
            sts32<4>(a + STATE_OFF, pack(issued, waited, phase));
        }
    }
};

/* layout of a warp's staging area at shared address a: token ring, offset ring, [literal ring,] mbarriers, state */
#define ST_OFF_LIT (2u * ST_RING)
#define ST_OFF_BAR (2u * ST_RING + (ZXC_STAGE_LIT ? ST_LIT_RING : 0u))
#define ST_OFF_STATE (ST_OFF_BAR + ST_BAR_BYTES)
typedef Stream<ST_SLOTS, ST_CH, 0u, ST_OFF_BAR, ST_OFF_STATE> TokStream;
typedef Stream<ST_SLOTS, ST_CH, ST_RING, ST_OFF_BAR + 8u * ST_SLOTS, ST_OFF_STATE + 8u> OffStream;
typedef Stream<ST_LIT_SLOTS, ST_LIT_CH, ST_OFF_LIT, ST_OFF_BAR + 16u * ST_SLOTS, ST_OFF_STATE + 16u> LitStream;

/* once per warp, before its first block */
__device__ __forceinline__ void st_init(u32 a, u32 lane) {
    if (lane == 0) {
        for (u32 s = 0; s < 2u * ST_SLOTS + (ZXC_STAGE_LIT ? ST_LIT_SLOTS : 0u); s++) st_mbar_init(a + ST_OFF_BAR + 8u * s);
        for (u32 k = 0; k < ST_STATE_BYTES; k += 4) sts32<0>(a + ST_OFF_STATE + k, 0u);
        st_init_fence();
    }
    __syncwarp();
}

/* generic pointer of the literal at coordinate x when [x - 8, x + n + 8) is resident and does not wrap; else 0 */
__device__ __forceinline__ const u8* lit_ptr(const StWindow& w, u32 x, u32 n) {
    const bool in = x >= w.lo + 8u && x + n + 8u <= w.hi && ((x - 8u) / ST_LIT_RING) == ((x + n + 7u) / ST_LIT_RING);
    return in ? st_generic(w.ring_s + (x & (ST_LIT_RING - 1u))) : (const u8*)0;
}
