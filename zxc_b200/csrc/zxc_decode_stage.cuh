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
 * No counters are kept: with the cursor at x_lo the chunks below issued(x_lo) = min(n_chunks, x_lo / CH + SLOTS) have
 * been issued, and a batch that reads up to x_hi has waited for the chunks below waited(x_hi) = min(n_chunks,
 * ceil(x_hi / CH)); the step from one batch to the next issues / waits for the difference.  The mbarriers are
 * re-initialised per block, so the parity of chunk c is (c / SLOTS) & 1.  A block that ends early waits for the
 * copies between those two marks (close()); a block that runs to its end has waited for all of them.
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
#ifndef ST_COLD
#define ST_COLD __forceinline__ /* __noinline__ measured: a call inside the decode loop costs far more than its size */
#endif
#define ST_SLOTS 2u
#define ST_CH 256u                       /* token / offset chunk: 8 GLO batches of tokens, 4 of offsets, 2 GHI batches */
#define ST_RING (ST_SLOTS * ST_CH)
#define ST_LIT_SLOTS 4u
#define ST_LIT_CH 512u
#define ST_LIT_RING (ST_LIT_SLOTS * ST_LIT_CH)
#define ST_BAR_BYTES 64u                 /* 2 + 2 + 4 mbarriers */
#define ST_STATE_BYTES 16u               /* x_clip per stream, the literal stream's counters */
#if ZXC_STAGE
#define STAGE_BYTES (2u * ST_RING + (ZXC_STAGE_LIT ? ST_LIT_RING : 0u) + ST_BAR_BYTES + ST_STATE_BYTES)
#else
#define STAGE_BYTES 0u
#endif

#ifdef __CUDACC__
__device__ __forceinline__ void st_mbar_init(u32 bar) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void st_mbar_reinit(u32 bar) {
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
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
static inline void st_mbar_reinit(u32 bar) {
    if (simt_bar_issued[simt_bar_idx(bar)] != simt_bar_waited[simt_bar_idx(bar)]) simt_stage_fail("re-init of a barrier with a copy in flight");
    st_mbar_init(bar);
}
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

/* One stream.  Nothing of it lives in registers across a batch: the shared addresses follow from the warp's ring
 * address (a warp's staging area sits right behind its output ring), x_clip is a word of that area, and the counters
 * are functions of the cursor (above).  The decode loop is register-bound; every value kept here would be a spill. */
template <u32 SLOTS, u32 CH>
struct StageStream {
    u32 ring_s; /* shared address of the ring (SLOTS * CH bytes); its SLOTS mbarriers sit at bar_s */
    u32 bar_s;
    u32 x_clip; /* coordinates below this are staged (a multiple of 16); 0 = stream not staged in this block */

    __device__ __forceinline__ u32 n_chunks() const { return (x_clip + CH - 1u) / CH; }
    __device__ __forceinline__ u32 issued_at(u32 x_lo) const { return min(n_chunks(), x_lo / CH + SLOTS); }
    __device__ __forceinline__ u32 waited_at(u32 x_hi) const { return min(n_chunks(), (x_hi + CH - 1u) / CH); }
    /* out of line: a chunk boundary is crossed once in several batches, and the decode loop has to stay small */
    __device__ ST_COLD void issue(const u8* g0, u32 from, u32 to, u32 lane) const {
        if (lane == 0) {
#pragma unroll 1
            for (u32 c = from; c < to; c++)
                st_issue(ring_s + (c % SLOTS) * CH, g0 + (size_t)c * CH, min(CH, x_clip - c * CH), bar_s + 8u * (c % SLOTS));
        }
    }
    __device__ ST_COLD void wait(u32 from, u32 to) const {
#pragma unroll 1
        for (u32 c = from; c < to; c++) st_wait(bar_s + 8u * (c % SLOTS), (c / SLOTS) & 1u);
    }
    /* Start a block: the stream is `bytes` long from p (coordinate p & 15); copies end at the last 16-byte boundary
     * at or below `lim`.  Issues the first chunks and waits for what the first batch reads (up to x_hi).  Returns
     * x_clip (the caller stores it). */
    __device__ __forceinline__ u32 open(const u8* p, u32 bytes, const u8* lim, bool on, u32 x_hi, u32 lane) {
        const u32 x0 = (u32)(reinterpret_cast<uintptr_t>(p) & 15u);
        const u8* g0 = p - x0;
        x_clip = 0;
        if (on && lim > g0) {
            const u64 room = (u64)(lim - g0) & ~15ull;
            const u64 want = ((u64)x0 + bytes + 15ull) & ~15ull;
            x_clip = (u32)(want < room ? want : room);
        }
        if (x_clip) {
            if (lane == 0) {
                for (u32 s = 0; s < SLOTS; s++) st_mbar_reinit(bar_s + 8u * s);
                st_init_fence();
            }
            __syncwarp();
            issue(g0, 0, issued_at(x0), lane);
            wait(0, waited_at(x_hi));
        }
        return x_clip;
    }
    /* from a batch at [lo, hi) to the next one at [nlo, nhi) */
    __device__ __forceinline__ void step(const u8* p, u32 lo, u32 hi, u32 nlo, u32 nhi, u32 lane) const {
        if (x_clip) {
            const u32 a = issued_at(lo), b = issued_at(nlo);
            if (a < b) issue(p - (reinterpret_cast<uintptr_t>(p) & 15u), a, b, lane);
            const u32 c = waited_at(hi), d = waited_at(nhi);
            if (c < d) wait(c, d);
        }
    }
    /* the block ends inside the batch at [lo, hi): nothing may stay in flight */
    __device__ __forceinline__ void close(u32 lo, u32 hi) const {
        if (x_clip) wait(waited_at(hi), issued_at(lo));
    }
    __device__ __forceinline__ bool staged(u32 x_hi) const { return x_hi <= x_clip; }
};

typedef StageStream<ST_SLOTS, ST_CH> SeqStream;
typedef StageStream<ST_LIT_SLOTS, ST_LIT_CH> LitStream;
/* layout of a warp's staging area at shared address a: token ring, offset ring, [literal ring,] mbarriers, state */
#define ST_OFF_BAR (2u * ST_RING + (ZXC_STAGE_LIT ? ST_LIT_RING : 0u))
#define ST_OFF_STATE (ST_OFF_BAR + ST_BAR_BYTES)
__device__ __forceinline__ SeqStream st_tok(u32 a) {
    SeqStream s;
    s.ring_s = a;
    s.bar_s = a + ST_OFF_BAR;
    s.x_clip = lds32(a + ST_OFF_STATE);
    return s;
}
__device__ __forceinline__ SeqStream st_off(u32 a) {
    SeqStream s;
    s.ring_s = a + ST_RING;
    s.bar_s = a + ST_OFF_BAR + 8u * ST_SLOTS;
    s.x_clip = lds32(a + ST_OFF_STATE + 4u);
    return s;
}
__device__ __forceinline__ LitStream st_lit(u32 a) {
    LitStream s;
    s.ring_s = a + 2u * ST_RING;
    s.bar_s = a + ST_OFF_BAR + 16u * ST_SLOTS;
    s.x_clip = lds32(a + ST_OFF_STATE + 8u);
    return s;
}
/* once per warp, before its first block */
__device__ __forceinline__ void st_init(u32 a, u32 lane) {
    if (lane == 0) {
        for (u32 s = 0; s < 2u * ST_SLOTS + (ZXC_STAGE_LIT ? ST_LIT_SLOTS : 0u); s++) st_mbar_init(a + ST_OFF_BAR + 8u * s);
        for (u32 k = 0; k < ST_STATE_BYTES; k += 4) sts32<0>(a + ST_OFF_STATE + k, 0u);
        st_init_fence();
    }
    __syncwarp();
}

#if ZXC_STAGE_LIT
/* The literal stream.  Its cursor moves by the batch's literal total -- anything from nothing to the whole block (a
 * giant run skips the ring altogether) -- so its counters are kept, packed into one shared word: chunks issued (12
 * bits), chunks waited for (12 bits), the parity the next wait on each of the four slots asks for (4 bits).  A lane's
 * literal run comes out of the ring when it lies, with the 8 bytes either side that the aligned word loads of the
 * copy helpers may touch, inside the resident chunks and does not wrap the ring; otherwise out of global memory. */
struct LitWindow {
    u32 ring_s;
    u32 lo, hi; /* coordinates [lo, hi) are resident */
};
__device__ __forceinline__ u32 lit_pack(u32 issued, u32 waited, u32 phase) { return issued | (waited << 12) | (phase << 24); }
__device__ __forceinline__ void lit_wait(const LitStream& s, u32& waited, u32 to, u32& phase) {
    for (; waited < to; waited++) {
        const u32 sl = waited % ST_LIT_SLOTS;
        st_wait(s.bar_s + 8u * sl, (phase >> sl) & 1u);
        phase ^= 1u << sl;
    }
}
/* start of a block */
__device__ __forceinline__ void lit_open(u32 a, const u8* p, u32 bytes, const u8* lim, bool on) {
    const u32 x0 = (u32)(reinterpret_cast<uintptr_t>(p) & 15u);
    const u8* g0 = p - x0;
    u32 clip = 0;
    if (on && lim > g0 && bytes >= 64u && bytes < 0x1FE000u) { /* 12-bit chunk counters */
        const u64 room = (u64)(lim - g0) & ~15ull;
        const u64 want = ((u64)x0 + bytes + 15ull) & ~15ull;
        clip = (u32)(want < room ? want : room);
    }
    sts32<8>(a + ST_OFF_STATE, clip);
    const u32 st = lds32(a + ST_OFF_STATE + 12u);
    __syncwarp(); /* every lane has read the word before any lane rewrites it */
    sts32<12>(a + ST_OFF_STATE, lit_pack(0u, 0u, st >> 24)); /* the barriers keep their phases from block to block */
}
/* a batch reads literals [x_lo, x_hi): bring the ring up to date, say what is resident */
__device__ __forceinline__ LitWindow lit_need(u32 a, const u8* p, u32 x_lo, u32 x_hi, u32 lane) {
    const LitStream s = st_lit(a);
    LitWindow w;
    w.ring_s = s.ring_s;
    w.lo = w.hi = 0;
    if (s.x_clip) {
        const u32 st = lds32(a + ST_OFF_STATE + 12u);
        __syncwarp(); /* every lane has read the word before any lane rewrites it */
        u32 issued = st & 0xFFFu, waited = (st >> 12) & 0xFFFu, phase = st >> 24;
        const u32 n = s.n_chunks();
        const u32 first = min(n, (x_lo >= 8u ? x_lo - 8u : 0u) / ST_LIT_CH); /* chunks below it are dead */
        lit_wait(s, waited, min(issued, first), phase);                      /* ... but land before their slot is reused */
        if (issued < first) issued = waited = first;                         /* never asked for: skipped */
        const u32 may = min(n, first + ST_LIT_SLOTS);
        if (issued < may) {
            if (lane == 0) {
                const u8* g0 = p - (reinterpret_cast<uintptr_t>(p) & 15u);
                for (u32 c = issued; c < may; c++)
                    st_issue(s.ring_s + (c % ST_LIT_SLOTS) * ST_LIT_CH, g0 + (size_t)c * ST_LIT_CH,
                             min(ST_LIT_CH, s.x_clip - c * ST_LIT_CH), s.bar_s + 8u * (c % ST_LIT_SLOTS));
            }
            issued = may;
        }
        lit_wait(s, waited, min(issued, (x_hi + 8u + ST_LIT_CH - 1u) / ST_LIT_CH), phase);
        sts32<12>(a + ST_OFF_STATE, lit_pack(issued, waited, phase));
        w.lo = first * ST_LIT_CH;
        w.hi = min(waited * ST_LIT_CH, s.x_clip);
    }
    return w;
}
/* end of a block, early or not: nothing stays in flight */
__device__ __forceinline__ void lit_close(u32 a) {
    const LitStream s = st_lit(a);
    if (s.x_clip) {
        const u32 st = lds32(a + ST_OFF_STATE + 12u);
        __syncwarp();
        u32 issued = st & 0xFFFu, waited = (st >> 12) & 0xFFFu, phase = st >> 24;
        lit_wait(s, waited, issued, phase);
        sts32<12>(a + ST_OFF_STATE, lit_pack(issued, waited, phase));
    }
}
/* generic pointer of the literal at coordinate x when [x - 8, x + n + 8) is resident and does not wrap; else 0 */
__device__ __forceinline__ const u8* lit_ptr(const LitWindow& w, u32 x, u32 n) {
    const bool in = x >= w.lo + 8u && x + n + 8u <= w.hi && ((x - 8u) / ST_LIT_RING) == ((x + n + 7u) / ST_LIT_RING);
    return in ? st_generic(w.ring_s + (x & (ST_LIT_RING - 1u))) : (const u8*)0;
}
#endif
