/*
 * zxc_decode_units.cuh -- output-centric LZ body of the warp-per-block decode kernel (device code only).
 *
 * The reference decodes a block sequence by sequence (zxc_decompress.c:847-1209 GLO, :1231-1469 GHI):
 * copy ll literals, copy ml match bytes, next.  Mapped naively onto a warp (lane = sequence) most of
 * the instructions go into making variable-length, unaligned copies line up.  Here the block is
 * decoded in two passes by ONE warp:
 *
 *   pass 1 (lane = sequence, 32 per step): ll / ml / off with the extras resolved (values of the whole
 *          extras section come from a segment-map scan, zxc_decode2_core.h; varint semantics of
 *          :51-88), running sums, validation in the reference's order, one 8-byte record per sequence
 *          (end, first match byte, distance, match bytes before it) and, per aligned 16-byte output unit,
 *          the index of the sequence that covers its first byte.  Both tables live in the warp's
 *          scratch in HBM (L2 resident).
 *   pass 2 (lane = 16-byte output unit): lane l owns units l, l+32, l+64, ... and walks the regions that
 *          cut its unit (a "piece" = literal run or match, clipped to the unit) one per iteration: an
 *          unaligned 16-byte gather, merged into four registers; a finished unit leaves as one 16-byte
 *          store.  A match piece waits until the units it reads are finished -- lanes publish how many
 *          units they have finished and readers fetch that with a shuffle, so lanes run ahead of each
 *          other as far as the data allows (no block barriers, no fences, no shared memory).
 *          Distances below 16 (runs, short periods) and pieces that straddle the dictionary boundary go
 *          byte by byte.
 *
 * Output is written exactly once, in whole aligned units; match sources are read back from the output
 * buffer (L1 / L2).  Used for blocks of at most 64 KiB decoded (16-bit record fields); larger blocks keep
 * the sequence-centric body in zxc_decode.cuh.
 */
#pragma once
#include "zxc_decode2_core.h"

#define UW_NOT_TAKEN (-1000) /* tables do not fit the scratch: caller uses the sequence-centric body */

__device__ __forceinline__ z2_rec_t uw_ld_rec(const z2_rec_t* rec, u32 i) {
    const uint2 v = *reinterpret_cast<const uint2*>(rec + i);
    z2_rec_t r;
    r.w0 = v.x;
    r.w1 = v.y;
    return r;
}

/* bytes sh .. sh+15 of the 32 bytes A:B (sh < 16): a two-stage word barrel and four funnel shifts */
__device__ __forceinline__ void uw_extract16(uint4 A, uint4 B, u32 sh, u32& w0, u32& w1, u32& w2, u32& w3) {
    const bool s1 = (sh & 4u) != 0u, s2 = (sh & 8u) != 0u;
    const u32 x0 = s1 ? A.y : A.x, x1 = s1 ? A.z : A.y, x2 = s1 ? A.w : A.z, x3 = s1 ? B.x : A.w, x4 = s1 ? B.y : B.x,
              x5 = s1 ? B.z : B.y, x6 = s1 ? B.w : B.z;
    const u32 y0 = s2 ? x2 : x0, y1 = s2 ? x3 : x1, y2 = s2 ? x4 : x2, y3 = s2 ? x5 : x3, y4 = s2 ? x6 : x4;
    const u32 bs = (sh & 3u) * 8u;
    w0 = __funnelshift_r(y0, y1, bs);
    w1 = __funnelshift_r(y1, y2, bs);
    w2 = __funnelshift_r(y2, y3, bs);
    w3 = __funnelshift_r(y3, y4, bs);
}
/* 128-bit value x3:x2:x1:x0 shifted left by k bytes (0..16), zeros shifted in */
__device__ __forceinline__ void uw_shl128(u32& x0, u32& x1, u32& x2, u32& x3, u32 k) {
    const u32 ws = k >> 2, bs = (k & 3u) * 8u;
    const u32 y3 = ws == 0 ? x3 : ws == 1 ? x2 : ws == 2 ? x1 : ws == 3 ? x0 : 0u;
    const u32 y2 = ws == 0 ? x2 : ws == 1 ? x1 : ws == 2 ? x0 : 0u;
    const u32 y1 = ws == 0 ? x1 : ws == 1 ? x0 : 0u;
    const u32 y0 = ws == 0 ? x0 : 0u;
    x3 = __funnelshift_l(y2, y3, bs);
    x2 = __funnelshift_l(y1, y2, bs);
    x1 = __funnelshift_l(y0, y1, bs);
    x0 = y0 << bs;
}

/* 16 bytes starting at byte address p (any alignment); only the vectors that hold bytes [lo, hi) of the
 * result are loaded, so nothing outside the source range is touched */
__device__ __forceinline__ void uw_read16(const u8* p, u32 lo, u32 hi, u32& w0, u32& w1, u32& w2, u32& w3) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const u32 sh = (u32)(a & 15u);
    const uint4* q = reinterpret_cast<const uint4*>(a - sh);
    uint4 A = make_uint4(0, 0, 0, 0), B = make_uint4(0, 0, 0, 0);
    if (sh + lo < 16u) A = q[0];
    if (sh + hi > 16u) B = q[1];
    uw_extract16(A, B, sh, w0, w1, w2, w3);
}

/* mask of the bytes of word j (bytes 4j..4j+3 of the unit) that lie below unit offset d */
__device__ __forceinline__ u32 uw_lowmask(u32 d, u32 j) {
    const i32 t = (i32)d - (i32)(4u * j);
    return t >= 4 ? 0xFFFFFFFFu : (t <= 0 ? 0u : ((1u << (8u * (u32)t)) - 1u));
}
__device__ __forceinline__ u32 uw_get_byte(u32 a0, u32 a1, u32 a2, u32 a3, u32 i) {
    const u32 w = i < 8u ? (i < 4u ? a0 : a1) : (i < 12u ? a2 : a3);
    return (w >> (8u * (i & 3u))) & 0xFFu;
}
__device__ __forceinline__ void uw_set_byte(u32& a0, u32& a1, u32& a2, u32& a3, u32 i, u32 v) {
    const u32 sh = 8u * (i & 3u), m = ~(0xFFu << sh), b = v << sh;
    if (i < 4u) a0 = (a0 & m) | b;
    else if (i < 8u) a1 = (a1 & m) | b;
    else if (i < 12u) a2 = (a2 & m) | b;
    else a3 = (a3 & m) | b;
}

/* Returns decoded bytes, a negative zxc_error_t, or UW_NOT_TAKEN.
 * tok/offs/ext/lit as prepared by parse_sections; tab: the warp's table area in HBM (8-byte aligned). */
__device__ __noinline__ int decode_lz_units(const u8* lit, u32 n_lit_avail, const u8* tok, const u8* offs, const u8* ext, u32 ext_end,
                               u32 n_seq, u32 enc_off, bool ghi, u8* out, u32 cap, const u8* dict, u32 dict_size, u8* tab,
                               u32 tab_bytes, u32 lane) {
    if (cap > 65536u || cap == 0u || n_seq > 0xFFF0u) return UW_NOT_TAKEN;
    const u32 n_units_cap = (cap + 15u) >> 4;
    const u32 rec_bytes = (n_seq + 4u) * 8u, uidx_bytes = ((n_units_cap + 2u) * 2u + 7u) & ~7u;
    if ((u64)rec_bytes + uidx_bytes + 4ull * ext_end + 16 > tab_bytes) return UW_NOT_TAKEN;
    z2_rec_t* rec = reinterpret_cast<z2_rec_t*>(tab);
    unsigned short* uidx = reinterpret_cast<unsigned short*>(tab + rec_bytes);
    u32* vals = reinterpret_cast<u32*>(tab + rec_bytes + uidx_bytes);
    const u32 esc = ghi ? 255u : 15u;
    const u32 lt_mask = (1u << lane) - 1u;

    /* ---- pass 1a: every varint value of the extras section ---- */
    u32 n_val = 0;
    if (ext_end) {
        const u32 seg = max(4u, (ext_end + 31u) / 32u);
        const u32 nseg = (ext_end + seg - 1u) / seg;
        const u32 lo = lane * seg, hi = min(ext_end, lo + seg);
        const u64 map = lane < nseg ? z2_seg_map(ext, lo, hi, ext_end) : Z2_MAP_ID;
        u64 inc = map;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const u64 o = __shfl_up_sync(FULL, inc, d);
            if (lane >= (u32)d) inc = z2_map_compose(o, inc);
        }
        u64 excl = __shfl_up_sync(FULL, inc, 1);
        if (lane == 0) excl = Z2_MAP_ID;
        n_val = z2_map_cnt(__shfl_sync(FULL, inc, 31), 0);
        if (lane < nseg) {
            const u32 ent = z2_map_exit(excl, 0);
            if (ent != 3u) z2_seg_values(ext, lo, hi, ext_end, ent, z2_map_cnt(excl, 0), vals);
        }
        __syncwarp();
    }

    /* ---- pass 1b: records and the per-unit sequence index ---- */
    u32 L = 0, O = 0, ord = 0;
    for (u32 base = 0; base < n_seq; base += 32u) {
        const u32 i = base + lane;
        const bool valid = i < n_seq;
        u32 ll = 0, ml = 0, off = 1;
        if (valid) {
            if (!ghi) {
                const u32 t = tok[i];
                ll = t >> 4;
                ml = t & 15u;
                off = (enc_off ? (u32)offs[i] : ld16(offs + 2u * (size_t)i)) + 1u;
            } else {
                const u32 w = ld32(tok + 4u * (size_t)i);
                ll = w >> 24;
                ml = (w >> 16) & 0xFFu;
                off = (w & 0xFFFFu) + 1u;
            }
        }
        const bool e_ll = valid && ll == esc, e_ml = valid && ml == esc;
        const u32 m_ll = __ballot_sync(FULL, e_ll), m_ml = __ballot_sync(FULL, e_ml);
        if (m_ll | m_ml) {
            u32 k = ord + __popc(m_ll & lt_mask) + __popc(m_ml & lt_mask);
            if (e_ll) {
                ll += k < n_val ? vals[k] : 0u;
                k++;
            }
            if (e_ml) ml += k < n_val ? vals[k] : 0u;
            ord += __popc(m_ll) + __popc(m_ml);
        }
        if (valid) ml += 5u;
        ll = min(ll, 0xFFFFu);
        ml = min(ml, 0xFFFFu);
        const u32 tot = ll + ml;
        const u32 s_ll = warp_incl_scan(ll, lane), s_tot = warp_incl_scan(tot, lane);
        const u32 ls = L + s_ll - ll, os = O + s_tot - tot;
        const u32 md = os + ll, E = md + ml;
        const bool ovf = valid && (ls + ll > n_lit_avail || E > cap);
        const bool bad = valid && (md + dict_size < off);
        const u32 m_err = __ballot_sync(FULL, ovf || bad);
        if (m_err) {
            const int code = ovf ? ZXC_ERROR_OVERFLOW : ZXC_ERROR_BAD_OFFSET;
            return __shfl_sync(FULL, code, __ffs(m_err) - 1);
        }
        if (valid) {
            const z2_rec_t r = z2_pack(E, md, off, os - ls);
            *reinterpret_cast<uint2*>(rec + i) = make_uint2(r.w0, r.w1);
        }
        /* units whose first byte this sequence covers */
        const u32 u_lo = valid ? (os + 15u) >> 4 : 0u, u_hi = valid ? (E + 15u) >> 4 : 0u;
        const bool longrun = u_hi - u_lo > 4u;
        if (!longrun)
            for (u32 q = u_lo; q < u_hi; q++) uidx[q] = (unsigned short)i;
        u32 m_long = __ballot_sync(FULL, longrun);
        while (m_long) { /* long literal runs / matches: the warp fills the range together */
            const int j = __ffs(m_long) - 1;
            m_long &= m_long - 1;
            const u32 a = __shfl_sync(FULL, u_lo, j), b = __shfl_sync(FULL, u_hi, j);
            for (u32 q = a + lane; q < b; q += 32u) uidx[q] = (unsigned short)(base + (u32)j);
        }
        L += __shfl_sync(FULL, s_ll, 31);
        O += __shfl_sync(FULL, s_tot, 31);
    }
    /* trailing literals (zxc_decompress.c:1198-1206): a virtual sequence without a match, then a sentinel */
    const u32 rem = n_lit_avail - L;
    if (rem > cap - O) return ZXC_ERROR_OVERFLOW;
    const u32 total = O + rem;
    if (total == 0u) return 0;
    const u32 n_units = (total + 15u) >> 4;
    if (lane == 0) {
        const z2_rec_t v = z2_pack(total, total & 0xFFFFu, 1u, (O - L) & 0xFFFFu);
        const z2_rec_t s = z2_pack(0x10000u, total & 0xFFFFu, 1u, 0u);
        *reinterpret_cast<uint2*>(rec + n_seq) = make_uint2(v.w0, v.w1);
        *reinterpret_cast<uint2*>(rec + n_seq + 1) = make_uint2(s.w0, s.w1);
        *reinterpret_cast<uint2*>(rec + n_seq + 2) = make_uint2(s.w0, s.w1);
    }
    for (u32 q = ((O + 15u) >> 4) + lane; q < n_units; q += 32u) uidx[q] = (unsigned short)n_seq;
    __syncwarp();

    /* ---- pass 2: units ---- */
    const bool out16 = (reinterpret_cast<uintptr_t>(out) & 15u) == 0;
    u32 u = lane, cnt = 0; /* current unit, units finished by this lane */
    bool active = u < n_units;
    u32 pos = 0, uend = 0, idx = 0;
    i32 cE = 0, cmd = 0, coff = 1, cM = 0;
    u32 a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#define UW_LOAD_SEQ()                                             \
    do {                                                          \
        const z2_seq_t _s = z2_unpack(uw_ld_rec(rec, idx));       \
        cE = idx >= n_seq ? (i32)total : _s.E;                    \
        cmd = idx >= n_seq ? Z2_MD_INF : _s.md;                   \
        coff = _s.off;                                            \
        cM = _s.M;                                                \
    } while (0)
    if (active) {
        pos = u << 4;
        uend = min(pos + 16u, total);
        idx = uidx[u];
        UW_LOAD_SEQ();
    }
    for (;;) {
        /* the piece in front of this lane */
        bool is_lit = false, slow = false, need = false, fold = false;
        u32 n = 0, d = 0, rot = 0;
        i32 s0 = 0;
        u32 ua = 0, ub = 0;
        if (active) {
            if ((i32)pos >= cE) {
                idx++;
                UW_LOAD_SEQ();
            }
            is_lit = (i32)pos < cmd;
            const u32 rend = is_lit ? (u32)min(cmd, (i32)uend) : (u32)min(cE, (i32)uend);
            n = rend - pos;
            d = pos & 15u;
            if (!is_lit) {
                s0 = (i32)pos - coff;
                const i32 kk = (i32)pos - cmd; /* offset into the match */
                if (kk >= coff && cmd < (i32)(u << 4) && cmd - coff >= 0) {
                    /* the plain source lies inside this very match, written by units next door: a run longer than a
                     * unit would make every unit wait for the one before it.  Byte-serial semantics make the match
                     * periodic, so read the period in front of the match instead (zxc_decompress.c:197-413 does the
                     * same with shuffle tables): position kk of the match equals position kk mod coff. */
                    rot = z2_mod((u32)kk, (u32)coff);
                    if (coff >= 16) {
                        n = min(n, (u32)coff - rot); /* a piece that would wrap ends at the period boundary */
                        s0 = cmd - coff + (i32)rot;
                    } else {
                        fold = out16; /* the register path below; the byte loop keeps the plain source (and its readiness) */
                    }
                }
                const i32 s1 = s0 + (i32)n - 1;
                slow = coff < 16 || (s0 < 0 && s1 >= 0);
                /* units this piece reads: byte-serial copies reach back `coff` bytes from pos */
                if (fold) {
                    need = true;
                    ua = (u32)(cmd - coff) >> 4;
                    ub = (u32)(cmd - 1) >> 4;
                } else if (s1 >= 0) {
                    need = true;
                    ua = (u32)max(s0, 0) >> 4;
                    ub = (u32)s1 >> 4;
                }
            }
        }
        const u32 ca = __shfl_sync(FULL, cnt, ua & 31u), cb = __shfl_sync(FULL, cnt, ub & 31u);
        const bool ready = !need || ((ua == u || ca > (ua >> 5)) && (ub == u || cb > (ub >> 5)));
        if (active && ready) {
            if (!slow) {
                const u8* sp = is_lit ? lit + ((i32)pos - cM) : (s0 >= 0 ? out + s0 : dict + ((i32)dict_size + s0));
                u32 w0, w1, w2, w3;
                uw_read16(sp - d, d, d + n, w0, w1, w2, w3);
                /* bytes below unit offset d keep what earlier pieces put there: a 128-bit mask from two 64-bit shifts */
                const u64 lo64 = d >= 8u ? ~0ull : ((1ull << (8u * d)) - 1ull);
                const u64 hi64 = d <= 8u ? 0ull : ((1ull << (8u * (d - 8u))) - 1ull);
                const u32 m0 = (u32)lo64, m1 = (u32)(lo64 >> 32), m2 = (u32)hi64, m3 = (u32)(hi64 >> 32);
                a0 = (a0 & m0) | (w0 & ~m0);
                a1 = (a1 & m1) | (w1 & ~m1);
                a2 = (a2 & m2) | (w2 & ~m2);
                a3 = (a3 & m3) | (w3 & ~m3);
            } else if ((fold || s0 >= 0) && out16) {
                /* distance below 16 inside the window: the piece repeats the `coff` bytes in front of it.  Take the
                 * 16 bytes that end at pos (previous unit : own registers) -- or, for a match that began before this
                 * unit, the 16 bytes that end at the match start, rotated to this position's phase -- keep the last
                 * `coff`, double the run until it covers a unit, and put it behind the bytes the unit already has. */
                const u32 u0 = u << 4;
                u32 h0, h1, h2, h3;
                if (fold) {
                    uw_read16(out + cmd - 16, 16u - (u32)coff, 16u, h0, h1, h2, h3);
                } else {
                    uint4 Pv = make_uint4(0, 0, 0, 0);
                    if (s0 < (i32)u0) Pv = *reinterpret_cast<const uint4*>(out + u0 - 16u);
                    uw_extract16(Pv, make_uint4(a0, a1, a2, a3), d, h0, h1, h2, h3);
                }
                /* pattern = bytes 16-coff .. 15 of h, moved to the bottom (a right shift by 16-coff bytes) */
                u32 x0, x1, x2, x3;
                uw_extract16(make_uint4(h0, h1, h2, h3), make_uint4(0, 0, 0, 0), 16u - (u32)coff, x0, x1, x2, x3);
                if (fold && rot) { /* rotate the period: byte i becomes byte (i + rot) mod coff */
                    u32 p0, p1, p2, p3;
                    uw_extract16(make_uint4(x0, x1, x2, x3), make_uint4(0, 0, 0, 0), rot, p0, p1, p2, p3);
                    uw_shl128(x0, x1, x2, x3, (u32)coff - rot);
                    const u32 c = (u32)coff;
                    const u64 l64 = c >= 8u ? ~0ull : ((1ull << (8u * c)) - 1ull);
                    const u64 g64 = c <= 8u ? 0ull : ((1ull << (8u * (c - 8u))) - 1ull);
                    x0 = (x0 | p0) & (u32)l64;
                    x1 = (x1 | p1) & (u32)(l64 >> 32);
                    x2 = (x2 | p2) & (u32)g64;
                    x3 = (x3 | p3) & (u32)(g64 >> 32);
                }
#pragma unroll 1
                for (u32 len = (u32)coff; len < 16u; len <<= 1) {
                    u32 y0 = x0, y1 = x1, y2 = x2, y3 = x3;
                    uw_shl128(y0, y1, y2, y3, len);
                    x0 |= y0;
                    x1 |= y1;
                    x2 |= y2;
                    x3 |= y3;
                }
                uw_shl128(x0, x1, x2, x3, d);
                const u64 lo64 = d >= 8u ? ~0ull : ((1ull << (8u * d)) - 1ull);
                const u64 hi64 = d <= 8u ? 0ull : ((1ull << (8u * (d - 8u))) - 1ull);
                a0 = (a0 & (u32)lo64) | x0;
                a1 = (a1 & (u32)(lo64 >> 32)) | x1;
                a2 = (a2 & (u32)hi64) | x2;
                a3 = (a3 & (u32)(hi64 >> 32)) | x3;
            } else {
                const u32 u0 = u << 4;
#pragma unroll 1
                for (u32 k = 0; k < n; k++) {
                    const i32 sb = s0 + (i32)k;
                    u32 v;
                    if (sb >= (i32)u0) v = uw_get_byte(a0, a1, a2, a3, (u32)sb - u0);
                    else if (sb >= 0) v = out[sb];
                    else v = dict[(i32)dict_size + sb];
                    uw_set_byte(a0, a1, a2, a3, d + k, v);
                }
            }
            pos += n;
            if (pos == uend) { /* unit finished: one store, then on to the lane's next unit */
                const u32 u0 = u << 4;
                if (uend - u0 == 16u && out16) {
                    *reinterpret_cast<uint4*>(out + u0) = make_uint4(a0, a1, a2, a3);
                } else {
                    for (u32 k = 0; k < uend - u0; k++) out[u0 + k] = (u8)uw_get_byte(a0, a1, a2, a3, k);
                }
                cnt++;
                u += 32u;
                active = u < n_units;
                if (active) {
                    pos = u << 4;
                    uend = min(pos + 16u, total);
                    idx = uidx[u];
                    UW_LOAD_SEQ();
                }
            }
        }
        __syncwarp(); /* finished units are visible to the lanes that fetch the counters next */
        if (!__any_sync(FULL, active)) break;
    }
#undef UW_LOAD_SEQ
    return (int)total;
}
