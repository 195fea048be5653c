// TEST INFRASTRUCTURE ONLY -- never linked into or called from libzxc.so.
//
// Sequential host replay of the block-cooperative decode kernel (zxc_b200/csrc/zxc_decode2.cuh)
// built on the SAME integer core the device code compiles (zxc_decode2_core.h): record packing,
// the word plan, the period fold, the extras segment maps.  It follows the kernel's two phases with
// the kernel's geometry (256 "threads" x 16 sequences, 512-byte groups, end-of-sequence bitmasks,
// literals staged behind the window) but runs them in order on one CPU thread, so it checks the
// arithmetic -- not the synchronisation.  tests/test_decode2_model.py diffs it against the reference.
//
// Scope mirrors the kernel: GLO / GHI blocks with raw literals and raw tokens
// (zxc_decompress.c:847-1209, :1231-1469).  Returns the decoded size, a negative zxc_error_t, or
// Z2M_DEFER for blocks the kernel leaves to the general kernel.
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../zxc_b200/csrc/zxc_decode2_core.h"

#define Z2M_DEFER (-1000)
#define ERR_OVERFLOW (-10)
#define ERR_BAD_OFFSET (-9)

static uint32_t ld32(const uint8_t* p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }

// round statistics of the last decoded block (dependency depth of the word gathers inside a group,
// sources in other groups counted as complete): sum over groups of rounds, and the number of groups
static uint64_t g_round_sum, g_group_cnt, g_round_hist[16];
extern "C" void z2_model_round_stats(uint64_t* sum, uint64_t* groups, uint64_t* hist16) {
    *sum = g_round_sum;
    *groups = g_group_cnt;
    memcpy(hist16, g_round_hist, sizeof g_round_hist);
}
extern "C" void z2_model_round_reset(void) {
    g_round_sum = g_group_cnt = 0;
    memset(g_round_hist, 0, sizeof g_round_hist);
}

// ---- "lane = 16-byte unit" iteration count (design exploration): every unit walks the regions that cut it, one
// piece per iteration; a match piece whose source lies in a unit of the same 512-byte group waits until that
// unit is finished.  Returns iterations per group summed over the block.
static uint64_t g_u16_iters, g_u16_groups, g_u16_pieces;
extern "C" void z2_model_u16_stats(uint64_t* it, uint64_t* gr, uint64_t* pc) { *it = g_u16_iters; *gr = g_u16_groups; *pc = g_u16_pieces; }
extern "C" void z2_model_u16_reset(void) { g_u16_iters = g_u16_groups = g_u16_pieces = 0; }
static void u16_count(const std::vector<z2_rec_t>& rec, uint32_t n_seq, uint32_t total) {
    // region list: boundaries (start, is_match, md, off) per sequence
    uint32_t idx = 0;
    for (uint32_t g0 = 0; g0 < total; g0 += 512) {
        struct P { uint32_t n; int32_t s0, s1; };
        std::vector<std::vector<P>> pieces(32);
        for (uint32_t l = 0; l < 32; l++) {
            const uint32_t u0 = g0 + 16 * l;
            if (u0 >= total) continue;
            const uint32_t uend = u0 + 16 < total ? u0 + 16 : total;
            uint32_t pos = u0;
            while (pos < uend) {
                z2_seq_t c = z2_unpack(rec[idx]);
                if (idx >= n_seq) { c.md = Z2_MD_INF; c.E = (int32_t)total; }
                if ((int32_t)pos >= c.E) { idx++; continue; }
                if ((int32_t)pos < c.md) {
                    const uint32_t e = (uint32_t)c.md < uend ? (uint32_t)c.md : uend;
                    pieces[l].push_back({e - pos, -1, -1});
                    pos = e;
                } else {
                    const uint32_t e = (uint32_t)c.E < uend ? (uint32_t)c.E : uend;
                    const int32_t k = (int32_t)pos - c.md;
                    int32_t s0 = k < c.off ? (int32_t)pos - c.off : c.md - c.off;
                    int32_t s1 = s0 + (int32_t)(e - pos) - 1;
                    if (s1 >= c.md) s1 = c.md - 1;
                    pieces[l].push_back({e - pos, s0, s1});
                    pos = e;
                }
            }
        }
        // iterations with blocking
        uint32_t ip[32] = {0}, done = 0, iters = 0, active = 0;
        for (uint32_t l = 0; l < 32; l++) { if (pieces[l].empty()) done |= 1u << l; else active++; g_u16_pieces += pieces[l].size(); }
        while (done != 0xFFFFFFFFu) {
            uint32_t newly = 0;
            for (uint32_t l = 0; l < 32; l++) {
                if ((done >> l) & 1u) continue;
                const P& q = pieces[l][ip[l]];
                bool ok = true;
                if (q.s0 >= 0 || q.s1 >= 0) {
                    for (int32_t sb = q.s0 < 0 ? 0 : q.s0; sb <= q.s1; sb += 16) {
                        const int32_t un = sb >> 4, my = (int32_t)((g0 >> 4) + l);
                        if (un >= (int32_t)(g0 >> 4) && un < my && !((done >> (un - (int32_t)(g0 >> 4))) & 1u)) ok = false;
                    }
                    const int32_t un = q.s1 >> 4, my = (int32_t)((g0 >> 4) + l);
                    if (q.s1 >= 0 && un >= (int32_t)(g0 >> 4) && un < my && !((done >> (un - (int32_t)(g0 >> 4))) & 1u)) ok = false;
                }
                if (ok && ++ip[l] == pieces[l].size()) newly |= 1u << l;
            }
            done |= newly;
            iters++;
            if (iters > 4096) break;
        }
        g_u16_iters += iters;
        g_u16_groups++;
    }
}

// ---- "lane = contiguous stripe" iteration count (design exploration): lane l decodes bytes [l*S, (l+1)*S) of
// the block piece by piece (a piece = region cut by 16-byte units); finished 16-byte units are what other lanes may
// read.  One iteration = every lane whose next piece is ready processes it.
static uint64_t g_st_iters, g_st_pieces, g_st_blocks;
extern "C" void z2_model_stripe_stats(uint64_t* it, uint64_t* pc, uint64_t* bl) { *it = g_st_iters; *pc = g_st_pieces; *bl = g_st_blocks; }
extern "C" void z2_model_stripe_reset(void) { g_st_iters = g_st_pieces = g_st_blocks = 0; }
static uint32_t g_stripe_lanes = 32;
extern "C" void z2_model_stripe_lanes(uint32_t n) { g_stripe_lanes = n; }
static void stripe_count(const std::vector<z2_rec_t>& rec, uint32_t n_seq, uint32_t total) {
    const uint32_t NL = g_stripe_lanes;
    const uint32_t S = ((total + NL - 1) / NL + 15) & ~15u;
    struct P { uint32_t end; int32_t s0, s1; };
    std::vector<std::vector<P>> pieces(NL);
    uint32_t idx = 0;
    uint64_t np = 0;
    for (uint32_t l = 0; l < NL; l++) {
        const uint32_t b0 = l * S, b1 = b0 + S < total ? b0 + S : total;
        for (uint32_t u0 = b0; u0 < b1; u0 += 16) {
            const uint32_t uend = u0 + 16 < b1 ? u0 + 16 : b1;
            uint32_t pos = u0;
            while (pos < uend) {
                z2_seq_t c = z2_unpack(rec[idx]);
                if (idx >= n_seq) { c.md = Z2_MD_INF; c.E = (int32_t)total; }
                if ((int32_t)pos >= c.E) { idx++; continue; }
                if ((int32_t)pos < c.md) {
                    const uint32_t e = (uint32_t)c.md < uend ? (uint32_t)c.md : uend;
                    pieces[l].push_back({e, -1, -1});
                    pos = e;
                } else {
                    const uint32_t e = (uint32_t)c.E < uend ? (uint32_t)c.E : uend;
                    const int32_t s0 = (int32_t)pos - c.off;
                    pieces[l].push_back({e, s0, s0 + (int32_t)(e - pos) - 1});
                    pos = e;
                }
                np++;
            }
        }
    }
    std::vector<uint32_t> ip(NL, 0), stored(NL, 0); // stored[l] = bytes of stripe l that are in finished units
    uint64_t iters = 0;
    for (;;) {
        bool any = false;
        std::vector<uint32_t> nstored = stored;
        for (uint32_t l = 0; l < NL; l++) {
            if (ip[l] >= pieces[l].size()) continue;
            any = true;
            const P& q = pieces[l][ip[l]];
            bool ok = true;
            if (q.s1 >= 0) {
                const int32_t a = q.s0 < 0 ? 0 : q.s0;
                const uint32_t la = (uint32_t)a / S, lb = (uint32_t)q.s1 / S;
                // own current unit counts as available (kept in registers); everything else must be stored
                const uint32_t my_u0 = (q.end - 1) & ~15u;
                if (!((uint32_t)q.s1 < la * S + stored[la] || ((uint32_t)a >= my_u0))) {
                    if (!((uint32_t)a < la * S + stored[la] && (uint32_t)q.s1 >= my_u0)) ok = (uint32_t)q.s1 < lb * S + stored[lb] && (uint32_t)a < la * S + stored[la];
                }
                if (la != lb && !((uint32_t)a < la * S + stored[la])) ok = false;
            }
            if (ok) {
                ip[l]++;
                const uint32_t done_to = q.end - l * S;
                if ((q.end & 15u) == 0 || ip[l] == pieces[l].size()) nstored[l] = done_to;
            }
        }
        if (!any) break;
        stored = nstored;
        iters++;
        if (iters > 1000000) break;
    }
    g_st_iters += iters;
    g_st_pieces += np;
    g_st_blocks++;
}

// ---- "lane owns units l, l+32, ... and runs ahead" (design exploration): a lane starts its next unit as soon as
// it finishes one; a match piece waits until the units it reads are finished (own unit: registers).
static uint64_t g_ra_iters, g_ra_pieces, g_ra_blocks;
extern "C" void z2_model_runahead_stats(uint64_t* it, uint64_t* pc, uint64_t* bl) { *it = g_ra_iters; *pc = g_ra_pieces; *bl = g_ra_blocks; }
extern "C" void z2_model_runahead_reset(void) { g_ra_iters = g_ra_pieces = g_ra_blocks = 0; }
static void runahead_count(const std::vector<z2_rec_t>& rec, uint32_t n_seq, uint32_t total) {
    const uint32_t n_units = (total + 15) / 16;
    struct P { uint32_t unit; bool last; int32_t s0, s1; };
    std::vector<std::vector<P>> pieces(32);
    uint32_t idx = 0;
    uint64_t np = 0;
    for (uint32_t u = 0; u < n_units; u++) {
        const uint32_t u0 = 16 * u, uend = u0 + 16 < total ? u0 + 16 : total;
        uint32_t pos = u0;
        while (pos < uend) {
            z2_seq_t c = z2_unpack(rec[idx]);
            if (idx >= n_seq) { c.md = Z2_MD_INF; c.E = (int32_t)total; }
            if ((int32_t)pos >= c.E) { idx++; continue; }
            uint32_t e;
            if ((int32_t)pos < c.md) {
                e = (uint32_t)c.md < uend ? (uint32_t)c.md : uend;
                pieces[u & 31].push_back({u, e == uend, -1, -1});
            } else {
                e = (uint32_t)c.E < uend ? (uint32_t)c.E : uend;
                const int32_t s0 = (int32_t)pos - c.off;
                pieces[u & 31].push_back({u, e == uend, s0, s0 + (int32_t)(e - pos) - 1});
            }
            pos = e;
            np++;
        }
    }
    std::vector<uint8_t> udone(n_units, 0);
    uint32_t ip[32] = {0};
    uint64_t iters = 0;
    for (;;) {
        bool any = false;
        std::vector<uint32_t> fin;
        for (uint32_t l = 0; l < 32; l++) {
            if (ip[l] >= pieces[l].size()) continue;
            any = true;
            const P& q = pieces[l][ip[l]];
            bool ok = true;
            if (q.s1 >= 0) {
                for (int32_t un = (q.s0 < 0 ? 0 : q.s0) >> 4; un <= (q.s1 >> 4); un++)
                    if ((uint32_t)un != q.unit && !udone[un]) ok = false;
            }
            if (ok) {
                ip[l]++;
                if (q.last) fin.push_back(q.unit);
            }
        }
        if (!any) break;
        for (uint32_t u : fin) udone[u] = 1;
        iters++;
        if (iters > 1000000) break;
    }
    g_ra_iters += iters;
    g_ra_pieces += np;
    g_ra_blocks++;
}

// ---- timing sketch of phase 2 (design exploration, not a test): W warps claim `claim` consecutive groups
// at a time in order; a step costs `fixed` cycles before its first round, a round costs `round` cycles, a word
// of another warp becomes visible `vis` cycles after the round that wrote it; `level` 0 = word-level
// completion tracking, 1 = group-level (a source counts as complete when its whole group is)
static struct { uint32_t W, claim, fixed, round, vis, level, guard; } g_sim = {8, 1, 300, 120, 200, 0, 8};
static uint64_t g_sim_cycles, g_sim_blocks;
extern "C" void z2_model_sim_config(uint32_t W, uint32_t claim, uint32_t fixed, uint32_t round, uint32_t vis,
                                    uint32_t level, uint32_t guard) {
    g_sim = {W, claim, fixed, round, vis, level, guard};
    g_sim_cycles = g_sim_blocks = 0;
}
extern "C" double z2_model_sim_result(void) { return g_sim_blocks ? (double)g_sim_cycles / (double)g_sim_blocks : 0.0; }

static uint64_t simulate(const std::vector<std::vector<uint32_t>>& deps, uint32_t n_words) {
    const uint32_t n_groups = (n_words + 127) / 128;
    std::vector<uint64_t> T(n_words, ~0ull), TG(n_groups, ~0ull), TR((n_words + 31) / 32, ~0ull), TP((n_words + 31) / 32, ~0ull);
    std::vector<uint32_t> owner(n_groups, 0);
    std::vector<uint64_t> freeat(g_sim.W, 0);
    uint32_t next = 0;
    uint64_t makespan = 0;
    // claims are handed out in order to whichever warp is free first
    while (next < n_groups) {
        uint32_t w = 0;
        for (uint32_t k = 1; k < g_sim.W; k++) if (freeat[k] < freeat[w]) w = k;
        uint64_t t = freeat[w];
        for (uint32_t c = 0; c < g_sim.claim && next < n_groups; c++, next++) {
            const uint32_t g = next;
            owner[g] = w;
            if (g >= g_sim.guard && TG[g - g_sim.guard] != ~0ull && TG[g - g_sim.guard] + g_sim.vis > t) t = TG[g - g_sim.guard] + g_sim.vis;
            t += g_sim.fixed;
            const uint32_t w0 = g * 128, w1 = w0 + 128 < n_words ? w0 + 128 : n_words;
            uint32_t pending = w1 - w0;
            std::vector<uint8_t> done(w1 - w0, 0);
            while (pending) {
                uint32_t did = 0;
                uint64_t soonest = ~0ull;
                for (uint32_t x = w0; x < w1; x++) {
                    if (done[x - w0]) continue;
                    uint64_t need = 0;
                    for (uint32_t d : deps[x]) {
                        const uint32_t dg = d / 128;
                        uint64_t td = g_sim.level == 1 && dg != g ? TG[dg] : g_sim.level == 2 && dg != g ? TR[d / 32] : g_sim.level == 3 && dg != g ? TP[d / 32] : T[d];
                        if (td != ~0ull && owner[dg] != w) td += g_sim.vis;
                        if (td > need) need = td;
                    }
                    if (need <= t) {
                        done[x - w0] = 2;
                        did++;
                    } else if (need < soonest) soonest = need;
                }
                if (did) {
                    t += g_sim.round;
                    for (uint32_t x = w0; x < w1; x++)
                        if (done[x - w0] == 2) { done[x - w0] = 1; T[x] = t; }
                    pending -= did;
                    for (uint32_t r = w0 / 32; r * 32 < w1; r++) { /* rows that just completed */
                        if (TR[r] != ~0ull) continue;
                        bool all = true;
                        for (uint32_t x = r * 32; x < r * 32 + 32 && x < w1; x++) all = all && done[x - w0] == 1;
                        if (all) TR[r] = t;
                    }
                    /* contiguous-prefix completion times (level 3): row r counts once every row <= r is complete */
                    for (uint32_t r = 0; r < TR.size(); r++) {
                        if (TR[r] == ~0ull) break;
                        if (TP[r] == ~0ull) TP[r] = r ? (TP[r - 1] > t ? TP[r - 1] : t) : t;
                    }
                } else {
                    if (soonest == ~0ull) return ~0ull; // cannot happen: sources are earlier words
                    t = soonest;
                }
            }
            TG[g] = t;
        }
        freeat[w] = t;
        if (t > makespan) makespan = t;
    }
    return makespan;
}

extern "C" int z2_model_decode_block(const uint8_t* blk, uint32_t src_len, uint8_t* out, uint32_t cap,
                                     const uint8_t* dict, uint32_t dict_size, uint32_t win_bytes, uint32_t gap,
                                     uint32_t threads, uint32_t* slow_words) {
    if (src_len < 20) return Z2M_DEFER;
    const uint32_t type = blk[0], comp = ld32(blk + 3);
    if ((uint64_t)src_len < 8ull + comp) return Z2M_DEFER;
    if (type != 1 && type != 2) return Z2M_DEFER;
    if (comp < 12 || cap > win_bytes) return Z2M_DEFER;
    const bool ghi = type == 2;
    const uint8_t* pay = blk + 8;
    const uint32_t n_seq = ld32(pay), n_lit = ld32(pay + 4), enc_lit = pay[8], enc_tok = pay[9], enc_off = pay[11];
    if (enc_lit || enc_tok || (!ghi && enc_off > 1)) return Z2M_DEFER;
    const uint32_t avail = comp - 12;
    const uint64_t seq_bytes = ghi ? (uint64_t)n_seq * 4 : (uint64_t)n_seq * (enc_off ? 2 : 3);
    const uint64_t consumed = (uint64_t)n_lit + seq_bytes;
    if (consumed > avail || avail - n_lit < 32) return Z2M_DEFER;
    if (n_lit > cap || n_seq > 0xFFF0u) return Z2M_DEFER;
    const uint32_t s_bytes = avail - n_lit, ext_len = avail - (uint32_t)consumed;
    const uint32_t lba = (win_bytes + gap - n_lit + 15u) & ~15u;
    if ((uint64_t)s_bytes + 48 + 4ull * ext_len + 16 > lba) return Z2M_DEFER;

    // window with the literal stream staged behind it (the kernel's two bulk copies)
    std::vector<uint8_t> win(win_bytes + gap + 128 + 64, 0xEE);
    const uint8_t* g_lit = pay + 12;
    const uint8_t* S = g_lit + n_lit;
    const int32_t lit_pos = (int32_t)lba + 5; // any shift 0..15 (global alignment of the literal section)
    memcpy(win.data() + lit_pos, g_lit, n_lit);
    const uint8_t* S_off = ghi ? S : S + n_seq;
    const uint8_t* S_ext = ghi ? S + 4u * n_seq : S_off + (enc_off ? n_seq : 2u * n_seq);
    const uint32_t esc = ghi ? 255u : 15u;

    // phase 1a: extras by segment maps
    std::vector<uint32_t> vals(ext_len + 4, 0);
    uint32_t n_val = 0;
    if (ext_len) {
        const uint32_t T = threads;
        const uint32_t seg = ext_len / T + 1 > 4 ? (ext_len + T - 1) / T : 4;
        const uint32_t nseg = (ext_len + seg - 1) / seg;
        std::vector<uint64_t> excl(nseg);
        uint64_t run = Z2_MAP_ID;
        for (uint32_t t = 0; t < nseg; t++) {
            const uint32_t lo = t * seg, hi = lo + seg < ext_len ? lo + seg : ext_len;
            excl[t] = run;
            run = z2_map_compose(run, z2_seg_map(S_ext, lo, hi, ext_len));
        }
        n_val = z2_map_cnt(run, 0);
        for (uint32_t t = 0; t < nseg; t++) {
            const uint32_t lo = t * seg, hi = lo + seg < ext_len ? lo + seg : ext_len;
            const uint32_t ent = z2_map_exit(excl[t], 0);
            if (ent != 3u) z2_seg_values(S_ext, lo, hi, ext_len, ent, z2_map_cnt(excl[t], 0), vals.data());
        }
    }
    // phase 1b: records
    std::vector<z2_rec_t> rec(n_seq + 4);
    const uint32_t n_groups_max = (cap + Z2_GROUP - 1) / Z2_GROUP;
    std::vector<uint16_t> gidx(n_groups_max + 2, (uint16_t)n_seq);
    uint32_t ord = 0, L = 0, O = 0, err = 0xFFFFFFFFu;
    for (uint32_t i = 0; i < n_seq; i++) {
        uint32_t ll, ml, off;
        if (!ghi) {
            ll = S[i] >> 4;
            ml = S[i] & 15u;
            off = enc_off ? S_off[i] : (uint32_t)S_off[2 * i] | ((uint32_t)S_off[2 * i + 1] << 8);
        } else {
            ll = S[4 * i + 3];
            ml = S[4 * i + 2];
            off = (uint32_t)S[4 * i] | ((uint32_t)S[4 * i + 1] << 8);
        }
        off += 1;
        if (ll == esc) { ll += ord < n_val ? vals[ord] : 0u; ord++; }
        if (ml == esc) { ml += ord < n_val ? vals[ord] : 0u; ord++; }
        ml += 5;
        if (ll > 0xFFFF) ll = 0xFFFF;
        if (ml > 0xFFFF) ml = 0xFFFF;
        const uint32_t md = O + ll, E = md + ml;
        const bool ovf = (L + ll > n_lit) || (E > cap);
        const bool bad = md + dict_size < off;
        if (ovf || bad) {
            const uint32_t key = (i << 1) | (ovf ? 0u : 1u);
            if (key < err) err = key;
        }
        rec[i] = z2_pack(E, md & 0xFFFFu, off, (O - L) & 0xFFFFu);
        const uint32_t Ec = E < cap ? E : cap;
        for (uint32_t g = (O + Z2_GROUP - 1) / Z2_GROUP; g < (Ec + Z2_GROUP - 1) / Z2_GROUP; g++) gidx[g] = (uint16_t)i;
        O = E;
        L += ll;
        if (O > (1u << 30)) O = 1u << 30;
        if (L > (1u << 30)) L = 1u << 30;
    }
    if (err != 0xFFFFFFFFu) return (err & 1u) ? ERR_BAD_OFFSET : ERR_OVERFLOW;
    const uint32_t rem = n_lit - L;
    if (rem > cap - O) return ERR_OVERFLOW;
    const uint32_t total = O + rem;
    if (total == 0) return 0;
    rec[n_seq] = z2_pack(total, total & 0xFFFFu, 1u, (O - L) & 0xFFFFu);
    rec[n_seq + 1] = z2_pack(0x10000u, total & 0xFFFFu, 1u, 0u);
    rec[n_seq + 2] = rec[n_seq + 1];

    // phase 2: words, group by group
    const uint32_t n_groups = (total + Z2_GROUP - 1) / Z2_GROUP;
    uint32_t n_slow = 0;
    std::vector<std::vector<uint32_t>> deps((total + 3) / 4);
    for (uint32_t g = 0; g < n_groups; g++) {
        const int32_t p0 = (int32_t)(g * Z2_GROUP);
        const uint32_t i0 = gidx[g];
        uint32_t M[4] = {0, 0, 0, 0};
        for (uint32_t jj = 0; jj < 4; jj++) {
            bool last_in = false;
            for (uint32_t lane = 0; lane < 32; lane++) {
                const uint32_t k = i0 + lane + 32 * jj;
                bool in = false;
                uint32_t cw = 0;
                if (k < n_seq) {
                    const int32_t rel = (int32_t)(rec[k].w0 & 0xFFFFu) + 1 - p0;
                    in = rel < (int32_t)Z2_GROUP;
                    cw = (uint32_t)(rel + 3) >> 2;
                }
                if (in && cw < 128) M[cw >> 5] |= 1u << (cw & 31);
                if (lane == 31) last_in = in;
            }
            if (!last_in) break;
        }
        uint32_t wround[128];
        uint32_t gmax = 0;
        for (uint32_t r = 0; r < 4; r++) {
            uint32_t pre = 0;
            for (uint32_t q = 0; q < r; q++) pre += (uint32_t)__builtin_popcount(M[q]);
            for (uint32_t lane = 0; lane < 32; lane++) {
                const int32_t p = p0 + (int32_t)(4 * (lane + 32 * r));
                if (p >= (int32_t)total) continue;
                const uint32_t le = (2u << lane) - 1u;
                const uint32_t idx = i0 + pre + (uint32_t)__builtin_popcount(M[r] & le);
                z2_seq_t c = z2_unpack(rec[idx]), n = z2_unpack(rec[idx + 1]);
                if (idx >= n_seq) c.md = Z2_MD_INF;
                if (idx + 1 >= n_seq) n.md = (int32_t)total;
                const z2_plan_t pl = z2_word_plan(p, c, n, lit_pos);
                uint8_t b4[4] = {0, 0, 0, 0};
                {   // dependency round of this word
                    uint32_t rd = 1;
                    auto dep = [&](int32_t sb) {
                        if (sb >= 0 && sb < p) deps[(uint32_t)p >> 2].push_back((uint32_t)sb >> 2);
                        if (sb >= p0 && sb < p) {
                            const uint32_t w = (uint32_t)(sb - p0) >> 2;
                            if (wround[w] + 1 > rd) rd = wround[w] + 1;
                        }
                    };
                    if (!(pl.flags & Z2_SLOW)) {
                        if (pl.flags & 1u) { dep(pl.srcX); dep(pl.srcX + 3); }
                        if (pl.flags & 2u) { dep(pl.srcY); dep(pl.srcY + 3); }
                        if (pl.flags & 4u) { dep(pl.srcZ); dep(pl.srcZ + 3); }
                    } else {
                        for (int32_t b = 0; b < 4 && p + b < (int32_t)total; b++) {
                            int im;
                            const int32_t sb = z2_byte_source(p + b, c, n, lit_pos, &im);
                            if (im) dep(sb);
                        }
                    }
                    wround[lane + 32 * r] = rd;
                    if (rd > gmax) gmax = rd;
                }
                if (!(pl.flags & Z2_SLOW)) {
                    for (uint32_t b = 0; b < 4; b++) {
                        const int32_t s = (b < pl.t ? pl.srcX : b < pl.t2 ? pl.srcY : pl.srcZ) + (int32_t)b;
                        if (s < 0 || s >= (int32_t)win.size()) return -2000; // model bug guard
                        if (b < pl.t ? (pl.flags & 1u) : b < pl.t2 ? (pl.flags & 2u) : (pl.flags & 4u)) {
                            if (s >= p) return -2001; // a fast gather must read completed words only
                        }
                        b4[b] = win[s];
                    }
                } else {
                    n_slow++;
                    for (int32_t b = 0; b < 4; b++) {
                        const int32_t q = p + b;
                        if (q >= (int32_t)total) break;
                        int is_match;
                        const int32_t s = z2_byte_source(q, c, n, lit_pos, &is_match);
                        if (is_match && s < 0) {
                            if (-s > (int32_t)dict_size) return -2002;
                            b4[b] = dict[(int32_t)dict_size + s];
                        } else if (is_match && s >= p) {
                            if (s >= q) return -2003;
                            b4[b] = b4[s - p];
                        } else {
                            b4[b] = win[s];
                        }
                    }
                }
                memcpy(win.data() + p, b4, 4);
            }
        }
        g_round_sum += gmax;
        g_group_cnt++;
        g_round_hist[gmax < 15 ? gmax : 15]++;
    }
    memcpy(out, win.data(), total);
    u16_count(rec, n_seq, total);
    runahead_count(rec, n_seq, total);
    if (g_sim.W) {
        const uint64_t c = simulate(deps, (total + 3) / 4);
        if (c != ~0ull) { g_sim_cycles += c; g_sim_blocks++; }
    }
    if (slow_words) *slow_words += n_slow;
    return (int)total;
}
