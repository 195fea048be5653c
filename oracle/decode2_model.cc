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
    }
    memcpy(out, win.data(), total);
    if (slow_words) *slow_words += n_slow;
    return (int)total;
}
