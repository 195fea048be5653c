/*
 * zxc_hufenc.h -- code-length construction for the level 6-7 literal / token Huffman sections:
 * length-limited package-merge, the flat/length "nudge", and exact PivCo section sizing.
 *
 * Pure integer functions of (histogram, cap) -> 256 code lengths.  They run ON THE DEVICE, one lane
 * per block, inside the encode kernel; the same source also compiles as plain C so that
 * tests/test_hufenc.py can check it on the CPU against the reference's own internals
 * (oracle/_ref internals shim).  No libc, no recursion, all scratch in a caller-provided work area.
 *
 * Behaviour restated (file:line in /root/reference/src/lib/zxc_huffman.c):
 *   leaf order (weight, symbol)            :100-143   zxh_sort_leaves
 *   boundary package-merge, capped depth   :172-311   zxh_build_code_lengths
 *   nudge: classes / prefix masses / cost  :343-430   zxh_classes, zxh_prefix_masses, zxh_eval
 *   nudge: greedy walk over level counts   :432-590   zxh_clamp, zxh_complete, zxh_walk
 *   nudge: grouped DP over level counts    :592-775   zxh_run_cost, zxh_dp_solve
 *   nudge: candidate selection             :803-945   zxh_nudge_code_lengths
 *   PivCo run sizes                        :1172-1249 zxh_calc_size (tree geometry as in zxc_huffman.cuh)
 *   literal cost estimate of the optimal parser  src/lib/zxc_compress.c:720-749  zxh_estimate_lit_bits
 */
#ifndef ZXC_B200_HUFENC_H
#define ZXC_B200_HUFENC_H

#include <stdint.h>

#ifdef __CUDACC__
#define ZXH __host__ __device__ static inline
#else
#define ZXH static inline
#endif

#define ZXH_NSYM 256
#define ZXH_LU 11          /* longest code (ULTRA) */
#define ZXH_LAMBDA_Q8 26u  /* ZXC_HUF_NUDGE_LAMBDA_Q8 */
#define ZXH_BITS_PERMIL 1015u
#define ZXH_MERGE_Q8 230u
#define ZXH_FLAT_SIMD_MAX 6
#define ZXH_DEEP_FLAT_PENALTY 24
#define ZXH_LEVEL_COST 64u
#define ZXH_U64MAX 0xFFFFFFFFFFFFFFFFull
#define ZXH_DP_M 64        /* most groups the nudge DP ever sees: ceil(256 / 4) */

typedef struct { uint32_t w; int16_t sym; } zxh_leaf_t;

typedef struct {
    uint32_t pm_weight[2][2 * ZXH_NSYM];        /* package-merge: item weights of the previous / current level */
    uint8_t pm_is_leaf[ZXH_LU][2 * ZXH_NSYM];   /* package-merge: 1 = leaf, 0 = package, per level and item */
    zxh_leaf_t leaves[ZXH_NSYM];
    zxh_leaf_t sort_tmp[ZXH_NSYM];
    uint64_t dp_a[(ZXH_DP_M + 1) * (ZXH_DP_M + 1)];
    uint64_t dp_b[(ZXH_DP_M + 1) * (ZXH_DP_M + 1)];
    uint16_t arrive[(ZXH_LU + 2) * (ZXH_DP_M + 1) * (ZXH_DP_M + 1)];
    uint64_t pf[ZXH_NSYM + 1];
    uint64_t pf_rank[ZXH_NSYM + 1];
    uint64_t pfg[ZXH_NSYM + 1];
    uint32_t val[ZXH_NSYM];
    int16_t sym_order[ZXH_NSYM];
    uint8_t cand[4][ZXH_NSYM];
    uint32_t node_count[2 * ZXH_NSYM];
    uint32_t cblc[ZXH_LU + 1]; /* grouped-DP result handed from the warp-wide solve to zxh_nudge_end */
} zxh_work_t;

ZXH unsigned zxh_log2(uint32_t v) { /* floor(log2(v)), v > 0 */
    unsigned r = 0;
    while (v >>= 1) r++;
    return r;
}

/* ascending (weight, symbol); the keys are unique so any correct sort gives the reference's order */
ZXH void zxh_sort_leaves(zxh_leaf_t* a, int n) {
    for (int gap = n / 2; gap > 0; gap /= 2) {
        for (int i = gap; i < n; i++) {
            const zxh_leaf_t key = a[i];
            int j = i;
            while (j >= gap && (a[j - gap].w > key.w || (a[j - gap].w == key.w && a[j - gap].sym > key.sym))) {
                a[j] = a[j - gap];
                j -= gap;
            }
            a[j] = key;
        }
    }
}

/* Length-limited Huffman code lengths by package-merge (same result as zxc_huf_build_code_lengths,
 * zxc_huffman.c:172-311, for every histogram and cap -- pinned by tests/test_hufenc.py), expressed without item
 * trees:
 *
 *   Level 0 is the sorted leaf list.  Level k is the merge, by weight, of the leaf list with the packages of level
 *   k-1 (item pairs 2p, 2p+1; a leaf goes first on equal weight).  Both inputs appear in the merge in their own
 *   order, so whatever prefix of a level is "taken" consists of a prefix of the leaves and a prefix of the packages,
 *   and taking p packages means taking the first 2p items of the level below.  The selection is therefore one
 *   number per level: m_top = min(2n-2, |top level|), then a_k = leaves among the first m_k items, m_(k-1) =
 *   2 (m_k - a_k); and a symbol's code length is the number of levels whose a_k covers its rank.
 *
 * Forward pass: item weights of the previous level only (two rolling rows) and one leaf/package flag per item.
 * Backward pass: L prefix counts.  No links, no traversal stack. */
ZXH int zxh_build_code_lengths(const uint32_t* freq, uint8_t* code_len, int max_code_len, zxh_work_t* W) {
    for (int i = 0; i < ZXH_NSYM; i++) code_len[i] = 0;
    zxh_leaf_t* leaves = W->leaves;
    int n = 0;
    for (int i = 0; i < ZXH_NSYM; i++) {
        if (freq[i]) {
            leaves[n].w = freq[i];
            leaves[n].sym = (int16_t)i;
            n++;
        }
    }
    if (n == 0) return -1;
    if (n == 1) {
        code_len[leaves[0].sym] = 1;
        return 0;
    }
    zxh_sort_leaves(leaves, n);
    int size[ZXH_LU]; /* items per level */
    uint32_t* prev = W->pm_weight[0];
    uint32_t* cur = W->pm_weight[1];
    for (int i = 0; i < n; i++) {
        prev[i] = leaves[i].w;
        W->pm_is_leaf[0][i] = 1;
    }
    size[0] = n;
    for (int k = 1; k < max_code_len; k++) {
        const int n_pack = size[k - 1] >> 1;
        uint8_t* flag = W->pm_is_leaf[k];
        int leaf = 0, pack = 0, out = 0;
        while (leaf < n && pack < n_pack) {
            const uint32_t pw = prev[2 * pack] + prev[2 * pack + 1];
            if (leaves[leaf].w <= pw) { /* equal weight: the leaf first */
                cur[out] = leaves[leaf++].w;
                flag[out++] = 1;
            } else {
                cur[out] = pw;
                flag[out++] = 0;
                pack++;
            }
        }
        for (; leaf < n; leaf++) {
            cur[out] = leaves[leaf].w;
            flag[out++] = 1;
        }
        for (; pack < n_pack; pack++) {
            cur[out] = prev[2 * pack] + prev[2 * pack + 1];
            flag[out++] = 0;
        }
        size[k] = out;
        uint32_t* t = prev;
        prev = cur;
        cur = t;
    }
    int m = 2 * n - 2;
    if (m > size[max_code_len - 1]) m = size[max_code_len - 1];
    for (int k = max_code_len - 1; k >= 0 && m > 0; k--) {
        const uint8_t* flag = W->pm_is_leaf[k];
        int a = 0;
        for (int i = 0; i < m; i++) a += flag[i];
        for (int j = 0; j < a; j++) code_len[leaves[j].sym]++;
        m = 2 * (m - a);
    }
    return 0;
}

/* ---- nudge ---------------------------------------------------------------------------------- */
typedef struct { uint64_t bits, touches; } zxh_cost_t;

ZXH int zxh_classes(const uint8_t* code_len, uint32_t* blc) {
    int n = 0;
    for (int l = 0; l <= ZXH_LU; l++) blc[l] = 0;
    for (int s = 0; s < ZXH_NSYM; s++) {
        if (code_len[s]) {
            blc[code_len[s]]++;
            n++;
        }
    }
    return n;
}

/* prefix masses of the frequencies laid out in canonical (length, symbol) order */
ZXH void zxh_prefix_masses(const uint8_t* code_len, const uint32_t* freq, const uint32_t* blc, uint64_t* pf, uint32_t* val) {
    uint32_t pos[ZXH_LU + 1];
    uint32_t acc = 0;
    for (int l = 1; l <= ZXH_LU; l++) {
        pos[l] = acc;
        acc += blc[l];
    }
    for (int s = 0; s < ZXH_NSYM; s++)
        if (code_len[s]) val[pos[code_len[s]]++] = freq[s];
    pf[0] = 0;
    for (uint32_t i = 0; i < acc; i++) pf[i + 1] = pf[i] + val[i];
}

/* decode-work weight of one maximal aligned sub-tree of D levels whose leaves have length lr */
ZXH int zxh_touch_weight(int lr, int D) {
    if (D == 0) return lr + 1;
    if (D == 1) return lr;
    return (lr - D) + 1 + (D > ZXH_FLAT_SIMD_MAX ? ZXH_DEEP_FLAT_PENALTY : 0);
}

/* bits and decoder "touches" of a code given its per-length leaf counts and prefix masses */
ZXH void zxh_eval(const uint32_t* blc, const uint64_t* pf, zxh_cost_t* out) {
    uint64_t bits = 0, touches = 0;
    int max_len = 0;
    uint32_t S = 0, base = 0; /* code-space cursor in 2^-LU slots; first leaf of the current length */
    for (int l = 1; l <= ZXH_LU; l++) {
        if (!blc[l]) continue;
        max_len = l;
        bits += (uint64_t)l * (pf[base + blc[l]] - pf[base]);
        const uint32_t w = 1u << (ZXH_LU - l);
        const uint32_t end = S + blc[l] * w;
        uint32_t x = S;
        while (x < end) { /* split [S, end) into maximal aligned power-of-two spans */
            const uint32_t wx = x ? (x & (0u - x)) : (1u << ZXH_LU);
            const uint32_t wr = 1u << zxh_log2(end - x);
            const uint32_t Wd = wx < wr ? wx : wr;
            const int D = (int)zxh_log2(Wd) - (ZXH_LU - l);
            const uint32_t i0 = base + ((x - S) >> (ZXH_LU - l));
            const uint64_t mass = pf[i0 + (Wd >> (ZXH_LU - l))] - pf[i0];
            touches += mass * (uint64_t)zxh_touch_weight(l, D);
            x += Wd;
        }
        S = end;
        base += blc[l];
    }
    touches += (uint64_t)ZXH_LEVEL_COST * (uint64_t)(max_len + 1);
    out->bits = bits;
    out->touches = touches;
}

ZXH uint64_t zxh_j(const zxh_cost_t* c) { return 256u * c->bits + (uint64_t)ZXH_LAMBDA_Q8 * c->touches; }

/* how many of the n_rem remaining leaves may sit at level l when s slots are open there */
ZXH uint32_t zxh_clamp(uint32_t want, uint32_t s, uint32_t n_rem, int l, int cap) {
    if (n_rem <= s || l >= cap) return n_rem;
    const uint64_t m = (uint64_t)1 << (cap - l);
    const uint32_t lo = (2 * s > n_rem) ? 2 * s - n_rem : 0;
    uint32_t hi = s - 1;
    const uint64_t cap_hi = ((uint64_t)s * m - n_rem) / (m - 1);
    if (cap_hi < hi) hi = (uint32_t)cap_hi;
    const uint32_t c = want < lo ? lo : want;
    return c > hi ? hi : c;
}

ZXH void zxh_complete(uint32_t* blc, int l, uint32_t s, uint32_t n_rem, const uint32_t* blc0, int cap) {
    for (int j = l + 1; j <= cap && n_rem; j++) {
        const uint32_t c = zxh_clamp(blc0[j], s, n_rem, j, cap);
        blc[j] = c;
        n_rem -= c;
        s = 2 * (s - c);
    }
}

/* greedy level-by-level choice of leaf counts, each level trying a handful of "flatter" shapes */
ZXH void zxh_walk(const uint32_t* blc0, const uint64_t* pf_rank, int n, int cap, uint32_t* out_blc) {
    for (int l = 0; l <= ZXH_LU; l++) out_blc[l] = 0;
    uint32_t s = 2, n_rem = (uint32_t)n;
    for (int l = 1; l <= cap && n_rem; l++) {
        uint32_t cand_c[6];
        int cand_flat[6];
        int n_cand = 0;
        const uint32_t c_base = zxh_clamp(blc0[l], s, n_rem, l, cap);
        cand_c[n_cand] = c_base;
        cand_flat[n_cand++] = 0;
        if (c_base < n_rem) {
            const uint32_t i_base = s - c_base;
            const uint32_t i_dn = 1u << zxh_log2(i_base);
            const uint32_t rest = i_base - i_dn;
            const uint32_t shapes[3] = {i_dn, i_dn << 1, i_dn | (rest ? 1u << zxh_log2(rest) : 0u)};
            for (int k = 0; k < 3; k++) {
                const uint32_t want = s > shapes[k] ? s - shapes[k] : 0;
                const uint32_t c = zxh_clamp(want, s, n_rem, l, cap);
                int dup = 0;
                for (int p = 0; p < n_cand; p++) dup |= (cand_c[p] == c);
                if (!dup) {
                    cand_c[n_cand] = c;
                    cand_flat[n_cand++] = 0;
                }
            }
            for (int d = 1; d <= cap - l && n_cand < 6; d++) { /* finish exactly as one flat run d levels down */
                const uint64_t den = ((uint64_t)1 << d) - 1;
                const int64_t num = (int64_t)((uint64_t)s << d) - (int64_t)n_rem;
                if (num < 0 || (uint64_t)num % den) continue;
                const uint64_t c64 = (uint64_t)num / den;
                if (c64 >= s || c64 >= n_rem) continue;
                cand_c[n_cand] = (uint32_t)c64;
                cand_flat[n_cand++] = d;
                break;
            }
        }
        uint64_t best_j = ZXH_U64MAX;
        uint32_t best_c = c_base;
        for (int k = 0; k < n_cand; k++) {
            uint32_t tmp[ZXH_LU + 1];
            for (int q = 0; q <= ZXH_LU; q++) tmp[q] = out_blc[q];
            tmp[l] = cand_c[k];
            const uint32_t rem = n_rem - cand_c[k];
            if (rem) {
                if (cand_flat[k]) tmp[l + cand_flat[k]] = rem;
                else zxh_complete(tmp, l, 2 * (s - cand_c[k]), rem, blc0, cap);
            }
            zxh_cost_t cc;
            zxh_eval(tmp, pf_rank, &cc);
            const uint64_t j = zxh_j(&cc);
            if (j < best_j) {
                best_j = j;
                best_c = cand_c[k];
            }
        }
        out_blc[l] = best_c;
        n_rem -= best_c;
        s = 2 * (s - best_c);
    }
}

/* cost of c groups placed at grouped level lc when s slots are open and k groups are already placed */
ZXH uint64_t zxh_run_cost(int lu, int lc, int g_log2, uint32_t s, uint32_t c, const uint64_t* pfg, uint32_t k) {
    if (!c) return 0;
    const int lr = lc + g_log2;
    const uint32_t w = 1u << (lu - lc);
    const uint32_t S = (1u << lu) - s * w;
    const uint32_t end = S + c * w;
    const uint64_t bits = (uint64_t)lr * (pfg[k + c] - pfg[k]);
    uint64_t touches = 0;
    uint32_t x = S;
    while (x < end) {
        const uint32_t wx = x ? (x & (0u - x)) : (1u << lu);
        const uint32_t wr = 1u << zxh_log2(end - x);
        const uint32_t Wd = wx < wr ? wx : wr;
        const int d = (int)zxh_log2(Wd >> (lu - lc)) + g_log2;
        const uint32_t i0 = k + ((x - S) >> (lu - lc));
        const uint64_t mass = pfg[i0 + (Wd >> (lu - lc))] - pfg[i0];
        touches += mass * (uint64_t)zxh_touch_weight(lr, d);
        x += Wd;
    }
    return 256u * bits + (uint64_t)ZXH_LAMBDA_Q8 * touches;
}

/* ---- grouped DP over (groups placed k, open slots s) per level (zxc_huffman.c:682-773) ------------
 * The reference pushes every state's transitions forward, keeping strict improvements in (k, s, c)
 * order.  A destination (k', s') is reached from at most one (s, c) per source row k -- c = k' - k,
 * s = s'/2 + c -- so pulling per destination over k = 0..k' visits its candidates in the same order
 * and keeps the same winner.  Destinations are independent: the device spreads them over the warp. */
typedef struct {
    uint64_t j;
    int l, k, s;
} zxh_dp_best_t;

/* best way into state (kd, sd) at level lc + 1, coming from level lc; *out_c is the arrival choice */
ZXH uint64_t zxh_dp_pull(const uint64_t* pfg, const uint64_t* jcur, int m, int cap_c, int lu, int g_log2, int lc, uint32_t kd,
                         uint32_t sd, uint32_t* out_c) {
    const uint32_t row = (uint32_t)(m + 1);
    const uint64_t mm = (uint64_t)1 << (cap_c - lc);
    uint64_t best = ZXH_U64MAX;
    uint32_t bc = 0;
    for (uint32_t k = 0; k <= kd && k < (uint32_t)m; k++) {
        const uint32_t c = kd - k, s = sd / 2 + c, n_rem = (uint32_t)m - k;
        if (s >= n_rem) continue; /* s > n_rem is no state; s == n_rem must finish at lc */
        const uint64_t j0 = jcur[k * row + s];
        if (j0 == ZXH_U64MAX) continue;
        const uint32_t lo = (2 * s > n_rem) ? 2 * s - n_rem : 0;
        uint32_t hi = s - 1;
        const uint64_t cap_hi = ((uint64_t)s * mm - n_rem) / (mm - 1);
        if (cap_hi < hi) hi = (uint32_t)cap_hi;
        if (c < lo || c > hi) continue;
        const uint64_t j = j0 + zxh_run_cost(lu, lc, g_log2, s, c, pfg, k);
        if (j < best) {
            best = j;
            bc = c;
        }
    }
    *out_c = bc;
    return best;
}

/* cost of closing the tree at level lc from row k (state s = m - k), or U64MAX */
ZXH uint64_t zxh_dp_finish(const uint64_t* pfg, const uint64_t* jcur, int m, int lu, int g_log2, int lc, uint32_t k) {
    const uint32_t s = (uint32_t)m - k;
    const uint64_t j0 = jcur[k * (uint32_t)(m + 1) + s];
    if (j0 == ZXH_U64MAX) return ZXH_U64MAX;
    return j0 + zxh_run_cost(lu, lc, g_log2, s, s, pfg, k) +
           (uint64_t)ZXH_LAMBDA_Q8 * (uint64_t)ZXH_LEVEL_COST * (uint64_t)(lc + g_log2 + 1);
}

ZXH int zxh_dp_backtrack(const uint16_t* arrive, int m, const zxh_dp_best_t* B, uint32_t* out_cblc) {
    if (B->j == ZXH_U64MAX) return 0;
    const uint32_t row = (uint32_t)(m + 1), plane = row * row;
    for (int l = 0; l <= ZXH_LU; l++) out_cblc[l] = 0;
    out_cblc[B->l] = (uint32_t)B->s;
    int k = B->k, s = B->s;
    for (int lc = B->l; lc > 1; lc--) {
        const uint32_t c = arrive[(uint32_t)lc * plane + (uint32_t)k * row + (uint32_t)s];
        out_cblc[lc - 1] = c;
        s = s / 2 + (int)c;
        k -= (int)c;
    }
    return (k == 0 && s == 2) ? 1 : 0;
}

/* one thread's version; 1 when a solution was written */
ZXH int zxh_dp_solve(const uint64_t* pfg, int m, int cap_c, int lu, int g_log2, uint32_t* out_cblc, zxh_work_t* W) {
    if (m < 2 || cap_c < 1 || m > ZXH_DP_M) return 0;
    const uint32_t row = (uint32_t)(m + 1), plane = row * row;
    uint64_t* jcur = W->dp_a;
    uint64_t* jnxt = W->dp_b;
    for (uint32_t i = 0; i < plane; i++) jcur[i] = ZXH_U64MAX;
    jcur[0 * row + 2] = 0;
    zxh_dp_best_t B;
    B.j = ZXH_U64MAX;
    B.l = B.k = B.s = 0;
    for (int lc = 1; lc <= cap_c; lc++) {
        for (uint32_t k = 0; k < (uint32_t)m; k++) {
            const uint64_t j = zxh_dp_finish(pfg, jcur, m, lu, g_log2, lc, k);
            if (j < B.j) {
                B.j = j;
                B.l = lc;
                B.k = (int)k;
                B.s = m - (int)k;
            }
        }
        if (lc == cap_c) break;
        for (uint32_t i = 0; i < plane; i++) jnxt[i] = ZXH_U64MAX;
        for (uint32_t kd = 0; kd <= (uint32_t)m; kd++) {
            for (uint32_t sd = 2; sd <= (uint32_t)m; sd += 2) {
                uint32_t c;
                const uint64_t j = zxh_dp_pull(pfg, jcur, m, cap_c, lu, g_log2, lc, kd, sd, &c);
                if (j != ZXH_U64MAX) {
                    jnxt[kd * row + sd] = j;
                    W->arrive[(uint32_t)(lc + 1) * plane + kd * row + sd] = (uint16_t)c;
                }
            }
        }
        uint64_t* t = jcur;
        jcur = jnxt;
        jnxt = t;
    }
    return zxh_dp_backtrack(W->arrive, m, &B, out_cblc);
}

/* The nudge in three steps so that the grouped DP in the middle can run warp-wide on the device:
 *   zxh_nudge_begin  baseline cost, rank order, the walk and reduced-cap candidates, DP inputs
 *   (DP)             zxh_dp_solve here; zxh_dp_pull / zxh_dp_finish spread over lanes on the device
 *   zxh_nudge_end    DP candidate -> lengths, guard rails, adoption */
typedef struct {
    int n, n_cand, do_dp, m, cap_c, g_log2;
    zxh_cost_t c0;
} zxh_nudge_t;

/* 0: alphabet too small to reshape (code_len stays); 1: continue with the DP (if S->do_dp) and _end */
ZXH int zxh_nudge_begin(const uint32_t* freq, const uint8_t* code_len, int max_code_len, zxh_work_t* W, zxh_nudge_t* S) {
    uint32_t blc0[ZXH_LU + 1];
    const int n = zxh_classes(code_len, blc0);
    S->n = n;
    S->n_cand = 0;
    S->do_dp = 0;
    if (n < 4) return 0;
    zxh_prefix_masses(code_len, freq, blc0, W->pf, W->val);
    zxh_eval(blc0, W->pf, &S->c0);

    /* symbols by descending (weight, symbol): the order lengths are handed out in */
    zxh_leaf_t* leaves = W->sort_tmp;
    int k = 0;
    for (int s = 0; s < ZXH_NSYM; s++) {
        if (!freq[s]) continue;
        leaves[k].w = freq[s];
        leaves[k].sym = (int16_t)s;
        k++;
    }
    zxh_sort_leaves(leaves, n);
    W->pf_rank[0] = 0;
    for (int r = 0; r < n; r++) {
        W->sym_order[r] = leaves[n - 1 - r].sym;
        W->pf_rank[r + 1] = W->pf_rank[r] + leaves[n - 1 - r].w;
    }

    int n_cand = 0;
    { /* candidate 1: greedy walk */
        uint32_t blc_w[ZXH_LU + 1];
        zxh_walk(blc0, W->pf_rank, n, max_code_len, blc_w);
        uint8_t* cl = W->cand[n_cand];
        for (int s = 0; s < ZXH_NSYM; s++) cl[s] = 0;
        int r = 0;
        for (int l = 1; l <= ZXH_LU; l++)
            for (uint32_t q = 0; q < blc_w[l]; q++) cl[W->sym_order[r++]] = (uint8_t)l;
        n_cand++;
    }
    int max_len0 = 0;
    for (int l = ZXH_LU; l >= 1; l--) {
        if (blc0[l]) {
            max_len0 = l;
            break;
        }
    }
    if (max_len0 >= 2) { /* candidates 2-3: plain package-merge with the cap lowered by 1 and 2 */
        for (int cut = 1; cut <= 2; cut++) {
            const int cap2 = max_len0 - cut;
            if (cap2 < 2 || (1u << cap2) < (uint32_t)n) break;
            if (zxh_build_code_lengths(freq, W->cand[n_cand], cap2, W) != 0) break;
            n_cand++;
        }
    }
    S->n_cand = n_cand;
    /* candidate 4: exact DP over groups of 1, 2 or 4 symbols */
    S->g_log2 = n <= 64 ? 0 : (n <= 128 ? 1 : 2);
    const int g = 1 << S->g_log2;
    S->m = (n + g - 1) / g;
    S->cap_c = max_code_len - S->g_log2;
    if (S->m >= 2 && S->cap_c >= 1 && S->m <= (1 << S->cap_c)) {
        for (int j2 = 0; j2 <= S->m; j2++) {
            int r = j2 * g;
            if (r > n) r = n;
            W->pfg[j2] = W->pf_rank[r];
        }
        S->do_dp = 1;
    }
    return 1;
}

/* dp_ok / cblc: result of the grouped DP (ignored unless S->do_dp); 1 if code_len was replaced */
ZXH int zxh_nudge_end(const uint32_t* freq, uint8_t* code_len, zxh_work_t* W, const zxh_nudge_t* S, int dp_ok,
                      const uint32_t* cblc) {
    const int n = S->n, g_log2 = S->g_log2, g = 1 << S->g_log2;
    int n_cand = S->n_cand;
    if (S->do_dp && dp_ok) {
        uint8_t* cl = W->cand[n_cand];
        for (int s = 0; s < ZXH_NSYM; s++) cl[s] = 0;
        int r = 0, ghosts = 0;
        uint8_t ghost_len = 0;
        for (int lc = 1; lc <= S->cap_c; lc++) {
            for (uint32_t q = 0; q < cblc[lc]; q++) {
                for (int e = 0; e < g; e++, r++) {
                    if (r < n) cl[W->sym_order[r]] = (uint8_t)(lc + g_log2);
                    else {
                        ghost_len = (uint8_t)(lc + g_log2);
                        ghosts++;
                    }
                }
            }
        }
        for (int s = 0; s < ZXH_NSYM && ghosts; s++) { /* pad the last group with absent symbols */
            if (freq[s] == 0 && cl[s] == 0) {
                cl[s] = ghost_len;
                ghosts--;
            }
        }
        n_cand++;
    }
    const uint64_t j0 = zxh_j(&S->c0);
    uint64_t best_j = j0;
    int best = -1;
    for (int ci = 0; ci < n_cand; ci++) {
        int valid = 1;
        for (int s = 0; s < ZXH_NSYM; s++) {
            if (freq[s] != 0 && W->cand[ci][s] == 0) {
                valid = 0;
                break;
            }
        }
        if (!valid) continue;
        uint32_t blc[ZXH_LU + 1];
        (void)zxh_classes(W->cand[ci], blc);
        zxh_prefix_masses(W->cand[ci], freq, blc, W->pf, W->val);
        zxh_cost_t c1;
        zxh_eval(blc, W->pf, &c1);
        if (c1.bits * 1000 > S->c0.bits * ZXH_BITS_PERMIL) continue;
        if (c1.touches * 256 > S->c0.touches * ZXH_MERGE_Q8) continue;
        const uint64_t j = zxh_j(&c1);
        if (j < best_j) {
            best_j = j;
            best = ci;
        }
    }
    if (best < 0) return 0;
    for (int s = 0; s < ZXH_NSYM; s++) code_len[s] = W->cand[best][s];
    return 1;
}

/* trades a few bytes of optimality for a flatter tree; 1 if code_len was replaced */
ZXH int zxh_nudge_code_lengths(const uint32_t* freq, uint8_t* code_len, int max_code_len, zxh_work_t* W) {
    zxh_nudge_t S;
    if (!zxh_nudge_begin(freq, code_len, max_code_len, W, &S)) return 0;
    uint32_t cblc[ZXH_LU + 1];
    int dp_ok = 0;
    if (S.do_dp) dp_ok = zxh_dp_solve(W->pfg, S.m, S.cap_c, ZXH_LU - S.g_log2, S.g_log2, cblc, W);
    return zxh_nudge_end(freq, code_len, W, &S, dp_ok, cblc);
}

/* ---- PivCo geometry of a canonical code (same (level, value) view as the decoder) ---------------- */
typedef struct {
    uint32_t cnt[ZXH_LU + 2], first[ZXH_LU + 2], lbase[ZXH_LU + 2], leafb[ZXH_LU + 2];
    uint32_t n_nodes;
    int single;
    uint8_t sorted[ZXH_NSYM]; /* symbols by (length, value) */
} zxh_geom_t;

/* 0 on success; -1 when the lengths are not a complete canonical code (zxc_huffman.c:1042-1084) */
ZXH int zxh_geometry(const uint8_t* code_len, zxh_geom_t* G) {
    uint32_t kraft = 0, present = 0;
    for (int l = 0; l <= ZXH_LU + 1; l++) G->cnt[l] = 0;
    for (int s = 0; s < ZXH_NSYM; s++) {
        const int l = code_len[s];
        if (l > ZXH_LU) return -1;
        if (l) {
            G->cnt[l]++;
            kraft += 1u << (ZXH_LU - l);
            present++;
        }
    }
    if (!present) return -1;
    G->single = (present == 1 && G->cnt[1] == 1 && kraft == (1u << (ZXH_LU - 1)));
    if (kraft != (1u << ZXH_LU) && !G->single) return -1;
    uint32_t pos[ZXH_LU + 2], acc = 0;
    for (int l = 1; l <= ZXH_LU; l++) {
        pos[l] = acc;
        acc += G->cnt[l];
    }
    for (int s = 0; s < ZXH_NSYM; s++)
        if (code_len[s]) G->sorted[pos[code_len[s]]++] = (uint8_t)s;
    uint32_t code = 0, nodes = 1, leaves = 0;
    G->first[0] = 0;
    G->lbase[0] = 0;
    G->leafb[0] = 0;
    for (int l = 1; l <= ZXH_LU; l++) {
        code = (code + G->cnt[l - 1]) << 1;
        G->first[l] = code;
        G->lbase[l] = nodes;
        G->leafb[l] = leaves;
        nodes += G->single ? (l == 1 ? 1u : 0u) : ((1u << l) - code);
        leaves += G->cnt[l];
    }
    G->first[ZXH_LU + 1] = 0;
    G->lbase[ZXH_LU + 1] = nodes;
    G->leafb[ZXH_LU + 1] = leaves;
    G->n_nodes = nodes;
    return nodes > 2 * ZXH_NSYM ? -1 : 0;
}

/* depth of the flat (perfect) subtree rooted at internal node (l, v): 0 = plain bitmap node */
ZXH uint32_t zxh_flat_depth(const zxh_geom_t* G, uint32_t l, uint32_t v) {
    for (uint32_t D = 1; l + D <= ZXH_LU; D++) {
        const uint32_t lo = v << D, hi = (v + 1) << D, ld = l + D;
        const uint32_t leaf_end = G->first[ld] + G->cnt[ld];
        if (hi <= leaf_end) return D >= 2 ? D : 0;
        if (lo < leaf_end) return 0;
    }
    return 0;
}

/* exact PivCo section size for these lengths (zxc_huffman.c:1219-1249); ZXH_U64MAX when not encodable.
 * node_count (>= 512 entries) receives the symbols routed through every node, by BFS index. */
ZXH uint64_t zxh_calc_size(const uint32_t* freq, const uint8_t* code_len, int with_header, zxh_geom_t* G, uint32_t* node_count) {
    if (zxh_geometry(code_len, G) != 0) return ZXH_U64MAX;
    for (int k = 0; k < ZXH_NSYM; k++)
        if (freq[k] != 0 && code_len[k] == 0) return ZXH_U64MAX;
    uint64_t total = 0;
    if (G->single) {
        const uint32_t c = freq[G->sorted[0]];
        node_count[0] = c;
        node_count[1] = c;
        total = ((uint64_t)c + 7) / 8;
        return total + (with_header ? 128u : 0u);
    }
    /* counts bottom-up: leaves first (they are the first cnt[l] nodes of each level) */
    for (int l = ZXH_LU; l >= 0; l--) {
        const uint32_t nn = l == 0 ? 1u : (1u << l) - G->first[l];
        for (uint32_t t = 0; t < nn; t++) {
            const uint32_t id = G->lbase[l] + t;
            if (l > 0 && t < G->cnt[l]) {
                node_count[id] = freq[G->sorted[G->leafb[l] + t]];
            } else {
                const uint32_t v = G->first[l] + t;
                const uint32_t cid = G->lbase[l + 1] + (2u * v - G->first[l + 1]);
                node_count[id] = node_count[cid] + node_count[cid + 1];
            }
        }
    }
    /* emitting nodes top-down, skipping everything below a flat root */
    for (uint32_t l = 0; l < ZXH_LU; l++) {
        const uint32_t nn = l == 0 ? 1u : (1u << l) - G->first[l];
        for (uint32_t t = (l == 0 ? 0u : G->cnt[l]); t < nn; t++) {
            const uint32_t v = G->first[l] + t;
            /* covered iff some proper ancestor is a flat root */
            int covered = 0;
            for (uint32_t a = 1; a <= l && !covered; a++) {
                const uint32_t al = l - a, av = v >> a;
                if (av < G->first[al] + (al ? G->cnt[al] : 0u)) break; /* ancestors are internal by construction */
                const uint32_t fd = zxh_flat_depth(G, al, av);
                if (fd >= 2 && fd >= a) covered = 1;
            }
            if (covered) continue;
            const uint32_t fd = zxh_flat_depth(G, l, v);
            const uint64_t c = node_count[G->lbase[l] + t];
            total += fd ? (c * fd + 7) / 8 : (c + 7) / 8;
        }
    }
    return total + (with_header ? 128u : 0u);
}

/* ceil(avg Huffman code length) over a strided sample, capped at 8 (zxc_compress.c:720-749);
 * hist is the sampled histogram, sampled its total */
ZXH uint32_t zxh_estimate_lit_bits(const uint32_t* hist, uint32_t sampled, uint8_t* code_len_tmp, zxh_work_t* W) {
    if (zxh_build_code_lengths(hist, code_len_tmp, 8, W) != 0) return 8;
    uint64_t total = 0;
    for (int k = 0; k < ZXH_NSYM; k++) total += (uint64_t)hist[k] * code_len_tmp[k];
    const uint32_t avg = (uint32_t)((total + sampled - 1) / sampled);
    return avg < 8 ? avg : 8;
}

#endif /* ZXC_B200_HUFENC_H */
