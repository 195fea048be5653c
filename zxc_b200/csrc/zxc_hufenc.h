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

/* ---- cost model of a code shape (what zxc_huffman.c:343-431 prices) ---------------------------------
 * A canonical code hands out its code space from the left, shorter codes first.  At length l the next free
 * codeword has index `at` counted in codewords of that length (at = 0 at length 1; at' = 2 (at + c) one level down
 * after c leaves), so the c leaves of one length are the index interval [at, at + c).  The decoder resolves an
 * aligned block of 2^D of them with one table step, and the interval splits in exactly one way into maximal aligned
 * blocks.  They are found here from the two ends: going up, every set bit of the left end whose block still fits is
 * peeled off; what remains is shorter than the alignment reached and splits by the set bits of its length, largest
 * first.  (The reference walks the same blocks left to right; integer sums do not care.)  Costs are two sums over
 * the symbol masses: code bits, and decoder work = mass x (steps to resolve a symbol of that block).
 * g > 0 is the grouped view of the DP: an item is a group of 2^g symbols, its level lc sits g above the leaves. */
typedef struct { uint64_t bits, work; } zxh_cost_t;

/* steps for a symbol of length len inside a flat block of depth D: walk len - D levels, then one lookup -- wide
 * lookups (D > 6) are charged extra */
ZXH uint64_t zxh_block_weight(int len, int D) { return (uint64_t)(len + 1 - D + (D > ZXH_FLAT_SIMD_MAX ? ZXH_DEEP_FLAT_PENALTY : 0)); }

/* add c items of level lc (length lc + g) that start at index `at`; their masses are pf[first .. first + c] */
ZXH void zxh_cost_add(zxh_cost_t* acc, int lc, int g, uint32_t at, uint32_t c, const uint64_t* pf, uint32_t first) {
    if (!c) return;
    const int len = lc + g;
    acc->bits += (uint64_t)len * (pf[first + c] - pf[first]);
    uint32_t lo = at, i = first;
    const uint32_t hi = at + c;
    int d = 0;
    for (; (1u << d) <= hi - lo; d++) { /* lo is a multiple of 2^d here */
        const uint32_t sz = 1u << d;
        if (lo & sz) {
            acc->work += (pf[i + sz] - pf[i]) * zxh_block_weight(len, d + g);
            lo += sz;
            i += sz;
        }
    }
    while (d-- > 0) { /* hi - lo < 2^(d+1) and lo is aligned that far: the rest by the bits of its length */
        const uint32_t sz = 1u << d;
        if ((hi - lo) & sz) {
            acc->work += (pf[i + sz] - pf[i]) * zxh_block_weight(len, d + g);
            lo += sz;
            i += sz;
        }
    }
}

ZXH uint64_t zxh_j(const zxh_cost_t* c) { return 256u * c->bits + (uint64_t)ZXH_LAMBDA_Q8 * c->work; }

/* cost of a whole shape: cnt[l] leaves of length l, masses in the order the lengths are handed out */
ZXH void zxh_shape_cost(const uint32_t* cnt, const uint64_t* pf, zxh_cost_t* out) {
    out->bits = out->work = 0;
    uint32_t at = 0, first = 0;
    int deepest = 0;
    for (int l = 1; l <= ZXH_LU; l++) {
        zxh_cost_add(out, l, 0, at, cnt[l], pf, first);
        if (cnt[l]) deepest = l;
        first += cnt[l];
        at = 2 * (at + cnt[l]);
    }
    out->work += (uint64_t)ZXH_LEVEL_COST * (uint64_t)(deepest + 1);
}

/* cost of a set of code lengths: leaves per length, masses laid out by (length, symbol) */
ZXH int zxh_lengths_cost(const uint8_t* code_len, const uint32_t* freq, zxh_work_t* W, uint32_t* cnt, zxh_cost_t* out) {
    uint32_t next[ZXH_LU + 2];
    int n = 0;
    for (int l = 0; l <= ZXH_LU; l++) cnt[l] = 0;
    for (int s = 0; s < ZXH_NSYM; s++) cnt[code_len[s]]++;
    n = ZXH_NSYM - (int)cnt[0];
    cnt[0] = 0;
    next[1] = 0;
    for (int l = 1; l <= ZXH_LU; l++) next[l + 1] = next[l] + cnt[l];
    for (int s = 0; s < ZXH_NSYM; s++)
        if (code_len[s]) W->val[next[code_len[s]]++] = freq[s];
    W->pf[0] = 0;
    for (int i = 0; i < n; i++) W->pf[i + 1] = W->pf[i] + W->val[i];
    zxh_shape_cost(cnt, W->pf, out);
    return n;
}

/* How many of the `left` remaining leaves can take length l when `open` codewords of that length are free and no
 * code may be longer than cap.  If they all fit, or this is the last length, they all go here.  Otherwise c of
 * them here leaves left - c for the 2 (open - c) codewords one level down: that needs at least one inner node
 * (c <= open - 1), must not strand code space (left - c >= 2 (open - c), i.e. c >= 2 open - left) and must still fit
 * at the cap (left - c <= (open - c) 2^(cap - l)).  `want` is pulled into that range. */
ZXH uint32_t zxh_fit(uint32_t want, uint32_t open, uint32_t left, int l, int cap) {
    if (left <= open || l >= cap) return left;
    const uint64_t room = (uint64_t)1 << (cap - l);
    const uint64_t most = ((uint64_t)open * room - left) / (room - 1);
    const uint32_t hi = most < open - 1 ? (uint32_t)most : open - 1;
    const uint32_t lo = 2 * open > left ? 2 * open - left : 0;
    return want < lo ? lo : (want > hi ? hi : want);
}

/* Greedy reshaping, level by level (zxc_huffman.c:432-590): at each length try the baseline count and a few counts
 * that leave a rounder number of inner nodes, finish each trial the baseline's way (or as one flat run), keep the
 * cheapest.  The trials of one level share everything above it, so only the part from this level down is priced:
 * the order of the totals is the order of these tails. */
ZXH void zxh_walk(const uint32_t* want, const uint64_t* pf, int n, int cap, uint32_t* cnt) {
    for (int l = 0; l <= ZXH_LU; l++) cnt[l] = 0;
    uint32_t open = 2, left = (uint32_t)n, first = 0;
    for (int l = 1; l <= cap && left; l++) {
        uint32_t opt[5];
        int run[5]; /* > 0: everything that is left goes run[] levels further down as one flat block */
        int k = 0;
        opt[0] = zxh_fit(want[l], open, left, l, cap);
        run[k++] = 0;
        if (opt[0] < left) {
            /* inner nodes kept by the baseline, rounded down to a power of two, that doubled, or its two leading bits */
            const uint32_t inner = open - opt[0];
            const uint32_t top = 1u << zxh_log2(inner);
            const uint32_t two = inner == top ? top : top | (1u << zxh_log2(inner - top));
            for (int t = 0; t < 3; t++) {
                const uint32_t keep = t == 0 ? top : (t == 1 ? 2 * top : two);
                const uint32_t c = zxh_fit(open > keep ? open - keep : 0, open, left, l, cap);
                int seen = 0;
                for (int q = 0; q < k; q++) seen |= opt[q] == c;
                if (!seen) {
                    opt[k] = c;
                    run[k++] = 0;
                }
            }
            /* the shallowest flat finish: (open - c) 2^d codewords for exactly left - c leaves */
            for (int d = 1; d <= cap - l; d++) {
                const uint64_t span = ((uint64_t)open << d), per = ((uint64_t)1 << d) - 1;
                if (span < left || (span - left) % per) continue;
                const uint64_t c = (span - left) / per;
                if (c >= open || c >= left) continue;
                opt[k] = (uint32_t)c;
                run[k++] = d;
                break;
            }
        }
        uint64_t best = ZXH_U64MAX;
        uint32_t pick = opt[0];
        for (int q = 0; q < k; q++) {
            zxh_cost_t tail;
            tail.bits = tail.work = 0;
            uint32_t at = (1u << l) - open, o = open, r = left, f = first;
            int lv = l, deepest = l;
            uint32_t c = opt[q];
            for (;;) {
                zxh_cost_add(&tail, lv, 0, at, c, pf, f);
                if (c) deepest = lv;
                r -= c;
                if (!r) break;
                f += c;
                if (run[q]) { /* one flat block run[q] levels down takes the rest */
                    at = (at + c) << run[q];
                    lv += run[q];
                    c = r;
                    continue;
                }
                o = 2 * (o - c);
                at = 2 * (at + c);
                lv++;
                c = zxh_fit(want[lv], o, r, lv, cap);
            }
            tail.work += (uint64_t)ZXH_LEVEL_COST * (uint64_t)(deepest + 1);
            const uint64_t j = zxh_j(&tail);
            if (j < best) {
                best = j;
                pick = opt[q];
            }
        }
        cnt[l] = pick;
        first += pick;
        left -= pick;
        open = 2 * (open - pick);
    }
}

/* cost of c groups placed at grouped level lc when s slots are open there and k groups are already placed */
ZXH uint64_t zxh_run_cost(int lu, int lc, int g_log2, uint32_t s, uint32_t c, const uint64_t* pfg, uint32_t k) {
    zxh_cost_t r;
    (void)lu;
    r.bits = r.work = 0;
    zxh_cost_add(&r, lc, g_log2, (1u << lc) - s, c, pfg, k);
    return zxh_j(&r);
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

/* The nudge (zxc_huffman.c:803-945) in three steps so that the grouped DP in the middle can run warp-wide on the
 * device:
 *   zxh_nudge_begin  baseline cost, rank order, the walk and reduced-cap candidates, DP inputs
 *   (DP)             zxh_dp_solve here; zxh_dp_pull / zxh_dp_finish spread over lanes on the device
 *   zxh_nudge_end    DP candidate -> lengths, guard rails, adoption */
typedef struct {
    int n, n_cand, do_dp, m, cap_c, g_log2;
    zxh_cost_t c0;
} zxh_nudge_t;

/* Lengths from a shape: ranks are served in order, cnt[lv] items of 2^g_log2 ranks each get length lv + g_log2.
 * Ranks past the alphabet (the DP pads its last group) fall on absent symbols, lowest first. */
ZXH void zxh_lengths_from_shape(const uint32_t* cnt, int levels, int g_log2, int n, const uint32_t* freq, const int16_t* order,
                                uint8_t* cl) {
    for (int s = 0; s < ZXH_NSYM; s++) cl[s] = 0;
    int r = 0, spare = 0;
    for (int lv = 1; lv <= levels; lv++) {
        const int stop = r + ((int)cnt[lv] << g_log2);
        for (; r < stop; r++) {
            if (r < n) {
                cl[order[r]] = (uint8_t)(lv + g_log2);
            } else { /* a padding rank: the next symbol that does not occur */
                while (spare < ZXH_NSYM && (freq[spare] != 0 || cl[spare] != 0)) spare++;
                if (spare < ZXH_NSYM) cl[spare] = (uint8_t)(lv + g_log2);
            }
        }
    }
}

/* 0: alphabet too small to reshape (code_len stays); 1: continue with the DP (if S->do_dp) and _end */
ZXH int zxh_nudge_begin(const uint32_t* freq, const uint8_t* code_len, int max_code_len, zxh_work_t* W, zxh_nudge_t* S) {
    uint32_t base[ZXH_LU + 1];
    S->n_cand = S->do_dp = 0;
    S->n = 0;
    for (int s = 0; s < ZXH_NSYM; s++) S->n += code_len[s] != 0;
    if (S->n < 4) return 0;
    const int n = zxh_lengths_cost(code_len, freq, W, base, &S->c0);

    /* ranks: symbols by falling (weight, symbol), the order in which a shape hands out its lengths */
    zxh_leaf_t* up = W->sort_tmp;
    for (int s = 0, k = 0; s < ZXH_NSYM; s++) {
        if (freq[s]) {
            up[k].w = freq[s];
            up[k++].sym = (int16_t)s;
        }
    }
    zxh_sort_leaves(up, n);
    W->pf_rank[0] = 0;
    for (int r = 0; r < n; r++) {
        W->sym_order[r] = up[n - 1 - r].sym;
        W->pf_rank[r + 1] = W->pf_rank[r] + up[n - 1 - r].w;
    }

    /* candidate: the greedy walk */
    uint32_t shape[ZXH_LU + 1];
    zxh_walk(base, W->pf_rank, n, max_code_len, shape);
    zxh_lengths_from_shape(shape, ZXH_LU, 0, n, freq, W->sym_order, W->cand[S->n_cand++]);

    /* candidates: package-merge again with the longest code one and two bits shorter, while the alphabet fits */
    int longest = ZXH_LU;
    while (longest > 0 && !base[longest]) longest--;
    for (int cap2 = longest - 1; cap2 >= longest - 2 && cap2 >= 2 && (1u << cap2) >= (uint32_t)n; cap2--) {
        if (zxh_build_code_lengths(freq, W->cand[S->n_cand], cap2, W) != 0) break;
        S->n_cand++;
    }

    /* candidate: the exact DP over groups of 1, 2 or 4 ranks (at most ZXH_DP_M of them) */
    S->g_log2 = n <= 64 ? 0 : (n <= 128 ? 1 : 2);
    S->m = (n + (1 << S->g_log2) - 1) >> S->g_log2;
    S->cap_c = max_code_len - S->g_log2;
    if (S->m >= 2 && S->cap_c >= 1 && S->m <= (1 << S->cap_c)) {
        for (int q = 0; q <= S->m; q++) W->pfg[q] = W->pf_rank[(q << S->g_log2) < n ? (q << S->g_log2) : n];
        S->do_dp = 1;
    }
    return 1;
}

/* dp_ok / cblc: result of the grouped DP (ignored unless S->do_dp); 1 if code_len was replaced */
ZXH int zxh_nudge_end(const uint32_t* freq, uint8_t* code_len, zxh_work_t* W, const zxh_nudge_t* S, int dp_ok,
                      const uint32_t* cblc) {
    int n_cand = S->n_cand;
    if (S->do_dp && dp_ok)
        zxh_lengths_from_shape(cblc, S->cap_c, S->g_log2, S->n, freq, W->sym_order, W->cand[n_cand++]);
    /* adopt the cheapest candidate that covers the alphabet, costs at most 1.5 % more bits and saves at least a
     * tenth of the decoder's work; the first one wins a tie */
    uint64_t best = zxh_j(&S->c0);
    const uint8_t* winner = 0;
    for (int q = 0; q < n_cand; q++) {
        const uint8_t* cl = W->cand[q];
        int holes = 0;
        for (int s = 0; s < ZXH_NSYM; s++) holes |= (freq[s] != 0) & (cl[s] == 0);
        if (holes) continue;
        uint32_t cnt[ZXH_LU + 1];
        zxh_cost_t c;
        (void)zxh_lengths_cost(cl, freq, W, cnt, &c);
        if (c.bits * 1000 > S->c0.bits * ZXH_BITS_PERMIL || c.work * 256 > S->c0.work * ZXH_MERGE_Q8) continue;
        if (zxh_j(&c) < best) {
            best = zxh_j(&c);
            winner = cl;
        }
    }
    if (!winner) return 0;
    for (int s = 0; s < ZXH_NSYM; s++) code_len[s] = winner[s];
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
