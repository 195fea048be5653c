/*
 * TEST INFRASTRUCTURE ONLY.  Compiles the product's host/device header zxc_b200/csrc/zxc_hufenc.h as
 * plain C so the CPU suite can compare it with the reference internals (oracle/ref_internals.c).
 */
#include <stdlib.h>
#include "../zxc_b200/csrc/zxc_hufenc.h"

#define EXPORT __attribute__((visibility("default")))

EXPORT int zxhh_build_code_lengths(const uint32_t* freq, uint8_t* code_len, int max_code_len) {
    zxh_work_t* W = malloc(sizeof *W);
    const int r = zxh_build_code_lengths(freq, code_len, max_code_len, W);
    free(W);
    return r;
}
EXPORT int zxhh_nudge_code_lengths(const uint32_t* freq, uint8_t* code_len, int max_code_len) {
    zxh_work_t* W = malloc(sizeof *W);
    const int r = zxh_nudge_code_lengths(freq, code_len, max_code_len, W);
    free(W);
    return r;
}
EXPORT uint64_t zxhh_calc_size(const uint32_t* freq, const uint8_t* code_len, int with_header) {
    zxh_geom_t G;
    uint32_t counts[2 * ZXH_NSYM + 2];
    return zxh_calc_size(freq, code_len, with_header, &G, counts);
}
EXPORT uint32_t zxhh_estimate_lit_bits(const uint32_t* hist, uint32_t sampled) {
    zxh_work_t* W = malloc(sizeof *W);
    uint8_t cl[256];
    const uint32_t r = zxh_estimate_lit_bits(hist, sampled, cl, W);
    free(W);
    return r;
}
EXPORT uint64_t zxhh_work_bytes(void) { return sizeof(zxh_work_t); }

/* zxh_cost_add on its own: bits and work of c items of level lc (group size 2^g) from index `at`, masses pf[first..] */
EXPORT void zxhh_cost_add(int lc, int g, uint32_t at, uint32_t c, const uint64_t* pf, uint32_t first, uint64_t* bits,
                          uint64_t* work) {
    zxh_cost_t acc;
    acc.bits = acc.work = 0;
    zxh_cost_add(&acc, lc, g, at, c, pf, first);
    *bits = acc.bits;
    *work = acc.work;
}
