/*
 * TEST INFRASTRUCTURE ONLY.  Exposes four internal (hidden-visibility) functions of the UNMODIFIED
 * reference, as compiled by oracle/Makefile into oracle/_ref/obj/, so that tests/test_hufenc.py can
 * pin zxc_b200/csrc/zxc_hufenc.h (code-length construction for levels 6-7) against the reference's
 * own behaviour.  Prototypes: /root/reference/src/lib/zxc_internal.h:1478, :1510, :1568, :1572.
 */
#include <stddef.h>
#include <stdint.h>

int zxc_huf_build_code_lengths_default(const uint32_t* freq, uint8_t* code_len, void* scratch, int max_code_len);
int zxc_huf_nudge_code_lengths(const uint32_t* freq, uint8_t* code_len, void* scratch, int max_code_len);
size_t zxc_huf_calc_size_default(const uint32_t* freq, const uint8_t* code_len, int with_header);
int zxc_huf_encode_section_default(const uint8_t* lit, size_t n, const uint32_t* freq, const uint8_t* code_len,
                                   uint8_t* dst, size_t dst_cap);

#define EXPORT __attribute__((visibility("default")))

EXPORT int zxri_build_code_lengths(const uint32_t* freq, uint8_t* code_len, int max_code_len) {
    return zxc_huf_build_code_lengths_default(freq, code_len, NULL, max_code_len);
}
EXPORT int zxri_nudge_code_lengths(const uint32_t* freq, uint8_t* code_len, int max_code_len) {
    return zxc_huf_nudge_code_lengths(freq, code_len, NULL, max_code_len);
}
EXPORT uint64_t zxri_calc_size(const uint32_t* freq, const uint8_t* code_len, int with_header) {
    return (uint64_t)zxc_huf_calc_size_default(freq, code_len, with_header);
}
EXPORT int zxri_encode_section(const uint8_t* lit, size_t n, const uint32_t* freq, const uint8_t* code_len, uint8_t* dst,
                               size_t dst_cap) {
    return zxc_huf_encode_section_default(lit, n, freq, code_len, dst, dst_cap);
}
