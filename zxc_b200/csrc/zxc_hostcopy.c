/*
 * zxc-b200: bulk host copies for the staging paths (caller memory <-> pinned bounce buffers).
 *
 * The staged pipeline (zxg_decode_staged, zxg_h2d / zxg_d2h in zxc_gpu.cu) is bound by host memory traffic,
 * not by PCIe or the GPU: every decoded byte is written by DMA into a pinned buffer, read back by a CPU copy
 * and written into the caller's buffer.  An ordinary store to a line that is not in cache costs a read for
 * ownership plus the write-back, so memcpy() moves three lines per line copied; glibc only switches to
 * non-temporal stores far above the 512 KiB slices the copy pool hands out.  This copy always streams: the
 * destination is written with non-temporal stores (no read for ownership, no cache pollution -- the caller
 * will not find these buffers in cache anyway, they are tens of MiB), which cuts the traffic to two lines.
 *
 * Plain C so that gcc's target attribute + <immintrin.h> are used as intended (not through nvcc's front end).
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#if defined(__x86_64__)
#include <immintrin.h>

__attribute__((target("avx2"))) static void stream_copy_avx2(uint8_t* d, const uint8_t* s, size_t n) {
    size_t head = (size_t)(-(uintptr_t)d) & 31u; /* stores must be 32-byte aligned */
    if (head > n) head = n;
    memcpy(d, s, head);
    d += head;
    s += head;
    n -= head;
    size_t blocks = n >> 7;
    while (blocks--) {
        const __m256i a = _mm256_loadu_si256((const __m256i*)(s));
        const __m256i b = _mm256_loadu_si256((const __m256i*)(s + 32));
        const __m256i c = _mm256_loadu_si256((const __m256i*)(s + 64));
        const __m256i e = _mm256_loadu_si256((const __m256i*)(s + 96));
        _mm256_stream_si256((__m256i*)(d), a);
        _mm256_stream_si256((__m256i*)(d + 32), b);
        _mm256_stream_si256((__m256i*)(d + 64), c);
        _mm256_stream_si256((__m256i*)(d + 96), e);
        s += 128;
        d += 128;
    }
    _mm_sfence();
    memcpy(d, s, n & 127u);
}

static int have_avx2(void) {
    static int cached = -1;
    if (cached < 0) {
        const char* e = getenv("ZXC_B200_STREAM_COPY"); /* =0: plain memcpy (for A/B measurements) */
        cached = (!e || atoi(e) != 0) && __builtin_cpu_supports("avx2") ? 1 : 0;
    }
    return cached;
}
#endif

/* Copies n bytes (regions must not overlap); streams the destination when that pays. */
void zxh_stream_copy(void* dst, const void* src, size_t n) {
#if defined(__x86_64__)
    if (n >= 4096 && have_avx2()) {
        stream_copy_avx2((uint8_t*)dst, (const uint8_t*)src, n);
        return;
    }
#endif
    memcpy(dst, src, n);
}
