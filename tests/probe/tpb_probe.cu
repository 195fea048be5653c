// Development probe: what could a thread-per-block LZ decoder reach on a B200?
// Each lane owns one 64 KiB output block and one compressed stream, and runs the memory/ALU pattern of a byte-aligned
// LZ decoder without the parsing: items of 1..13 bytes (mean ~6.5, the bench corpus' mean literal run / match),
// alternately taken from the lane's sequential "literal" stream and from its own earlier output at a random distance
// (up to 64 KiB back), appended through a 16-byte register window, stored as aligned 16-byte words.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tpb_probe tpb_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
typedef uint32_t u32; typedef uint64_t u64; typedef uint8_t u8;

__device__ __forceinline__ void shr128(u32 (&w)[8], u32 sh_bytes, u32 (&o)[4]) { // o = (w >> 8*sh) low 128 bits, w = 256 bits
    const u32 ws = sh_bytes >> 2, bs = (sh_bytes & 3) * 8;
    u32 t[5];
#pragma unroll
    for (int k = 0; k < 5; k++) {
        u32 v = w[k];
        if (ws == 1) v = w[k + 1]; else if (ws == 2) v = w[k + 2]; else if (ws == 3) v = w[k + 3];
        t[k] = v;
    }
#pragma unroll
    for (int k = 0; k < 4; k++) o[k] = __funnelshift_r(t[k], t[k + 1], bs);
}

__global__ void __launch_bounds__(128) tpb_kernel(const u8* __restrict__ lit, u8* out, u32 n_blocks, u32 lit_stride,
                                                  u32 max_dist, unsigned long long* counter) {
    for (;;) {
        u32 b;
        { // one block per lane, claimed 32 at a time by the warp
            unsigned long long base = 0;
            if ((threadIdx.x & 31) == 0) base = atomicAdd(counter, 32ull);
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            if (base >= n_blocks) return;
            b = (u32)base + (threadIdx.x & 31);
        }
        const bool live = b < n_blocks;
        const u8* ls = lit + (size_t)(live ? b : 0) * lit_stride;
        u8* dst = out + (size_t)(live ? b : 0) * 65536u;
        u32 rng = b * 2654435761u + 12345u;
        u32 op = 0, lp = 0, fill = 0;
        u32 acc[4] = {0, 0, 0, 0};
        bool is_lit = true;
        while (live && op + fill < 65536u - 32u) {
            rng = rng * 1664525u + 1013904223u;
            const u32 n = 1u + ((rng >> 20) % 13u);
            const u8* sp;
            const u32 pos = op + fill;
            if (is_lit || pos < 64u) { sp = ls + lp; lp += n; if (lp + 64u > lit_stride) lp = 0; }
            else {
                u32 lim = pos - 16u < max_dist ? pos - 16u : max_dist;
                const u32 d = 16u + ((rng >> 8) % lim);           // >= 16 back: bytes already stored (window flushed below if needed)
                if (d < fill + 16u) {                             // source overlaps the unflushed window: flush it partially
                    *reinterpret_cast<uint4*>(dst + op) = make_uint4(acc[0], acc[1], acc[2], acc[3]);
                }
                sp = dst + pos - d;
            }
            is_lit = !is_lit;
            // unaligned 16-byte read: two aligned 16-byte loads + shift
            const uintptr_t a = reinterpret_cast<uintptr_t>(sp);
            const uint4 x = *reinterpret_cast<const uint4*>(a & ~(uintptr_t)15);
            const uint4 y = *reinterpret_cast<const uint4*>((a & ~(uintptr_t)15) + 16);
            u32 w[8] = {x.x, x.y, x.z, x.w, y.x, y.y, y.z, y.w};
            u32 v[4];
            shr128(w, (u32)(a & 15), v);
            // append n bytes of v at byte position fill of the window
            u32 z[8] = {0, 0, 0, 0, v[0], v[1], v[2], v[3]};      // (v << 128) >> (128 - 8*fill) == v << 8*fill
            u32 sh[4];
            shr128(z, 16u - fill, sh);
            const u32 keep = fill;                                // bytes of acc that stay
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int lo = 4 * k;                             // byte index of word k
                u32 m = 0xFFFFFFFFu;                              // mask of bytes in word k below `keep`
                if ((int)keep <= lo) m = 0; else if ((int)keep < lo + 4) m = (1u << (8 * (keep - lo))) - 1u;
                acc[k] = (acc[k] & m) | (sh[k] & ~m);
            }
            if (fill + n >= 16u) {
                *reinterpret_cast<uint4*>(dst + op) = make_uint4(acc[0], acc[1], acc[2], acc[3]);
                op += 16u;
                // the bytes of v that did not fit: v >> 8*(16 - fill)
                u32 z2[8] = {v[0], v[1], v[2], v[3], 0, 0, 0, 0};
                u32 r[4];
                shr128(z2, 16u - fill, r);
                if (fill == 0) { r[0] = r[1] = r[2] = r[3] = 0; }
                acc[0] = r[0]; acc[1] = r[1]; acc[2] = r[2]; acc[3] = r[3];
                fill = fill + n - 16u;
            } else fill += n;
        }
        if (live) *reinterpret_cast<uint4*>(dst + op) = make_uint4(acc[0], acc[1], acc[2], acc[3]);
    }
}

int main(int argc, char** argv) {
    const u32 gib = argc > 1 ? atoi(argv[1]) : 4;
    const u32 n_blocks = gib * 16384u;
    const u32 lit_stride = 28672;
    u8 *lit, *out; unsigned long long* ctr;
    cudaMalloc(&lit, (size_t)n_blocks * lit_stride + 64); cudaMalloc(&out, (size_t)n_blocks * 65536 + 64); cudaMalloc(&ctr, 8);
    cudaMemset(lit, 0x5A, (size_t)n_blocks * lit_stride + 64);
    cudaMemset(out, 0, (size_t)n_blocks * 65536 + 64);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const u32 dists[3] = {65536u, 4096u, 1024u};
    for (int ctas = 2; ctas <= 16; ctas *= 2) for (int di = 0; di < 3; di++) {
        int grid = 148 * ctas;
        float best = 1e9f;
        for (int rep = 0; rep < 3; rep++) {
            cudaMemset(ctr, 0, 8);
            cudaEventRecord(e0);
            tpb_kernel<<<grid, 128>>>(lit, out, n_blocks, lit_stride, dists[di], ctr);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        cudaError_t err = cudaGetLastError();
        printf("ctas/SM %d (%d warps/SM) max_dist %6u: %.3f ms  %.1f GB/s out  (%s)\n", ctas, ctas * 4, dists[di], best,
               (double)n_blocks * 65536 / best / 1e6, cudaGetErrorString(err));
    }
    return 0;
}
