"""Encode throughput per level through zxc_compress (host buffers), GPU build vs the reference on all
host threads, frames compared byte for byte.   python tests/enc_levels_bench.py [MiB] [levels] [block]"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, __file__.rsplit("/", 1)[0])
import zxc_corpus as zc  # noqa: E402
import zxc_ctypes as z  # noqa: E402


def main():
    mib = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    levels = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [6, 7]
    bs = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
    prod, ref = z.ZxcLib(z.PRODUCT_SO), z.ZxcLib(z.REF_SO)
    data = zc.silesia_shaped(mib << 20, seed=3)
    n = data.size
    cap = int(prod.lib.zxc_compress_bound(n))
    out = np.zeros(cap, np.uint8)
    th = zc.host_threads()
    for level in levels:
        o = z.CompressOpts(level=level, block_size=bs, seekable=1)
        r = prod.lib.zxc_compress(data.ctypes.data, n, out.ctypes.data, cap, C.byref(o))
        t = time.perf_counter()
        r = prod.lib.zxc_compress(data.ctypes.data, n, out.ctypes.data, cap, C.byref(o))
        dt = time.perf_counter() - t
        assert r > 0, z.ERR.get(r, r)
        if os.environ.get("ZXC_BENCH_NOREF"):
            print(f"L{level} bs={bs} {mib} MiB: GPU e2e {n / dt / 1e9:.3f} GB/s ({dt * 1e3:.0f} ms), ratio {r / n:.4f}", flush=True)
            continue
        t = time.perf_counter()
        rf = zc.compress_ref_mt(ref, data, level=level, block_size=bs)
        rdt = time.perf_counter() - t
        same = r == rf.size and np.array_equal(out[:r], rf)
        print(f"L{level} bs={bs} {mib} MiB: GPU e2e {n / dt / 1e9:.3f} GB/s ({dt * 1e3:.0f} ms), reference {th} threads "
              f"{n / rdt / 1e9:.3f} GB/s, ratio {r / n:.4f}, identical={same}", flush=True)


if __name__ == "__main__":
    main()
