"""zxc_decompress with ordinary (pageable) host buffers: GB/s of decoded bytes, next to the reference on all
host threads.   python tests/pageable_e2e.py [GiB[,GiB...]] [reps]"""
import sys
import time

import numpy as np

sys.path.insert(0, __file__.rsplit("/", 1)[0])
import zxc_corpus as zc  # noqa: E402
import zxc_ctypes as z  # noqa: E402

sizes = [float(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "1").split(",")]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
prod, ref = z.ZxcLib(z.PRODUCT_SO), z.ZxcLib(z.REF_SO)
full = zc.silesia_shaped(int(max(sizes) * (1 << 30)), seed=1)
for gib in sizes:
    data = full[: int(gib * (1 << 30))]
    frame = zc.compress_ref_mt(ref, data, level=3, block_size=65536)
    out = np.zeros(data.size, np.uint8)
    out[::4096] = 1  # touch every page: the timed calls should not be the ones that fault the buffer in
    rates = []
    for rep in range(reps + 1):
        t = time.perf_counter()
        r = prod.lib.zxc_decompress(frame.ctypes.data, frame.size, out.ctypes.data, out.size, None)
        dt = time.perf_counter() - t
        assert r == data.size, r
        rates.append(data.size / dt / 1e9)
    print(f"GPU zxc_decompress, pageable buffers, {gib:g} GiB: first call {rates[0]:.2f}, then "
          + " ".join(f"{x:.2f}" for x in rates[1:]) + " GB/s", flush=True)
    assert np.array_equal(out, data)
    h = ref.lib.zxc_seekable_open(frame.ctypes.data, frame.size)
    t = time.perf_counter()
    r = ref.lib.zxc_seekable_decompress_range_mt(h, out.ctypes.data, out.size, 0, out.size, zc.host_threads())
    print(f"reference range_mt, {zc.host_threads()} threads, {gib:g} GiB: {data.size / (time.perf_counter() - t) / 1e9:.2f} GB/s", flush=True)
