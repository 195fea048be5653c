"""One zxc_decompress call over several devices (ZXC_B200_DEVICES): GB/s of decoded bytes host to host, page-locked and
ordinary buffers, for 1..N devices.   python tests/multi_dev_e2e.py [GiB] [reps]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, __file__.rsplit("/", 1)[0])
import zxc_corpus as zc  # noqa: E402
import zxc_ctypes as z  # noqa: E402

gib = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
prod, ref = z.ZxcLib(z.PRODUCT_SO), z.ZxcLib(z.REF_SO)
nd = prod.lib.zxc_b200_device_count()
data = zc.silesia_shaped(int(gib * (1 << 30)), seed=1)
frame = zc.compress_ref_mt(ref, data, level=3, block_size=65536)
out = np.zeros(data.size, np.uint8)
out[::4096] = 1
p_frame = torch.from_numpy(frame).pin_memory()
p_out = torch.zeros(data.size, dtype=torch.uint8).pin_memory()
devs = [d for d in (1, 2, 4, 8) if d <= nd]
for d in devs:
    os.environ["ZXC_B200_DEVICES"] = str(d)
    for name, src, dst in (("page-locked", p_frame.data_ptr(), p_out.data_ptr()), ("pageable", frame.ctypes.data, out.ctypes.data)):
        rates = []
        for rep in range(reps + 1):
            t = time.perf_counter()
            r = prod.lib.zxc_decompress(src, frame.size, dst, data.size, None)
            dt = time.perf_counter() - t
            assert r == data.size, r
            rates.append(data.size / dt / 1e9)
        print(f"{d} device(s), {name}, {gib:g} GiB: first {rates[0]:.1f}, then " + " ".join(f"{x:.1f}" for x in rates[1:]) + " GB/s", flush=True)
    assert np.array_equal(out, data) and np.array_equal(p_out.numpy(), data)
h = ref.lib.zxc_seekable_open(frame.ctypes.data, frame.size)
t = time.perf_counter()
ref.lib.zxc_seekable_decompress_range_mt(h, out.ctypes.data, out.size, 0, out.size, zc.host_threads())
print(f"reference range_mt, {zc.host_threads()} threads: {data.size / (time.perf_counter() - t) / 1e9:.1f} GB/s")
