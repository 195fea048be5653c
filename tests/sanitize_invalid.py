"""compute-sanitizer driver: hostile inputs through the decode kernel (reject vectors + random
frame mutations) -- memcheck must stay silent.  Usage on the GPU box:
    compute-sanitizer --tool memcheck python tests/sanitize_invalid.py"""
import glob
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import zxc_corpus as zc  # noqa: E402
import zxc_ctypes as z  # noqa: E402

prod, ref = z.ZxcLib(z.PRODUCT_SO), z.ZxcLib(z.REF_SO)
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
n = 0
for p in sorted(glob.glob(os.path.join(G, "invalid", "*.zxc")) + glob.glob(os.path.join(G, "valid", "*.zxc"))):
    fr = open(p, "rb").read()
    if len(fr) >= 28:
        prod.decompress(fr, 1 << 20, checksum=1)
        n += 1
data = zc.silesia_shaped(1 << 20, seed=5)[:150000]
rng = np.random.default_rng(9)
for level, bs in ((3, 4096), (1, 4096), (6, 65536), (7, 65536)):
    frame = ref.compress(data, level=level, block_size=bs, checksum=0)
    for t in range(4 if os.environ.get("ZXC_SANITIZE_QUICK") else 60):
        f = frame.copy()
        for _ in range(int(rng.integers(1, 4))):
            f[int(rng.integers(16, f.size - 12))] = int(rng.integers(0, 256))
        prod.decompress(f, data.size)
        n += 1
print("sanitize_invalid: ran", n, "hostile decodes")
