"""Small driver for profiling the encode kernel under ncu (256 MiB, level 3, 64 KiB blocks)."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import zxc_corpus as zc  # noqa: E402
import zxc_ctypes as z  # noqa: E402

prod = z.ZxcLib(z.PRODUCT_SO)
n = 256 << 20
d = zc.silesia_shaped(n, seed=1)
cap = int(prod.lib.zxc_compress_bound(n))
out = np.zeros(cap, np.uint8)
o = z.CompressOpts(level=int(sys.argv[1]) if len(sys.argv) > 1 else 3, block_size=65536, seekable=1)
for _ in range(2):
    r = prod.lib.zxc_compress(d.ctypes.data, n, out.ctypes.data, cap, C.byref(o))
print("compressed", r)
