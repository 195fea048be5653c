"""compute-sanitizer driver: the encode kernels (all levels, edge sizes, dictionaries) -- memcheck must
stay silent and every frame must equal the reference's.  Usage on the GPU box:
    compute-sanitizer --tool memcheck python tests/sanitize_encode.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import zxc_corpus as zc  # noqa: E402
import zxc_ctypes as z  # noqa: E402
from test_oracle import GC_DICT, golden_dicts, make_case  # noqa: E402

prod, ref = z.ZxcLib(z.PRODUCT_SO), z.ZxcLib(z.REF_SO)
n = bad = 0
cases = [("tiny", 1), ("small", 8), ("small", 9), ("small", 37), ("text", 1023), ("text", 1024), ("text", 4097), ("random", 5000),
         ("zeros", 9000), ("period7", 20000), ("runs", 30000), ("binrec", 40000), ("text", 70000), ("silesia", 140000)]
dicts = [(None, None), (GC_DICT, None)] + [(c, h) for c, h in list(golden_dicts().values())[:1]]
if os.environ.get("ZXC_SANITIZE_QUICK"):  # racecheck / synccheck are slow: a handful of cases
    cases = [("small", 37), ("text", 4097), ("runs", 30000), ("silesia", 70000)]
for kind, size in cases:
    data = make_case(kind, size)
    for level in (1, 3, 5, 6, 7):
        for bs in (4096, 65536, 0):
            for d, h in dicts:
                if d is not None and (level not in (5, 6, 7) or bs == 0):
                    continue
                a = ref.compress(data, level=level, block_size=bs, checksum=1, seekable=1, dict=d, dict_huf=h)
                b = prod.compress(data, level=level, block_size=bs, checksum=1, seekable=1, dict=d, dict_huf=h)
                n += 1
                if isinstance(b, int) or a.size != b.size or not np.array_equal(a, b):
                    bad += 1
                    print("MISMATCH", kind, size, level, bs, d is not None, h is not None)
print("sanitize_encode: ran", n, "encodes, mismatches:", bad)
