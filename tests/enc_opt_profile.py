"""Phase profile of the level 6-7 parser (library built with NVEXTRA=-DZXC_OPT_PROFILE)."""
import sys
import time
sys.path.insert(0, __file__.rsplit("/", 1)[0])
import zxc_ctypes as z  # noqa: E402
from test_oracle import make_case  # noqa: E402
prod = z.ZxcLib(z.PRODUCT_SO)
for kind, n in (("text", 65536), ("silesia", 65536), ("text", 300000), ("silesia", 512 * 1024)):
    data = make_case(kind, n)
    for level in (6, 7):
        prod.compress(data[:1000], level=level)
        t = time.perf_counter()
        fr = prod.compress(data, level=level, block_size=0)
        print(kind, n, "L", level, f"{(time.perf_counter() - t) * 1e3:.0f} ms", fr.size, flush=True)
