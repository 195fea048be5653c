"""Debug probe (GPU box): levels 6-7 frames from the product vs the reference, case by case.
    python tests/enc_opt_probe.py [quick]"""
import sys
import time

import numpy as np

sys.path.insert(0, __file__.rsplit("/", 1)[0])
import zxc_corpus as zc  # noqa: E402
import zxc_ctypes as z  # noqa: E402
from test_oracle import CASES, make_case  # noqa: E402


def block_sizes(frame, bs_hint):
    """(type, comp_size) of each block by walking headers"""
    out, p = [], 16
    while p + 8 <= frame.size:
        t = int(frame[p])
        cs = int(frame[p + 3]) | int(frame[p + 4]) << 8 | int(frame[p + 5]) << 16 | int(frame[p + 6]) << 24
        out.append((t, cs, p))
        if t == 255:
            break
        p += 8 + cs
    return out


def main():
    quick = len(sys.argv) > 1
    prod, ref = z.ZxcLib(z.PRODUCT_SO), z.ZxcLib(z.REF_SO)
    cases = [("small", 37), ("tiny", 1), ("text", 20000), ("period7", 30000), ("random", 20000), ("zeros", 50000),
             ("runs", 60000), ("numeric", 80000), ("binrec", 100000), ("period300", 90000), ("text", 300000)]
    if not quick:
        cases += [("silesia", 1 << 20)]
    bad = 0
    for kind, n in cases:
        data = make_case(kind, n)
        for level in (6, 7):
            for bs, cks in ((65536, 0), (4096, 1), (0, 1)):
                a = ref.compress(data, level=level, block_size=bs, checksum=cks, seekable=1)
                t = time.perf_counter()
                b = prod.compress(data, level=level, block_size=bs, checksum=cks, seekable=1)
                dt = time.perf_counter() - t
                if isinstance(b, int):
                    print(f"FAIL {kind} n={n} L{level} bs={bs}: error {z.ERR.get(b, b)}")
                    bad += 1
                    continue
                same = a.size == b.size and np.array_equal(a, b)
                msg = "ok  " if same else "DIFF"
                extra = ""
                if not same:
                    bad += 1
                    m = min(a.size, b.size)
                    first = int(np.argmax(a[:m] != b[:m])) if m and (a[:m] != b[:m]).any() else m
                    ba, bb = block_sizes(a, bs), block_sizes(b, bs)
                    blk = max([i for i, (_, _, p) in enumerate(ba) if p <= first] or [0])
                    pa = ba[blk][2]
                    hdr_a = a[pa:pa + 28].tobytes().hex()
                    hdr_b = b[pa:pa + 28].tobytes().hex()
                    extra = f" sizes {a.size}/{b.size} first diff {first} in block {blk} (@{pa}, rel {first - pa})\n      ref {hdr_a}\n      got {hdr_b}"
                    r, out = ref.decompress(b, data.size, checksum=cks)
                    extra += f"\n      reference decodes ours: {r == data.size and np.array_equal(out, data)} ({r})"
                print(f"{msg} {kind} n={n} L{level} bs={bs} cks={cks} {dt * 1e3:.0f} ms{extra}", flush=True)
    print("bad:", bad)


if __name__ == "__main__":
    main()
