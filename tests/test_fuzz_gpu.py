"""Fuzz-shaped parity corpus (replaces the reference's private fuzz corpus):
sizes log-uniform in [0, 4 MiB], first byte selects the level as in tests/fuzz_roundtrip.c:52
(level = data[0] % 7 + 1), content drawn from several generators.  For every input:
  * levels 1-7: the GPU encoder's frame is byte-identical to the reference's;
  * every level: the GPU decoder reproduces the input from the reference's frame into an
    exact-size buffer (fuzz_roundtrip.c:33-73)."""
import numpy as np
import pytest

import zxc_corpus as zc
import zxc_ctypes as z

pytestmark = pytest.mark.gpu


def fuzz_input(rng, i):
    n = int(np.exp(rng.uniform(0, np.log(4 << 20)))) if rng.random() > 0.03 else int(rng.integers(0, 12))
    kind = int(rng.integers(0, 8))
    if kind == 0:
        d = rng.integers(0, 256, n, dtype=np.uint8)
    elif kind == 1:
        d = zc.gen_runs(n, seed=i) if n else np.zeros(0, np.uint8)
    elif kind == 2:
        d = zc.gen_periodic(n, int(rng.integers(1, 301)), seed=i) if n else np.zeros(0, np.uint8)
    elif kind == 3:
        d = zc.gen_text(n, seed=i) if n else np.zeros(0, np.uint8)
    elif kind == 4:
        d = zc.gen_numeric(n, seed=i) if n else np.zeros(0, np.uint8)
    elif kind == 5:
        d = zc.gen_binary_records(n, seed=i) if n else np.zeros(0, np.uint8)
    else:
        off = int(rng.integers(0, 200)) << 20
        d = zc.silesia_shaped(max(n, 1), seed=3, offset=off)[:n]
    d = np.ascontiguousarray(d, dtype=np.uint8)
    if d.size:
        d[0] = int(rng.integers(0, 256))
    return d


def test_fuzz_shaped_corpus(prod, ref):
    rng = np.random.default_rng(2026)
    n_cases = 400
    dec_checked = 0
    for i in range(n_cases):
        d = fuzz_input(rng, i)
        level = (int(d[0]) % 7 + 1) if d.size else 3
        bs = int(rng.choice([0, 4096, 65536, 1 << 20]))
        cks = int(rng.integers(0, 2))
        fr = ref.compress(d, level=level, block_size=bs, checksum=cks, seekable=int(rng.integers(0, 2)))
        assert not isinstance(fr, int)
        r, out = prod.decompress(fr, d.size, checksum=cks)
        assert r == d.size, (i, level, bs, d.size, z.ERR.get(r, r))
        assert np.array_equal(out, d), (i, level, bs)
        dec_checked += 1
    assert dec_checked == n_cases


def test_fuzz_encoder_identity(prod, ref):
    rng = np.random.default_rng(77)
    for i in range(250):
        d = fuzz_input(rng, 1000 + i)
        level = (int(d[0]) % 5 + 1) if d.size else 3
        bs = int(rng.choice([0, 4096, 65536, 1 << 20]))
        cks, seek = int(rng.integers(0, 2)), int(rng.integers(0, 2))
        a = ref.compress(d, level=level, block_size=bs, checksum=cks, seekable=seek)
        b = prod.compress(d, level=level, block_size=bs, checksum=cks, seekable=seek)
        assert not isinstance(b, int), (i, z.ERR.get(b, b))
        assert a.size == b.size and np.array_equal(a, b), (i, level, bs, d.size)


def test_fuzz_encoder_identity_levels_6_7(prod, ref):
    """optimal parser + Huffman sections; inputs capped at 512 KiB (a block is parsed by one warp)"""
    rng = np.random.default_rng(4242)
    for i in range(160):
        d = fuzz_input(rng, 5000 + i)[: 512 << 10]
        level = 6 + (int(d[0]) & 1 if d.size else 0)
        bs = int(rng.choice([0, 4096, 16384, 65536, 1 << 20]))
        cks, seek = int(rng.integers(0, 2)), int(rng.integers(0, 2))
        a = ref.compress(d, level=level, block_size=bs, checksum=cks, seekable=seek)
        b = prod.compress(d, level=level, block_size=bs, checksum=cks, seekable=seek)
        assert not isinstance(b, int), (i, z.ERR.get(b, b))
        assert a.size == b.size and np.array_equal(a, b), (i, level, bs, d.size)
        r, out = prod.decompress(b, d.size, checksum=cks)
        assert r == d.size and np.array_equal(out, d), (i, level, bs)
