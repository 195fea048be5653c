"""zxc_b200/csrc/zxc_hufenc.h (package-merge, nudge, PivCo sizing -- the level 6-7 entropy-stage
decisions the encode kernel makes on the device) compiled as host C and pinned against the
UNMODIFIED reference's internals (oracle/_ref/libzxc_ref_internals.so; reference
src/lib/zxc_huffman.c:172-311, :803-945, :1219-1249)."""
import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST_SO = os.path.join(ROOT, "oracle", "libzxc_hufenc_host.so")
REFI_SO = os.path.join(ROOT, "oracle", "_ref", "libzxc_ref_internals.so")

pytestmark = pytest.mark.skipif(not (os.path.exists(HOST_SO) and os.path.exists(REFI_SO)),
                                reason="oracle/ not built (python -c 'import __graft_entry__ as g; g.build()')")


@pytest.fixture(scope="module")
def libs():
    h, r = C.CDLL(HOST_SO), C.CDLL(REFI_SO)
    for lib, pre in ((h, "zxhh"), (r, "zxri")):
        for name in ("build_code_lengths", "nudge_code_lengths"):
            f = getattr(lib, f"{pre}_{name}")
            f.restype = C.c_int
            f.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        f = getattr(lib, f"{pre}_calc_size")
        f.restype = C.c_uint64
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    return h, r


def histograms():
    rng = np.random.default_rng(1234)
    out = []
    for n_sym in (1, 2, 3, 4, 5, 7, 16, 17, 33, 64, 65, 100, 128, 129, 200, 255, 256):
        for kind in range(6):
            f = np.zeros(256, np.uint32)
            syms = rng.choice(256, n_sym, replace=False)
            if kind == 0:
                f[syms] = rng.integers(1, 50, n_sym)
            elif kind == 1:
                f[syms] = (rng.pareto(1.2, n_sym) * 40 + 1).astype(np.uint32)
            elif kind == 2:
                f[syms] = (2.0 ** rng.uniform(0, 17, n_sym)).astype(np.uint32)
            elif kind == 3:
                f[syms] = 1
            elif kind == 4:
                f[syms] = np.sort(rng.geometric(0.02, n_sym)).astype(np.uint32)
            else:
                f[syms] = (np.arange(n_sym) ** 2 + 1).astype(np.uint32)
            out.append(f)
    # text-like and binary-like byte histograms
    text = rng.choice(np.frombuffer(b"etaoin shrdlucmfwypvbgkqjxz ETAOIN.,;\n0123456789", np.uint8), 40000,
                      p=None)
    out.append(np.bincount(text, minlength=256).astype(np.uint32))
    out.append(np.bincount((rng.normal(128, 20, 60000).clip(0, 255)).astype(np.uint8), minlength=256).astype(np.uint32))
    out.append(np.bincount((rng.exponential(12, 65536).clip(0, 255)).astype(np.uint8), minlength=256).astype(np.uint32))
    return out


@pytest.mark.parametrize("cap", [8, 9, 11])
def test_package_merge_matches_reference(libs, cap):
    h, r = libs
    for f in histograms():
        a, b = np.zeros(256, np.uint8), np.zeros(256, np.uint8)
        ra = h.zxhh_build_code_lengths(f.ctypes.data, a.ctypes.data, cap)
        rb = r.zxri_build_code_lengths(f.ctypes.data, b.ctypes.data, cap)
        assert (ra == 0) == (rb == 0)
        assert np.array_equal(a, b)


@pytest.mark.parametrize("cap", [8, 11])
def test_nudge_and_size_match_reference(libs, cap):
    h, r = libs
    adopted = 0
    for f in histograms():
        base = np.zeros(256, np.uint8)
        assert r.zxri_build_code_lengths(f.ctypes.data, base.ctypes.data, cap) == 0
        a, b = base.copy(), base.copy()
        ra = h.zxhh_nudge_code_lengths(f.ctypes.data, a.ctypes.data, cap)
        rb = r.zxri_nudge_code_lengths(f.ctypes.data, b.ctypes.data, cap)
        assert ra == rb
        assert np.array_equal(a, b)
        adopted += rb
        for hdr in (0, 1):
            assert h.zxhh_calc_size(f.ctypes.data, a.ctypes.data, hdr) == r.zxri_calc_size(f.ctypes.data, b.ctypes.data, hdr)
            assert h.zxhh_calc_size(f.ctypes.data, base.ctypes.data, hdr) == r.zxri_calc_size(f.ctypes.data, base.ctypes.data, hdr)
    assert adopted > 5  # the comparison exercised the adoption path, not only rejections


def test_calc_size_rejects_what_the_reference_rejects(libs):
    h, r = libs
    f = np.zeros(256, np.uint32)
    f[:4] = [5, 3, 2, 1]
    cl = np.zeros(256, np.uint8)
    cl[:3] = [1, 2, 2]  # symbol 3 present but uncoded
    assert h.zxhh_calc_size(f.ctypes.data, cl.ctypes.data, 1) == r.zxri_calc_size(f.ctypes.data, cl.ctypes.data, 1) == 2 ** 64 - 1
    cl[:4] = [1, 2, 3, 4]  # Kraft-incomplete
    assert h.zxhh_calc_size(f.ctypes.data, cl.ctypes.data, 1) == r.zxri_calc_size(f.ctypes.data, cl.ctypes.data, 1) == 2 ** 64 - 1


def test_randomised_histograms_match_reference(libs):
    """wide sweep: alphabet sizes 1..256, flat / heavy-tailed / power-of-two / tie-rich weights up to 2^21"""
    h, r = libs
    rng = np.random.default_rng(99)
    adopted = 0
    for t in range(700):
        n_sym = int(rng.integers(1, 257))
        f = np.zeros(256, np.uint32)
        syms = rng.choice(256, n_sym, replace=False)
        kind = t % 7
        if kind == 0:
            f[syms] = rng.integers(1, 4, n_sym)
        elif kind == 1:
            f[syms] = (rng.pareto(0.8, n_sym) * 100 + 1).clip(1, 2_000_000).astype(np.uint32)
        elif kind == 2:
            f[syms] = (2.0 ** rng.uniform(0, 21, n_sym)).astype(np.uint32)
        elif kind == 3:
            f[syms] = int(rng.integers(1, 1000))
        elif kind == 4:
            f[syms] = np.sort(rng.geometric(0.001, n_sym)).astype(np.uint32)
        elif kind == 5:
            f[syms] = rng.integers(1, 70000, n_sym)
        else:
            f[syms] = 1
            f[syms[:max(1, n_sym // 8)]] = rng.integers(1000, 100000, max(1, n_sym // 8))
        for cap in (8, 11):
            base = np.zeros(256, np.uint8)
            if r.zxri_build_code_lengths(f.ctypes.data, base.ctypes.data, cap) != 0:
                continue
            a = np.zeros(256, np.uint8)
            assert h.zxhh_build_code_lengths(f.ctypes.data, a.ctypes.data, cap) == 0
            assert np.array_equal(a, base), (t, cap)
            b = base.copy()
            ra = h.zxhh_nudge_code_lengths(f.ctypes.data, a.ctypes.data, cap)
            rb = r.zxri_nudge_code_lengths(f.ctypes.data, b.ctypes.data, cap)
            assert ra == rb and np.array_equal(a, b), (t, cap, n_sym, kind)
            adopted += rb
            assert h.zxhh_calc_size(f.ctypes.data, a.ctypes.data, 1) == r.zxri_calc_size(f.ctypes.data, b.ctypes.data, 1)
    assert adopted > 100


def test_package_merge_wide_sweep(libs):
    """The prefix-count package-merge (zxc_hufenc.h) against the reference's item-tree one: 4000 histograms biased
    towards equal weights (where the leaf-before-package rule decides), every cap the encoder uses."""
    h, r = libs
    rng = np.random.default_rng(2024)
    for t in range(4000):
        n_sym = int(rng.integers(2, 257)) if t % 5 else int(rng.integers(2, 12))
        f = np.zeros(256, np.uint32)
        syms = rng.choice(256, n_sym, replace=False)
        kind = t % 4
        if kind == 0:
            f[syms] = rng.integers(1, 3, n_sym)                      # almost everything ties
        elif kind == 1:
            f[syms] = (1 << rng.integers(0, 12, n_sym)).astype(np.uint32)  # package sums collide with leaf weights
        elif kind == 2:
            f[syms] = rng.integers(1, 40, n_sym)
        else:
            f[syms] = (rng.pareto(1.2, n_sym) * 20 + 1).clip(1, 1_000_000).astype(np.uint32)
        for cap in (8, 9, 10, 11):
            a, b = np.zeros(256, np.uint8), np.zeros(256, np.uint8)
            ra = h.zxhh_build_code_lengths(f.ctypes.data, a.ctypes.data, cap)
            rb = r.zxri_build_code_lengths(f.ctypes.data, b.ctypes.data, cap)
            assert ra == rb and np.array_equal(a, b), (t, cap, n_sym, kind)


def test_block_decomposition_matches_left_to_right_greedy(libs):
    """zxh_cost_add splits an index interval into maximal aligned power-of-two blocks from its two ends; the plain
    statement of the same split (zxc_huffman.c:343-431 walks it left to right: at x take the largest aligned block
    that fits) must give the same sums for every interval of a 2^7 code space, grouped or not."""
    h, _ = libs
    h.zxhh_cost_add.restype = None
    h.zxhh_cost_add.argtypes = [C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(5)
    mass = rng.integers(1, 1000, 300).astype(np.uint64)
    pf = np.concatenate([[0], np.cumsum(mass)]).astype(np.uint64)

    def weight(length, d):
        return length + 1 - d + (24 if d > 6 else 0)

    def greedy(lc, g, at, c, first):
        bits = (lc + g) * int(pf[first + c] - pf[first])
        work, x, i, end = 0, at, first, at + c
        while x < end:
            low = (x & -x) if x else 1 << 30
            fit = 1 << ((end - x).bit_length() - 1)
            sz = min(low, fit)
            work += int(pf[i + sz] - pf[i]) * weight(lc + g, (sz.bit_length() - 1) + g)
            x += sz
            i += sz
        return bits, work

    bits, work = C.c_uint64(), C.c_uint64()
    checked = 0
    for lc, g in ((7, 0), (5, 2), (6, 1), (3, 0)):
        n = 1 << lc
        for at in range(0, n):
            for c in range(0, n - at + 1):
                first = int(rng.integers(0, 300 - c)) if c < 300 else 0
                if first + c > 300:
                    continue
                h.zxhh_cost_add(lc, g, at, c, pf.ctypes.data, first, C.byref(bits), C.byref(work))
                assert (bits.value, work.value) == greedy(lc, g, at, c, first), (lc, g, at, c)
                checked += 1
    assert checked > 10000
