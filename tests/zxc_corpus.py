"""Synthetic inputs for tests and bench (no Silesia, no network in this sandbox).

* silesia_shaped(n, seed): oracle/zxc_corpus.c generator, multi-threaded, any size.
* small pattern generators modelled on the reference's tests/test_common.c:35-135
  (random, lorem-like text, numeric, binary records, offset-8 / offset-16 repeats).
* compress_ref_mt(): compress with the UNMODIFIED reference (oracle/_ref) in parallel
  slices and stitch the slices into one seekable frame.  Blocks are independent by
  format (docs/FORMAT.md:651-662), so the stitched frame is
  byte-identical to a single zxc_compress call -- tests/test_oracle.py checks that.
"""
import ctypes as C
import os
import struct
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CORPUS_SO = os.path.join(ROOT, "oracle", "libzxc_corpus.so")
CHUNK = 1 << 20
_lib = None


def _corpus():
    global _lib
    if _lib is None:
        _lib = C.CDLL(CORPUS_SO)
        _lib.zxcorp_fill.restype = C.c_int
        _lib.zxcorp_fill.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64]
        _lib.zxcorp_init()
    return _lib


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def silesia_shaped(n, seed=1, out=None, threads=None, offset=0):
    """n bytes of the Silesia-shaped stream starting at `offset` (1 MiB aligned)."""
    lib = _corpus()
    if out is None:
        out = np.empty(n, dtype=np.uint8)
    threads = threads or host_threads()
    per = 4 * CHUNK  # small work items: the data classes differ a lot in cost
    base = out.ctypes.data

    def work(i):
        lo = i * per
        if lo >= n:
            return 0
        ln = min(per, n - lo)
        return lib.zxcorp_fill(base + lo, offset + lo, ln, seed)

    with ThreadPoolExecutor(threads) as ex:
        rc = list(ex.map(work, range((n + per - 1) // per)))
    assert all(r == 0 for r in rc)
    return out


def records(n, record_size=4096, seed=7):
    """n fixed-size JSON-ish records (BASELINE.json configs[3]), generated in parallel."""
    lib = _corpus()
    lib.zxcorp_records.restype = C.c_int
    lib.zxcorp_records.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint64]
    out = np.empty(n * record_size, np.uint8)
    th = host_threads()
    per = (n + th - 1) // th

    def work(i):
        lo = i * per
        if lo < n:
            lib.zxcorp_records(out.ctypes.data + lo * record_size, lo, min(per, n - lo), record_size, seed)

    with ThreadPoolExecutor(th) as ex:
        list(ex.map(work, range(th)))
    return out


def train_dict_ref(ref, data, record_size=4096, n_samples=4096, cap=16384):
    """dictionary bytes from the reference's own trainer (zxc_train_dict) over the first n_samples records"""
    ref.lib.zxc_train_dict.restype = C.c_int64
    ref.lib.zxc_train_dict.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
    ns = min(n_samples, data.size // record_size)
    ptrs = (C.c_void_p * ns)(*[data.ctypes.data + i * record_size for i in range(ns)])
    sizes = (C.c_size_t * ns)(*([record_size] * ns))
    dbuf = np.zeros(cap, np.uint8)
    dsz = ref.lib.zxc_train_dict(ptrs, sizes, ns, dbuf.ctypes.data, dbuf.size)
    assert dsz > 0, dsz
    return dbuf[:dsz].tobytes()


# ---- small generators (reference tests/test_common.c:35-135 equivalents) -------
def gen_random(n, seed=42):
    return np.random.default_rng(seed).integers(0, 256, n, dtype=np.uint8)


def gen_text(n, seed=1):
    words = ("lorem ipsum dolor sit amet consectetur adipiscing elit sed do eiusmod tempor incididunt ut "
             "labore et dolore magna aliqua ut enim ad minim veniam quis nostrud exercitation ullamco").split()
    rng = np.random.default_rng(seed)
    idx = rng.zipf(1.3, size=n // 3 + 16) % len(words)
    s = " ".join(words[i] for i in idx).encode()
    return np.frombuffer(s[:n].ljust(n, b" "), dtype=np.uint8).copy()


def gen_numeric(n, seed=3):
    rng = np.random.default_rng(seed)
    v = np.cumsum(rng.integers(0, 50, n // 4 + 1)).astype("<u4")
    return v.view(np.uint8)[:n].copy()


def gen_binary_records(n, seed=4):
    rng = np.random.default_rng(seed)
    rec = rng.integers(0, 256, 64, dtype=np.uint8)
    out = np.tile(rec, n // 64 + 1)[:n].copy()
    pos = rng.integers(0, n, n // 20)
    out[pos] = rng.integers(0, 256, pos.size, dtype=np.uint8)
    return out


def gen_periodic(n, period, seed=5):
    rng = np.random.default_rng(seed)
    return np.tile(rng.integers(0, 256, period, dtype=np.uint8), n // period + 1)[:n].copy()


def gen_runs(n, seed=6):
    rng = np.random.default_rng(seed)
    out = np.empty(n, dtype=np.uint8)
    p = 0
    while p < n:
        L = int(rng.integers(1, 400))
        out[p:p + L] = rng.integers(0, 256)
        p += L
    return out


# ---- reference-compressed frames, in parallel ---------------------------------
def compress_ref_mt(ref, data, level=3, block_size=65536, checksum=0, threads=None, slice_bytes=None):
    """Seekable frame of `data` produced by the reference library `ref` (ZxcLib)."""
    from zxc_ctypes import CompressOpts
    data = np.ascontiguousarray(data).view(np.uint8).reshape(-1)
    n = data.size
    threads = threads or host_threads()
    if slice_bytes is None:
        slice_bytes = max(block_size, ((n + threads * 4 - 1) // (threads * 4) + block_size - 1) // block_size * block_size)
        slice_bytes = min(slice_bytes, 64 << 20)
    assert slice_bytes % block_size == 0
    nsl = max(1, (n + slice_bytes - 1) // slice_bytes)
    L = ref.lib

    def work(i):
        lo = i * slice_bytes
        ln = min(slice_bytes, n - lo)
        cap = int(L.zxc_compress_bound(ln))
        out = np.empty(cap, dtype=np.uint8)
        o = CompressOpts(level=level, block_size=block_size, checksum_enabled=checksum, seekable=1)
        r = L.zxc_compress(data.ctypes.data + lo, ln, out.ctypes.data, cap, C.byref(o))
        assert r > 0, r
        return out[:r]

    with ThreadPoolExecutor(threads) as ex:
        parts = list(ex.map(work, range(nsl)))
    if nsl == 1:
        return parts[0].copy()
    return stitch_frames(ref, parts, n, block_size, checksum)


def stitch_frames(ref, parts, total, block_size, checksum):
    """Concatenate the data blocks of seekable frames; rebuild EOF, SEK table, footer."""
    hdr = parts[0][:16]
    bodies, sizes = [], []
    ghash = 0
    for fr in parts:
        tot = struct.unpack("<Q", fr[-12:-4].tobytes())[0]
        nb = (tot + block_size - 1) // block_size
        sek = fr.size - 12 - (8 + 4 * nb)
        ent = np.frombuffer(fr[sek + 8: sek + 8 + 4 * nb].tobytes(), dtype="<u4")
        end = 16 + int(ent.sum())
        assert end + 8 == sek, (end, sek)
        bodies.append(fr[16:end])
        sizes.append(ent)
        if checksum:
            offs = 16 + np.cumsum(ent.astype(np.int64))
            for e in offs:
                bh = struct.unpack("<I", fr[e - 4:e].tobytes())[0]
                ghash = (((ghash << 1) | (ghash >> 31)) & 0xFFFFFFFF) ^ bh
    sizes = np.concatenate(sizes).astype("<u4")
    # EOF header bytes are constant: take them from the first part
    tot0 = struct.unpack("<Q", parts[0][-12:-4].tobytes())[0]
    nb0 = (tot0 + block_size - 1) // block_size
    sek0 = parts[0].size - 12 - (8 + 4 * nb0)
    eof = parts[0][sek0 - 8:sek0]
    tab = np.empty(8 + 4 * sizes.size, dtype=np.uint8)
    wst = ref.lib.zxc_write_seek_table
    wst.restype = C.c_int64
    wst.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint32]
    r = wst(tab.ctypes.data, tab.size, sizes.ctypes.data, sizes.size)
    assert r == tab.size, r
    footer = np.frombuffer(struct.pack("<QI", total, ghash if checksum else 0), dtype=np.uint8)
    return np.concatenate([hdr] + bodies + [eof, tab, footer])
