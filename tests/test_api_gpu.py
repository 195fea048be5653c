"""GPU tests of the drop-in boundary beyond the one-shot buffer calls: reader- and FILE*-backed seekable
handles (include/zxc_seekable.h:96-140), zxc_seekable_set_dict semantics (src/lib/zxc_seekable.c:1144-1174),
the caller-workspace contexts, the stateless buffer API under concurrent callers (docs/API.md:1528-1538),
and BASELINE.json configs[3]'s shape (trained dictionary, 4 KiB records, level 5) through the seekable API."""
import ctypes as C
import os
import threading

import numpy as np
import pytest

import zxc_corpus as zc
import zxc_ctypes as z

pytestmark = pytest.mark.gpu

READ_AT = C.CFUNCTYPE(C.c_int64, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint64)


class Reader(C.Structure):
    _fields_ = [("read_at", READ_AT), ("ctx", C.c_void_p), ("size", C.c_uint64)]


def make_reader(frame, calls):
    buf = np.ascontiguousarray(frame)
    lock = threading.Lock()

    def read_at(ctx, dst, ln, off):  # thread-safe: positional, no shared cursor
        if off + ln > buf.size:
            return -11
        C.memmove(dst, buf.ctypes.data + off, ln)
        with lock:
            calls.append((off, ln))
        return ln
    cb = READ_AT(read_at)
    return Reader(cb, None, buf.size), (cb, buf)


def test_reader_backed_handle(prod, ref):
    data = zc.silesia_shaped(80 << 20, seed=4)  # > 32 MiB: the staged reader pipeline
    frame = zc.compress_ref_mt(ref, data, level=3, block_size=65536)
    calls = []
    rd, keep = make_reader(frame, calls)
    L = prod.lib
    L.zxc_seekable_open_reader.restype = C.c_void_p
    L.zxc_seekable_open_reader.argtypes = [C.c_void_p]
    h = L.zxc_seekable_open_reader(C.byref(rd))
    assert h
    assert L.zxc_seekable_get_decompressed_size(h) == data.size
    assert L.zxc_seekable_get_num_blocks(h) == data.size // 65536
    out = np.zeros(data.size, np.uint8)
    n0 = len(calls)
    assert L.zxc_seekable_decompress_range_mt(h, out.ctypes.data, out.size, 0, out.size, 8) == data.size
    assert np.array_equal(out, data)
    nblk = data.size // 65536
    body = frame.size - 16 - 8 - (8 + 4 * nblk) - 12  # header, EOF block, SEK block, footer
    assert len(calls) > n0 and sum(ln for _, ln in calls[n0:]) >= body  # the blocks came through read_at
    for off, ln in ((0, 1), (65535, 2), (123457, 700001), (data.size - 5, 5)):
        o = np.zeros(ln, np.uint8)
        assert L.zxc_seekable_decompress_range(h, o.ctypes.data, ln, off, ln) == ln
        assert np.array_equal(o, data[off:off + ln]), (off, ln)
    L.zxc_seekable_free(h)
    # the reference through the same reader object gives the same bytes
    R = ref.lib
    R.zxc_seekable_open_reader.restype = C.c_void_p
    R.zxc_seekable_open_reader.argtypes = [C.c_void_p]
    hr = R.zxc_seekable_open_reader(C.byref(rd))
    o2 = np.zeros(1 << 20, np.uint8)
    assert R.zxc_seekable_decompress_range_mt(hr, o2.ctypes.data, o2.size, 4097, o2.size, 4) == o2.size
    assert np.array_equal(o2, data[4097:4097 + o2.size])
    R.zxc_seekable_free(hr)


def test_file_backed_handle(prod, ref, tmp_path):
    libc = C.CDLL(None)
    libc.fopen.restype = C.c_void_p
    libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
    libc.fclose.argtypes = [C.c_void_p]
    data = zc.silesia_shaped(6 << 20, seed=8)
    frame = ref.compress(data, level=5, block_size=32768, seekable=1)
    p = tmp_path / "a.zxc"
    p.write_bytes(frame.tobytes())
    f = libc.fopen(str(p).encode(), b"rb")
    assert f
    L = prod.lib
    L.zxc_seekable_open_file.restype = C.c_void_p
    L.zxc_seekable_open_file.argtypes = [C.c_void_p]
    h = L.zxc_seekable_open_file(f)
    assert h
    out = np.zeros(data.size, np.uint8)
    assert L.zxc_seekable_decompress_range_mt(h, out.ctypes.data, out.size, 0, out.size, 4) == data.size
    assert np.array_equal(out, data)
    o = np.zeros(99999, np.uint8)
    assert L.zxc_seekable_decompress_range(h, o.ctypes.data, o.size, 1234567, o.size) == o.size
    assert np.array_equal(o, data[1234567:1234567 + o.size])
    L.zxc_seekable_free(h)
    libc.fclose(f)


def test_set_dict_semantics_match_reference(prod, ref):
    rec = zc.records(2048)
    d = zc.train_dict_ref(ref, rec)
    frame = ref.compress(rec, level=5, block_size=4096, seekable=1, dict=d)
    other = bytes(reversed(d))
    out = np.zeros(rec.size, np.uint8)
    for lib in (ref.lib, prod.lib):
        h = lib.zxc_seekable_open(frame.ctypes.data, frame.size)
        assert h
        assert lib.zxc_seekable_decompress_range(h, out.ctypes.data, out.size, 0, out.size) == -15  # DICT_REQUIRED
        assert lib.zxc_seekable_set_dict(h, None, 0, None) == -12       # NULL_INPUT (zxc_seekable.c:1146)
        assert lib.zxc_seekable_set_dict(h, d, 0, None) == -12
        assert lib.zxc_seekable_set_dict(h, d, (64 << 10) + 1, None) == -17  # DICT_TOO_LARGE
        assert lib.zxc_seekable_set_dict(h, other, len(other), None) == -16  # DICT_MISMATCH
        assert lib.zxc_seekable_set_dict(h, d, len(d), None) == 0
        # rejected calls leave the installed dictionary alone (validation comes first, :1147-1150)
        assert lib.zxc_seekable_set_dict(h, other, len(other), None) == -16
        assert lib.zxc_seekable_set_dict(h, None, 0, None) == -12
        out[:] = 0
        assert lib.zxc_seekable_decompress_range_mt(h, out.ctypes.data, out.size, 0, out.size, 4) == rec.size
        assert np.array_equal(out, rec)
        lib.zxc_seekable_free(h)


def test_dictionary_records_config4_shape(prod, ref):
    """configs[3] at reduced count: 64 Ki records x 4 KiB, 16 KiB-class trained dictionary, level 5"""
    n = 1 << 16
    rec = zc.records(n)
    d = zc.train_dict_ref(ref, rec)
    assert 8192 <= len(d) <= 16384
    frame = prod.compress(rec, level=5, block_size=4096, seekable=1, dict=d)
    sub = rec[: 4096 * 4096]
    assert np.array_equal(ref.compress(sub, level=5, block_size=4096, seekable=1, dict=d),
                          prod.compress(sub, level=5, block_size=4096, seekable=1, dict=d))
    L = prod.lib
    h = L.zxc_seekable_open(frame.ctypes.data, frame.size)
    assert h and L.zxc_seekable_get_num_blocks(h) == n
    assert L.zxc_seekable_set_dict(h, d, len(d), None) == 0
    out = np.zeros(rec.size, np.uint8)
    assert L.zxc_seekable_decompress_range_mt(h, out.ctypes.data, out.size, 0, out.size, 16) == rec.size
    assert np.array_equal(out, rec)
    rng = np.random.default_rng(5)
    for _ in range(20):  # single records and ragged spans
        a = int(rng.integers(0, rec.size - 1))
        ln = int(min(rec.size - a, rng.integers(1, 20000)))
        o = np.zeros(ln, np.uint8)
        assert L.zxc_seekable_decompress_range(h, o.ctypes.data, ln, a, ln) == ln
        assert np.array_equal(o, rec[a:a + ln])
    L.zxc_seekable_free(h)
    # the reference decodes the GPU-encoded frame with the same dictionary
    r, o = ref.decompress(frame, rec.size, dict=d)
    assert r == rec.size and np.array_equal(o, rec)


def test_static_workspace_contexts(prod, ref):
    L = prod.lib
    L.zxc_static_dctx_workspace_size.restype = C.c_size_t
    L.zxc_static_dctx_workspace_size.argtypes = [C.c_size_t]
    L.zxc_init_static_dctx.restype = C.c_void_p
    L.zxc_init_static_dctx.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t]
    L.zxc_static_cctx_workspace_size.restype = C.c_size_t
    L.zxc_static_cctx_workspace_size.argtypes = [C.c_size_t, C.c_int]
    L.zxc_init_static_cctx.restype = C.c_void_p
    L.zxc_init_static_cctx.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
    data = zc.silesia_shaped(1 << 20, seed=12)[:700001]
    assert L.zxc_static_dctx_workspace_size(12345) == 0  # not a valid block size
    need_d = L.zxc_static_dctx_workspace_size(65536)
    need_c = L.zxc_static_cctx_workspace_size(65536, 3)
    assert need_d > 0 and need_c > 0
    ws_d = np.zeros(need_d + 64, np.uint8)
    ws_c = np.zeros(need_c + 64, np.uint8)
    base_d = (ws_d.ctypes.data + 63) & ~63
    base_c = (ws_c.ctypes.data + 63) & ~63
    assert L.zxc_init_static_dctx(base_d, need_d - 1, 65536) is None
    dctx = L.zxc_init_static_dctx(base_d, need_d, 65536)
    o = z.CompressOpts(level=3, block_size=65536, checksum_enabled=1)
    cctx = L.zxc_init_static_cctx(base_c, need_c, C.byref(o))
    assert dctx and cctx
    cap = int(L.zxc_compress_bound(data.size))
    fr = np.zeros(cap, np.uint8)
    r = L.zxc_compress_cctx(cctx, data.ctypes.data, data.size, fr.ctypes.data, cap, C.byref(o))
    assert r > 0
    want = ref.compress(data, level=3, block_size=65536, checksum=1)
    assert r == want.size and np.array_equal(fr[:r], want)
    out = np.zeros(data.size, np.uint8)
    do = z.DecompressOpts(checksum_enabled=1)
    for _ in range(2):  # the context is reusable
        assert L.zxc_decompress_dctx(dctx, fr.ctypes.data, r, out.ctypes.data, out.size, C.byref(do)) == data.size
        assert np.array_equal(out, data)
    L.zxc_free_dctx(dctx)  # releases device resources, not the caller's workspace
    L.zxc_free_cctx(cctx)


def test_buffer_api_is_callable_concurrently(prod, ref):
    """zxc_compress / zxc_decompress are stateless and thread-safe (docs/API.md:1528-1538): 8 threads, each with
    its own input, hammer both calls; every frame must equal the reference's and decode back."""
    n_threads, rounds = 8, 6
    errs = []

    def work(t):
        try:
            rng = np.random.default_rng(100 + t)
            for k in range(rounds):
                n = int(rng.integers(1, 3 << 20))
                data = zc.silesia_shaped(4 << 20, seed=20 + t)[k * 1000: k * 1000 + n].copy()
                level = int(rng.integers(1, 6))
                bs = int(rng.choice([4096, 65536, 262144]))
                fr = prod.compress(data, level=level, block_size=bs, checksum=k & 1)
                want = ref.compress(data, level=level, block_size=bs, checksum=k & 1)
                assert not isinstance(fr, int), fr
                assert fr.size == want.size and np.array_equal(fr, want), ("encode", t, k)
                r, out = prod.decompress(want, n, checksum=k & 1)
                assert r == n and np.array_equal(out, data), ("decode", t, k, r)
        except Exception as e:  # noqa: BLE001
            errs.append(repr(e))
    th = [threading.Thread(target=work, args=(t,)) for t in range(n_threads)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs


def test_large_unaligned_ranges_through_the_pipelines(prod, ref):
    """Ranges of >= 32 MiB decoded take the overlapped routes: memory-backed handles with pageable
    buffers (staged, clipped at both ends), page-locked block-aligned ranges (straight into the caller's memory), and a
    reader-backed handle whose read_at fills the pinned input slots."""
    import torch
    data = zc.silesia_shaped(160 << 20, seed=6)
    bs = 65536
    frame = zc.compress_ref_mt(ref, data, level=3, block_size=bs)
    L = prod.lib
    h = L.zxc_seekable_open(frame.ctypes.data, frame.size)
    assert h
    rng = np.random.default_rng(8)
    spans = [(1, data.size - 2), (bs * 3 + 17, (100 << 20) + 12345), (data.size - (40 << 20) - 7, (40 << 20) + 7)]
    spans += [(int(rng.integers(0, 60 << 20)), int(rng.integers(33 << 20, 90 << 20))) for _ in range(3)]
    for off, ln in spans:
        out = np.zeros(ln, np.uint8)
        r = L.zxc_seekable_decompress_range_mt(h, out.ctypes.data, ln, off, ln, 8)
        assert r == ln, (off, ln, z.ERR.get(r, r))
        assert np.array_equal(out, data[off:off + ln]), (off, ln)
    # page-locked and block-aligned: the direct pipeline
    h_frame = torch.from_numpy(frame).pin_memory()
    hp = L.zxc_seekable_open(h_frame.data_ptr(), h_frame.numel())
    ln = 96 << 20
    pout = torch.zeros(ln, dtype=torch.uint8).pin_memory()
    assert L.zxc_seekable_decompress_range_mt(hp, pout.data_ptr(), ln, 16 * bs, ln, 4) == ln
    assert np.array_equal(pout.numpy(), data[16 * bs:16 * bs + ln])
    # page-locked but ragged: staged with clipping
    assert L.zxc_seekable_decompress_range(hp, pout.data_ptr(), ln - 5, 16 * bs + 3, ln - 5) == ln - 5
    assert np.array_equal(pout.numpy()[:ln - 5], data[16 * bs + 3:16 * bs + 3 + ln - 5])
    L.zxc_seekable_free(hp)
    L.zxc_seekable_free(h)
    # reader-backed, ragged
    calls = []
    rd, keep = make_reader(frame, calls)
    L.zxc_seekable_open_reader.restype = C.c_void_p
    L.zxc_seekable_open_reader.argtypes = [C.c_void_p]
    hr = L.zxc_seekable_open_reader(C.byref(rd))
    off, ln = 5 * bs + 1000, (70 << 20) + 3
    out = np.zeros(ln, np.uint8)
    assert L.zxc_seekable_decompress_range_mt(hr, out.ctypes.data, ln, off, ln, 8) == ln
    assert np.array_equal(out, data[off:off + ln])
    L.zxc_seekable_free(hr)


def test_one_call_over_several_devices(prod, ref):
    """ZXC_B200_DEVICES: one zxc_decompress / zxc_seekable_decompress_range_mt call fans its block stripes out over
    the visible devices (zxc_api.c decode_multi), the reference's fork-join (zxc_seekable.c:999-1108) with devices for
    threads.  Same bytes, same verdicts as the single-device route."""
    import os
    import torch
    if prod.lib.zxc_b200_device_count() < 2:
        pytest.skip("needs two devices")
    data = zc.silesia_shaped(400 << 20, seed=12)
    bs = 65536
    frame = zc.compress_ref_mt(ref, data, level=3, block_size=bs, checksum=1)
    L = prod.lib
    os.environ["ZXC_B200_DEVICES"] = "all"
    try:
        n0 = L.zxc_b200_launch_count()
        r, out = prod.decompress(frame, data.size, checksum=1)  # pageable: staged stripes
        assert r == data.size and np.array_equal(out, data)
        assert L.zxc_b200_launch_count() - n0 >= 2
        h_frame = torch.from_numpy(frame).pin_memory()  # page-locked: direct stripes
        h_out = torch.zeros(data.size, dtype=torch.uint8).pin_memory()
        o = z.DecompressOpts(checksum_enabled=1)
        assert L.zxc_decompress(h_frame.data_ptr(), h_frame.numel(), h_out.data_ptr(), data.size, C.byref(o)) == data.size
        assert np.array_equal(h_out.numpy(), data)
        f2 = frame.copy()  # damage in the last stripe: the reference's code
        f2[int(f2.size * 0.93)] ^= 0x21
        r0, _ = ref.decompress(f2, data.size, checksum=1)
        r1, _ = prod.decompress(f2, data.size, checksum=1)
        assert r0 == r1 < 0, (r0, r1)
        h = L.zxc_seekable_open(frame.ctypes.data, frame.size)
        for off, ln in ((3, data.size - 5), (7 * bs + 11, (300 << 20) + 1)):
            rout = np.zeros(ln, np.uint8)
            assert L.zxc_seekable_decompress_range_mt(h, rout.ctypes.data, ln, off, ln, 8) == ln
            assert np.array_equal(rout, data[off:off + ln]), (off, ln)
        L.zxc_seekable_free(h)
        calls = []
        rd, keep = make_reader(frame, calls)  # reader-backed: read_at called from every stripe's thread
        L.zxc_seekable_open_reader.restype = C.c_void_p
        L.zxc_seekable_open_reader.argtypes = [C.c_void_p]
        hr = L.zxc_seekable_open_reader(C.byref(rd))
        off, ln = 2 * bs + 5, (350 << 20) + 9
        rout = np.zeros(ln, np.uint8)
        assert L.zxc_seekable_decompress_range_mt(hr, rout.ctypes.data, ln, off, ln, 8) == ln
        assert np.array_equal(rout, data[off:off + ln])
        L.zxc_seekable_free(hr)
    finally:
        del os.environ["ZXC_B200_DEVICES"]
