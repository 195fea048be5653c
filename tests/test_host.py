"""CPU-only: the drop-in boundary.  libzxc.so.4 loads, exports every symbol include/*.h
declares (and the reference's 67), and the host-side logic (bounds, probes, dict container,
SEK parsing, frame planning, header-level rejects) matches the reference.  No compute calls."""
import ctypes as C
import glob
import json
import os
import re
import subprocess

import numpy as np
import pytest

import zxc_corpus as zc
import zxc_ctypes as z
from conftest import has_cuda

ROOT = z.ROOT
G = os.path.join(ROOT, "tests", "golden")


def declared_symbols():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        src = open(h).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        for m in re.finditer(r"ZXC_EXPORT\s+[^;(]*?\b(zxc_\w+)\s*\(", src):
            names.add(m.group(1))
    return names


def exported(path):
    out = subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True).stdout
    return {l.split()[-1] for l in out.splitlines() if l.strip()}


def test_library_loads_and_exports_declared_symbols(prod):
    decl = declared_symbols()
    exp = exported(z.PRODUCT_SO)
    assert len(decl) >= 67
    assert decl <= exp, sorted(decl - exp)
    for n in decl:
        getattr(prod.lib, n)  # resolvable through the loader
    so = subprocess.run(["readelf", "-d", z.PRODUCT_SO], capture_output=True, text=True).stdout
    assert "libzxc.so.4" in so  # SONAME of the reference (CMakeLists.txt:64-71)


def test_exports_superset_of_reference(prod, ref):
    refsyms = {s for s in exported(z.REF_SO) if s.startswith("zxc_")}
    assert len(refsyms) == 67
    assert refsyms <= exported(z.PRODUCT_SO)


def test_info_and_bounds_match_reference(prod, ref):
    assert prod.lib.zxc_version_string() == ref.lib.zxc_version_string() == b"0.13.3"
    for f in ("zxc_min_level", "zxc_max_level", "zxc_default_level"):
        assert getattr(prod.lib, f)() == getattr(ref.lib, f)()
    for lib in (prod.lib, ref.lib):
        lib.zxc_compress_opts_size.restype = C.c_size_t
        lib.zxc_decompress_opts_size.restype = C.c_size_t
        lib.zxc_seek_table_size.restype = C.c_size_t
        lib.zxc_seek_table_size.argtypes = [C.c_uint32]
    assert prod.lib.zxc_compress_opts_size() == ref.lib.zxc_compress_opts_size() == C.sizeof(z.CompressOpts)
    assert prod.lib.zxc_decompress_opts_size() == ref.lib.zxc_decompress_opts_size() == C.sizeof(z.DecompressOpts)
    for n in (0, 1, 4095, 4096, 4097, 65536, 1 << 20, (1 << 21) + 1, 1 << 33):
        assert prod.lib.zxc_compress_bound(n) == ref.lib.zxc_compress_bound(n)
        assert prod.lib.zxc_compress_block_bound(n) == ref.lib.zxc_compress_block_bound(n)
        assert prod.lib.zxc_decompress_block_bound(n) == ref.lib.zxc_decompress_block_bound(n)
    for n in (0, 1, 1000, 1 << 20):
        assert prod.lib.zxc_seek_table_size(n) == ref.lib.zxc_seek_table_size(n)
    for code in list(range(-19, 1)):
        assert prod.lib.zxc_error_name(code) == ref.lib.zxc_error_name(code)


def test_probes_and_dict_container_match_reference(prod, ref):
    rng = np.random.default_rng(1)
    for p in sorted(glob.glob(os.path.join(G, "valid", "*.zxc")) + glob.glob(os.path.join(G, "invalid", "*.zxc"))
                    + glob.glob(os.path.join(G, "format", "*.zxc"))):
        b = open(p, "rb").read()
        if not b:
            continue
        assert prod.lib.zxc_get_decompressed_size(b, len(b)) == ref.lib.zxc_get_decompressed_size(b, len(b)), p
        assert prod.lib.zxc_get_dict_id(b, len(b)) == ref.lib.zxc_get_dict_id(b, len(b)), p
    for n in (1, 5, 16, 17, 100, 113, 1000, 65535):
        d = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        huf = rng.integers(0, 12, 128, dtype=np.uint8).tobytes()
        assert prod.lib.zxc_dict_id(d, n, None) == ref.lib.zxc_dict_id(d, n, None)
        assert prod.lib.zxc_dict_id(d, n, huf) == ref.lib.zxc_dict_id(d, n, huf)
        for lib in (prod.lib, ref.lib):
            lib.zxc_dict_save.restype = C.c_int64
            lib.zxc_dict_save.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t]
        a = np.zeros(n + 200, np.uint8)
        bb = np.zeros(n + 200, np.uint8)
        ra = prod.lib.zxc_dict_save(d, n, huf, a.ctypes.data, a.size)
        rb = ref.lib.zxc_dict_save(d, n, huf, bb.ctypes.data, bb.size)
        assert ra == rb == n + 144 and np.array_equal(a, bb)
        rc, content, h2, did = prod.dict_load(a[:ra])
        assert rc == 0 and content == d and h2 == huf and did == ref.lib.zxc_dict_id(d, n, huf)
    for p in glob.glob(os.path.join(G, "valid", "*.zxd")):
        b = open(p, "rb").read()
        assert prod.dict_load(b) == ref.dict_load(b)
        prod.lib.zxc_dict_get_id.restype = C.c_uint32
        ref.lib.zxc_dict_get_id.restype = C.c_uint32
        assert prod.lib.zxc_dict_get_id(b, len(b)) == ref.lib.zxc_dict_get_id(b, len(b))


def test_seek_table_and_plan(prod, ref, orc):
    data = zc.silesia_shaped(1 << 20, seed=4)[:700001]
    for cks in (0, 1):
        frame = ref.compress(data, level=3, block_size=65536, checksum=cks, seekable=1)
        fb = frame.tobytes()
        hp = prod.lib.zxc_seekable_open(fb, len(fb))
        hr = ref.lib.zxc_seekable_open(fb, len(fb))
        assert hp and hr
        nb = ref.lib.zxc_seekable_get_num_blocks(hr)
        assert prod.lib.zxc_seekable_get_num_blocks(hp) == nb == 11
        assert prod.lib.zxc_seekable_get_decompressed_size(hp) == ref.lib.zxc_seekable_get_decompressed_size(hr)
        for i in range(nb + 2):
            assert prod.lib.zxc_seekable_get_block_comp_size(hp, i) == ref.lib.zxc_seekable_get_block_comp_size(hr, i)
            assert prod.lib.zxc_seekable_get_block_decomp_size(hp, i) == ref.lib.zxc_seekable_get_block_decomp_size(hr, i)
        # job table from the sequential walk agrees with the SEK table
        class Job(C.Structure):
            _fields_ = [("src_off", C.c_uint64), ("dst_off", C.c_uint64), ("src_len", C.c_uint32), ("dst_cap", C.c_uint32)]
        class Info(C.Structure):
            _fields_ = [("decoded_size", C.c_uint64), ("block_size", C.c_uint32), ("n_blocks", C.c_uint32),
                        ("dict_id", C.c_uint32), ("has_checksum", C.c_int), ("seekable", C.c_int), ("global_hash", C.c_uint32)]
        prod.lib.zxc_b200_plan_frame.restype = C.c_int64
        prod.lib.zxc_b200_plan_frame.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]
        jobs = (Job * nb)()
        info = Info()
        assert prod.lib.zxc_b200_plan_frame(fb, len(fb), jobs, nb, C.byref(info)) == nb
        assert info.decoded_size == data.size and info.block_size == 65536 and info.seekable == 1
        assert info.has_checksum == cks
        off = 16
        for i in range(nb):
            assert jobs[i].src_off == off and jobs[i].src_len == ref.lib.zxc_seekable_get_block_comp_size(hr, i)
            assert jobs[i].dst_off == i * 65536
            assert jobs[i].dst_cap == ref.lib.zxc_seekable_get_block_decomp_size(hr, i)
            off += jobs[i].src_len
        prod.lib.zxc_seekable_free(hp)
        ref.lib.zxc_seekable_free(hr)
        # not-seekable / damaged SEK -> NULL handle in both
        plain = ref.compress(data, level=3, block_size=65536, checksum=cks, seekable=0).tobytes()
        assert not prod.lib.zxc_seekable_open(plain, len(plain)) and not ref.lib.zxc_seekable_open(plain, len(plain))
        bad = bytearray(fb)
        bad[-20] ^= 0x55
        bad = bytes(bad)
        assert bool(prod.lib.zxc_seekable_open(bad, len(bad))) == bool(ref.lib.zxc_seekable_open(bad, len(bad)))


def test_walk_ignores_what_a_forged_seek_table_says(prod, ref):
    """zxw_walk uses a seek table at the tail only to prefetch block headers ahead of the sequential walk
    (zxc_frame.c); the plan must not depend on the entries: forged ones (zero, huge, random) give the same job table."""
    class Job(C.Structure):
        _fields_ = [("src_off", C.c_uint64), ("dst_off", C.c_uint64), ("src_len", C.c_uint32), ("dst_cap", C.c_uint32)]
    prod.lib.zxc_b200_plan_frame.restype = C.c_int64
    prod.lib.zxc_b200_plan_frame.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]
    data = zc.silesia_shaped(1 << 20, seed=9)
    rng = np.random.default_rng(5)
    for bs, cks in ((4096, 0), (4096, 1), (65536, 1)):
        fb = ref.compress(data, level=3, block_size=bs, checksum=cks, seekable=1).tobytes()
        nb = (data.size + bs - 1) // bs
        def plan(b):
            jobs = (Job * nb)()
            assert prod.lib.zxc_b200_plan_frame(b, len(b), jobs, nb, None) == nb
            return [(j.src_off, j.dst_off, j.src_len, j.dst_cap) for j in jobs]
        want = plan(fb)
        ent0 = len(fb) - 12 - 4 * nb
        assert fb[ent0 - 8] == 254  # the SEK block header
        for kind in range(4):
            b = bytearray(fb)
            if kind == 0:
                b[ent0:ent0 + 4 * nb] = bytes(4 * nb)
            elif kind == 1:
                b[ent0:ent0 + 4 * nb] = b"\xff" * (4 * nb)
            elif kind == 2:
                b[ent0:ent0 + 4 * nb] = rng.integers(0, 256, 4 * nb, dtype=np.uint8).tobytes()
            else:
                b[ent0:ent0 + 4 * nb] = np.full(nb, len(fb) // 2, dtype="<u4").tobytes()
            assert plan(bytes(b)) == want


def test_write_seek_table_matches_reference(prod, ref):
    sizes = np.array([22, 4000, 70000, 8, 123456], dtype="<u4")
    for lib in (prod.lib, ref.lib):
        lib.zxc_write_seek_table.restype = C.c_int64
        lib.zxc_write_seek_table.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint32]
    a = np.zeros(64, np.uint8)
    b = np.zeros(64, np.uint8)
    assert prod.lib.zxc_write_seek_table(a.ctypes.data, 64, sizes.ctypes.data, 5) == 28
    assert ref.lib.zxc_write_seek_table(b.ctypes.data, 64, sizes.ctypes.data, 5) == 28
    assert np.array_equal(a, b)
    assert prod.lib.zxc_write_seek_table(a.ctypes.data, 27, sizes.ctypes.data, 5) == -2


HOST_LEVEL = ["zero_length", "too_short_4bytes", "truncated_header_only", "all_0xff_garbage", "bad_magic",
              "magic_then_zeros", "bad_version", "bad_header_crc", "bad_checksum_algo", "bad_block_size_field"]


@pytest.mark.parametrize("name", HOST_LEVEL)
def test_header_level_rejects_need_no_device(prod, name):
    exp = json.load(open(os.path.join(G, "invalid", "expected.json")))[name]
    frame = open(os.path.join(G, "invalid", name + ".zxc"), "rb").read()
    out = np.zeros(1 << 16, np.uint8)
    o = z.DecompressOpts(checksum_enabled=1)
    r = prod.lib.zxc_decompress(frame if frame else b"\0", len(frame), out.ctypes.data, out.size, C.byref(o))
    assert r == exp, z.ERR.get(r)


def test_argument_checks(prod):
    out = np.zeros(64, np.uint8)
    assert prod.lib.zxc_decompress(None, 100, out.ctypes.data, 64, None) == -12
    assert prod.lib.zxc_decompress(b"x" * 10, 10, out.ctypes.data, 64, None) == -3
    empty = open(os.path.join(G, "valid", "empty.zxc"), "rb").read()
    assert prod.lib.zxc_decompress(empty, len(empty), None, 0, None) == 0
    one = open(os.path.join(G, "valid", "one_byte.zxc"), "rb").read()
    assert prod.lib.zxc_decompress(one, len(one), None, 0, None) == -2
    assert prod.lib.zxc_compress(b"abc", 3, None, 0, None) == -12
    o = z.CompressOpts(block_size=12345)
    assert prod.lib.zxc_compress(b"abc", 3, out.ctypes.data, 64, C.byref(o)) == -14
    # FILE* entry point: argument checks come before any device work (zxc_driver.c:1035-1050)
    prod.lib.zxc_stream_compress.restype = C.c_int64
    prod.lib.zxc_stream_compress.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    assert prod.lib.zxc_stream_compress(None, None, None) == -12
    libc = C.CDLL(None)
    libc.fopen.restype = C.c_void_p
    libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
    libc.fclose.argtypes = [C.c_void_p]
    f = libc.fopen(os.path.join(G, "valid", "empty.zxc").encode(), b"rb")
    assert prod.lib.zxc_stream_compress(f, None, C.byref(o)) == -14
    libc.fclose(f)


@pytest.mark.skipif(has_cuda(), reason="only meaningful without a GPU")
def test_no_device_fails_loudly(prod):
    """There is no CPU codec behind the API: a valid frame on a GPU-less host is an error."""
    frame = open(os.path.join(G, "valid", "text_64k_level3.zxc"), "rb").read()
    r, _ = prod.decompress(frame, 65536)
    assert r == -100 and prod.lib.zxc_error_name(r) == b"ZXC_B200_ERROR_NO_DEVICE"
    assert prod.lib.zxc_b200_device_count() == 0
    for level in (1, 3, 6, 7):  # every level encodes on the GPU only
        assert prod.compress(np.frombuffer(b"abcdefgh" * 64, np.uint8), level=level) == -100


def test_host_parsers_fuzz_against_reference(prod, ref):
    """Frame-level parsing needs no device: damaged / truncated / extended seekable frames through
    zxc_seekable_open (+ the per-block queries), zxc_get_decompressed_size and zxc_get_dict_id give what the
    reference gives (zxc_seekable.c:270-396, zxc_dispatch.c:1010-1075).  Damage is aimed at the regions the host
    reads: file header, block headers, SEK block, EOF block, footer."""
    data = zc.silesia_shaped(1 << 20, seed=9)[:400000]
    rng = np.random.default_rng(41)
    for L in (prod.lib, ref.lib):
        L.zxc_seekable_open.restype = C.c_void_p
        L.zxc_seekable_open.argtypes = [C.c_void_p, C.c_size_t]
        L.zxc_seekable_free.argtypes = [C.c_void_p]
        for fn in ("zxc_seekable_get_num_blocks", "zxc_seekable_get_block_comp_size", "zxc_seekable_get_block_decomp_size"):
            getattr(L, fn).restype = C.c_uint32
        L.zxc_seekable_get_num_blocks.argtypes = [C.c_void_p]
        L.zxc_seekable_get_block_comp_size.argtypes = [C.c_void_p, C.c_uint32]
        L.zxc_seekable_get_block_decomp_size.argtypes = [C.c_void_p, C.c_uint32]
        L.zxc_seekable_get_decompressed_size.restype = C.c_uint64
        L.zxc_seekable_get_decompressed_size.argtypes = [C.c_void_p]

    def view(L, buf):
        p, n = buf.ctypes.data, buf.size
        out = [int(L.zxc_get_decompressed_size(p, n)), int(L.zxc_get_dict_id(p, n))]
        h = L.zxc_seekable_open(p, n)
        out.append(bool(h))
        if h:
            nb = L.zxc_seekable_get_num_blocks(h)
            out += [nb, int(L.zxc_seekable_get_decompressed_size(h))]
            out += [(L.zxc_seekable_get_block_comp_size(h, i), L.zxc_seekable_get_block_decomp_size(h, i)) for i in range(min(nb, 40) + 1)]
            L.zxc_seekable_free(h)
        return out

    opened = 0
    for bs, cks in ((4096, 0), (16384, 1), (65536, 0)):
        frame = ref.compress(data, level=3, block_size=bs, checksum=cks, seekable=1)
        nb = -(-data.size // bs)
        tail = 8 + 4 * nb + 8 + 12 + 64  # SEK block + EOF block + footer and a little of the last data block
        for t in range(220):
            f = frame.copy()
            kind = t % 5
            if kind == 0:    # file header
                f[int(rng.integers(0, 16))] = int(rng.integers(0, 256))
            elif kind == 1:  # SEK / EOF / footer region
                for _ in range(int(rng.integers(1, 3))):
                    f[f.size - 1 - int(rng.integers(0, tail))] ^= int(rng.integers(1, 256))
            elif kind == 2:  # truncated
                f = f[:f.size - int(rng.integers(1, tail + 200))].copy()
            elif kind == 3:  # trailing garbage
                f = np.concatenate([f, rng.integers(0, 256, int(rng.integers(1, 40)), dtype=np.uint8)])
            else:            # a block header somewhere in the body
                f[16 + int(rng.integers(0, 8))] ^= int(rng.integers(1, 256))
            a, b = view(prod.lib, f), view(ref.lib, f)
            assert a == b, (bs, cks, t, kind, a[:5], b[:5])
            opened += a[2]
    assert opened > 20  # some damage leaves a frame the SEK parser still accepts: those were compared in full


def test_dict_container_fuzz_against_reference(prod, ref):
    """Damaged / truncated .zxd containers: zxc_dict_load and zxc_dict_get_id agree with the reference
    (zxc_dict.c container checks: magic, version, sizes, table validity, checksum)."""
    rng = np.random.default_rng(5)
    for L in (prod.lib, ref.lib):
        L.zxc_dict_save.restype = C.c_int64
        L.zxc_dict_save.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t]
        L.zxc_dict_get_id.restype = C.c_uint32
    same_ok = 0
    for n in (1, 40, 700, 9000):
        d = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        for huf in (bytes([0x88] * 128), rng.integers(0, 12, 128, dtype=np.uint8).tobytes()):  # a flat table / arbitrary nibbles
            box = np.zeros(n + 200, np.uint8)
            sz = ref.lib.zxc_dict_save(d, n, huf, box.ctypes.data, box.size)
            assert sz > 0
            good = box[:sz].copy()
            for t in range(120):
                b = good.copy()
                k = t % 3
                if k == 0:
                    b[int(rng.integers(0, min(b.size, 160)))] ^= int(rng.integers(1, 256))
                elif k == 1:
                    b[int(rng.integers(0, b.size))] = int(rng.integers(0, 256))
                else:
                    b = b[:int(rng.integers(0, b.size))].copy()
                raw = b.tobytes()
                a, r = prod.dict_load(raw), ref.dict_load(raw)
                assert a == r, (n, t, a[0], r[0])
                assert prod.lib.zxc_dict_get_id(raw, len(raw)) == ref.lib.zxc_dict_get_id(raw, len(raw))
                same_ok += a[0] == 0
    assert same_ok > 0
