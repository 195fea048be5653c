"""GPU parity for the encoder: zxc_compress through the C ABI must emit frames that are
BYTE-IDENTICAL to the unmodified reference encoder's (levels 1-7, with and without dictionaries),
and the reference's golden encoder KAT (tests/format/golden.sha256) must reproduce."""
import hashlib
import os

import numpy as np
import pytest

import zxc_corpus as zc
import zxc_ctypes as z
from test_oracle import CASES, G, make_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind,n", CASES + [("silesia_big", 9 << 20)])
@pytest.mark.parametrize("level", [1, 2, 3, 4, 5])
def test_emitted_frame_identical_to_reference(prod, ref, kind, n, level):
    data = zc.silesia_shaped(n, seed=23, offset=60 << 20) if kind == "silesia_big" else make_case(kind, n)
    for bs, cks, seek in ((65536, 0, 1), (4096, 1, 0), (0, 1, 1), (2 << 20, 0, 0)):
        a = ref.compress(data, level=level, block_size=bs, checksum=cks, seekable=seek)
        b = prod.compress(data, level=level, block_size=bs, checksum=cks, seekable=seek)
        assert not isinstance(b, int), (kind, level, bs, z.ERR.get(b, b))
        if a.size != b.size or not np.array_equal(a, b):
            first = int(np.argmax(a[:min(a.size, b.size)] != b[:min(a.size, b.size)])) if a.size and b.size else 0
            pytest.fail(f"{kind} L{level} bs={bs}: sizes {a.size} vs {b.size}, first diff at {first}")


OPT_CASES = [("silesia", 1 << 20), ("text", 150000), ("random", 40000), ("numeric", 120000), ("binrec", 90000),
             ("period1", 100000), ("period7", 100000), ("period300", 90000), ("runs", 80000), ("tiny", 1), ("small", 37),
             ("zeros", 200000)]


@pytest.mark.parametrize("kind,n", OPT_CASES)
@pytest.mark.parametrize("level", [6, 7])
def test_optimal_levels_identical_to_reference(prod, ref, kind, n, level):
    """levels 6-7: optimal parser + Huffman literal (and, at 7, token) sections"""
    data = make_case(kind, n)
    for bs, cks, seek in ((65536, 0, 1), (4096, 1, 0), (0, 1, 1)):
        if bs == 0 and n > 200000:
            continue  # one 512 KiB block parsed by a single warp: seconds; covered by the smaller cases
        a = ref.compress(data, level=level, block_size=bs, checksum=cks, seekable=seek)
        b = prod.compress(data, level=level, block_size=bs, checksum=cks, seekable=seek)
        assert not isinstance(b, int), (kind, level, bs, z.ERR.get(b, b))
        assert a.size == b.size and np.array_equal(a, b), (kind, level, bs)


def test_optimal_levels_use_huffman_sections(prod, ref, orc):
    """the parity above is not vacuous: these inputs do select enc_lit = 2 / enc_tok = 2"""
    data = make_case("text", 150000)
    for level in (6, 7):
        fr = prod.compress(data, level=level, block_size=65536)
        rc, st = orc.stats(fr)
        assert rc == 0 and st["huf_blocks"] > 0
        r, out = prod.decompress(fr, data.size)
        assert r == data.size and np.array_equal(out, data)


def test_optimal_levels_with_dictionary_and_shared_table(prod, ref):
    """dictionary-seeded optimal parse, and the dictionary's shared literal table (enc_lit = 3)"""
    import ctypes as C
    from test_oracle import GC_DICT, golden_dicts
    rng = np.random.default_rng(5)
    lines = []
    for _ in range(400):
        lines.append(b"GET /api/v1/users/%d/profile?session=%08x&page=%d HTTP/1.1\r\nHost: api.example.com\r\n"
                     b"Accept: application/json\r\nUser-Agent: zxc-client\r\n\r\n"
                     % (int(rng.integers(0, 100000)), int(rng.integers(0, 1 << 32)), int(rng.integers(0, 64))))
    data = np.frombuffer(b"".join(lines), np.uint8)
    # the reference's own trainer makes the shared table (the product ships no trainer)
    ref.lib.zxc_train_dict_huf.restype = C.c_int
    ref.lib.zxc_train_dict_huf.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]
    sample = data[:4000].copy()
    ptrs = (C.c_void_p * 1)(sample.ctypes.data)
    sizes = (C.c_size_t * 1)(sample.size)
    huf = np.zeros(128, np.uint8)
    assert ref.lib.zxc_train_dict_huf(ptrs, sizes, 1, GC_DICT, len(GC_DICT), huf.ctypes.data) == 0
    tables = [(GC_DICT, None), (GC_DICT, huf.tobytes())] + [(c, h) for c, h in golden_dicts().values()]
    used3 = 0
    for d, h in tables:
        for level, bs in ((6, 4096), (7, 4096), (6, 65536), (7, 0)):
            a = ref.compress(data, level=level, block_size=bs, seekable=1, checksum=1, dict=d, dict_huf=h)
            b = prod.compress(data, level=level, block_size=bs, seekable=1, checksum=1, dict=d, dict_huf=h)
            assert not isinstance(a, int), a
            assert not isinstance(b, int), (level, bs, z.ERR.get(b, b))
            assert a.size == b.size and np.array_equal(a, b), (level, bs, len(d), h is not None)
            r, out = prod.decompress(b, data.size, checksum=1, dict=d, dict_huf=h)
            assert r == data.size and np.array_equal(out, data)
            # GLO sub-header byte 8 = enc_lit (docs/FORMAT.md 5.2) of the first data block
            used3 += int(b[16] == 1 and b[16 + 8 + 8] == 3)
    assert used3 > 0
    bad = bytes([0xFF] * 128)  # lengths of 15: not a valid table -> the call fails like the reference's
    assert prod.compress(data, level=6, dict=GC_DICT, dict_huf=bad) == ref.compress(data, level=6, dict=GC_DICT, dict_huf=bad)


def test_golden_encoder_kat_levels_6_7(prod, ref):
    """golden cases 05 (level 6 Huffman), 13 (level 7, 11-bit codes), 12 (shared dictionary table)"""
    import ctypes as C
    from test_oracle import GC_DICT
    sha = {l.split()[1]: l.split()[0] for l in open(os.path.join(G, "format", "golden.sha256"))}

    def lcg(seed):
        s = seed
        while True:
            s = (s * 1103515245 + 12345) & 0xFFFFFFFF
            yield s

    alpha = b"aaaaaabbbbccdefg"
    g = lcg(0x0BADF00D)
    huff = np.array([alpha[(next(g) >> 16) & 15] for _ in range(16384)], np.uint8)
    g = lcg(0x0C0FFEE1)
    wide = np.empty(16384, np.uint8)
    for i in range(16384):
        u = (next(g) >> 16) & 0xFFFF
        u2 = (u * u) >> 16
        u4 = (u2 * u2) >> 16
        wide[i] = (u4 * 220) >> 16
    fr = prod.compress(huff, level=6)
    assert hashlib.sha256(fr.tobytes()).hexdigest() == sha["05_block_glo_huffman.zxc"]
    fr = prod.compress(wide, level=7)
    assert hashlib.sha256(fr.tobytes()).hexdigest() == sha["13_glo_huffman_wide.zxc"]
    # 12: payload with LCG-varied request lines, table from the reference's trainer on that payload
    g = lcg(0x5EEDCAFE)
    buf = b""
    while len(buf) + 160 < 4096:
        uid, sess, page = next(g) % 100000, next(g), next(g) % 64
        buf += (b"GET /api/v1/users/%d/profile?session=%08x&page=%d HTTP/1.1\r\nHost: api.example.com\r\n"
                b"Accept: application/json\r\nUser-Agent: zxc-client\r\n\r\n" % (uid, sess, page))
    payload = np.frombuffer(buf, np.uint8).copy()
    ref.lib.zxc_train_dict_huf.restype = C.c_int
    ref.lib.zxc_train_dict_huf.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]
    ptrs = (C.c_void_p * 1)(payload.ctypes.data)
    sizes = (C.c_size_t * 1)(payload.size)
    huf = np.zeros(128, np.uint8)
    assert ref.lib.zxc_train_dict_huf(ptrs, sizes, 1, GC_DICT, len(GC_DICT), huf.ctypes.data) == 0
    fr = prod.compress(payload, level=6, dict=GC_DICT, dict_huf=huf.tobytes())
    assert hashlib.sha256(fr.tobytes()).hexdigest() == sha["12_glo_huffman_dict.zxc"]
    golden = np.fromfile(os.path.join(G, "format", "12_glo_huffman_dict.zxc"), np.uint8)
    r, out = prod.decompress(golden, payload.size, dict=GC_DICT, dict_huf=huf.tobytes())
    assert r == payload.size and np.array_equal(out, payload)


def test_golden_encoder_kat(prod):
    """tests/format/gen_golden.c cases reproducible without Huffman / dictionaries."""
    sha = {}
    for line in open(os.path.join(G, "format", "golden.sha256")):
        h, name = line.split()
        sha[name] = h
    phrase = (b"the quick brown fox jumps over the lazy dog. ZXC compresses repeated "
              b"patterns efficiently and decompresses them very fast. ")
    text = lambda n: np.frombuffer(bytes(phrase[i % len(phrase)] for i in range(n)), np.uint8)

    def lcg(seed):
        s = seed
        while True:
            s = (s * 1103515245 + 12345) & 0xFFFFFFFF
            yield s

    def raw4096():
        g = lcg(0x1234567)
        return np.array([next(g) >> 24 for _ in range(4096)], np.uint8)

    def offset16():
        g = lcg(0x5EED1234)
        per = np.array([next(g) >> 24 for _ in range(1024)], np.uint8)
        return np.tile(per, 8)

    def rle_lits():
        g = lcg(0x1357BD13)
        out = np.full(16384, 0xAA, np.uint8)
        for i in range(0, 16384, 5):
            out[i] = next(g) >> 24
        return out

    multiblock = text(5 * 4096 + 777)
    cases = {  # name: (data, level, block_size, checksum, seekable)   tests/format/golden_cases.h
        "01_empty_eof_only.zxc": (np.zeros(0, np.uint8), 1, 0, 0, 0),
        "02_block_raw.zxc": (raw4096(), 1, 0, 0, 0),
        "03_block_ghi.zxc": (text(8192), 1, 0, 0, 0),
        "04_block_glo.zxc": (text(8192), 3, 0, 0, 0),
        "06_checksum_per_block.zxc": (text(8192), 3, 0, 1, 0),
        "07_multiple_blocks.zxc": (multiblock, 3, 4096, 1, 0),
        "08_seekable_table.zxc": (multiblock, 3, 4096, 1, 1),
        "10_glo_offset16.zxc": (offset16(), 3, 0, 0, 0),
        "11_glo_rle.zxc": (rle_lits(), 3, 0, 0, 0),
    }
    got = {}
    for name, (data, level, bs, cks, seek) in cases.items():
        fr = prod.compress(data if data.size else np.zeros(1, np.uint8)[:0], level=level, block_size=bs, checksum=cks, seekable=seek)
        assert not isinstance(fr, int), (name, fr)
        got[name] = hashlib.sha256(fr.tobytes()).hexdigest()
    for name in cases:
        assert got[name] == sha[name], name


def test_roundtrip_through_own_decoder(prod):
    data = zc.silesia_shaped(16 << 20, seed=41)
    for level in (1, 3, 5):
        fr = prod.compress(data, level=level, block_size=65536, checksum=1, seekable=1)
        assert not isinstance(fr, int)
        r, out = prod.decompress(fr, data.size, checksum=1)
        assert r == data.size and np.array_equal(out, data)


def test_dictionary_frames_identical_to_reference(prod, ref):
    from test_oracle import GC_DICT
    rng = np.random.default_rng(12)
    words = [b'"user_id":', b'"timestamp":', b'"status":"ok"', b'"payload":{', b'"region":"eu-west"', b'},{']
    dict16k = b"".join(words[i % len(words)] + b"," for i in range(2000))[:16384]
    recs = b"".join(b"{" + b",".join(words[int(k)] + str(int(v)).encode() for k, v in zip(rng.integers(0, 6, 20), rng.integers(0, 1 << 20, 20))) + b"}\n"
                    for _ in range(3000))
    data = np.frombuffer(recs, np.uint8)
    for d in (dict16k, GC_DICT, dict16k[:7], dict16k[:5], bytes(rng.integers(0, 256, 65535, dtype=np.uint8))):
        for level, bs in ((5, 4096), (3, 65536), (1, 4096), (2, 8192), (4, 1 << 20)):
            a = ref.compress(data, level=level, block_size=bs, seekable=1, checksum=1, dict=d)
            b = prod.compress(data, level=level, block_size=bs, seekable=1, checksum=1, dict=d)
            assert not isinstance(b, int), (level, bs, len(d), z.ERR.get(b, b))
            assert a.size == b.size and np.array_equal(a, b), (level, bs, len(d))
    # golden 09_block_dict
    req = (b"GET /api/v1/users/4242/profile HTTP/1.1\r\nHost: api.example.com\r\n"
           b"Accept: application/json\r\nUser-Agent: zxc-client\r\n\r\n")
    payload = np.frombuffer(bytes(req[i % len(req)] for i in range(4096)), np.uint8)
    fr = prod.compress(payload, level=3, dict=GC_DICT)
    want = [l.split()[0] for l in open(os.path.join(G, "format", "golden.sha256")) if "09_block_dict" in l][0]
    assert hashlib.sha256(fr.tobytes()).hexdigest() == want


def test_block_api_compress_identical(prod, ref):
    import ctypes as C
    data = zc.silesia_shaped(4 << 20, seed=77)
    rc_ref = ref.lib.zxc_create_cctx(None)
    rc_prod = prod.lib.zxc_create_cctx(None)
    dctx = prod.lib.zxc_create_dctx()
    for n, level, cks in ((4096, 5, 0), (65536, 3, 1), (100000, 1, 0), (700, 3, 1), (1 << 20, 4, 0), (9, 3, 0), (1, 2, 1),
                          (20000, 6, 1), (50000, 7, 0), (300, 6, 0), (8, 7, 1)):
        src = data[1000:1000 + n].copy()
        cap = int(ref.lib.zxc_compress_block_bound(n))
        a = np.zeros(cap, np.uint8)
        b = np.zeros(cap, np.uint8)
        o = z.CompressOpts(level=level, checksum_enabled=cks)
        ra = ref.lib.zxc_compress_block(rc_ref, src.ctypes.data, n, a.ctypes.data, cap, C.byref(o))
        rb = prod.lib.zxc_compress_block(rc_prod, src.ctypes.data, n, b.ctypes.data, cap, C.byref(o))
        assert ra == rb > 0, (n, level, ra, rb)
        assert np.array_equal(a[:ra], b[:rb]), (n, level)
        out = np.zeros(n, np.uint8)
        do = z.DecompressOpts(checksum_enabled=cks)
        assert prod.lib.zxc_decompress_block(dctx, b.ctypes.data, rb, out.ctypes.data, n, C.byref(do)) == n
        assert np.array_equal(out, src)
    prod.lib.zxc_free_cctx(rc_prod)
    ref.lib.zxc_free_cctx(rc_ref)
    prod.lib.zxc_free_dctx(dctx)


def test_stream_compress_file_api(prod, ref, tmp_path):
    """zxc_stream_compress (FILE* -> FILE*): same frame as the reference's streaming engine, dry run = size"""
    import ctypes as C
    libc = C.CDLL(None)
    libc.fopen.restype = C.c_void_p
    libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
    libc.fclose.argtypes = [C.c_void_p]
    data = zc.silesia_shaped(3 << 20, seed=4)[:2500000]
    src = tmp_path / "in.bin"
    src.write_bytes(data.tobytes())
    for lib in (prod.lib, ref.lib):
        lib.zxc_stream_compress.restype = C.c_int64
        lib.zxc_stream_compress.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    for level, bs, cks, seek in ((3, 65536, 1, 1), (5, 0, 0, 0), (1, 4096, 1, 0), (6, 65536, 0, 1)):
        o = z.CompressOpts(level=level, block_size=bs, checksum_enabled=cks, seekable=seek, n_threads=2)
        outs = []
        for name, lib in (("prod", prod.lib), ("ref", ref.lib)):
            dst = tmp_path / f"{name}.zxc"
            fi, fo = libc.fopen(str(src).encode(), b"rb"), libc.fopen(str(dst).encode(), b"wb")
            r = lib.zxc_stream_compress(fi, fo, C.byref(o))
            libc.fclose(fi)
            libc.fclose(fo)
            got = np.fromfile(dst, np.uint8)
            assert r == got.size > 0, (name, level, r)
            outs.append(got)
        assert outs[0].size == outs[1].size and np.array_equal(outs[0], outs[1]), (level, bs)
        fi = libc.fopen(str(src).encode(), b"rb")
        assert prod.lib.zxc_stream_compress(fi, None, C.byref(o)) == outs[0].size  # dry run
        libc.fclose(fi)
    assert prod.lib.zxc_stream_compress(None, None, None) == -12  # ZXC_ERROR_NULL_INPUT
