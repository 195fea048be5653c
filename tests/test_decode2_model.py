"""CPU-only: the host replay of the block-cooperative decode kernel (oracle/decode2_model.cc, built on
the same zxc_decode2_core.h the device code compiles) against the oracle / the unmodified reference,
block by block: record packing, the word plan, the period fold for overlapped matches, the
end-of-sequence bitmasks, the extras segment maps with the reference's varint failure behaviour
(zxc_decompress.c:51-88).  The synchronisation of the real kernel is covered by the -m gpu tests."""
import ctypes as C
import os

import numpy as np
import pytest

import zxc_corpus as zc
import zxc_ctypes as z

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MODEL_SO = os.path.join(ROOT, "oracle", "libzxc_decode2_model.so")
DEFER = -1000

pytestmark = pytest.mark.skipif(not (os.path.exists(MODEL_SO) and z.have_ref()), reason="model / reference not built")


@pytest.fixture(scope="module")
def model():
    L = C.CDLL(MODEL_SO)
    L.z2_model_decode_block.restype = C.c_int
    L.z2_model_decode_block.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32,
                                        C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32)]
    return L


def walk(frame):
    """(offset, on-disk length) of every data block of a frame without checksums"""
    b = frame.tobytes() if isinstance(frame, np.ndarray) else bytes(frame)
    p, out = 16, []
    while True:
        t = b[p]
        comp = int.from_bytes(b[p + 3:p + 7], "little")
        if t in (254, 255):
            break
        out.append((p, 8 + comp))
        p += 8 + comp
    return b, out


def run_blocks(model, orc, frame, data_len, block_size, dict_=None, threads=256):
    b, blocks = walk(frame)
    win = (block_size + 511) & ~511
    gap = min(4096, win)
    n_def, slow = 0, C.c_uint32(0)
    buf = np.frombuffer(b, np.uint8)
    d = np.frombuffer(dict_, np.uint8) if dict_ else None
    for i, (off, ln) in enumerate(blocks):
        cap = min(block_size, data_len - i * block_size)
        out = np.zeros(cap + 8, np.uint8)
        r = model.z2_model_decode_block(buf.ctypes.data + off, ln, out.ctypes.data, cap,
                                        d.ctypes.data if d is not None else None, len(dict_) if dict_ else 0,
                                        win, gap, threads, C.byref(slow))
        exp = np.zeros(cap + 8, np.uint8)
        e = orc.lib.zxo_decode_block(buf.ctypes.data + off, ln, exp.ctypes.data, cap,
                                     d.ctypes.data if d is not None else None, len(dict_) if dict_ else 0, None, 0)
        if r == DEFER:
            n_def += 1
            continue
        assert r == e, (i, r, e)
        if r > 0:
            assert np.array_equal(out[:r], exp[:r]), (i, int(np.argmax(out[:r] != exp[:r])))
    return len(blocks), n_def, slow.value


KINDS = {
    "silesia": lambda n: zc.silesia_shaped(max(n, 1 << 20), seed=5)[:n],
    "text": lambda n: zc.gen_text(n),
    "numeric": lambda n: zc.gen_numeric(n),
    "records": lambda n: zc.gen_binary_records(n),
    "period1": lambda n: zc.gen_periodic(n, 1),
    "period3": lambda n: zc.gen_periodic(n, 3),
    "period7": lambda n: zc.gen_periodic(n, 7),
    "period300": lambda n: zc.gen_periodic(n, 300),
    "runs": lambda n: zc.gen_runs(n),
    "random": lambda n: zc.gen_random(n),
}


@pytest.mark.parametrize("kind", sorted(KINDS))
@pytest.mark.parametrize("level", [1, 2, 3, 5])
def test_model_matches_reference(model, orc, ref, kind, level):
    n = 300000
    data = KINDS[kind](n)
    taken = 0
    for bs in (4096, 16384, 65536):
        fr = ref.compress(data, level=level, block_size=bs)
        nb, nd, _ = run_blocks(model, orc, fr, n, bs, threads=64 if bs < 16384 else 256)
        taken += nb - nd
    assert taken > 0 or kind in ("random", "numeric")


def test_model_silesia_large(model, orc, ref):
    n = 24 << 20
    data = zc.silesia_shaped(n, seed=1)
    fr = zc.compress_ref_mt(ref, data, level=3, block_size=65536)
    nb, nd, slow = run_blocks(model, orc, fr, n, 65536)
    assert nb == n // 65536 and nd < nb // 4
    print("blocks", nb, "deferred", nd, "slow words per block", slow / max(1, nb - nd))


def test_model_dictionary(model, orc, ref):
    rng = np.random.default_rng(3)
    words = [bytes(rng.integers(97, 123, int(rng.integers(3, 12)), dtype=np.uint8)) for _ in range(300)]
    def text(k, seed):
        r = np.random.default_rng(seed)
        return b" ".join(words[int(i)] for i in r.integers(0, 300, k))
    dict_ = text(2500, 1)[:16384]
    data = np.frombuffer(text(40000, 2), np.uint8)[:200000].copy()
    for bs in (4096, 65536):
        fr = ref.compress(data, level=5, block_size=bs, dict=dict_)
        nb, nd, slow = run_blocks(model, orc, fr, data.size, bs, dict_=dict_, threads=64 if bs == 4096 else 256)
        assert nd == 0 and slow > 0  # dictionary sources take the byte-wise path


def test_model_mutations(model, orc, ref):
    """damaged sequence sections: same verdict (or same bytes) as the oracle, block by block"""
    data = zc.silesia_shaped(1 << 20, seed=9)[:262144]
    rng = np.random.default_rng(11)
    for level in (1, 3):
        fr = ref.compress(data, level=level, block_size=65536)
        b, blocks = walk(fr)
        for trial in range(150):
            bi = int(rng.integers(0, len(blocks)))
            off, ln = blocks[bi]
            blk = bytearray(b[off:off + ln])
            if blk[0] not in (1, 2):
                continue
            n_lit = int.from_bytes(blk[12:16], "little")
            lo = 8 + 12 + n_lit  # only the sequence sections (the header stays valid)
            if lo >= ln:
                continue
            for _ in range(int(rng.integers(1, 4))):
                blk[int(rng.integers(lo, ln))] = int(rng.integers(0, 256))
            arr = np.frombuffer(bytes(blk), np.uint8)
            cap = 65536
            out = np.zeros(cap + 8, np.uint8)
            exp = np.zeros(cap + 8, np.uint8)
            r = model.z2_model_decode_block(arr.ctypes.data, ln, out.ctypes.data, cap, None, 0, 65536, 4096, 256, None)
            e = orc.lib.zxo_decode_block(arr.ctypes.data, ln, exp.ctypes.data, cap, None, 0, None, 0)
            if r == DEFER:
                continue
            assert r == e, (level, trial, r, e)
            if r > 0:
                assert np.array_equal(out[:r], exp[:r])
