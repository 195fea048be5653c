"""CPU tool (not collected by pytest): tests/simt_fuzz.py for dictionary frames -- records of 1-16 KiB, a dictionary
from the reference's trainer (2-16 KiB), levels 1/3/5, both decode bodies, then single byte flips, on the warp
emulator against the unmodified reference.  python tests/simt_fuzz_dict.py SEED SECONDS"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import zxc_ctypes as z, zxc_corpus as zc, zxc_simt as zs
ref = z.ZxcLib(z.REF_SO); prod = z.ZxcLib(z.PRODUCT_SO)
seed = int(sys.argv[1]); budget = float(sys.argv[2])
rng = np.random.default_rng(seed)
t0 = time.time(); n_ok = n_bad = n_skip = 0
while time.time() - t0 < budget:
    rs = int(rng.choice([1024, 4096, 16384]))
    nrec = int(rng.integers(4, 64))
    recs = zc.records(nrec, record_size=rs, seed=int(rng.integers(1, 1000)))
    cap = int(rng.choice([2048, 4096, 8192, 16384]))
    d = zc.train_dict_ref(ref, recs, record_size=rs, n_samples=min(nrec, 32), cap=cap)
    level = int(rng.choice([1, 3, 5]))
    bs = int(rng.choice([4096, 16384, 65536]))
    frame = ref.compress(recs, level=level, block_size=bs, checksum=0, seekable=1, dict=d)
    for units in (0, 1):
        st, out, oob, _ = zs.decode_frame(prod, frame, dict=d, units=units, seed=int(rng.integers(1, 1 << 30)))
        assert oob == 0 and all(s >= 0 for s in st) and np.array_equal(out, recs), ("valid", rs, nrec, cap, level, bs, units)
    n_ok += 1
    fb = bytearray(frame.tobytes())
    nb = (recs.size + bs - 1) // bs
    caps = [min(bs, recs.size - i * bs) for i in range(nb)]
    for _ in range(8):
        b = bytearray(fb)
        pos = int(rng.integers(32, len(b) - 24))
        b[pos] ^= int(rng.integers(1, 256))
        r_ref, out_ref = ref.decompress(bytes(b), recs.size, dict=d)
        try:
            st, out, oob, _ = zs.decode_frame(prod, bytes(b), dict=d, seed=int(rng.integers(1, 1 << 30)))
        except AssertionError:
            n_skip += 1; continue
        assert oob == 0
        if len(st) != nb: n_skip += 1; continue
        bad = [s for s, c in zip(st, caps) if s < 0 or s != c]
        if r_ref == recs.size and bad and all(s2 == -10 or s2 >= 0 for s2 in bad): n_skip += 1; continue
        if r_ref == recs.size:
            assert not bad and np.array_equal(out[:recs.size], out_ref), ("ref accepts", rs, nrec, cap, level, bs, [z.ERR.get(s, s) for s in bad][:3])
        else:
            assert bad, ("ref rejects, kernel accepts", rs, nrec, cap, level, bs, r_ref)
        n_bad += 1
print("dict: valid frames", n_ok, "damaged compared", n_bad, "skipped", n_skip, "in %.0f s" % (time.time() - t0))
