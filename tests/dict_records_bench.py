"""BASELINE.json configs[3]: dictionary decode, 16 KiB shared dictionary, N x 4 KiB records, level 5,
one frame with block_size = 4096 (one record per block).  Run on the GPU box:
    python tests/dict_records_bench.py [records]
Prints decode-only GB/s (HBM resident), e2e GB/s, and the reference's CPU numbers beside them."""
import ctypes as C
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

sys.path.insert(0, __file__.rsplit("/", 1)[0])
import zxc_corpus as zc  # noqa: E402
import zxc_ctypes as z  # noqa: E402

REC = 4096


def records(n, seed=7):
    lib = zc._corpus()
    lib.zxcorp_records.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint64]
    out = np.empty(n * REC, np.uint8)
    th = zc.host_threads()
    per = (n + th - 1) // th

    def work(i):
        lo = i * per
        if lo < n:
            lib.zxcorp_records(out.ctypes.data + lo * REC, lo, min(per, n - lo), REC, seed)

    with ThreadPoolExecutor(th) as ex:
        list(ex.map(work, range(th)))
    return out


def main():
    import torch
    n = int(sys.argv[1]) if len(sys.argv) > 1 else (1 << 20)
    prod, ref = z.ZxcLib(z.PRODUCT_SO), z.ZxcLib(z.REF_SO)
    data = records(n)
    # dictionary: the reference's own trainer on the first 4096 records
    ref.lib.zxc_train_dict.restype = C.c_int64
    ref.lib.zxc_train_dict.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
    ns = 4096
    ptrs = (C.c_void_p * ns)(*[data.ctypes.data + i * REC for i in range(ns)])
    sizes = (C.c_size_t * ns)(*([REC] * ns))
    dbuf = np.zeros(16384, np.uint8)
    dsz = ref.lib.zxc_train_dict(ptrs, sizes, ns, dbuf.ctypes.data, dbuf.size)
    assert dsz > 0, dsz
    dict_bytes = dbuf[:dsz].tobytes()
    t = time.perf_counter()
    frame = prod.compress(data, level=5, block_size=REC, seekable=1, dict=dict_bytes)  # GPU encoder
    enc_dt = time.perf_counter() - t
    assert not isinstance(frame, int), frame
    sub = data[: 4096 * REC]
    assert np.array_equal(ref.compress(sub, level=5, block_size=REC, seekable=1, dict=dict_bytes),
                          prod.compress(sub, level=5, block_size=REC, seekable=1, dict=dict_bytes))
    print(f"records {n}, dict {dsz} B, ratio {frame.size / data.size:.4f}, GPU encode e2e {data.size / enc_dt / 1e9:.2f} GB/s "
          f"(byte-identical to the reference on the first 4096 records)")

    lib = prod.lib
    lib.zxc_b200_plan_frame.restype = C.c_int64
    lib.zxc_b200_plan_frame.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.zxc_b200_decode_scratch_size.restype = C.c_size_t
    lib.zxc_b200_decode_scratch_size.argtypes = [C.c_uint32]
    lib.zxc_b200_decode_blocks.restype = C.c_int
    lib.zxc_b200_decode_blocks.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32,
                                           C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32, C.c_int, C.c_void_p]
    lib.zxc_b200_reduce_status.restype = C.c_int64
    lib.zxc_b200_reduce_status.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
    nb = lib.zxc_b200_plan_frame(frame.ctypes.data, frame.size, None, 0, None)
    jobs = np.zeros(nb * 24, np.uint8)
    assert lib.zxc_b200_plan_frame(frame.ctypes.data, frame.size, jobs.ctypes.data, nb, None) == nb == n
    dev = torch.device("cuda", 0)
    d_src = torch.from_numpy(frame).to(dev)
    d_dst = torch.empty(data.size, dtype=torch.uint8, device=dev)
    d_jobs = torch.from_numpy(jobs).to(dev)
    d_status = torch.empty(nb, dtype=torch.int32, device=dev)
    d_dict = torch.from_numpy(np.frombuffer(dict_bytes, np.uint8).copy()).to(dev)
    ss = lib.zxc_b200_decode_scratch_size(REC)
    d_scr = torch.empty(ss, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream(dev)

    def step():
        assert lib.zxc_b200_decode_blocks(d_src.data_ptr(), d_dst.data_ptr(), d_jobs.data_ptr(), nb, d_status.data_ptr(),
                                          d_dict.data_ptr(), dsz, None, d_scr.data_ptr(), ss, REC, 0, st.cuda_stream) == 0
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    assert lib.zxc_b200_reduce_status(d_status.data_ptr(), d_jobs.data_ptr(), nb, st.cuda_stream) == data.size
    assert np.array_equal(d_dst.cpu().numpy(), data)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"GPU decode (HBM resident): {data.size / ms / 1e6:.1f} GB/s  ({ms:.3f} ms/step, {nb} blocks);"
          f" (C+U)/t = {(data.size + frame.size) / ms / 1e6:.1f} GB/s")
    out = np.zeros(data.size, np.uint8)
    o = z.DecompressOpts(dict=C.cast(C.c_char_p(dict_bytes), C.c_void_p), dict_size=dsz)
    prod.lib.zxc_decompress(frame.ctypes.data, frame.size, out.ctypes.data, out.size, C.byref(o))
    t = time.perf_counter()
    r = prod.lib.zxc_decompress(frame.ctypes.data, frame.size, out.ctypes.data, out.size, C.byref(o))
    dt = time.perf_counter() - t
    assert r == data.size and np.array_equal(out, data)
    print(f"GPU e2e zxc_decompress (pageable host buffers): {data.size / dt / 1e9:.2f} GB/s")
    h = ref.lib.zxc_seekable_open(frame.ctypes.data, frame.size)
    assert ref.lib.zxc_seekable_set_dict(h, dict_bytes, dsz, None) == 0
    th = zc.host_threads()
    best = 1e9
    for _ in range(3):
        t = time.perf_counter()
        r = ref.lib.zxc_seekable_decompress_range_mt(h, out.ctypes.data, out.size, 0, out.size, th)
        best = min(best, time.perf_counter() - t)
    assert r == data.size
    print(f"reference CPU zxc_seekable_decompress_range_mt, {th} threads: {data.size / best / 1e9:.2f} GB/s")


if __name__ == "__main__":
    main()
