"""Development probe (GPU): decode-only rate of the dictionary-records workload (BASELINE configs[3] shape) for several
builds of the library; ZXC_B200_UNITS=0/1 in the environment forces the sequence-centric / output-centric body.
python tests/dict_variant_probe.py RECORDS lib1.so [lib2.so ...]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np, torch
import zxc_corpus as zc, zxc_ctypes as z

n = int(sys.argv[1])
REC = 4096
ref = z.ZxcLib(z.REF_SO)
prod = z.ZxcLib(z.PRODUCT_SO)
data = zc.records(n, record_size=REC, seed=7)
d = zc.train_dict_ref(ref, data, record_size=REC, n_samples=4096, cap=16384)
frame = prod.compress(data, level=5, block_size=REC, seekable=1, dict=d)
assert not isinstance(frame, int), frame
dev = torch.device("cuda", 0)
d_src = torch.from_numpy(frame).to(dev)
d_ref = torch.from_numpy(data).to(dev)
d_dst = torch.zeros(data.size, dtype=torch.uint8, device=dev)
d_dict = torch.from_numpy(np.frombuffer(d, np.uint8).copy()).to(dev)
st = torch.cuda.current_stream(dev)
for path in sys.argv[2:]:
    lib = C.CDLL(os.path.abspath(path))
    lib.zxc_b200_plan_frame.restype = C.c_int64
    lib.zxc_b200_plan_frame.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.zxc_b200_decode_scratch_size.restype = C.c_size_t
    lib.zxc_b200_decode_scratch_size.argtypes = [C.c_uint32]
    lib.zxc_b200_decode_blocks.restype = C.c_int
    lib.zxc_b200_decode_blocks.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p,
                                           C.c_uint32, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32, C.c_int, C.c_void_p]
    nb = lib.zxc_b200_plan_frame(frame.ctypes.data, frame.size, None, 0, None)
    jobs = np.zeros(nb * 24, dtype=np.uint8)
    assert lib.zxc_b200_plan_frame(frame.ctypes.data, frame.size, jobs.ctypes.data, nb, None) == nb
    d_jobs = torch.from_numpy(jobs).to(dev)
    d_status = torch.zeros(nb, dtype=torch.int32, device=dev)
    ss = lib.zxc_b200_decode_scratch_size(REC)
    d_scr = torch.empty(ss, dtype=torch.uint8, device=dev)

    def step():
        rc = lib.zxc_b200_decode_blocks(d_src.data_ptr(), d_dst.data_ptr(), d_jobs.data_ptr(), nb, d_status.data_ptr(),
                                        d_dict.data_ptr(), len(d), None, d_scr.data_ptr(), ss, REC, 0, st.cuda_stream)
        assert rc == 0, rc
    d_dst.zero_()
    step(); torch.cuda.synchronize()
    ok = bool(torch.equal(d_dst, d_ref))
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): step()
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 5)
    print("dict records %d  UNITS=%s  %-44s identical=%s  %.3f ms  %.1f GB/s" % (n, os.environ.get("ZXC_B200_UNITS", "rule"), path[-44:], ok, best, data.size / best / 1e6), flush=True)
    del d_scr, d_jobs, d_status
