"""Development probe (GPU): time zxc_b200_decode_blocks of several builds of the library on the same frame.
python tests/variant_probe.py MiB lib1.so lib2.so ...   (build variants: make OUT=... NVEXTRA=-D...)"""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np, torch
import zxc_ctypes as z, zxc_corpus as zc

mib = int(sys.argv[1])
libs = sys.argv[2:]
bs = 65536
ref = z.ZxcLib(z.REF_SO)
n = mib << 20
data = zc.silesia_shaped(n, seed=1)
frame = zc.compress_ref_mt(ref, data, level=3, block_size=bs)
dev = torch.device("cuda", 0)
d_src = torch.from_numpy(frame).to(dev)
d_ref = torch.from_numpy(data).to(dev)
d_dst = torch.zeros(n, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream(dev)
for path in libs:
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
    ss = lib.zxc_b200_decode_scratch_size(bs)
    d_scr = torch.empty(ss, dtype=torch.uint8, device=dev)
    def step():
        rc = lib.zxc_b200_decode_blocks(d_src.data_ptr(), d_dst.data_ptr(), d_jobs.data_ptr(), nb, d_status.data_ptr(),
                                        None, 0, None, d_scr.data_ptr(), ss, bs, 0, st.cuda_stream)
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
    print("%-40s identical=%s  %.3f ms  %.1f GB/s" % (os.path.basename(path), ok, best, n / best / 1e6), flush=True)
    del d_scr, d_jobs, d_status
