"""Development probe (GPU): decode GB/s per data class of the Silesia-shaped corpus, for the kernel selection the
environment currently asks for (run once with ZXC_B200_UNITS=0 and once with =1).  python tests/class_probe.py"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np, torch
import zxc_ctypes as z, zxc_corpus as zc

class Job(C.Structure):
    _fields_ = [("src_off", C.c_uint64), ("dst_off", C.c_uint64), ("src_len", C.c_uint32), ("dst_cap", C.c_uint32)]

ref = z.ZxcLib(z.REF_SO)
lib = C.CDLL(z.PRODUCT_SO)
lib.zxc_b200_plan_frame.restype = C.c_int64
lib.zxc_b200_plan_frame.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]
lib.zxc_b200_decode_scratch_size.restype = C.c_size_t
lib.zxc_b200_decode_scratch_size.argtypes = [C.c_uint32]
lib.zxc_b200_decode_blocks.restype = C.c_int
lib.zxc_b200_decode_blocks.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p,
                                       C.c_uint32, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32, C.c_int, C.c_void_p]
# segment order and sizes (MB) of oracle/zxc_corpus.c, SURVEY 8(d)-2
SEG = [("dickens", 10.2), ("mozilla", 51.2), ("mr", 10.0), ("nci", 33.6), ("ooffice", 6.2), ("osdb", 10.1), ("reymont", 6.6),
       ("samba", 21.6), ("sao", 7.3), ("webster", 41.5), ("xml", 5.3), ("x-ray", 8.5)]
bs = 65536
dev = torch.device("cuda", 0)
ss = lib.zxc_b200_decode_scratch_size(bs)
d_scr = torch.empty(ss, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream(dev)
pos = 0.0
for name, mb in SEG:
    lo = int((pos + 0.5) * 1e6) >> 20 << 20  # 1 MiB aligned, inside the segment
    pos += mb
    piece = zc.silesia_shaped(4 << 20, seed=1, offset=lo)
    data = np.tile(piece, 64)  # 256 MiB of one class (identical tiles are fine: blocks are independent)
    frame = zc.compress_ref_mt(ref, data, level=3, block_size=bs)
    nb = lib.zxc_b200_plan_frame(frame.ctypes.data, frame.size, None, 0, None)
    jobs = np.zeros(nb * C.sizeof(Job), dtype=np.uint8)
    lib.zxc_b200_plan_frame(frame.ctypes.data, frame.size, jobs.ctypes.data, nb, None)
    d_src = torch.from_numpy(frame).to(dev)
    d_dst = torch.zeros(data.size, dtype=torch.uint8, device=dev)
    d_jobs = torch.from_numpy(jobs).to(dev)
    d_status = torch.zeros(nb, dtype=torch.int32, device=dev)
    def step():
        assert lib.zxc_b200_decode_blocks(d_src.data_ptr(), d_dst.data_ptr(), d_jobs.data_ptr(), nb, d_status.data_ptr(), None, 0, None,
                                          d_scr.data_ptr(), ss, bs, 0, st.cuda_stream) == 0
    step(); torch.cuda.synchronize()
    assert np.array_equal(d_dst.cpu().numpy(), data), name
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("%-8s ratio %.3f  %7.1f GB/s" % (name, frame.size / data.size, data.size / ms / 1e6), flush=True)
