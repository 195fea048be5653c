"""Development probe (GPU): decode a Silesia-shaped frame through zxc_b200_decode_blocks, check bytes, time it.
python tests/d2_quick.py [MiB] [level] [block_size] [iters]"""
import ctypes as C, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np, torch
import zxc_ctypes as z, zxc_corpus as zc

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 256
level = int(sys.argv[2]) if len(sys.argv) > 2 else 3
bs = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 10

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
n = mib << 20
data = zc.silesia_shaped(n, seed=1)
frame = zc.compress_ref_mt(ref, data, level=level, block_size=bs)
nb = lib.zxc_b200_plan_frame(frame.ctypes.data, frame.size, None, 0, None)
jobs = np.zeros(nb * C.sizeof(Job), dtype=np.uint8)
assert lib.zxc_b200_plan_frame(frame.ctypes.data, frame.size, jobs.ctypes.data, nb, None) == nb
dev = torch.device("cuda", 0)
d_src = torch.from_numpy(frame).to(dev)
d_dst = torch.zeros(n, dtype=torch.uint8, device=dev)
d_jobs = torch.from_numpy(jobs).to(dev)
d_status = torch.zeros(nb, dtype=torch.int32, device=dev)
ss = lib.zxc_b200_decode_scratch_size(bs)
d_scr = torch.empty(ss, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream(dev)
def step():
    rc = lib.zxc_b200_decode_blocks(d_src.data_ptr(), d_dst.data_ptr(), d_jobs.data_ptr(), nb, d_status.data_ptr(),
                                    None, 0, None, d_scr.data_ptr(), ss, bs, 0, st.cuda_stream)
    assert rc == 0, rc
step(); torch.cuda.synchronize()
stt = d_status.cpu().numpy()
jv = jobs.view(np.dtype([("src_off", "<u8"), ("dst_off", "<u8"), ("src_len", "<u4"), ("dst_cap", "<u4")]))
bad = np.nonzero(stt != jv["dst_cap"].astype(np.int64))[0]
usable = (ss - 32) & ~7
tail = d_scr[usable:usable + 32].cpu().numpy().view(np.uint32)
print("blocks", nb, "bad status", bad.size, stt[bad[:8]] if bad.size else "", "deferred by the block-cooperative kernel:", int(tail[4]))
got = d_dst.cpu().numpy()
if not np.array_equal(got, data):
    w = np.nonzero(got != data)[0]
    print("MISMATCH bytes", w.size, "first", w[:10], "blocks", np.unique(w // bs)[:20])
    sys.exit(1)
print("bytes identical")
def timed():
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): step()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
ms = timed()
print("ms/step %.3f  decode GB/s %.1f  (C+U) GB/s %.1f" % (ms, n / ms / 1e6, (n + frame.size) / ms / 1e6))
for x in os.environ.get("D2_EXP_LIST", "").split(","):  # development switches (ZXC_B200_EXP), A/B in one process
    if not x: continue
    os.environ["ZXC_B200_EXP"] = x
    step(); torch.cuda.synchronize()
    ok = bool(torch.equal(d_dst, torch.from_numpy(data).to(dev)))
    ms = min(timed(), timed())
    print("EXP=%s: identical=%s ms/step %.3f  decode GB/s %.1f" % (x, ok, ms, n / ms / 1e6))
