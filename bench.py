#!/usr/bin/env python3
"""bench.py -- ZXC block decode on B200: decompress GB/s (uncompressed bytes) and HBM roofline.

One "step" = one pass of the hot path (decode every block of the shard) over synthetic input.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--gib G]

Workload (BASELINE.json configs[1]): seekable ZXC frame, 64 KiB blocks, level 3, over the
Silesia-shaped synthetic corpus (oracle/zxc_corpus.c), G GiB decoded per GPU (default 4).
Frames are produced by the UNMODIFIED reference encoder (oracle/_ref);
at N > 1 every rank owns the contiguous block range [rank*G GiB, (rank+1)*G GiB) of the frame
(weak scaling: independent seekable blocks, no data-path collective).

  value        decode-only, compressed input and output resident in HBM, CUDA events on the
               launching stream, max over ranks; inputs (G GiB + its frame) are far larger than L2.
  e2e          same metric through the reference-facing C ABI call zxc_decompress() with HOST
               (pinned) buffers: H2D of the frame and D2H of the output inside the timed region.
  roofline     algorithmic bytes per launch (C + U: on-disk block bytes read once + decoded bytes
               written once) / average launch time, vs MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline the reference's own SIMD CPU path (zxc_seekable_decompress_range_mt, all host
               threads; zxc_decompress 1 thread) on the same frame, same run, rank 0 at N=1.
  --impl reference   times only that CPU path, same metric / config.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

BLOCK = 65536
LEVEL = 3


class Job(C.Structure):
    _fields_ = [("src_off", C.c_uint64), ("dst_off", C.c_uint64), ("src_len", C.c_uint32), ("dst_cap", C.c_uint32)]


class Info(C.Structure):
    _fields_ = [("decoded_size", C.c_uint64), ("block_size", C.c_uint32), ("n_blocks", C.c_uint32),
                ("dict_id", C.c_uint32), ("has_checksum", C.c_int), ("seekable", C.c_int), ("global_hash", C.c_uint32)]


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock and throttle reasons during the timed region (B200_PROFILING.md recipe).  NVML is polled
    every 2 ms from a thread (the timed region is ~0.15 s, too short for nvidia-smi's 100 ms loop);
    nvidia-smi is the fallback when the NVML binding is missing."""

    REASONS = (("hw_slowdown", 0x8), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40), ("sw_power_cap", 0x4))

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []  # one (sm_mhz, reasons_bitmask) per sample
        self.max_mhz = None
        self.stop_flag = False
        self.nvml = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv = self.nvml
        while not self.stop_flag:
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                except Exception:
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                self.lines.append((mhz, mask))
            except Exception:
                pass
            time.sleep(0.002)

    def _pump(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                mhz = float(f[0])
                self.max_mhz = float(f[1])
            except ValueError:
                continue
            mask = 0
            for (name, bit), v in zip((("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20),
                                       ("sw_power_cap", 0x4)), f[3:7]):
                if v.lower().startswith("active"):
                    mask |= bit
            self.lines.append((mhz, mask))

    def stop(self):
        self.stop_flag = True
        if self.nvml is None and not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no NVML / nvidia-smi"]}
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        else:
            self.t.join(timeout=1)
        sm = [m for m, _ in self.lines]
        mask = 0
        for _, k in self.lines:
            mask |= k
        reasons = sorted(name for name, bit in self.REASONS if mask & bit)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(sm), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def build_shard(ref, gib, rank, seed=1):
    """(data ndarray, frame ndarray) for this rank's slice of the Silesia-shaped stream."""
    import zxc_corpus as zc
    n = int(gib * (1 << 30))
    n -= n % (1 << 20)
    t0 = time.time()
    data = zc.silesia_shaped(n, seed=seed, offset=rank * n)
    t1 = time.time()
    frame = zc.compress_ref_mt(ref, data, level=LEVEL, block_size=BLOCK, checksum=0)
    t2 = time.time()
    return data, frame, {"gen_s": round(t1 - t0, 2), "ref_compress_s": round(t2 - t1, 2)}


def cpu_reference_decode(ref, frame, n, threads, reps):
    """best-of-reps GB/s of zxc_seekable_decompress_range_mt over the whole frame (output pre-faulted)."""
    out = np.zeros(n, dtype=np.uint8)
    h = ref.lib.zxc_seekable_open(frame.ctypes.data, frame.size)
    assert h
    best = None
    for _ in range(reps):
        t = time.perf_counter()
        r = ref.lib.zxc_seekable_decompress_range_mt(h, out.ctypes.data, n, 0, n, threads)
        dt = time.perf_counter() - t
        assert r == n, r
        best = dt if best is None or dt < best else best
    ref.lib.zxc_seekable_free(h)
    return n / best / 1e9, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--gib", type=float, default=4.0, help="decoded GiB per GPU")
    ap.add_argument("--no-verify", action="store_true")
    ap.add_argument("--decode-only", action="store_true", help="development: skip the cpu_baseline and encode legs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    import zxc_corpus as zc
    import zxc_ctypes as z

    if not os.path.exists(z.REF_SO):
        raise SystemExit("oracle/_ref/libzxc_ref.so missing: run __graft_entry__.build() where /root/reference exists")
    ref = z.ZxcLib(z.REF_SO)
    threads = zc.host_threads()
    config = {"workload": f"seekable decode, {args.gib:g} GiB/GPU Silesia-shaped synthetic, 64 KiB blocks, level 3",
              "block_size": BLOCK, "level": LEVEL, "gib_per_gpu": args.gib, "sharding": f"block-range x{world}",
              "l2_policy": "inputs (frame + output) >> 126 MB L2, no flush needed"}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        data, frame, prep = build_shard(ref, args.gib, 0)
        n = data.size
        out = np.zeros(n, dtype=np.uint8)
        h = ref.lib.zxc_seekable_open(frame.ctypes.data, frame.size)
        for _ in range(args.warmup):
            ref.lib.zxc_seekable_decompress_range_mt(h, out.ctypes.data, n, 0, n, threads)
        t = time.perf_counter()
        for _ in range(args.steps):
            r = ref.lib.zxc_seekable_decompress_range_mt(h, out.ctypes.data, n, 0, n, threads)
            assert r == n
        dt = (time.perf_counter() - t) / args.steps
        ref.lib.zxc_seekable_free(h)
        assert np.array_equal(out, data)
        gbs = n / dt / 1e9
        line = {"impl": "reference", "metric": "decompress GB/s (uncompressed)", "value": round(gbs, 3), "unit": "GB/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt * 1e3, 3),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": config,
                "cpu_baseline": {"value": round(gbs, 3), "unit": "GB/s", "cores": threads, "kind": "reference",
                                 "sample": f"whole {args.gib:g} GiB frame per step, zxc_seekable_decompress_range_mt, {threads} threads"},
                "e2e": {"value": round(gbs, 3), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0, "prep": prep}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ our arm (GPU)
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    lib = C.CDLL(z.PRODUCT_SO)  # fails loudly if the CUDA library is missing
    prod = z.ZxcLib(z.PRODUCT_SO)
    lib.zxc_b200_plan_frame.restype = C.c_int64
    lib.zxc_b200_plan_frame.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.zxc_b200_decode_scratch_size.restype = C.c_size_t
    lib.zxc_b200_decode_scratch_size.argtypes = [C.c_uint32]
    lib.zxc_b200_decode_blocks.restype = C.c_int
    lib.zxc_b200_decode_blocks.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p,
                                           C.c_uint32, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32, C.c_int, C.c_void_p]
    lib.zxc_b200_reduce_status.restype = C.c_int64
    lib.zxc_b200_reduce_status.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
    lib.zxc_b200_launch_count.restype = C.c_uint64

    data, frame, prep = build_shard(ref, args.gib, rank)
    n = data.size
    info = Info()
    nb = lib.zxc_b200_plan_frame(frame.ctypes.data, frame.size, None, 0, C.byref(info))
    assert nb > 0 and info.decoded_size == n, (nb, info.decoded_size)
    jobs = np.zeros(nb * C.sizeof(Job), dtype=np.uint8)
    assert lib.zxc_b200_plan_frame(frame.ctypes.data, frame.size, jobs.ctypes.data, nb, None) == nb
    jv = jobs.view(np.dtype([("src_off", "<u8"), ("dst_off", "<u8"), ("src_len", "<u4"), ("dst_cap", "<u4")]))
    comp_bytes = int(jv["src_len"].astype(np.int64).sum())
    algo_bytes = comp_bytes + n  # C + U per launch

    h_frame = torch.from_numpy(frame).pin_memory()
    d_src = h_frame.to(dev, non_blocking=True)
    d_dst = torch.empty(n, dtype=torch.uint8, device=dev)
    d_jobs = torch.from_numpy(jobs).to(dev)
    d_status = torch.empty(nb, dtype=torch.int32, device=dev)
    scratch_size = lib.zxc_b200_decode_scratch_size(BLOCK)
    d_scratch = torch.empty(scratch_size, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev)

    def step():
        rc = lib.zxc_b200_decode_blocks(d_src.data_ptr(), d_dst.data_ptr(), d_jobs.data_ptr(), nb, d_status.data_ptr(),
                                        None, 0, None, d_scratch.data_ptr(), scratch_size, BLOCK, 0, stream.cuda_stream)
        assert rc == 0, rc

    sampler = ClockSampler(local_rank)
    sampler.start()  # nvidia-smi takes ~0.5 s to deliver its first line: start it ahead of the warm-up
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize(dev)
    total = lib.zxc_b200_reduce_status(d_status.data_ptr(), d_jobs.data_ptr(), nb, stream.cuda_stream)
    assert total == n, f"decode verdict {total} != {n}"
    if not args.no_verify:
        got = d_dst.cpu().numpy()
        assert np.array_equal(got, data), "decoded bytes differ from the original"
        del got

    t_wait = time.time()
    while len(sampler.lines) < 2 and time.time() - t_wait < 3.0:
        step()  # keep the GPU under load until the sampler is live (untimed)
        torch.cuda.synchronize(dev)
    sampler.lines.clear()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    launches0 = lib.zxc_b200_launch_count()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    evs[0].record(stream)
    for i in range(args.steps):
        step()
        evs[i + 1].record(stream)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    launches = int(lib.zxc_b200_launch_count() - launches0)
    total_ms = evs[0].elapsed_time(evs[-1])
    per_launch_ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)]
    clocks = sampler.stop()

    t_ms = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(t_ms.item()) / args.steps
    value = (n * world) / (ms_per_step * 1e-3) / 1e9

    # ---- e2e through the C ABI with host (pinned) buffers: H2D + decode + D2H every step
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    e2e_steps = max(2, min(args.steps, 5))
    for _ in range(2):
        r = prod.lib.zxc_decompress(h_frame.data_ptr(), h_frame.numel(), h_out.data_ptr(), n, None)
        assert r == n, r
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        r = prod.lib.zxc_decompress(h_frame.data_ptr(), h_frame.numel(), h_out.data_ptr(), n, None)
    torch.cuda.synchronize(dev)
    e2e_dt = (time.perf_counter() - t0) / e2e_steps
    assert r == n
    if not args.no_verify:
        assert np.array_equal(h_out.numpy(), data), "e2e output differs"
    t_e = torch.tensor([e2e_dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_value = (n * world) / float(t_e.item()) / 1e9

    # ---- supplementary: NVLink gather of decoded output (N > 1), bounded slice
    gather = None
    if world > 1:
        sl = min(n, 1 << 30)
        outs = torch.empty(sl * world, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(outs, d_dst[:sl])
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier()
        e0.record()
        for _ in range(3):
            dist.all_gather_into_tensor(outs, d_dst[:sl])
        e1.record()
        torch.cuda.synchronize(dev)
        g_ms = torch.tensor([e0.elapsed_time(e1) / 3], dtype=torch.float64, device=dev)
        dist.all_reduce(g_ms, op=dist.ReduceOp.MAX)
        gather = {"collective": "nccl all_gather of decoded ranges", "bytes_per_rank": sl,
                  "gbs_per_rank_in": round(sl * (world - 1) / (float(g_ms.item()) * 1e-3) / 1e9, 1)}
        del outs

    if rank == 0:
        peak, peak_src = measured_peak()
        avg_launch_ms = float(np.mean(per_launch_ms))
        achieved = algo_bytes / (avg_launch_ms * 1e-3) / 1e9
        line = {"metric": "decompress GB/s (uncompressed)", "value": round(value, 2), "unit": "GB/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": config,
                "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                             "frac": round(achieved / peak, 4), "traffic": None, "peak_source": peak_src,
                             "algorithmic_bytes_per_launch": algo_bytes, "compressed_bytes": comp_bytes,
                             "decoded_bytes": n, "kernel": "zxc_decode_kernel", "avg_launch_ms": round(avg_launch_ms, 4),
                             "decoded_only_frac": round((n / (avg_launch_ms * 1e-3) / 1e9) / peak, 4)},
                "e2e": {"value": round(e2e_value, 2), "unit": "GB/s", "h2d_bytes_per_step": int(frame.size),
                        "d2h_bytes_per_step": int(n), "api": "zxc_decompress(host frame, host dst), pinned host buffers",
                        "steps": e2e_steps},
                "gpu_launches": launches, "clocks": clocks, "ratio": round(frame.size / n, 4), "blocks_per_gpu": int(nb),
                "prep": prep}
        if gather:
            line["gather"] = gather
        if world == 1 and not args.decode_only:
            reps = 3
            mt, out = cpu_reference_decode(ref, frame, n, threads, reps)
            sample_n = min(n, 256 << 20)
            sf = zc.compress_ref_mt(ref, data[:sample_n], level=LEVEL, block_size=BLOCK)
            o1 = np.zeros(sample_n, dtype=np.uint8)
            t = time.perf_counter()
            r1 = ref.lib.zxc_decompress(sf.ctypes.data, sf.size, o1.ctypes.data, sample_n, None)
            st = sample_n / (time.perf_counter() - t) / 1e9
            assert r1 == sample_n
            line["cpu_baseline"] = {"value": round(mt, 3), "unit": "GB/s", "cores": threads, "kind": "reference",
                                    "sample": f"whole {args.gib:g} GiB frame, zxc_seekable_decompress_range_mt best of {reps}",
                                    "single_thread_gbs": round(st, 3)}
            # ---- supplementary: the encoder, BASELINE.json configs[2] (level 6, 1 GiB, frame bit-exact vs the
            # reference) and the same input at level 3; through zxc_compress with host buffers
            enc_n = min(n, 1 << 30)
            src_v = data[:enc_n]
            cap = int(prod.lib.zxc_compress_bound(enc_n))
            enc_out = np.zeros(cap, dtype=np.uint8)

            def encode_leg(level):
                o = z.CompressOpts(level=level, block_size=BLOCK, seekable=1)
                r_enc = prod.lib.zxc_compress(src_v.ctypes.data, enc_n, enc_out.ctypes.data, cap, C.byref(o))  # warm-up
                t = time.perf_counter()
                r_enc = prod.lib.zxc_compress(src_v.ctypes.data, enc_n, enc_out.ctypes.data, cap, C.byref(o))
                enc_dt = time.perf_counter() - t
                t = time.perf_counter()
                ref_frame = zc.compress_ref_mt(ref, src_v, level=level, block_size=BLOCK)
                ref_dt = time.perf_counter() - t
                return {"level": level, "bytes_in": int(enc_n), "gbs_in_e2e": round(enc_n / enc_dt / 1e9, 3),
                        "identical_to_reference": bool(r_enc == ref_frame.size and np.array_equal(enc_out[:r_enc], ref_frame)),
                        "ratio": round(r_enc / enc_n, 4),
                        "cpu_reference_gbs_in": round(enc_n / ref_dt / 1e9, 3), "cpu_threads": threads}

            line["encode"] = encode_leg(6)
            line["encode"]["note"] = "configs[2]: optimal parser + Huffman sections on the GPU; levels 1-7 all encode on the GPU"
            line["encode"]["level3"] = encode_leg(LEVEL)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
