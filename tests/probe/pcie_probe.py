"""Development probe (GPU): what the PCIe link of this box gives page-locked copies -- the ceiling of bench.py's e2e leg.
D2H alone, H2D alone, both directions at once (the pipeline's steady state), whole-buffer and in 64 MiB pieces.
python tests/probe/pcie_probe.py [GiB out] [GiB in]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
from bench import bind_to_gpu_numa

out_b = int(float(sys.argv[1]) * (1 << 30)) if len(sys.argv) > 1 else 4 << 30
in_b = int(float(sys.argv[2]) * (1 << 30)) if len(sys.argv) > 2 else int(1.62 * (1 << 30))
torch.cuda.set_device(0)
print("numa:", bind_to_gpu_numa(0))
dev = torch.device("cuda", 0)
h_out = torch.empty(out_b, dtype=torch.uint8).pin_memory()
h_in = torch.empty(in_b, dtype=torch.uint8).pin_memory()
h_in.zero_(); h_out.zero_()
d_out = torch.empty(out_b, dtype=torch.uint8, device=dev)
d_in = torch.empty(in_b, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
CH = 64 << 20


def run(name, fn, nbytes, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        for s in (s1, s2):
            torch.cuda.current_stream().wait_stream(s)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print("%-44s %8.2f ms  %6.1f GB/s" % (name, best, nbytes / best / 1e6), flush=True)


def d2h_whole():
    s1.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s1):
        h_out.copy_(d_out, non_blocking=True)


def d2h_chunks():
    s1.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s1):
        for o in range(0, out_b, CH):
            h_out[o:o + CH].copy_(d_out[o:o + CH], non_blocking=True)


def h2d_whole():
    s2.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s2):
        d_in.copy_(h_in, non_blocking=True)


def both():
    d2h_chunks()
    s2.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s2):
        ci = CH * in_b // out_b
        for o in range(0, in_b, ci):
            d_in[o:o + ci].copy_(h_in[o:o + ci], non_blocking=True)


run("D2H whole buffer", d2h_whole, out_b)
run("D2H 64 MiB pieces", d2h_chunks, out_b)
run("H2D whole buffer", h2d_whole, in_b)
run("D2H + H2D together (rate of the D2H bytes)", both, out_b)
