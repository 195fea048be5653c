"""Instruction / stall share per source region of zxc_decode.cuh from an ncu capture taken with --import-source on.
python profiles/ncu_regions.py report.ncu-rep"""
import csv, sys, subprocess, collections
rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
REG = [(1, 160, "helpers ld16/ld32/scan/varint"), (161, 217, "checksum"), (218, 241, "rle"), (242, 357, "parse_sections"),
       (358, 378, "window_byte"), (379, 405, "ring_flush"), (406, 438, "warp_copy_words (unused)"), (439, 468, "slow match paths"),
       (469, 515, "lane_copy_words"), (516, 568, "group4_copy_words"), (569, 645, "block prologue/extras prepass"),
       (646, 668, "unpack tok/off"), (669, 691, "escapes"), (692, 703, "scans"), (704, 731, "giant"), (732, 746, "validation"),
       (747, 764, "classify"), (765, 786, "depmask"), (787, 808, "pass loop head"), (809, 809, "call lane_copy"),
       (810, 811, "grp ballot/call"), (812, 824, "slow dispatch"), (825, 829, "pass loop tail"), (830, 849, "advance/flush call"),
       (850, 858, "trailing literals"), (859, 888, "decode_job"), (889, 943, "kernel loop"), (944, 2000, "other")]
cur_file = None
agg = collections.Counter(); st = collections.Counter(); other = collections.Counter(); ost = collections.Counter()
hdr = None
for x in rows:
    if not x: continue
    if x[0] == "File Path": cur_file = x[1]; continue
    if x[0] == "Line No": hdr = x; iE = hdr.index("Instructions Executed"); iW = hdr.index("Warp Stall Sampling (All Samples)"); continue
    if hdr is None or len(x) <= iE: continue
    try: ln = int(x[0]); n = int(x[iE]); w = int(x[iW])
    except ValueError: continue
    if cur_file and cur_file.endswith("zxc_decode.cuh"):
        for lo, hi, name in REG:
            if lo <= ln <= hi: agg[name] += n; st[name] += w; break
    else:
        k = (cur_file or "?").split("/")[-1]
        other[k] += n; ost[k] += w
tot = sum(agg.values()) + sum(other.values()); tots = (sum(st.values()) + sum(ost.values())) or 1
print("total inst", tot)
for lo, hi, name in REG:
    if agg[name]: print(f"{100*agg[name]/tot:5.1f}% inst {100*st[name]/tots:5.1f}% stall | {name} (L{lo}-{hi})")
for k, n in other.most_common(10): print(f"{100*n/tot:5.1f}% inst {100*ost[k]/tots:5.1f}% stall | file {k}")
