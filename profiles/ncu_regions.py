"""Instruction / stall share per source region of zxc_decode.cuh from an ncu capture taken with --import-source on.
python profiles/ncu_regions.py report.ncu-rep"""
import csv, sys, subprocess, collections
rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
import re, os
SRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "zxc_b200", "csrc", "zxc_decode.cuh")
marks = []  # (line, name): function definitions and the "---- section ----" comments of decode_lz_block
for i, line in enumerate(open(SRC), 1):
    m = re.match(r"(?:template <[^>]*>\s*)?(?:__device__|__global__|static inline)[^(]*?(\w+)\(", line)
    if m and not line.startswith(" "): marks.append((i, m.group(1)))
    m = re.match(r"\s+/\* ---- (.*?)(?: ----|:|\(|$)", line)
    if m: marks.append((i, "· " + m.group(1).strip()[:40]))
marks.sort()
REG = [(1, marks[0][0] - 1, "preamble")] + [(a, (marks[k + 1][0] - 1) if k + 1 < len(marks) else 10 ** 6, n) for k, (a, n) in enumerate(marks)]
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
seen = set()
for lo, hi, name in REG:
    if agg[name] and name not in seen: print(f"{100*agg[name]/tot:5.1f}% inst {100*st[name]/tots:5.1f}% stall | {name} (L{lo}-)"); seen.add(name)
for k, n in other.most_common(10): print(f"{100*n/tot:5.1f}% inst {100*ost[k]/tots:5.1f}% stall | file {k}")
