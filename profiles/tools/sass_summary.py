"""Opcode histogram per kernel of libhvi_b200.so (cuobjdump -sass): which kernels carry tcgen05 (UTCHMMA / UTCQMMA,
LDTM / STTM), TMA (UTMALDG / UTMASTG), legacy warp MMA (HMMA), packed FP32 (FFMA2 / FADD2 / FMUL2), cp.async (LDGSTS).
Usage: python profiles/tools/sass_summary.py [path to .so] > profiles/sass_summary.txt"""
import os
import re
import subprocess
import sys
from collections import Counter

so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))),
                                                         "hybridvariationalinference.jl_b200", "csrc", "libhvi_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "SYNCS", "HMMA", "MOVM", "FFMA2", "FADD2", "FMUL2",
        "FFMA", "DFMA", "MUFU", "LDGSTS", "SHFL", "REDUX"]
rows = []
for blk in re.split(r"\n\s*Function : ", out)[1:]:
    name = blk.split("\n", 1)[0].strip()
    c = Counter()
    n = 0
    for line in blk.split("\n"):
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            c[m.group(1)] += 1
            n += 1
    dem = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
    dem = re.sub(r"\(hvi::Cfg.*$", "", dem).replace("void hvi::", "").replace("hvi::", "").replace("(anonymous namespace)::", "")
    dem = re.sub(r"\((unsigned int|int|bool)\)", "", dem)
    rows.append((dem, n, c))
print(f"# static SASS opcode counts per kernel of {os.path.basename(so)} (cuobjdump -sass); {len(rows)} kernels")
print("# kernel | instructions | " + " ".join(KEYS))
agg = Counter()
for dem, n, c in sorted(rows):
    for k in KEYS:
        agg[k] += c[k]
    cells = " ".join(f"{k}={c[k]}" for k in KEYS if c[k])
    print(f"{dem[:110]:110s} {n:6d}  {cells}")
print("# totals: " + " ".join(f"{k}={agg[k]}" for k in KEYS))
