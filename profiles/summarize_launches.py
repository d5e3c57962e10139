"""Per-kernel totals of the LAST step in an `ncu --metrics gpu__time_duration.sum --csv` launch list.
Usage: python profiles/summarize_launches.py gpurun_out/launches.csv [first-kernel-of-a-step, default k_prologue]"""
import collections
import csv
import re
import sys

path = sys.argv[1]
marker = sys.argv[2] if len(sys.argv) > 2 else "k_prologue"
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rows = list(csv.DictReader(lines))
names = [x["Kernel Name"] for x in rows]
idx = [i for i, n in enumerate(names) if marker in n]
start = idx[-1] if idx else 0
tot, cnt = collections.OrderedDict(), collections.Counter()
for x in rows[start:]:
    n = re.sub(r"\(.*", "", x["Kernel Name"])
    n = re.sub(r"^void |hvi::", "", n)
    v = float(x["Metric Value"])
    u = x["Metric Unit"]
    v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
    tot[n] = tot.get(n, 0) + v
    cnt[n] += 1
s = sum(tot.values())
print(f"{len(rows)} launches captured; last step starts at launch {start}; {sum(cnt.values())} launches, {s:.1f} us")
print(f"{'total us':>10} {'n':>5} {'avg us':>8} {'share':>6}  kernel")
for n, v in sorted(tot.items(), key=lambda t: -t[1]):
    print(f"{v:10.1f} {cnt[n]:5d} {v / cnt[n]:8.1f} {100 * v / s:5.1f}%  {n[:100]}")
