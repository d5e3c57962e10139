"""Extracts the judged numbers from an `ncu --set full` report into a small text summary.
Usage: python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep > profiles/<name>.txt"""
import csv
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
KEYS = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warps_eligible.avg.per_cycle_active"]
for r in rows[2:]:
    print("=" * 100)
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"{k:75s} {r[i]} {units[i]}")
    st = []
    for i, k in enumerate(hdr):
        if "smsp__average_warps_issue_stalled" in k and "per_issue_active" in k and "not_issued" not in k:
            try:
                st.append((float(r[i]), k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
            except ValueError:
                pass
    print("warp stall reasons (warps per issue-active cycle): " + ", ".join(f"{n}={v:.2f}" for v, n in sorted(st, reverse=True)[:8]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(src.splitlines()))
try:
    h = next(r for r in srows if "Instructions Executed" in r)
    iS, iE = h.index("Source"), h.index("Instructions Executed")
    c = Counter()
    for r in srows[srows.index(h) + 1:]:
        if len(r) > iE and r[iE].isdigit():
            op = (r[iS].split()[1] if r[iS].startswith("@") else r[iS].split()[0]).split(".")[0]
            c[op] += int(r[iE])
    tot = sum(c.values())
    print("dynamic instruction mix of the first profiled launch (warp instructions, % of total):")
    print("  total", tot, " ".join(f"{op}={100 * n / tot:.1f}%" for op, n in c.most_common(14)))
except StopIteration:
    pass
