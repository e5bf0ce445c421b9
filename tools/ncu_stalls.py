#!/usr/bin/env python
"""Per-launch digest of an `ncu --page raw --csv` dump: time, the warp stall reasons (per issue-active), memory-pipe
counters.  python tools/ncu_stalls.py gpurun_out/ncu_X_raw.csv [substring-filter ...]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
kn = hdr.index("Kernel Name")
pats = sys.argv[2:]


def f(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return None


for r in rows[2:]:
    name = r[kn].split("(")[0]
    print("==", name, "grid", r[hdr.index("launch__grid_size")] if "launch__grid_size" in hdr else "?")
    st = []
    for i, h in enumerate(hdr):
        if "issue_stalled" in h and h.endswith("_per_issue_active.ratio"):
            v = f(r[i])
            if v:
                st.append((v, h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
    st.sort(reverse=True)
    print("  stalls (warps per issue-active cycle):", ", ".join(f"{n} {v:.2f}" for v, n in st[:9]))
    for i, h in enumerate(hdr):
        if any(p in h for p in pats) or h in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
                                               "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
                                               "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_active",
                                               "sm__warps_active.avg.pct_of_peak_sustained_active"):
            print(f"  {h} = {r[i]} {units[i]}")
