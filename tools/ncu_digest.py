#!/usr/bin/env python
"""Digest an `ncu -i X.ncu-rep --page raw --csv` dump into a small markdown table (one row per launch).

    python tools/ncu_digest.py gpurun_out/prof_X_raw.csv [kernel-substring] > profiles/rN_ncu_X.md

Values are converted to fixed units (ms, GB, %) using the CSV's own unit row."""
import csv
import sys

COLS = [  # (metric, header, target unit)
    ("gpu__time_duration.sum", "time [ms]", "ms"),
    ("dram__bytes_read.sum", "DRAM rd [GB]", "GB"),
    ("dram__bytes_write.sum", "DRAM wr [GB]", "GB"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %", None),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %", None),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %", None),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %", None),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe %", None),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %", None),
    ("launch__registers_per_thread", "regs", None),
    ("launch__grid_size", "grid", None),
    ("smsp__inst_executed.sum", "warp instr", None),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/instr", None),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %", None),
    ("lts__t_sector_hit_rate.pct", "L2 hit %", None),
]
TIME = {"nsecond": 1e-6, "ns": 1e-6, "usecond": 1e-3, "us": 1e-3, "msecond": 1.0, "ms": 1.0, "second": 1e3, "s": 1e3}
BYTES = {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0, "Tbyte": 1e3}


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return None


def main():
    path = sys.argv[1]
    sub = sys.argv[2] if len(sys.argv) > 2 else ""
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    kn = hdr.index("Kernel Name")
    idx = [(m, h, u, hdr.index(m)) for m, h, u in COLS if m in hdr]
    print("| kernel | " + " | ".join(h for _, h, _, _ in idx) + " |")
    print("|---|" + "---|" * len(idx))
    for r in rows[2:]:
        if sub and sub not in r[kn]:
            continue
        name = r[kn].split("(")[0].replace("void ", "").replace("svb::<unnamed>::", "").replace("<unnamed>::", "")
        cells = []
        for m, h, u, i in idx:
            v = num(r[i])
            if v is None:
                cells.append(r[i])
                continue
            if u == "ms":
                v *= TIME.get(units[i], 1.0)
                cells.append(f"{v:.3f}")
            elif u == "GB":
                v *= BYTES.get(units[i], 1.0)
                cells.append(f"{v:.3f}")
            elif h in ("regs", "grid", "warp instr"):
                cells.append(f"{v:.0f}")
            else:
                cells.append(f"{v:.1f}")
        print(f"| `{name}` | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main()
