#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (markdown table)."""
import collections
import csv
import sys

with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    name = row["Kernel Name"].split("(")[0].replace("void ", "").replace("svb::<unnamed>::", "").replace("<unnamed>::", "")
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    ms = v / 1e6 if unit in ("ns", "nsecond") else (v / 1e3 if unit in ("us", "usecond") else v)
    a = agg.setdefault(name, [0, 0.0, 0.0])
    a[0] += 1
    a[1] += ms
    a[2] = max(a[2], ms)
tot = sum(a[1] for a in agg.values())
n = sum(a[0] for a in agg.values())
print("| kernel | launches | total ms | max ms | share |\n|---|---|---|---|---|")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {a[0]} | {a[1]:.3f} | {a[2]:.3f} | {100 * a[1] / tot:.1f}% |")
print(f"\nTotal {tot:.1f} ms over {n} launches (per-launch times under ncu are cold-cache and serialised: compare shares).")
