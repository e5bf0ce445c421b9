#!/bin/bash
# ncu launch list of one full 16K^3 city build, aggregated per kernel (gpurun -- tools/gpu_launch_list.sh)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cat > /tmp/ncu_city.py <<'PY'
import sys
sys.path.insert(0, ".")
import numpy as np
import __graft_entry__ as g
pkg = g._pkg()
tris = pkg.meshgen.city(256)
v = tris.reshape(-1, 3)
bbox = (v.min(axis=0).astype(np.float64), v.max(axis=0).astype(np.float64))
t = pkg.GeomOctree(tris)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
for it in range(n):
    st = t.build(14, 4, bbox=bbox); t.to_sdag()
print(st["nTotalVoxels"], st["msTotal"], st["msVoxelize"], st["nKernelLaunches"], st["nExactTests"], st["nPairsTotal"])
PY
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_city16k.csv python /tmp/ncu_city.py 1 > gpurun_out/ncu_city_list.log 2>&1
tail -2 gpurun_out/ncu_city_list.log
python - <<'PY'
import csv, collections
with open("gpurun_out/launches_city16k.csv") as f:
    lines = [l for l in f if not l.startswith("==")]
r = csv.DictReader(lines)
agg = collections.OrderedDict()
for row in r:
    name = row["Kernel Name"].split("(")[0]
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    ms = v / 1e6 if unit in ("ns", "nsecond") else (v / 1e3 if unit in ("us", "usecond") else v)
    a = agg.setdefault(name, [0, 0.0, 0.0]); a[0] += 1; a[1] += ms; a[2] = max(a[2], ms)
tot = sum(a[1] for a in agg.values())
with open("gpurun_out/launches_city16k_summary.md", "w") as o:
    o.write("| kernel | launches | total ms | max ms | share |\n|---|---|---|---|---|\n")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        o.write(f"| `{k}` | {a[0]} | {a[1]:.3f} | {a[2]:.3f} | {100*a[1]/tot:.1f}% |\n")
    o.write(f"\nTotal {tot:.1f} ms\n")
print(open("gpurun_out/launches_city16k_summary.md").read())
PY
