#!/bin/bash
# ncu launch list of one build of a bench workload (default terrain_4k), aggregated per kernel
W=${1:-terrain_4k}; NB=${2:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cat > /tmp/ncu_wl.py <<PY
import sys
sys.path.insert(0, ".")
import numpy as np
import bench
pkg = bench.load_pkg()
wl = bench.WORKLOADS["$W"]
tris = pkg.meshgen.make_mesh(wl["mesh"], **wl["kw"])
v = tris.reshape(-1, 3)
bbox = (v.min(axis=0).astype(np.float64), v.max(axis=0).astype(np.float64))
t = pkg.GeomOctree(tris)
for it in range($NB):
    st = t.build(wl["levels"], wl["step"], bbox=bbox); sd = t.to_sdag()
print(st["nTotalVoxels"], st["msTotal"], st["msVoxelize"], st["msDedup"], st["msFinalize"], sd["msSdag"], st["nKernelLaunches"], t.level_sizes())
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$W.csv python /tmp/ncu_wl.py > gpurun_out/ncu_${W}_list.log 2>&1
tail -1 gpurun_out/ncu_${W}_list.log
python - <<PY
import csv, collections
with open("gpurun_out/launches_$W.csv") as f:
    lines = [l for l in f if not l.startswith("==")]
rows = list(csv.DictReader(lines))
rows = rows[len(rows) // 2:] if $NB == 2 else rows         # second build only (warm pool)
agg = collections.OrderedDict()
for row in rows:
    name = row["Kernel Name"].split("(")[0]
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    ms = v / 1e6 if unit in ("ns", "nsecond") else (v / 1e3 if unit in ("us", "usecond") else v)
    a = agg.setdefault(name, [0, 0.0, 0.0]); a[0] += 1; a[1] += ms; a[2] = max(a[2], ms)
tot = sum(a[1] for a in agg.values())
print("| kernel | launches | total ms | max ms | share |\n|---|---|---|---|---|")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
    print(f"| \`{k}\` | {a[0]} | {a[1]:.3f} | {a[2]:.3f} | {100*a[1]/tot:.1f}% |")
print(f"\nTotal {tot:.2f} ms in {sum(a[0] for a in agg.values())} launches")
PY
