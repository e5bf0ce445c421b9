#!/bin/bash
# ncu --set full of the deepest k_emit / k_children / k_flat_leaves launches of the first tile batch (city 16K^3); raw CSV -> gpurun_out/
TAG=${1:-r1e}
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
st = t.build(14, 4, bbox=bbox)
print(st["nTotalVoxels"], st["msTotal"], st["msVoxelize"], st["nKernelLaunches"])
PY
cap() {  # name regex skip count
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -o gpurun_out/tmp_$1 python /tmp/ncu_city.py > gpurun_out/ncu_$1_${TAG}.log 2>&1
  ncu -i gpurun_out/tmp_$1.ncu-rep --page raw --csv > gpurun_out/ncu_$1_${TAG}_raw.csv 2>/dev/null
  rm -f gpurun_out/tmp_$1.ncu-rep
  tail -2 gpurun_out/ncu_$1_${TAG}.log
}
cap emit "k_emit|k_children" ${2:-40} ${3:-3}
cap flat "k_flat_leaves" 0 1
du -sh gpurun_out
