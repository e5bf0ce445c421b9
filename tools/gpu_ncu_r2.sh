#!/bin/bash
# ncu --set full of the round-2 kernels on the THIRD tile batch of the 16K^3 city build (a "later" batch: the common case)
TAG=${1:-r2h}
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
t.set_profiling(True)
st = t.build(14, 4, bbox=bbox)
print(st["nTotalVoxels"], st["msTotal"], st["msVoxelize"], st["nKernelLaunches"], st["nBatches"])
for i, r in enumerate(t.profile()):
    print("PROFREC", i, r["name"], r["level"], r["n_in"], r["n_out"], r["bytes"], round(r["ms"], 4))
PY
cap() {  # name regex skip count
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -o gpurun_out/tmp_$1 python /tmp/ncu_city.py > gpurun_out/ncu_$1_${TAG}.log 2>&1
  ncu -i gpurun_out/tmp_$1.ncu-rep --page raw --csv > gpurun_out/ncu_$1_${TAG}_raw.csv 2>/dev/null
  rm -f gpurun_out/tmp_$1.ncu-rep
  tail -1 gpurun_out/ncu_$1_${TAG}.log
}
# emit / children launches per batch: 8 levels x (children, emit slow [, emit flat]); batches 0..2 are the warm-up ones + base.
# capture a window of 40 launches well inside the 4th batch
cap emit "k_emit_warp|k_children" 60 40
cap dedup "k_leaf_known|k_insert|k_winner|k_convert|k_k64_query" 40 24
du -sh gpurun_out
