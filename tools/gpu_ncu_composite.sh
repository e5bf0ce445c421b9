#!/bin/bash
# ncu --set full of the general-triangle classify (FLATONLY = 0) deep inside the 65536^3 composite build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cat > /tmp/ncu_comp.py <<'PY'
import sys
sys.path.insert(0, ".")
import numpy as np
import bench
pkg = bench.load_pkg()
wl = bench.WORKLOADS["composite_64k"]
tris = pkg.meshgen.make_mesh(wl["mesh"], **wl["kw"])
v = tris.reshape(-1, 3)
bbox = (v.min(axis=0).astype(np.float64), v.max(axis=0).astype(np.float64))
t = pkg.GeomOctree(tris)
st = t.build(wl["levels"], wl["step"], bbox=bbox)
print(st["nTotalVoxels"], st["msTotal"], st["nBatches"])
PY
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_classify_filtered" -s 400 -c 8 -o gpurun_out/tmp_comp python /tmp/ncu_comp.py > gpurun_out/ncu_composite_classify.log 2>&1
ncu -i gpurun_out/tmp_comp.ncu-rep --page raw --csv > gpurun_out/ncu_composite_classify_raw.csv 2>/dev/null
rm -f gpurun_out/tmp_comp.ncu-rep
tail -2 gpurun_out/ncu_composite_classify.log
