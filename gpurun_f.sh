#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
cat > /tmp/ncu_city.py <<'PY'
import sys
sys.path.insert(0, "/root/repo")
import numpy as np
import __graft_entry__ as g
pkg = g._pkg()
tris = pkg.meshgen.city(256)
v = tris.reshape(-1, 3)
bbox = (v.min(axis=0).astype(np.float64), v.max(axis=0).astype(np.float64))
t = pkg.GeomOctree(tris)
st = t.build(14, 4, bbox=bbox)
print(st["nTotalVoxels"], st["msTotal"], st["msVoxelize"], st["nKernelLaunches"], st["nExactTests"], st["nPairsTotal"])
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_flat_leaves" -s 3 -c 2 -o gpurun_out/prof_flatleaves python /tmp/ncu_city.py > gpurun_out/ncu_f1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_classify_filtered" -s 28 -c 8 -o gpurun_out/prof_slow python /tmp/ncu_city.py > gpurun_out/ncu_f2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_emit" -s 26 -c 9 -o gpurun_out/prof_emit python /tmp/ncu_city.py > gpurun_out/ncu_f3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_classify_fast|k_children" -s 24 -c 9 -o gpurun_out/prof_fast python /tmp/ncu_city.py > gpurun_out/ncu_f4.log 2>&1
for f in prof_flatleaves prof_slow prof_emit prof_fast; do
  ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/${f}_raw.csv 2>/dev/null
  ncu -i gpurun_out/$f.ncu-rep --page source --csv --print-source sass > gpurun_out/${f}_sass.csv 2>/dev/null
  rm -f gpurun_out/$f.ncu-rep
done
gzip -f gpurun_out/*_sass.csv
ls -la gpurun_out | tail -12
du -sh gpurun_out
