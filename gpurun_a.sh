#!/bin/bash
# scratch: round-1 session-2 first GPU call: gpu tests, bench, voxelizer stats, ncu full for dedup kernels
cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err
tail -c 3000 gpurun_out/bench_a.json
cat > /tmp/stats_target.py <<'PY'
import sys
sys.path.insert(0, "/root/repo")
import __graft_entry__ as g
pkg = g._pkg()
import numpy as np
tris = pkg.meshgen.city(256)
v = tris.reshape(-1, 3)
bbox = (v.min(axis=0).astype(np.float64), v.max(axis=0).astype(np.float64))
t = pkg.GeomOctree(tris)
st = t.build(14, 4, bbox=bbox)
print(st["msTotal"], st["msVoxelize"])
PY
SVB_VX_STATS=1 timeout 600 python /tmp/stats_target.py > gpurun_out/vxstats_city256.log 2>&1
cat > /tmp/ncu_target.py <<'PY'
import sys
sys.path.insert(0, "/root/repo")
import __graft_entry__ as g
pkg = g._pkg()
tris = pkg.meshgen.city(64)
t = pkg.GeomOctree(tris)
for it in range(2):
    st = t.build(12, 3); t.to_sdag()
print(st["nTotalVoxels"], st["msTotal"], st["msVoxelize"], st["nKernelLaunches"], st["nExactTests"], st["nPairsTotal"])
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_leaf_min|k_insert|k_convert|k_assign_k64|k_winner" -c 24 -o gpurun_out/prof_dedup_a python /tmp/ncu_target.py > gpurun_out/ncu_dedup_a.log 2>&1
ncu -i gpurun_out/prof_dedup_a.ncu-rep --page raw --csv > gpurun_out/prof_dedup_a_raw.csv 2>/dev/null
ls -la gpurun_out | tail -8
