#!/bin/bash
# scratch: ncu launch list + full capture of the top kernels on a mid-size city build
cd /root/repo
mkdir -p gpurun_out
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
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r1b.csv python /tmp/ncu_target.py > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_classify_filtered" -s 22 -c 2 -o gpurun_out/prof_classify2 python /tmp/ncu_target.py > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_emit|k_leaf_min|k_children" -s 36 -c 6 -o gpurun_out/prof_rest2 python /tmp/ncu_target.py > gpurun_out/ncu_full2.log 2>&1
tail -n 2 gpurun_out/ncu_list.log
