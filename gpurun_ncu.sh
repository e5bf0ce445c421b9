#!/bin/bash
# scratch: ncu launch list + full capture of the top kernels on a mid-size city build
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
print(st["nTotalVoxels"], st["msTotal"], st["msVoxelize"], st["nKernelLaunches"])
PY
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python gpurun_probe.py city256 2>&1 | tee gpurun_out/probe4.log | grep -E "it[0-9]|mesh"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r1.csv python /tmp/ncu_target.py > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_classify_filtered" -s 16 -c 8 -o gpurun_out/prof_classify python /tmp/ncu_target.py > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_emit|k_leaf_min|k_insert|k_convert|k_children|k_scan_apply" -s 60 -c 30 -o gpurun_out/prof_rest python /tmp/ncu_target.py > gpurun_out/ncu_full2.log 2>&1
tail -n 3 gpurun_out/ncu_list.log gpurun_out/ncu_full.log
