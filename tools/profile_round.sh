#!/bin/bash
# Round profile: run on the GPU box (gpurun -- tools/profile_round.sh [tag]).  Produces, under gpurun_out/:
#   launches_bench_<tag>.csv/.md   ncu launch list (gpu__time_duration only) of the bench command
#   ncu_<kernel>_<tag>_raw.csv     `ncu --set full` raw pages of the dominant kernels (the .ncu-rep files are deleted:
#                                  gpurun only copies back 64 MiB)
# Copy the digests you want judged into profiles/ (tools/ncu_digest.py).
TAG=${1:-r1}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
# 1. launch list of the bench command itself (short: 1 warm-up + 1 step, no CPU baseline)
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench_${TAG}.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
python tools/launch_summary.py gpurun_out/launches_bench_${TAG}.csv > gpurun_out/launches_bench_${TAG}.md
head -20 gpurun_out/launches_bench_${TAG}.md
# 2. full captures of the dominant kernels on the same workload
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
print(st["nTotalVoxels"], st["msTotal"], st["msVoxelize"], st["nKernelLaunches"])
for r in t.profile():
    if r["name"].startswith("dedup_leaf") or r["name"].startswith("dedup_k64"):
        print("PROFREC", r["name"], r["level"], r["n_in"], r["n_out"])
PY
cap() {  # name regex skip count
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -o gpurun_out/tmp_$1 python /tmp/ncu_city.py > gpurun_out/ncu_$1_${TAG}.log 2>&1
  ncu -i gpurun_out/tmp_$1.ncu-rep --page raw --csv > gpurun_out/ncu_$1_${TAG}_raw.csv 2>/dev/null
  rm -f gpurun_out/tmp_$1.ncu-rep
}
cap dedup "k_leaf_min|k_insert|k_convert|k_assign_k64|k_winner" 0 14
cap classify "k_classify_filtered" 28 4
cap flat "k_flat_leaves|k_classify_fast" 10 6
cap emit "k_emit|k_children" 40 6
du -sh gpurun_out
