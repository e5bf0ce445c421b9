#!/bin/bash
# Round 2, call A: GPU tests, the bench line, the e2e phase breakdown, and `ncu --set full` of the default-path dedup kernels
# of the SECOND tile batch (a "later" batch, like 18 of the 20) + the final classify instantiation.
TAG=${1:-r2a}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_${TAG}.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
tail -c 400 gpurun_out/bench_${TAG}.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_${TAG}.json"))
print("value", d["value"], "ms", d["ms_per_step"], "e2e_s", d["e2e"]["seconds_per_step"], "roof", d["roofline"]["frac"], "parity", d["parity"])
print({k: round(v["ms_per_step"], 1) for k, v in d["kernels"].items()})
PY
timeout 300 python tools/e2e_breakdown.py city_16k 5 2>&1 | tail -1 | tee gpurun_out/e2e_breakdown_${TAG}.json
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
    if r["name"].startswith("dedup"):
        print("PROFREC", r["name"], r["level"], r["n_in"], r["n_out"], r["ms"])
PY
cap() {  # name regex skip count
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -o gpurun_out/tmp_$1 python /tmp/ncu_city.py > gpurun_out/ncu_$1_${TAG}.log 2>&1
  ncu -i gpurun_out/tmp_$1.ncu-rep --page raw --csv > gpurun_out/ncu_$1_${TAG}_raw.csv 2>/dev/null
  rm -f gpurun_out/tmp_$1.ncu-rep
  tail -2 gpurun_out/ncu_$1_${TAG}.log
}
# batch 0: k_leaf_lazy, k_insert_k64, 7 x (k_insert, k_winner, k_convert) = 23 launches; capture batches 1 and 2
cap dedup "k_leaf_lazy|k_leaf_known|k_leaf_query|k_insert|k_winner|k_convert" 23 46
cap classify "k_classify_filtered" 28 4
du -sh gpurun_out
