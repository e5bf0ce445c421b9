#!/bin/bash
# End-of-round GPU call: full GPU test suite, the default bench line (+ reference arm), and the library's own per-launch
# records of the emit / children kernels (second argument `ncu`: under an ncu --set full capture of launches 41-43).
TAG=${1:-r1e}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_${TAG}.log
timeout 600 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
tail -c 600 gpurun_out/bench_${TAG}.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_${TAG}.json"))
print("value", d["value"], "ms", d["ms_per_step"], "e2e_s", d["e2e"]["seconds_per_step"], "roof", d["roofline"]["frac"], d["roofline"]["achieved"])
print({k: round(v["ms_per_step"], 1) for k, v in d["kernels"].items()})
print(d.get("dedup_effective"))
PY
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_${TAG}.json 2>> gpurun_out/bench_${TAG}.err
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
k = 0
for r in t.profile():
    if r["name"] in ("emit", "children"):
        if k < 60:
            print("PROFREC", k, r["name"], r["level"], r["n_in"], r["n_out"], r["bytes"], r["ms"])
        k += 1
PY
if [ "$2" = "ncu" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_emit|k_children" -s 40 -c 3 -o gpurun_out/tmp_emit python /tmp/ncu_city.py > gpurun_out/ncu_emit_${TAG}.log 2>&1
  ncu -i gpurun_out/tmp_emit.ncu-rep --page raw --csv > gpurun_out/ncu_emit_${TAG}_raw.csv 2>/dev/null
  rm -f gpurun_out/tmp_emit.ncu-rep
else
  timeout 300 python /tmp/ncu_city.py > gpurun_out/profrec_${TAG}.log 2>&1
fi
grep -h PROFREC gpurun_out/ncu_emit_${TAG}.log gpurun_out/profrec_${TAG}.log 2>/dev/null | sed -n 38,46p
