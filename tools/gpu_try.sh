#!/bin/bash
# run an arbitrary bench workload once on the GPU box: gpurun -- tools/gpu_try.sh <workload> [extra bench args]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
W=${1:-composite_64k}; shift
timeout 1500 python bench.py --workload $W --steps 1 --warmup 0 --no-cpu-baseline --no-e2e "$@" > gpurun_out/try_$W.json 2> gpurun_out/try_$W.err
tail -c 1500 gpurun_out/try_$W.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/try_$W.json"))
    print({k: d.get(k) for k in ("value", "ms_per_step", "voxels", "triangles", "nodes", "tiles", "batches", "pairs", "error")})
    print({k: round(v["ms_per_step"], 1) for k, v in (d.get("kernels") or {}).items()})
except Exception as e:
    print("no json:", e, open("gpurun_out/try_$W.json").read()[-500:])
PY
