#!/bin/bash
# every bench workload once (short), with the parity block of each line
TAG=${1:-wl}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for W in spongeball_1k terrain_4k city_16k_c composite_64k; do
  timeout 900 python bench.py --workload $W --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_${TAG}_$W.json 2> gpurun_out/bench_${TAG}_$W.err
  tail -c 300 gpurun_out/bench_${TAG}_$W.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_${TAG}_$W.json").read().strip().splitlines()[-1])
    print("$W", "ms", round(d["ms_per_step"], 2), "e2e_ms", round(d["e2e"]["seconds_per_step"] * 1e3, 2), "Gvox/s", round(d["value"], 2), "batches", d["batches"], "parity", d["parity"].get("ok"), d["parity"].get("checks"), d["parity"].get("cross_merge"))
except Exception as e:
    print("$W failed", e)
PY
done
