#!/bin/bash
# Round 2, call C (1 GPU): GPU tests (device encoders, bin-once root pairs, 4^3 level without first touches), bench, e2e breakdown
TAG=${1:-r2c}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_${TAG}.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_${TAG}_n1.json 2> gpurun_out/bench_${TAG}_n1.err
tail -c 600 gpurun_out/bench_${TAG}_n1.err
timeout 300 python tools/e2e_breakdown.py city_16k 5 2>&1 | tail -1 | tee gpurun_out/e2e_breakdown_${TAG}.json
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_${TAG}_n1.json").read().strip().splitlines()[-1])
print("value", round(d["value"], 1), "ms", round(d["ms_per_step"], 1), "e2e_ms", round(d["e2e"]["seconds_per_step"] * 1e3, 1), "parity", d["parity"]["ssvdag_sha256"][:16], d["parity"]["ok"])
print({k: round(v["ms_per_step"], 1) for k, v in d["kernels"].items()})
PY
SVB_VX_STATS=1 timeout 300 python tools/e2e_breakdown.py city_16k 1 2>&1 | grep "4^3 level" | tail -25
