#!/bin/bash
# GPU box check: gpu tests + a short bench line (gpurun -- tools/gpu_check.sh)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_b.json"))
print("value", d["value"], "ms", d["ms_per_step"], "e2e_s", d["e2e"]["seconds_per_step"], "roof", d["roofline"]["frac"])
print({k: round(v["ms_per_step"], 1) for k, v in d["kernels"].items()}, "exact", d["exact_retests"])
PY
tail -3 gpurun_out/bench_b.err
