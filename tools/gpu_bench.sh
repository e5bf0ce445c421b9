#!/bin/bash
# the driver's two bench arms at N=1, default flags: gpurun -- tools/gpu_bench.sh
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python bench.py --impl reference > gpurun_out/bench_ref_n1.json 2> gpurun_out/bench_ref_n1.err
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -n 2 gpurun_out/bench_n1.err
python - <<'PY'
import json
for f in ("bench_ref_n1", "bench_n1"):
    d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
    print(f, {k: d.get(k) for k in ("impl", "value", "unit", "ms_per_step", "steps", "warmup", "gpu_launches")})
    print("   e2e", d.get("e2e"), "\n   roofline", {k: d["roofline"][k] for k in ("achieved", "peak", "frac", "traffic")} if d.get("roofline") else None, "\n   cpu", d.get("cpu_baseline", {}).get("value"), d.get("clocks"))
PY
